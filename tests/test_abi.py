"""CPU-only checks of the boundary: the C-ABI library loads and exports exactly what include/cntt_b200.h
declares, host-only entry points work without a GPU, and compute entry points fail loudly (no fallback)."""
import ctypes
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cntt_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cntt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(cntt):
    lib_path = cntt._lib.LIB_PATH
    assert os.path.exists(lib_path), "build the library first (__graft_entry__.build())"
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    decl = declared_symbols()
    assert len(decl) > 40
    missing = [s for s in decl if s not in exported]
    assert not missing, missing
    # nothing but the declared ABI leaks out of the library
    extra = sorted(s for s in exported if s.startswith("cntt_") and s not in decl)
    assert not extra, extra
    # and the ctypes table covers every declared symbol
    assert sorted(cntt._lib.SIGNATURES) == decl


def test_host_only_entry_points(cntt, oracle):
    assert "sm_100a" in cntt.version()
    for n in list(range(0, 200)) + [1062862849, 1062862851, 0xFFFFFFFF00000001, 2**64 - 59, 2**64 - 1]:
        assert cntt.prime.is_prime64(n) == oracle.is_prime64(n)
    f, g = cntt.prime.largest_prime_in_arithmetic_progression64, oracle.largest_prime_in_arithmetic_progression64
    for args in [(0, 2, 1, 4), (0, 2, 2, 1), (1, 0, 14, 16), (2, 1, 14, 17), (6, 5, 0, 2**64 - 1), (6, 1, 0, 2**64 - 1),
                 (1 << 16, 1, 1 << 29, 1 << 30), (1 << 16, 1, 1 << 63, 2**64 - 1)]:
        assert f(*args) == g(*args), args
    # the product's own root finder must pick the reference's root (it decides every fwd output)
    for p, deg in [(1062862849, 64), (1062862849, 2048), (1062862849, 1 << 17), (1063059457, 1 << 16),
                   (0xFFFFFFFF00000001, 4096), (0xFFFFFFFF00000001, 1 << 20), (1073479681, 2048), (4293918721, 2048)]:
        assert cntt.roots.find_primitive_root64(p, deg) == oracle.find_primitive_root64(p, deg), (p, deg)
    assert cntt.roots.find_primitive_root64(1063059457, 1 << 18) is None


def test_validation_precedes_cuda(cntt):
    """try_new's None / panic outcomes are decided on the host before any CUDA call."""
    assert cntt.prime32.Plan.try_new(16, 1062862849) is None
    assert cntt.prime32.Plan.try_new(1000, 1062862849) is None
    assert cntt.prime32.Plan.try_new(32, 15) is None
    assert cntt.prime64.Plan.try_new(2048, 1024) is None
    assert cntt.native64.Plan32.try_new(65536) is None
    with pytest.raises(cntt.ReferencePanic):
        cntt.prime64.Plan.try_new(64, 1)


def test_no_cpu_fallback(cntt):
    """Without a usable CUDA device a valid plan request must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on the CPU box")
    with pytest.raises(cntt.CnttError) as e:
        cntt.prime32.Plan.try_new(1024, 1062862849)
    assert "CUDA" in str(e.value)


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under concrete-ntt_b200/ may reference it."""
    pkg = os.path.join(ROOT, "concrete-ntt_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(d, f)
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libcntt_b200.so")], text=True)
    assert "oracle" not in out


def test_fastdiv_and_prime_helpers(cntt, oracle):
    """fastdiv::{Div32,Div64} and prime::{mul,exp}_mod (src/fastdiv.rs:152-208 tests: div/rem agree with `/`, `%`)."""
    g = np.random.Generator(np.random.PCG64(11))
    D32, D64 = cntt.fastdiv.Div32, cntt.fastdiv.Div64
    for _ in range(200):
        d = int(g.integers(2, 2**32))
        n = int(g.integers(0, 2**64, dtype=np.uint64))
        dv = D32.new(d)
        assert dv.divisor() == d
        assert D32.div(n & 0xFFFFFFFF, dv) == (n & 0xFFFFFFFF) // d and D32.rem(n & 0xFFFFFFFF, dv) == (n & 0xFFFFFFFF) % d
        assert D32.div_u64(n, dv) == n // d and D32.rem_u64(n, dv) == n % d
        d = int(g.integers(2, 2**64, dtype=np.uint64))
        big = n * int(g.integers(0, 2**64, dtype=np.uint64))
        dv = D64.new(d)
        assert D64.div(n, dv) == n // d and D64.rem(n, dv) == n % d
        assert D64.div_u128(big, dv) == big // d and D64.rem_u128(big, dv) == big % d
    for bad in (0, 1):
        with pytest.raises(cntt.ReferencePanic):
            D32.new(bad)
        with pytest.raises(cntt.ReferencePanic):
            D64.new(bad)
    p = 0xFFFFFFFF00000001
    assert cntt.prime.exp_mod64(D64.new(p), 7, p - 1) == 1 == oracle.exp_mod64(p, 7, p - 1)
    assert cntt.prime.mul_mod64(D64.new(p), p - 1, p - 1) == 1
    assert cntt.prime.exp_mod32(D32.new(1062862849), 5, 1062862848) == 1
    assert cntt.prime.mul_mod32(1062862849, 1062862848, 2) == 1062862847


def test_rust_ffi_declares_the_whole_header():
    """rust/src/ffi.rs cannot be compiled here (no toolchain); at least keep its extern block in step with the header:
    same symbol set, same argument counts."""
    src = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    rust = dict((m.group(1), m.group(2)) for m in re.finditer(r"pub fn (cntt_[a-z0-9_]+)\(([^)]*)\)", src))
    hdr_src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    hdr = dict((m.group(1), m.group(2)) for m in re.finditer(r"\b(cntt_[a-z0-9_]+)\s*\(([^)]*)\)", hdr_src))
    assert sorted(rust) == sorted(hdr), sorted(set(rust) ^ set(hdr))
    for name, args in hdr.items():
        n_c = 0 if args.strip() in ("", "void") else len(args.split(","))
        n_r = 0 if not rust[name].strip() else len(rust[name].split(","))
        assert n_c == n_r, (name, args, rust[name])
