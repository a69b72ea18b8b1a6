"""The AVX-512 / AVX2 ports of the reference's SIMD Solinas path (oracle/cntt_simd.c, the CPU baseline of bench.py) return
exactly what the scalar restatement returns -- the reference asserts the same of its own SIMD paths (src/prime64.rs:1211-1267
runs every available ISA against the schoolbook product)."""
import numpy as np
import pytest

from conftest import rng, rand_mod

P = 0xFFFFFFFF00000001


def isas(oracle):
    have = oracle.simd_isa()
    return [i for i in ("avx2", "avx512") if i == "avx2" or have == "avx512"] if have != "scalar" else []


@pytest.mark.parametrize("n", [16, 32, 64, 128, 1024, 2048, 4096, 8192, 32768])
def test_simd_equals_scalar(oracle, n):
    if not isas(oracle):
        pytest.skip("host has neither AVX2 nor AVX-512")
    g = rng(n)
    plan = oracle.Plan64.try_new(n, P)
    a = rand_mod(g, P, (3, n), np.uint64)
    a[0, :4] = [0, 1, P - 1, P - 2]                     # edge residues
    a[1, :] = P - 1
    ref_f = plan.fwd(a.copy())
    ref_i = plan.inv(ref_f.copy())
    for isa in isas(oracle):
        f = plan.fwd_simd(a.copy(), isa)
        assert (f == ref_f).all(), (isa, n, "fwd")
        assert (plan.inv_simd(f.copy(), isa) == ref_i).all(), (isa, n, "inv")


def test_batch_entry_points_use_simd_and_agree(oracle):
    n, batch = 2048, 64
    g = rng(5)
    plan = oracle.Plan64.try_new(n, P)
    a = rand_mod(g, P, (batch, n), np.uint64)
    ref = plan.fwd(a.copy())
    for isa in ["scalar"] + isas(oracle) + ["best"]:
        oracle.set_batch_isa(isa)
        b = a.copy()
        plan.fwd_batch(b, 4)
        assert (b == ref).all(), isa
        plan.inv_batch(b, 4)
        assert (b == plan.inv(ref.copy())).all(), isa
    oracle.set_batch_isa("best")


def test_other_classes_fall_back_to_scalar(oracle):
    f = oracle.largest_prime_in_arithmetic_progression64
    p = f(1 << 16, 1, 1 << 61, 1 << 62)
    plan = oracle.Plan64.try_new(64, p)
    a = rand_mod(rng(1), p, (2, 64), np.uint64)
    assert (plan.fwd_simd(a.copy()) == plan.fwd(a.copy())).all()


@pytest.mark.parametrize("n", [32, 64, 128, 512, 2048, 4096, 16384])
def test_simd32_equals_scalar(oracle, n):
    """oracle/cntt_simd32.c: the 16- and 8-lane Shoup paths of prime32 (p < 2^30 and p < 2^31) against the scalar restatement;
    p >= 2^31 and n < 64 fall back to the scalar path and must agree trivially."""
    if not isas(oracle):
        pytest.skip("host has neither AVX2 nor AVX-512")
    f = oracle.largest_prime_in_arithmetic_progression64
    primes = [1062862849, f(1 << 16, 1, 1 << 29, 1 << 30), f(1 << 16, 1, 1 << 30, 1 << 31), f(1 << 16, 1, 1 << 31, 1 << 32)]
    g = rng(n + 32)
    for p in primes:
        plan = oracle.Plan32.try_new(n, p)
        assert plan is not None
        a = rand_mod(g, p, (3, n), np.uint32)
        a[0, :4] = [0, 1, p - 1, p - 2]
        a[1, :] = p - 1
        ref_f = plan.fwd(a.copy())
        ref_i = plan.inv(ref_f.copy())
        for isa in isas(oracle):
            fv = plan.fwd_simd(a.copy(), isa)
            assert (fv == ref_f).all(), (isa, n, p, "fwd")
            assert (plan.inv_simd(fv.copy(), isa) == ref_i).all(), (isa, n, p, "inv")


def test_batch32_entry_points_agree(oracle):
    n, batch, p = 1024, 32, 1062862849
    plan = oracle.Plan32.try_new(n, p)
    a = rand_mod(rng(9), p, (batch, n), np.uint32)
    ref = plan.fwd(a.copy())
    for isa in ["scalar"] + isas(oracle) + ["best"]:
        oracle.set_batch_isa(isa)
        b = a.copy()
        plan.fwd_batch(b, 4)
        assert (b == ref).all(), isa
        plan.inv_batch(b, 4)
        assert (b == plan.inv(ref.copy())).all(), isa
    oracle.set_batch_isa("best")
