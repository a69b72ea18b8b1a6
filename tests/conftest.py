import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/cntt_oracle.c) -- the checker, never the thing under test."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def cntt():
    """The product package; fails loudly if libcntt_b200.so has not been built."""
    return importlib.import_module("concrete-ntt_b200")


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def rand_mod(g, p, shape, dtype):
    """uniform in [0, p)"""
    if p <= 2**63:
        return g.integers(0, p, size=shape, dtype=np.uint64).astype(dtype)
    # p > 2^63: rejection on full 64-bit draws
    out = g.integers(0, 2**64, size=shape, dtype=np.uint64)
    bad = out >= np.uint64(p)
    while bad.any():
        out[bad] = g.integers(0, 2**64, size=int(bad.sum()), dtype=np.uint64)
        bad = out >= np.uint64(p)
    return out.astype(dtype)


def rand_words(g, bits, shape):
    if bits == 32:
        return g.integers(0, 2**32, size=shape, dtype=np.uint64).astype(np.uint32)
    if bits == 64:
        return g.integers(0, 2**64, size=shape, dtype=np.uint64)
    return g.integers(0, 2**64, size=tuple(shape) + (2,), dtype=np.uint64)


# one prime per bit class of the reference's dispatch (prime32.rs:713-754, prime64.rs:812-864),
# chosen the way the reference's tests and benches do (largest k*2^16+1 / k*2^?+1 prime in a window)
PRIMES32 = {
    "lt30": 1062862849,           # primes32::P0
    "lt31": None,                 # filled by primes32()/primes64() below
    "ge31": None,
}


def primes32(O):
    f = O.largest_prime_in_arithmetic_progression64
    return {
        "lt30_P0": 1062862849,
        "lt30": f(1 << 16, 1, 1 << 29, 1 << 30),   # benches/ntt.rs:89-91
        "lt31": f(1 << 16, 1, 1 << 30, 1 << 31),
        "ge31": f(1 << 16, 1, 1 << 31, 1 << 32),
    }


def primes64(O):
    f = O.largest_prime_in_arithmetic_progression64
    return {
        "lt50": f(1 << 16, 1, 1 << 49, 1 << 50),   # benches/ntt.rs:115-122
        "lt51": f(1 << 16, 1, 1 << 50, 1 << 51),
        "lt62": f(1 << 16, 1, 1 << 61, 1 << 62),
        "lt63": f(1 << 16, 1, 1 << 62, 1 << 63),
        "solinas": 0xFFFFFFFF00000001,
        "ge63": f(1 << 16, 1, 1 << 63, (1 << 64) - 1),
    }
