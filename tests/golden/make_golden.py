"""Generates tests/golden/golden_v1.npz from the CPU oracle (seeded).  The reference holds no golden vectors
for this path and cannot be executed here (Rust, no toolchain), so these vectors freeze the *pinned oracle*
(tests/test_oracle.py) against regressions and give the GPU tests fixed byte-level targets.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from conftest import rng, rand_mod, rand_words, primes32, primes64  # noqa: E402


def main():
    out = {}
    g = rng(0xC0FFEE)
    for name, p in primes32(O).items():
        for n in (32, 256):
            plan = O.Plan32.try_new(n, p)
            a = rand_mod(g, p, (2, n), np.uint32)
            b = rand_mod(g, p, (2, n), np.uint32)
            key = "p32_%s_%d" % (name, n)
            out[key + "_p"] = np.array([p], np.uint64)
            out[key + "_in"] = a
            out[key + "_in2"] = b
            f = plan.fwd(a.copy())
            out[key + "_fwd"] = f
            out[key + "_inv"] = plan.inv(f.copy())
            out[key + "_man"] = plan.mul_assign_normalize(a.copy(), b)
            out[key + "_nrm"] = plan.normalize(a.copy())
            out[key + "_mac"] = plan.mul_accumulate(f.copy(), a, b)
    for name, p in primes64(O).items():
        for n in (16, 256):
            plan = O.Plan64.try_new(n, p)
            a = rand_mod(g, p, (2, n), np.uint64)
            b = rand_mod(g, p, (2, n), np.uint64)
            key = "p64_%s_%d" % (name, n)
            out[key + "_p"] = np.array([p], np.uint64)
            out[key + "_in"] = a
            out[key + "_in2"] = b
            f = plan.fwd(a.copy())
            out[key + "_fwd"] = f
            out[key + "_inv"] = plan.inv(f.copy())
            out[key + "_man"] = plan.mul_assign_normalize(a.copy(), b)
            out[key + "_nrm"] = plan.normalize(a.copy())
            out[key + "_mac"] = plan.mul_accumulate(f.copy(), a, b)
    for bits in (32, 64, 128):
        for binary in (False, True):
            n = 64
            plan = O.Native.try_new(n, bits, binary=binary)
            lhs = rand_words(g, bits, (2, n))
            rhs = rand_words(g, bits, (2, n))
            if binary:
                rhs = rhs & rhs.dtype.type(1)
                if bits == 128:
                    rhs[..., 1] = 0
            key = "nat%d%s_%d" % (bits, "b" if binary else "", n)
            out[key + "_lhs"] = lhs
            out[key + "_rhs"] = rhs
            out[key + "_prod"] = plan.negacyclic_polymul(lhs, rhs)
            out[key + "_planes"] = np.stack([plan.fwd(v) for v in lhs], axis=1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
