#!/usr/bin/env python3
"""What oracle/ref_fixtures/src/main.rs must print, computed with the CPU oracle instead of the reference crate.
TEST INFRASTRUCTURE.  `python oracle/ref_fixtures/predict.py > tests/golden/ref_v1_predicted.jsonl`.
The day the Rust generator is run (`make -C oracle ref-fixtures`, needs cargo), `diff tests/golden/ref_v1.jsonl
tests/golden/ref_v1_predicted.jsonl` must be empty: that pins the oracle -- and through it the CUDA library -- against
outputs of the reference itself.  `records(backend)` is also what tests/test_ref_fixtures.py drives with the CUDA library."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M64 = (1 << 64) - 1


class Rng:
    """splitmix64, as in main.rs"""

    def __init__(self, seed):
        self.s = seed & M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)


def fnv(words):
    h = 0xcbf29ce484222325
    for w in words:
        h ^= int(w)
        h = (h * 0x100000001b3) & M64
    return h


def field(rec, name, words, full):
    words = [int(w) for w in words]
    rec[name + "_hash"] = str(fnv(words))
    rec[name + "_head"] = [str(w) for w in words[:8]]
    if full:
        rec[name + "_full"] = [str(w) for w in words]


def widen(a):
    """u32 / u64 arrays -> list of ints; 128-bit words (n, 2) -> lo, hi, lo, hi ..."""
    return [int(v) for v in np.asarray(a).reshape(-1)]


class OracleBackend:
    """the CPU oracle (oracle/oracle.py)"""

    def __init__(self):
        from oracle import oracle as O
        O.build()
        self.O = O

    def lpap(self, *a):
        return self.O.largest_prime_in_arithmetic_progression64(*a)

    def prime(self, bits, n, p, a, b, c):
        O = self.O
        plan = (O.Plan32 if bits == 32 else O.Plan64).try_new(n, p)
        if plan is None:
            return None
        f = plan.fwd(a.copy())
        return {"fwd": f, "inv": plan.inv(f.copy()), "mul_assign_normalize": plan.mul_assign_normalize(a.copy(), b),
                "normalize": plan.normalize(a.copy()), "mul_accumulate": plan.mul_accumulate(c.copy(), a, b)}

    def polymul(self, bits, binary, n, lhs, rhs):
        plan = self.O.Native.try_new(n, bits, binary)
        return None if plan is None else plan.negacyclic_polymul(lhs, rhs)

    def native64_split(self, n, value):
        plan = self.O.Native.try_new(n, 64)
        planes = plan.fwd(value)
        return planes.copy(), plan.inv(planes)

    def product(self, n, p0, p1, std):
        O = self.O
        plan = O.Product.try_new(n, p0 * p1, [p0, p1])
        ntt = np.zeros(plan.ntt_domain_len(), np.uint64)
        plan.fwd(ntt, std, plan.GENERIC)
        dom = ntt.copy()
        back = np.zeros(n, np.uint64)
        plan.inv(back, ntt, plan.REPLACE)
        return dom, back


def records(B):
    """the record stream of main.rs, computed through backend B"""
    yield {"kind": "header", "crate": "concrete-ntt", "version": "0.2.0", "format": 1}
    p32 = [1062862849, B.lpap(1 << 16, 1, 1 << 29, 1 << 30), B.lpap(1 << 16, 1, 1 << 30, 1 << 31), B.lpap(1 << 16, 1, 1 << 31, 1 << 32)]
    p64 = [B.lpap(1 << 16, 1, 1 << 49, 1 << 50), B.lpap(1 << 16, 1, 1 << 50, 1 << 51), B.lpap(1 << 16, 1, 1 << 61, 1 << 62),
           B.lpap(1 << 16, 1, 1 << 62, 1 << 63), 0xFFFFFFFF00000001, B.lpap(1 << 16, 1, 1 << 63, M64)]

    def prime_case(bits, n, p, seed):
        r = Rng(seed)
        dt = np.uint32 if bits == 32 else np.uint64
        a, b, c = (np.array([r.next() % p for _ in range(n)], dtype=dt) for _ in range(3))
        out = B.prime(bits, n, p, a, b, c)
        rec = {"kind": "prime%d" % bits, "n": n, "p": str(p)}
        if out is None:
            rec["none"] = True
            return rec
        rec["seed"] = str(seed)
        for name in ("fwd", "inv", "mul_assign_normalize", "normalize", "mul_accumulate"):
            field(rec, name, widen(out[name]), n == 32 and name in ("fwd", "inv"))
        return rec

    seed = 0xC0FFEE
    for n in (32, 1024, 4096, 32768):
        for p in p32:
            seed += 1
            yield prime_case(32, n, p, seed)
        for p in p64:
            seed += 1
            yield prime_case(64, n, p, seed)
    yield prime_case(64, 16, 0xFFFFFFFF00000001, 0xBEEF)

    def polymul_case(kind, bits, n, seed, binary):
        r = Rng(seed)

        def word(bin_):
            if bits == 128:
                lo = r.next()
                hi = r.next()
                return (lo & 1, 0) if bin_ else (lo, hi)
            v = r.next()
            v = v & 1 if bin_ else v
            return v & 0xFFFFFFFF if bits == 32 else v
        dt = np.uint32 if bits == 32 else np.uint64
        lhs = np.array([word(False) for _ in range(n)], dtype=dt)
        rhs = np.array([word(binary) for _ in range(n)], dtype=dt)
        prod = B.polymul(bits, binary, n, lhs, rhs)
        rec = {"kind": kind, "n": n, "bits": bits}
        if prod is None:
            rec["none"] = True
            return rec
        rec["seed"] = str(seed)
        field(rec, "prod", widen(prod), n == 32)
        return rec

    for n in (32, 1024):
        for kind, binary in (("native", False), ("native_binary", True)):
            for bits in (32, 64, 128):
                seed += 1
                yield polymul_case(kind, bits, n, seed, binary)

    n, s = 2048, 0xABCD
    r = Rng(s)
    value = np.array([r.next() for _ in range(n)], dtype=np.uint64)
    planes, back = B.native64_split(n, value)
    rec = {"kind": "native64_split", "n": n, "bits": 64, "seed": str(s)}
    field(rec, "planes", widen(planes), False)
    field(rec, "inv", widen(back), False)
    yield rec

    for n, s in ((1024, 0x1234), (2048, 0x1235)):
        p0 = B.lpap(2 * n, 1, 0, 1 << 31)
        p1 = B.lpap(2 * n, 1, 0, p0 - 1)
        r = Rng(s)
        std = np.array([r.next() % (p0 * p1) for _ in range(n)], dtype=np.uint64)
        dom, back = B.product(n, p0, p1, std)
        rec = {"kind": "product", "n": n, "p0": str(p0), "p1": str(p1), "seed": str(s)}
        field(rec, "fwd", widen(dom), False)
        field(rec, "inv", widen(back), False)
        yield rec


def dumps(rec):
    """the exact text main.rs prints (key order and spacing included)"""
    return json.dumps(rec, separators=(",", ":"))


if __name__ == "__main__":
    for rec in records(OracleBackend()):
        print(dumps(rec))
