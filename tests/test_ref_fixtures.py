"""Parity against outputs of the reference crate itself (tests/golden/ref_v1.jsonl, written by `make -C oracle ref-fixtures`
on a machine with cargo).  The image has no Rust toolchain, so the file may be absent: then the CPU test SKIPS with the word
"unpinned", and the remaining checks run against tests/golden/ref_v1_predicted.jsonl -- the same records computed by the CPU
oracle (oracle/ref_fixtures/predict.py), committed so that a later `diff` against the real thing is a one-liner."""
import importlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "ref_fixtures"))
import predict  # noqa: E402

REF = os.path.join(HERE, "golden", "ref_v1.jsonl")
PRED = os.path.join(HERE, "golden", "ref_v1_predicted.jsonl")


def load(path):
    return [json.loads(line) for line in open(path) if line.strip()]


def test_predicted_file_is_current():
    """the committed prediction is what the oracle computes today (freezes the oracle against regressions)"""
    want = [predict.dumps(r) for r in predict.records(predict.OracleBackend())]
    have = [line.rstrip("\n") for line in open(PRED)]
    assert have == want


def test_reference_fixtures_match_oracle():
    if not os.path.exists(REF):
        pytest.skip("parity UNPINNED against the reference binary: tests/golden/ref_v1.jsonl absent (no cargo in this image; "
                    "run `make -C oracle ref-fixtures` where a Rust toolchain exists)")
    assert load(REF) == load(PRED), "the reference crate and the oracle disagree"


class GpuBackend(predict.OracleBackend):
    """the same record stream computed by the CUDA library through the C ABI (plan-time helpers included)"""

    def __init__(self):
        import torch
        self.torch = torch
        self.cntt = importlib.import_module("concrete-ntt_b200")

    def lpap(self, *a):
        return self.cntt.prime.largest_prime_in_arithmetic_progression64(*a)

    def dev(self, a):
        sd = np.int32 if a.dtype.itemsize == 4 else np.int64
        return self.torch.from_numpy(np.ascontiguousarray(a).view(sd).copy()).cuda()

    def host(self, t, dt):
        return t.cpu().numpy().view(dt)

    def prime(self, bits, n, p, a, b, c):
        mod = self.cntt.prime32 if bits == 32 else self.cntt.prime64
        plan = mod.Plan.try_new(n, p)
        if plan is None:
            return None
        dt = a.dtype
        d = self.dev(a)
        plan.fwd(d)
        out = {"fwd": self.host(d, dt).copy()}
        plan.inv(d)
        out["inv"] = self.host(d, dt).copy()
        da, db, dc = self.dev(a), self.dev(b), self.dev(c)
        plan.mul_assign_normalize(da, db)
        out["mul_assign_normalize"] = self.host(da, dt).copy()
        da = self.dev(a)
        plan.normalize(da)
        out["normalize"] = self.host(da, dt).copy()
        plan.mul_accumulate(dc, self.dev(a), db)
        out["mul_accumulate"] = self.host(dc, dt).copy()
        return out

    def polymul(self, bits, binary, n, lhs, rhs):
        mod = getattr(self.cntt, ("native_binary%d" if binary else "native%d") % bits)
        plan = mod.Plan32.try_new(n)
        if plan is None:
            return None
        dl, dr = self.dev(lhs), self.dev(rhs)
        dp = self.torch.empty_like(dl)
        plan.negacyclic_polymul(dp, dl, dr)
        return self.host(dp, lhs.dtype)

    def native64_split(self, n, value):
        plan = self.cntt.native64.Plan32.try_new(n)
        planes = self.torch.empty((plan.num_primes(), n), dtype=self.torch.int32, device="cuda")
        dv = self.dev(value)
        plan.fwd(dv, planes)
        fwd = self.host(planes, np.uint32).copy()
        out = self.torch.empty_like(dv)
        plan.inv(out, planes)
        return fwd, self.host(out, np.uint64)

    def product(self, n, p0, p1, std):
        plan = self.cntt.product.Plan.try_new(n, p0 * p1, [p0, p1])
        dom = self.torch.zeros(plan.ntt_domain_len(), dtype=self.torch.int64, device="cuda")
        ds = self.dev(std)
        plan.fwd(dom, ds)
        fwd = self.host(dom, np.uint64).copy()
        back = self.torch.zeros(n, dtype=self.torch.int64, device="cuda")
        plan.inv(back, dom)
        return fwd, self.host(back, np.uint64)


@pytest.mark.gpu
def test_cuda_library_reproduces_the_fixtures():
    """every record, recomputed by the CUDA library, equals the reference crate's (when ref_v1.jsonl exists) or the oracle's"""
    want = load(REF) if os.path.exists(REF) else load(PRED)
    got = [json.loads(predict.dumps(r)) for r in predict.records(GpuBackend())]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g == w, (w.get("kind"), w.get("n"), w.get("p", w.get("bits")))
