"""Small driver for ncu captures: launches one hot-path kernel a few times on a device-resident batch.
usage: python tools/prof_driver.py {ntt64|ntt64shoup|ntt32|pointwise32|pointwise64|polymul64|polymul128|polymulb64|polymul32|split64} [batch] [n]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cntt = importlib.import_module("concrete-ntt_b200")

which = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
g = torch.Generator(device="cuda").manual_seed(1)
reps = 4
if which == "ntt64":
    plan = cntt.prime64.Plan.try_new(n, cntt.prime64.Solinas.P)
    d = torch.randint(0, 2**62, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    for _ in range(reps):
        plan.fwd(d)
        plan.inv(d)
elif which == "ntt64shoup":
    p = cntt.prime.largest_prime_in_arithmetic_progression64(1 << 17, 1, 1 << 61, 1 << 62)
    plan = cntt.prime64.Plan.try_new(n, p)
    d = torch.randint(0, 2**61, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    for _ in range(reps):
        plan.fwd(d)
        plan.inv(d)
elif which in ("pointwise32", "pointwise64"):
    if which == "pointwise32":
        plan = cntt.prime32.Plan.try_new(n, 1062862849)
        mk = lambda: torch.randint(0, 1062862849, (batch, n), dtype=torch.int32, device="cuda", generator=g)
    else:
        plan = cntt.prime64.Plan.try_new(n, cntt.prime64.Solinas.P)
        mk = lambda: torch.randint(0, 2**62, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    a, b, c = mk(), mk(), mk()
    for _ in range(reps):
        plan.mul_assign_normalize(a, b)
        plan.normalize(a)
        plan.mul_accumulate(c, a, b)
elif which == "ntt32":
    plan = cntt.prime32.Plan.try_new(n, 1062862849)
    d = torch.randint(0, 1062862849, (batch, n), dtype=torch.int32, device="cuda", generator=g)
    for _ in range(reps):
        plan.fwd(d)
        plan.inv(d)
elif which in ("polymul64", "polymulb64", "polymul32"):
    mod = {"polymul64": cntt.native64, "polymulb64": cntt.native_binary64, "polymul32": cntt.native32}[which]
    plan = mod.Plan32.try_new(n) or mod.Plan32.try_new_extended(n)
    dt = torch.int32 if which == "polymul32" else torch.int64
    hi = 2**31 - 1 if which == "polymul32" else 2**63 - 1
    lhs = torch.randint(-hi - 1, hi, (batch, n), dtype=dt, device="cuda", generator=g)
    rhs = torch.randint(-hi - 1, hi, (batch, n), dtype=dt, device="cuda", generator=g)
    if which == "polymulb64":
        rhs &= 1
    prod = torch.empty_like(lhs)
    for _ in range(reps):
        plan.negacyclic_polymul(prod, lhs, rhs)
elif which == "split64":   # Plan32::fwd / inv on device residue planes
    plan = cntt.native64.Plan32.try_new(n)
    val = torch.randint(-2**63, 2**63 - 1, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    planes = torch.empty((plan.num_primes(), batch, n), dtype=torch.int32, device="cuda")
    for _ in range(reps):
        plan.fwd(val, planes)
elif which == "split64inv":   # Plan32::inv on device residue planes: nprimes inverse transforms + k_native_crt; odd sizes use k_native_reduce
    plan = cntt.native64.Plan32.try_new(n)
    val = torch.randint(-2**63, 2**63 - 1, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    planes = torch.empty((plan.num_primes(), batch, n), dtype=torch.int32, device="cuda")
    small = cntt.native64.Plan32.try_new(128)     # N < 256: unfused k_native_reduce
    sval = torch.randint(-2**63, 2**63 - 1, (batch * n // 128, 128), dtype=torch.int64, device="cuda", generator=g)
    splanes = torch.empty((5, batch * n // 128, 128), dtype=torch.int32, device="cuda")
    for _ in range(reps):
        plan.fwd(val, planes)
        plan.inv(val, planes)
        small.fwd(sval, splanes)
elif which == "plan52":       # Plan52 twins: k_native52_reduce + prime64 transforms + k_native52_crt
    plan = cntt.native64.Plan52.try_new(n)
    val = torch.randint(-2**63, 2**63 - 1, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    planes = torch.empty((plan.num_primes(), batch, n), dtype=torch.int64, device="cuda")
    for _ in range(reps):
        plan.fwd(val, planes)
        plan.inv(val, planes)
elif which in ("product", "product_generic"):
    f = cntt.prime.largest_prime_in_arithmetic_progression64
    if which == "product":    # two u32 primes of one class: the fused kernels
        p1 = f(1 << 17, 1, 1 << 30, 1 << 31); p2 = f(1 << 17, 1, 1 << 30, p1 - 1)
    else:                     # a u32 and a u64 prime: k_product_reduce / k_product_crt around the prime kernels
        p1 = f(1 << 17, 1, 1 << 20, 1 << 21); p2 = f(1 << 17, 1, 1 << 41, 1 << 42)
    plan = cntt.product.Plan.try_new(n, p1 * p2, [p1, p2])
    std = torch.randint(0, p1 * p2, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    a = torch.empty((batch, plan.ntt_domain_len()), dtype=torch.int64, device="cuda")
    acc = torch.zeros_like(a)
    for _ in range(reps):
        plan.fwd(a, std)
        plan.mul_accumulate(acc, a, a)
        plan.inv(std, a)
elif which == "polymul128":
    plan = cntt.native128.Plan32.try_new(n)
    lhs = torch.randint(-2**63, 2**63 - 1, (batch, n, 2), dtype=torch.int64, device="cuda", generator=g)
    rhs = torch.randint(-2**63, 2**63 - 1, (batch, n, 2), dtype=torch.int64, device="cuda", generator=g)
    prod = torch.empty_like(lhs)
    for _ in range(reps):
        plan.negacyclic_polymul(prod, lhs, rhs)
torch.cuda.synchronize()
print("done", which, batch, n)
