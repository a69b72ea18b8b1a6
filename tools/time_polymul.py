"""Times device-resident polymul / NTT launches with CUDA events (quick A/B experiments)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cntt = importlib.import_module("concrete-ntt_b200")
def bench(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
g = torch.Generator(device="cuda").manual_seed(1)
cases = sys.argv[1:] or ["native64:2048:65536"]
for c in cases:
    kind, n, batch = c.split(":"); n = int(n); batch = int(batch)
    if kind == "product":   # product::Plan, modulus = two primes just below 2^31 (the tfhe-rs NTT-PBS shape)
        f = cntt.prime.largest_prime_in_arithmetic_progression64
        p1 = f(1 << 17, 1, 1 << 30, 1 << 31); p2 = f(1 << 17, 1, 1 << 30, p1 - 1)
        plan = cntt.product.Plan.try_new(n, p1 * p2, [p1, p2])
        std = torch.randint(0, p1 * p2, (batch, n), dtype=torch.int64, device="cuda", generator=g)
        a = torch.empty((batch, plan.ntt_domain_len()), dtype=torch.int64, device="cuda"); b = torch.empty_like(a); acc = torch.zeros_like(a)
        plan.fwd(b, std)
        tf = bench(lambda: plan.fwd(a, std)); tm = bench(lambda: plan.mul_accumulate(acc, a, b))
        ti = bench(lambda: plan.inv(std, a))
        gb = lambda t, words: words * 8 * batch / t / 1e6
        print("product(2 x 31-bit primes) n=%d batch=%d: fwd %.3f ms (%.1f M/s)  mul_accumulate %.3f ms (%.0f GB/s)  inv %.3f ms (%.1f M/s)"
              % (n, batch, tf, batch / tf / 1e3, tm, gb(tm, 4 * plan.ntt_domain_len()), ti, batch / ti / 1e3))
        continue
    if kind.startswith("split"):   # Plan32::fwd / inv on device residue planes (the NTT-domain-key workflow)
        bits = int(kind[5:])
        plan = getattr(cntt, "native%d" % bits).Plan32.try_new(n)
        dt = torch.int32 if bits == 32 else torch.int64
        hi = 2**31 - 1 if bits == 32 else 2**63 - 1
        shape = (batch, n, 2) if bits == 128 else (batch, n)
        val = torch.randint(-hi - 1, hi, shape, dtype=dt, device="cuda", generator=g)
        planes = torch.empty((plan.num_primes(), batch, n), dtype=torch.int32, device="cuda")
        tf = bench(lambda: plan.fwd(val, planes))
        ti = bench(lambda: plan.inv(val, planes))
        print("native%d split n=%d batch=%d: fwd %.3f ms (%.1f M/s)  inv %.3f ms (%.1f M/s)" % (bits, n, batch, tf, batch / tf / 1e3, ti, batch / ti / 1e3))
        continue
    if kind.startswith("pre"):      # polymul with the rhs already transformed (cntt_native_polymul_ntt_rhs), one key per product
        binary = kind.startswith("preb")   # preb64: native_binary64, the {0,1} operand is the pre-transformed key
        bits = int(kind[4:] if binary else kind[3:])
        plan = getattr(cntt, ("native_binary%d" if binary else "native%d") % bits).Plan32.try_new(n)
        dt = torch.int32 if bits == 32 else torch.int64
        hi = 2**31 - 1 if bits == 32 else 2**63 - 1
        shape = (batch, n, 2) if bits == 128 else (batch, n)
        lhs = torch.randint(-hi - 1, hi, shape, dtype=dt, device="cuda", generator=g)
        rhs = torch.randint(-hi - 1, hi, shape, dtype=dt, device="cuda", generator=g)
        planes = torch.empty((plan.num_primes(), batch, n), dtype=torch.int32, device="cuda")
        if binary:
            rhs &= 1
            if bits == 128: rhs[..., 1] = 0
            plan.fwd_binary(rhs, planes)
        else:
            plan.fwd(rhs, planes)
        key = planes[:, :1].contiguous()
        prod = torch.empty_like(lhs)
        ms = bench(lambda: plan.negacyclic_polymul_ntt_rhs(prod, lhs, planes))
        ms1 = bench(lambda: plan.negacyclic_polymul_ntt_rhs(prod, lhs, key))
        print("%s%d polymul, rhs pre-transformed n=%d batch=%d: %.3f ms  %.2f M polymul/s (one key per product)   %.3f ms  %.2f M polymul/s (shared key)"
              % ("native_binary" if binary else "native", bits, n, batch, ms, batch / ms / 1e3, ms1, batch / ms1 / 1e3))
        continue
    if kind.startswith("native") or kind.startswith("binary"):
        bits = int(kind.replace("native", "").replace("binary", ""))
        mod = getattr(cntt, ("native_binary%d" if kind.startswith("binary") else "native%d") % bits)
        plan = mod.Plan32.try_new(n) or mod.Plan32.try_new_extended(n)
        shape = (batch, n, 2) if bits == 128 else (batch, n)
        dt = torch.int32 if bits == 32 else torch.int64
        hi = 2**31 - 1 if bits == 32 else 2**63 - 1
        lhs = torch.randint(-hi - 1, hi, shape, dtype=dt, device="cuda", generator=g)
        rhs = torch.randint(-hi - 1, hi, shape, dtype=dt, device="cuda", generator=g)
        if kind.startswith("binary"):
            rhs &= 1
            if bits == 128: rhs[..., 1] = 0
        prod = torch.empty_like(lhs)
        ms = bench(lambda: plan.negacyclic_polymul(prod, lhs, rhs))
        print("%s n=%d batch=%d: %.3f ms  %.2f M polymul/s" % (kind, n, batch, ms, batch / ms / 1e3))
    else:
        if kind == "p32":
            plan = cntt.prime32.Plan.try_new(n, 1062862849); d = torch.randint(0, 1062862849, (batch, n), dtype=torch.int32, device="cuda", generator=g)
        elif kind == "p64s":
            plan = cntt.prime64.Plan.try_new(n, cntt.prime64.Solinas.P); d = torch.randint(0, 2**62, (batch, n), dtype=torch.int64, device="cuda", generator=g)
        else:
            p = cntt.prime.largest_prime_in_arithmetic_progression64(1 << 17, 1, 1 << 61, 1 << 62)
            plan = cntt.prime64.Plan.try_new(n, p); d = torch.randint(0, 2**61, (batch, n), dtype=torch.int64, device="cuda", generator=g)
        f = bench(lambda: plan.fwd(d)); i = bench(lambda: plan.inv(d))
        wb = 4 if kind == "p32" else 8
        print("%s n=%d batch=%d: fwd %.3f ms (%.1f M NTT/s, %.0f GB/s)  inv %.3f ms (%.1f M NTT/s)" % (kind, n, batch, f, batch / f / 1e3, 2 * n * wb * batch / f / 1e6, i, batch / i / 1e3))
