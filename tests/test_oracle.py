"""Pins the CPU oracle (oracle/cntt_oracle.c) against everything the reference's own tests hold for the
hot path (SURVEY.md sections 4 and 8c).  The reference has no golden vectors; its tests are properties
plus a handful of literal known answers -- all of them are restated here.  CPU only.
"""
import numpy as np
import pytest

from conftest import rng, rand_mod, rand_words, primes32, primes64


# ---- literal known answers -----------------------------------------------------------------------------
def test_prime_kats(oracle):
    """src/prime.rs:187-222"""
    O = oracle
    small = [n for n in range(1000) if O.is_prime64(n)]
    sieve = [n for n in range(2, 1000) if all(n % d for d in range(2, int(n ** 0.5) + 1))]
    assert small == sieve
    assert O.is_prime64(0xFFFFFFFF00000001)
    f = O.largest_prime_in_arithmetic_progression64
    M = 2**64 - 1
    assert f(0, 2, 1, 4) == 2
    assert f(0, 2, 2, 2) == 2
    assert f(0, 2, 2, 1) is None
    assert f(1, 0, 14, 16) is None
    assert f(1, 0, 14, 17) == 17
    assert f(1, 0, 17, 18) == 17
    assert f(2, 1, 14, 16) is None
    assert f(2, 1, 14, 17) == 17
    assert f(2, 1, 17, 18) == 17
    assert f(6, 5, 0, M) == 18446744073709551557
    assert f(6, 1, 0, M) == 18446744073709551427


def test_builtin_primes(oracle):
    """src/lib.rs:453-462 (values re-derived in SURVEY.md section 2)"""
    expect = [1062862849, 1063059457, 1064697857, 1065484289, 1068236801,
              1068433409, 1068564481, 1069219841, 1071513601, 1073479681]
    assert [oracle.primes32(i) for i in range(10)] == expect
    assert all(oracle.is_prime64(p) for p in expect)


def test_primitive_root_order(oracle):
    """src/roots.rs:119-130"""
    O = oracle
    deg = 1 << 10
    p = O.largest_prime_in_arithmetic_progression64(deg, 1, 0, 2**64 - 1)
    root = O.find_primitive_root64(p, deg)
    assert O.exp_mod64(p, root, deg) == 1
    assert O.exp_mod64(p, root, deg // 2) == p - 1     # order exactly deg (deg is a power of two)
    # psi for config 1 as derived independently in SURVEY.md section 8(a1)
    assert O.find_primitive_root64(1062862849, 2048) == 306208274
    assert O.Plan32.try_new(1024, 1062862849).psi() == 306208274


def test_try_new_rejections(oracle):
    """src/prime32.rs:635-641, src/prime64.rs:709-713, regression src/prime64.rs:1879-1882"""
    O = oracle
    assert O.Plan64.try_new(2048, 1024) is None
    assert O.Plan32.try_new(16, 1062862849) is None          # n < 32
    assert O.Plan32.try_new(48, 1062862849) is None          # not a power of two
    assert O.Plan32.try_new(32, 1062862849 + 2) is None      # composite
    assert O.Plan64.try_new(8, 0xFFFFFFFF00000001) is None   # n < 16
    assert O.Plan64.try_new(16, 0xFFFFFFFF00000001) is not None
    assert O.Plan32.try_new(65536, 1062862849) is not None   # v2(P0 - 1) = 17
    assert O.Plan32.try_new(131072, 1062862849) is None      # no 2^18-th root
    assert O.Plan32.try_new(65536, 1063059457) is None       # v2(P1 - 1) = 16
    with pytest.raises(O.ReferencePanic):
        O.Plan32.try_new(32, 1)
    with pytest.raises(O.ReferencePanic):
        O.Plan64.try_new(32, 0)
    # native plans stop at 32768 because of P1
    assert O.Native.try_new(32768, 64) is not None
    assert O.Native.try_new(65536, 64) is None


def test_readme_example(oracle):
    """README.md:30-51"""
    plan = oracle.Plan32.try_new(32, 1062862849)
    data = np.arange(32, dtype=np.uint32)
    f = plan.fwd(data.copy())
    back = plan.inv(f.copy())
    assert (back == data * 32).all()


# ---- transform definition, independent of the stage drivers ---------------------------------------------
@pytest.mark.parametrize("n", [32, 64, 256])
def test_fwd_is_evaluation_at_odd_powers(oracle, n):
    """fwd(a)[j] == sum_i a_i psi^((2 brv(j) + 1) i): pins output order and psi."""
    O = oracle
    g = rng(1)
    for p in list(primes32(O).values()):
        plan = O.Plan32.try_new(n, p)
        a = rand_mod(g, p, n, np.uint32)
        got = plan.fwd(a.copy())
        ref = O.direct_fwd64(n, p, plan.psi(), a.astype(np.uint64))
        assert (got.astype(np.uint64) == ref).all(), p
    for name, p in primes64(O).items():
        plan = O.Plan64.try_new(n, p)
        a = rand_mod(g, p, n, np.uint64)
        got = plan.fwd(a.copy())
        ref = O.direct_fwd64(n, p, plan.psi(), a)
        assert (got == ref).all(), name


def test_twiddle_tables(oracle):
    """twid[brv(k)] = psi^k, inv_twid[brv((n-k)%n)] = -psi^k (src/prime32.rs:248-282)"""
    O = oracle
    n, p = 64, 1062862849
    plan = O.Plan32.try_new(n, p)
    psi = plan.psi()
    tw, itw = plan.twid(), plan.inv_twid()
    brv = lambda i: int(format(i, "06b")[::-1], 2)
    for k in range(n):
        assert tw[brv(k)] == pow(psi, k, p)
        assert itw[brv((n - k) % n)] == (1 if k == 0 else p - pow(psi, k, p))


# ---- the reference's property tests ---------------------------------------------------------------------
@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024])
def test_prime32_product(oracle, n):
    """src/prime32.rs:1007-1053: canonical range; inv(fwd a . fwd b) == n (a * b); normalised variant."""
    O = oracle
    g = rng(n)
    for p in primes32(O).values():
        plan = O.Plan32.try_new(n, p)
        a = rand_mod(g, p, n, np.uint32)
        b = rand_mod(g, p, n, np.uint32)
        conv = O.schoolbook32(p, a, b)
        fa, fb = plan.fwd(a.copy()), plan.fwd(b.copy())
        assert (fa < p).all() and (fb < p).all()
        prod = ((fa.astype(object) * fb.astype(object)) % p).astype(np.uint32)
        back = plan.inv(prod.copy())
        assert (back < p).all()
        assert (back.astype(object) == (conv.astype(object) * n) % p).all()
        fa2 = plan.mul_assign_normalize(fa.copy(), fb)
        assert (plan.inv(fa2.copy()) == conv).all()


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024])
def test_prime64_product(oracle, n):
    """src/prime64.rs:1211-1267"""
    O = oracle
    g = rng(n + 7)
    for name, p in primes64(O).items():
        plan = O.Plan64.try_new(n, p)
        a = rand_mod(g, p, n, np.uint64)
        b = rand_mod(g, p, n, np.uint64)
        conv = O.schoolbook64(p, a, b)
        fa, fb = plan.fwd(a.copy()), plan.fwd(b.copy())
        assert (fa < np.uint64(p)).all()
        prod = np.array([(int(x) * int(y)) % p for x, y in zip(fa, fb)], dtype=np.uint64)
        back = plan.inv(prod.copy())
        assert (back < np.uint64(p)).all(), name
        assert [int(v) for v in back] == [(int(c) * n) % p for c in conv], name
        fa2 = plan.mul_assign_normalize(fa.copy(), fb)
        assert (plan.inv(fa2.copy()) == conv).all(), name


def test_depth_first_recursion_matches_definition(oracle):
    """n above RECURSION_THRESHOLD (2048 / 1024) runs the depth-first drivers, which the reference's own
    tests barely reach; pin them with the round trip and a direct evaluation."""
    O = oracle
    g = rng(99)
    for n, p, P in [(8192, 1062862849, O.Plan32), (4096, 0xFFFFFFFF00000001, O.Plan64),
                    (4096, primes64(O)["lt62"], O.Plan64)]:
        plan = P.try_new(n, p)
        dt = np.uint32 if P is O.Plan32 else np.uint64
        a = rand_mod(g, p, n, dt)
        fa = plan.fwd(a.copy())
        back = plan.inv(fa.copy())
        assert [int(v) for v in back] == [(int(x) * n) % p for x in a]
        # spot-check 8 outputs against the definition
        psi = plan.psi()
        nb = n.bit_length() - 1
        for j in g.integers(0, n, 8):
            e = 2 * int(format(int(j), "0%db" % nb)[::-1], 2) + 1
            x = pow(psi, e, p)
            acc, xp = 0, 1
            for v in a:
                acc = (acc + int(v) * xp) % p
                xp = xp * x % p
            assert int(fa[j]) == acc


def test_pointwise_vs_modulo(oracle):
    """src/prime32.rs:1055-1248, src/prime64.rs:1269-1463: pointwise ops == % arithmetic, n = 128."""
    O = oracle
    g = rng(5)
    n = 128
    for P, primes, dt in [(O.Plan32, primes32(O), np.uint32), (O.Plan64, primes64(O), np.uint64)]:
        for p in primes.values():
            plan = P.try_new(n, p)
            ninv = pow(n, p - 2, p)
            a, b, c = (rand_mod(g, p, n, dt) for _ in range(3))
            A, B, Cc = [int(v) for v in a], [int(v) for v in b], [int(v) for v in c]
            assert [int(v) for v in plan.mul_assign_normalize(a.copy(), b)] == [x * y * ninv % p for x, y in zip(A, B)]
            assert [int(v) for v in plan.normalize(a.copy())] == [x * ninv % p for x in A]
            assert [int(v) for v in plan.mul_accumulate(c.copy(), a, b)] == [(z + x * y) % p for x, y, z in zip(A, B, Cc)]


def _to_int(a, bits):
    if bits == 128:
        return [int(lo) | (int(hi) << 64) for lo, hi in a.reshape(-1, 2)]
    return [int(v) for v in a.reshape(-1)]


@pytest.mark.parametrize("bits", [32, 64, 128])
@pytest.mark.parametrize("n", [32, 64, 256, 1024])
def test_native_polymul(oracle, bits, n):
    """src/native32.rs:507-531, native64.rs:1176-1215, native128.rs:394-447"""
    O = oracle
    g = rng(bits * 10000 + n)
    plan = O.Native.try_new(n, bits)
    lhs, rhs = rand_words(g, bits, (n,)), rand_words(g, bits, (n,))
    got = plan.negacyclic_polymul(lhs, rhs)
    if bits == 32:
        ref = O.schoolbook32(0, lhs, rhs)
    elif bits == 64:
        ref = O.schoolbook64(0, lhs, rhs)
    else:
        ref = O.schoolbook128(lhs, rhs)
    assert (got == ref).all()
    # inv(fwd(v)) == v * n (wrapping)
    planes = plan.fwd(lhs)
    assert all((planes[k] < O.primes32(k)).all() for k in range(plan.nprimes))
    back = plan.inv(planes)
    assert _to_int(back, bits) == [(v * n) % (1 << bits) for v in _to_int(lhs, bits)]


@pytest.mark.parametrize("bits", [32, 64, 128])
@pytest.mark.parametrize("n", [32, 256, 1024])
def test_native_binary_polymul(oracle, bits, n):
    """src/native_binary32.rs:333-347, native_binary64.rs:532-545, native_binary128.rs:208-221"""
    O = oracle
    g = rng(bits * 777 + n)
    plan = O.Native.try_new(n, bits, binary=True)
    lhs = rand_words(g, bits, (n,))
    rhs = rand_words(g, bits, (n,))
    rhs = rhs & np.uint64(1) if bits != 32 else rhs & np.uint32(1)
    if bits == 128:
        rhs[:, 1] = 0
    got = plan.negacyclic_polymul(lhs, rhs)
    ref = O.schoolbook32(0, lhs, rhs) if bits == 32 else O.schoolbook64(0, lhs, rhs) if bits == 64 else O.schoolbook128(lhs, rhs)
    assert (got == ref).all()


def test_crt_sign_rule(oracle):
    """Garner lift is the centred representative: x in (-M/2, M/2) round-trips through its residues.
    (The reference pins SIMD == scalar on arbitrary residues, src/native64.rs:1245-1293; with a single
    scalar restatement the observable property is the centred lift.)"""
    O = oracle
    g = rng(2024)
    P = [O.primes32(i) for i in range(10)]
    for x in [0, 1, -1, 12345678901234567, -98765432109876543, 2**62, -(2**62)] + [int(v) for v in g.integers(-2**63, 2**63, 50)]:
        r = [x % p for p in P]
        assert O.lib().o_reconstruct_32bit_01234_u64(*r[:5]) == x % 2**64
        assert O.lib().o_reconstruct_32bit_012_u64(*r[:3]) == x % 2**64
        assert O.lib().o_reconstruct_32bit_012_u32(*r[:3]) == x % 2**32
        if abs(x) < P[0] * P[1] // 2:
            assert O.lib().o_reconstruct_32bit_01(*r[:2]) == x % 2**32


# ---- product::Plan (src/product.rs:969-1170) -------------------------------------------------------------
def _product_cases(O, n):
    """the prime sets of the reference's product tests (src/product.rs:976-1155), n = 256 there"""
    f = O.largest_prime_in_arithmetic_progression64
    d = 2 * n
    u64x1 = [f(d, 1, 0, 2**64 - 1)]
    u32x1 = [f(d, 1, 0, 2**32 - 1)]
    p0 = f(d, 1, 0, 2**32 - 1)
    u32x2 = [p0, f(d, 1, 0, p0 - 1)]
    q0 = f(d, 1, 0, 1 << 30)
    u30x2 = [q0, f(d, 1, 0, q0 - 1)]
    r = [f(d, 1, 0, 2**16 - 1)]
    for _ in range(3):
        nxt = f(d, 1, 0, r[-1] - 1)
        if nxt is None:                                  # n > 256: fewer than four such primes below 2^16
            break
        r.append(nxt)
    if n <= 256:
        s1 = f(d, 1, 0, 1 << 15)
        mixed = [f(d, 1, 0, 1 << 33), s1, f(d, 1, 0, s1 - 1)]
    else:                                                # larger n: same shape, windows that still hold primes
        s1 = f(d, 1, 0, 1 << 16)
        s2 = f(d, 1, 0, s1 - 1)
        mixed = [f(d, 1, 1 << 32, (2**64 - 1) // (s1 * s2)), s1, s2]
    return {"u64x1": u64x1, "u32x1": u32x1, "u32x2": u32x2, "u30x2": u30x2, "u32x4": r, "u32x2_u64x1": mixed}


def _prod(ps):
    m = 1
    for p in ps:
        m *= p
    return m


@pytest.mark.parametrize("case", ["u64x1", "u32x1", "u32x2", "u30x2", "u32x4", "u32x2_u64x1"])
def test_product_roundtrip(oracle, case):
    """test_product_* (src/product.rs:976-1155): inv(fwd(x)) * n^-1 == x mod the composite modulus, both
    InvModes (Accumulate onto zeros)."""
    n = 256
    ps = _product_cases(oracle, n)[case]
    p = _prod(ps)
    assert p < 2**64
    plan = oracle.Product.try_new(n, p, ps)
    assert plan is not None and plan.ntt_size() == n
    g = rng(sum(case.encode()))
    standard = rand_mod(g, p, (n,), np.uint64)
    n_inv = pow(n, -1, p)
    for mode in (plan.REPLACE, plan.ACCUMULATE):
        ntt = np.zeros(plan.ntt_domain_len(), dtype=np.uint64)
        rt = np.zeros(n, dtype=np.uint64)
        plan.fwd(ntt, standard, plan.GENERIC)
        plan.inv(rt, ntt, mode)
        got = [int(x) * n_inv % p for x in rt]
        assert got == [int(x) for x in standard], (case, mode)


def test_product_failures(oracle):
    """test_plan_failure_zero / _dup (src/product.rs:1155-1170) and the other None outcomes of try_new"""
    O = oracle
    n = 256
    f = O.largest_prime_in_arithmetic_progression64
    p0, p1 = f(2 * n, 1, 0, 1 << 33), f(2 * n, 1, 0, 1 << 15)
    assert O.Product.try_new(n, 0, [p0, 0]) is None
    assert O.Product.try_new(n, p0 * p1 * p1 % 2**64, [p1, p0, p1]) is None
    assert O.Product.try_new(n, p0 * p1 + 1, [p0, p1]) is None          # product != modulus
    assert O.Product.try_new(n, p0 * p1, [p0, p1, 1]) is not None       # 1s are skipped
    assert O.Product.try_new(n, 15 * p1, [15, p1]) is None              # a non-prime factor
    big = f(2 * n, 1, 0, 2**64 - 1)
    assert O.Product.try_new(n, (big * p0) % 2**64, [big, p0]) is None  # checked_mul overflow


def test_product_bounded_and_polymul(oracle):
    """FwdMode::Bounded on the two-u32-prime plan gives the residues of the centred value (product.rs:305-322),
    and mul_assign_normalize + inv is the negacyclic product mod p0*p1."""
    n = 64
    f = oracle.largest_prime_in_arithmetic_progression64
    p0 = f(2 * n, 1, 0, 1 << 31)
    p1 = f(2 * n, 1, 0, p0 - 1)
    p = p0 * p1
    plan = oracle.Product.try_new(n, p, [p0, p1])
    g = rng(77)
    bound = 1 << 20
    small = g.integers(-bound + 1, bound, size=n)
    standard = np.array([int(x) % p for x in small], dtype=np.uint64)
    a = np.zeros(plan.ntt_domain_len(), dtype=np.uint64)
    b = np.zeros_like(a)
    plan.fwd(a, standard, plan.bounded(bound))
    plan.fwd(b, standard, plan.GENERIC)
    assert (a == b).all()
    rhs = rand_mod(g, p, (n,), np.uint64)
    c = np.zeros_like(a)
    plan.fwd(c, rhs, plan.GENERIC)
    plan.mul_assign_normalize(a, c)
    out = np.zeros(n, dtype=np.uint64)
    plan.inv(out, a, plan.REPLACE)
    ref = [0] * n
    for i in range(n):
        for j in range(n):
            t = int(standard[i]) * int(rhs[j])
            if i + j < n:
                ref[i + j] = (ref[i + j] + t) % p
            else:
                ref[i + j - n] = (ref[i + j - n] - t) % p
    assert [int(x) for x in out] == ref
    # Accumulate adds the same lift modulo p (product.rs:749-753)
    acc0 = rand_mod(g, p, (n,), np.uint64)
    acc = acc0.copy()
    plan.fwd(a, standard, plan.GENERIC)
    plan.mul_assign_normalize(a, c)
    plan.inv(acc, a, plan.ACCUMULATE)
    assert [int(x) for x in acc] == [(int(s) + r) % p for s, r in zip(acc0, ref)]


def test_fast_wrapping_schoolbook_equals_reference_test_oracle(oracle):
    """o_negacyclic_wrapping (used to check the N = 65536 extension) == the reference's own test oracle
    (wrapping schoolbook, src/native64.rs:1176-1215 / native128.rs:359-372) at sizes both can run."""
    g = np.random.Generator(np.random.PCG64(77))
    for n in (32, 100, 256):
        a = g.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        b = g.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        assert (oracle.negacyclic_wrapping(32, a, b) == oracle.schoolbook32(0, a, b)).all()
        a = g.integers(0, 2**64, size=n, dtype=np.uint64)
        b = g.integers(0, 2**64, size=n, dtype=np.uint64)
        assert (oracle.negacyclic_wrapping(64, a, b) == oracle.schoolbook64(0, a, b)).all()
        a = g.integers(0, 2**64, size=(n, 2), dtype=np.uint64)
        b = g.integers(0, 2**64, size=(n, 2), dtype=np.uint64)
        a[::3] = 0
        assert (oracle.negacyclic_wrapping(128, a, b) == oracle.schoolbook128(a, b)).all()
    # and against the oracle's NTT-based native64 polymul at the reference's largest size
    n = 32768
    a = g.integers(0, 2**64, size=n, dtype=np.uint64)
    b = g.integers(0, 2**64, size=n, dtype=np.uint64)
    assert (oracle.negacyclic_wrapping(64, a, b) == oracle.Native.try_new(n, 64).negacyclic_polymul(a, b)).all()


@pytest.mark.parametrize("bits,binary", [(32, False), (64, False), (32, True), (64, True)])
def test_plan52_restatement_is_a_polymul(oracle, bits, binary):
    """Native52 (the Plan52 twins, restated from the AVX-512-only reference code): fwd, per-prime
    mul_assign_normalize, inv == wrapping negacyclic schoolbook -- the reference's own Plan52 tests
    (src/native64.rs:1217-1243)."""
    g = np.random.Generator(np.random.PCG64(520 + bits + int(binary)))
    for n in (32, 128):
        op = oracle.Native52.try_new(n, bits, binary)
        dt = np.uint32 if bits == 32 else np.uint64
        a = g.integers(0, 2**bits, size=n, dtype=np.uint64).astype(dt)
        b = g.integers(0, 2 if binary else 2**bits, size=n, dtype=np.uint64).astype(dt)
        A, B = op.fwd(a), op.fwd(b, binary_copy=binary)
        for k, pl in enumerate(op.plans):
            pl.mul_assign_normalize(A[k], B[k])
        ref = (oracle.schoolbook32 if bits == 32 else oracle.schoolbook64)(0, a, b)
        assert (op.inv(A) == ref).all()
