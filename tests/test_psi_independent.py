"""The 2n-th root of unity psi fixes the ORDER and the VALUES of every `fwd` output, and no reference test inspects them
(SURVEY.md 8c): the product's root search (csrc/host_math.hpp) and the oracle's (oracle/cntt_oracle.c) were both restated from
src/roots.rs.  This file is a third, independently written search -- Python big integers, straight from the textbook statement of
Tonelli-Shanks (least quadratic non-residue z, Q 2^S = p - 1, the classical loop on (M, c, t, R)) applied the way the crate's doc
comment describes: start from -1 and take square roots until the order is `degree` -- and compares all three for every prime the
test-suite, the native plans and the bench use.  It also checks the defining properties, which no shared misreading could fake:
psi^degree = 1, psi^(degree/2) = -1."""
import importlib

import pytest

P0_9 = [0b0011_1111_0101_1010_0000_0000_0000_0001, 0b0011_1111_0101_1101_0000_0000_0000_0001, 0b0011_1111_0111_0110_0000_0000_0000_0001,
        0b0011_1111_1000_0010_0000_0000_0000_0001, 0b0011_1111_1010_1100_0000_0000_0000_0001, 0b0011_1111_1010_1111_0000_0000_0000_0001,
        0b0011_1111_1011_0001_0000_0000_0000_0001, 0b0011_1111_1011_1011_0000_0000_0000_0001, 0b0011_1111_1101_1110_0000_0000_0000_0001,
        0b0011_1111_1111_1100_0000_0000_0000_0001]                      # primes32::P0..P9, src/lib.rs:453-462
P52 = [0b0011_1111_1111_1111_1111_1111_1110_0111_0111_0000_0000_0000_0001, 0b0011_1111_1111_1111_1111_1111_1110_1011_1001_0000_0000_0000_0001,
       0b0011_1111_1111_1111_1111_1111_1110_1100_1000_0000_0000_0000_0001, 0b0011_1111_1111_1111_1111_1111_1111_1000_1011_0000_0000_0000_0001,
       0b0011_1111_1111_1111_1111_1111_1111_1011_1000_0000_0000_0000_0001, 0b0011_1111_1111_1111_1111_1111_1111_1100_0111_0000_0000_0000_0001]  # primes52, src/lib.rs:601-606


def is_prime(n):
    """deterministic Miller-Rabin for n < 2^64 (first twelve primes as witnesses)"""
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for q in small:
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def tonelli_shanks(n, p, z):
    """a square root of n modulo the odd prime p (None if n is a non-residue); z = a quadratic non-residue.
    Textbook form: https://en.wikipedia.org/wiki/Tonelli-Shanks_algorithm#The_algorithm"""
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    m, c, t, r = s, pow(z, q, p), pow(n, q, p), pow(n, (q + 1) // 2, p)
    while True:
        if t == 0:
            return 0
        if t == 1:
            return r
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
            if i == m:
                return None
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p


def psi_independent(p, degree):
    """primitive `degree`-th root of unity: -1 has order 2; each square root doubles the order"""
    z = next(a for a in range(2, p) if pow(a, (p - 1) // 2, p) == p - 1)   # least quadratic non-residue
    root, order = p - 1, 2
    while order < degree:
        root = tonelli_shanks(root, p, z)
        if root is None:
            return None
        order *= 2
    return root


def all_primes(O):
    f = O.largest_prime_in_arithmetic_progression64
    ps = set(P0_9) | set(P52) | {0xFFFFFFFF00000001}
    for step in (1 << 16, 1 << 17, 1 << 18):
        for lo, hi in ((1 << 29, 1 << 30), (1 << 30, 1 << 31), (1 << 31, 1 << 32), (1 << 49, 1 << 50), (1 << 50, 1 << 51),
                       (1 << 61, 1 << 62), (1 << 62, 1 << 63), (1 << 63, (1 << 64) - 1)):
            ps.add(f(step, 1, lo, hi))
    # the nine extended native primes k 2^17 + 1 below 2^30 and the product-plan bench primes
    k, found = (1 << 30) // (1 << 17), []
    while len(found) < 9:
        k -= 1
        if is_prime(k * (1 << 17) + 1):
            found.append(k * (1 << 17) + 1)
    ps |= set(found)
    for n in (1024, 2048):
        p0 = f(2 * n, 1, 0, 1 << 31)
        ps |= {p0, f(2 * n, 1, 0, p0 - 1)}
    return sorted(ps)


def test_miller_rabin_agrees(oracle):
    cntt = importlib.import_module("concrete-ntt_b200")
    for p in all_primes(oracle):
        assert is_prime(p) and oracle.is_prime64(p) and cntt.prime.is_prime64(p)
    for c in (1, 4, 561, 1062862849 + 2, 0xFFFFFFFF00000001 + 2, 3215031751, 3825123056546413051):
        assert is_prime(c) == oracle.is_prime64(c) == cntt.prime.is_prime64(c)


def test_psi_three_ways(oracle):
    cntt = importlib.import_module("concrete-ntt_b200")      # host-only entry point of the C ABI: no GPU needed
    checked = 0
    for p in all_primes(oracle):
        v2 = ((p - 1) & -(p - 1)).bit_length() - 1
        for logd in sorted({5, 6, 11, 12, 13, 16, 17, 18, min(v2, 27), v2 + 1}):
            degree = 1 << logd
            mine = psi_independent(p, degree)
            assert (mine is None) == (logd > v2), (p, logd)
            assert oracle.find_primitive_root64(p, degree) == mine, (p, logd)
            assert cntt.roots.find_primitive_root64(p, degree) == mine, (p, logd)
            if mine is not None:
                assert pow(mine, degree, p) == 1 and pow(mine, degree // 2, p) == p - 1
                checked += 1
    assert checked > 250


def test_known_value():
    """SURVEY.md section 8 derived psi(P0, 2048) = 306208274 independently"""
    assert psi_independent(1062862849, 2048) == 306208274


@pytest.mark.parametrize("bits,n", [(32, 32), (32, 1024), (64, 16), (64, 2048)])
def test_twiddle_tables_follow_from_psi(oracle, bits, n):
    """twid[brv(k)] = psi^k, inv_twid[brv(k)] = psi^-k (src/prime32.rs:248-282) -- the oracle's tables against big-int powers of the
    independently found psi"""
    f = oracle.largest_prime_in_arithmetic_progression64
    primes = [1062862849, f(1 << 16, 1, 1 << 31, 1 << 32)] if bits == 32 else [0xFFFFFFFF00000001, f(1 << 16, 1, 1 << 61, 1 << 62)]
    logn = n.bit_length() - 1
    for p in primes:
        psi = psi_independent(p, 2 * n)
        plan = (oracle.Plan32 if bits == 32 else oracle.Plan64).try_new(n, p)
        assert plan.psi() == psi
        tw, itw = plan.twid(), plan.inv_twid()
        for k in range(n):
            r = int(format(k, "0%db" % logn)[::-1], 2)
            assert int(tw[r]) == pow(psi, k, p)
            assert int(itw[r]) == pow(psi, -k, p)
