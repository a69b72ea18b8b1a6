// host_math.hpp -- plan-time number theory of the product library (host only, runs once per plan).
//
// What must match the reference bit for bit is the *choice* of the primitive 2N-th root psi
// (every fwd output depends on it) and the acceptance rules of try_new.  The procedure is the
// reference's: psi = (log2(2N) - 1)-fold Tonelli-Shanks square root of -1, with the smallest
// quadratic non-residue as the Tonelli-Shanks generator (roots.rs:17-28, 31-66, 68-91), and the
// deterministic 12-base Miller-Rabin of prime.rs:76-126.  Everything else (how tables are laid out,
// which reduction the kernels use) is this library's own.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace cntt {
namespace host {

typedef unsigned __int128 u128;

struct Fp {
    uint64_t p;
    explicit Fp(uint64_t p_) : p(p_) {}
    uint64_t mul(uint64_t a, uint64_t b) const { return (uint64_t)(((u128)a * b) % p); }
    uint64_t pow(uint64_t b, uint64_t e) const
    {
        uint64_t r = 1 % p;
        b %= p;
        while (e) {
            if (e & 1) r = mul(r, b);
            b = mul(b, b);
            e >>= 1;
        }
        return r;
    }
    uint64_t inv(uint64_t a) const { return pow(a, p - 2); }
};

inline bool is_prime_u64(uint64_t n)
{
    static const uint64_t bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return false;
    for (uint64_t b : bases)
        if (n % b == 0) return n == b;
    uint64_t d = n - 1;
    int s = 0;
    while ((d & 1) == 0) { d >>= 1; ++s; }
    Fp f(n);
    for (uint64_t a : bases) {
        uint64_t x = f.pow(a, d);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int r = 1; r < s; ++r) {
            x = f.mul(x, x);
            if (x == n - 1) { witness = false; break; }
        }
        if (witness) return false;
    }
    return true;
}

// Largest prime of the form factor*x + offset in [lo, hi]; prime.rs:130-180.
inline bool largest_prime_in_arithmetic_progression(uint64_t factor, uint64_t offset, uint64_t lo, uint64_t hi, uint64_t* out)
{
    if (lo > hi || offset > hi) return false;
    if (factor == 0) {
        if (lo <= offset && is_prime_u64(offset)) { *out = offset; return true; }
        return false;
    }
    const uint64_t start = lo > offset ? lo : offset;
    uint64_t x_lo = (start - offset) / factor + ((start - offset) % factor != 0);
    uint64_t x = (hi - offset) / factor;
    for (;; --x) {
        const uint64_t v = factor * x + offset;
        if (is_prime_u64(v)) { *out = v; return true; }
        if (x == x_lo) return false;
    }
}

// Tonelli-Shanks with the reference's generator and its choice between the two roots.
inline bool ts_sqrt(const Fp& f, uint64_t q, uint64_t s, uint64_t z, uint64_t a, uint64_t* root)
{
    uint64_t m = s, c = f.pow(z, q), t = f.pow(a, q), r = f.pow(a, (q + 1) / 2);
    while (true) {
        if (t == 0) { *root = 0; return true; }
        if (t == 1) { *root = r; return true; }
        uint64_t i = 0, tp = t;
        while (i < m) {
            tp = f.mul(tp, tp);
            ++i;
            if (tp == 1) break;
        }
        if (i == m) return false;
        const uint64_t b = f.pow(c, (uint64_t)1 << (m - i - 1));
        m = i;
        c = f.mul(b, b);
        t = f.mul(t, c);
        r = f.mul(r, b);
    }
}

// Primitive `degree`-th root of unity (degree = 2N, a power of two > 1) or false.
inline bool primitive_root_pow2(uint64_t p, uint64_t degree, uint64_t* psi)
{
    Fp f(p);
    uint64_t q = p - 1, s = 0;
    while ((q & 1) == 0) { q >>= 1; ++s; }
    uint64_t z = 0;
    for (uint64_t c = 2; c < p; ++c)
        if (f.pow(c, (p - 1) / 2) == p - 1) { z = c; break; }
    if (z == 0) return false;
    int lg = 0;
    while (((uint64_t)1 << lg) < degree) ++lg;
    uint64_t root = p - 1;
    for (int i = 0; i + 1 < lg; ++i)
        if (!ts_sqrt(f, q, s, z, root, &root)) return false;
    *psi = root;
    return true;
}

inline size_t brv(int bits, size_t i)
{
    size_t r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}

// Heap-ordered negacyclic twiddles: fwd[brv(k)] = psi^k, inv[brv((n-k) mod n)] = -psi^k (k != 0),
// inv[0] = 1  (prime32.rs:248-282, prime64.rs:183-218).
inline void negacyclic_twiddles(uint64_t p, int logn, uint64_t psi, std::vector<uint64_t>& fwd, std::vector<uint64_t>& inv)
{
    const size_t n = (size_t)1 << logn;
    fwd.assign(n, 0);
    inv.assign(n, 0);
    Fp f(p);
    uint64_t wk = 1;
    for (size_t k = 0; k < n; ++k) {
        fwd[brv(logn, k)] = wk;
        inv[brv(logn, (n - k) % n)] = k == 0 ? wk : p - wk;
        wk = f.mul(wk, psi);
    }
}

inline int ilog2(uint64_t x) { return 63 - __builtin_clzll(x); }

} // namespace host
} // namespace cntt
