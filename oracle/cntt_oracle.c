/*
 * cntt_oracle.c -- CPU ORACLE for the concrete-ntt hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the *scalar* code paths of zama-ai/concrete-ntt v0.2.0
 * (the reference at /root/reference).  It is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker and as the CPU baseline.
 * It is NEVER linked into, imported by, or called from the product library (libcntt_b200.so).
 *
 * The reference is Rust and no Rust toolchain exists in this image, so the reference itself can
 * not be compiled or run here.  The reference's own tests hold no golden
 * vectors for this path; they pin results by properties only.  This oracle is pinned against
 * every one of those properties and the few literal known-answer values the reference has
 * (tests/test_oracle.py): README.md:30-51 round trip, src/prime.rs:187-222 prime KATs,
 * src/prime64.rs:1879-1882 try_new regression, src/roots.rs:111-131 root order, the per-module
 * property tests, and the independent identity
 * fwd(a)[j] == sum_i a_i psi^((2 brv(j)+1) i).
 *
 * Every function cites the reference file:line it follows.  All paths are relative to
 * /root/reference/.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Plan-time scalar number theory: src/prime.rs, src/roots.rs, src/fastdiv.rs
 * Div32/Div64 (src/fastdiv.rs:28-150) are exact Lemire division/remainder; here they are the
 * plain C `/` and `%` operators on u64/u128, which return the same values for every input.
 * ------------------------------------------------------------------------------------------ */

/* src/prime.rs:7-10 */
static inline u64 mul_mod64(u64 p, u64 x, u64 y) { return (u64)(((u128)x * y) % p); }

/* src/prime.rs:31-48 (same loop structure; any correct square-and-multiply gives the same value) */
static u64 exp_mod64(u64 p, u64 base, u64 pw)
{
    if (pw == 0) return 1;
    u64 y = 1, x = base;
    while (pw > 1) {
        if (pw % 2 == 1) y = mul_mod64(p, x, y);
        x = mul_mod64(p, x, x);
        pw /= 2;
    }
    return mul_mod64(p, x, y);
}

/* src/prime.rs:50-66 */
static int miller_rabin_iter(u64 n, u64 s, u64 d, u64 a)
{
    u64 x = exp_mod64(n, a, d);
    u64 n_minus_1 = n - 1;
    if (x == 1 || x == n_minus_1) return 1;
    for (u64 count = 0; count + 1 < s; count++) {
        x = mul_mod64(n, x, x);
        if (x == n_minus_1) return 1;
    }
    return 0;
}

/* src/prime.rs:76-126 */
EXPORT int o_is_prime64(u64 n)
{
    static const u64 small[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (int i = 0; i < 12; i++)
        if (n % small[i] == 0) return n == small[i];
    u64 s = 0, d = n - 1;
    while (d % 2 == 0) { s++; d /= 2; }
    for (int i = 0; i < 12; i++)
        if (!miller_rabin_iter(n, s, d, small[i])) return 0;
    return 1;
}

/* src/prime.rs:130-180.  Returns 1 and *out on Some(..), 0 on None. */
EXPORT int o_largest_prime_in_arithmetic_progression64(u64 factor, u64 offset, u64 lo, u64 hi, u64 *out)
{
    if (lo > hi) return 0;
    u64 a = factor, b = offset;
    if (b > hi) return 0;
    if (a == 0) {
        if (lo <= b && b <= hi && o_is_prime64(b)) { *out = b; return 1; }
        return 0;
    }
    u64 mx = lo > b ? lo : b;
    u64 x_lo = (mx - b) / a;
    if ((mx - b) % a != 0) x_lo += 1;
    u64 x_hi = (hi - b) / a;
    u64 x = x_hi;
    for (;;) {
        u64 val = a * x + b;
        if (o_is_prime64(val)) { *out = val; return 1; }
        if (x == x_lo) break;
        x -= 1;
    }
    return 0;
}

/* src/roots.rs:17-28: smallest quadratic non-residue >= 2.  Returns 0 for None. */
static int get_z64(u64 p, u64 *z)
{
    for (u64 n = 2; n < p; n++) {
        if (exp_mod64(p, n, (p - 1) / 2) == p - 1) { *z = n; return 1; }
    }
    return 0;
}

/* src/roots.rs:31-66: Tonelli-Shanks, returning exactly the root the reference returns. */
static int sqrt_mod_ex64(u64 p, u64 q, u64 s, u64 z, u64 n, u64 *out)
{
    u64 m = s;
    u64 c = exp_mod64(p, z, q);
    u64 t = exp_mod64(p, n, q);
    u64 r = exp_mod64(p, n, (q + 1) / 2);
    for (;;) {
        if (t == 0) { *out = 0; return 1; }
        if (t == 1) { *out = r; return 1; }
        u64 i = 0;
        u64 t_pow = t;
        while (i < m) {
            t_pow = mul_mod64(p, t_pow, t_pow);
            i += 1;
            if (t_pow == 1) break;
        }
        if (i == m) return 0;
        u64 b = exp_mod64(p, c, (u64)1 << (m - i - 1));
        m = i;
        c = mul_mod64(p, b, b);
        t = mul_mod64(p, t, c);
        r = mul_mod64(p, r, b);
    }
}

/* src/roots.rs:68-91.  degree is a power of two > 1.  Returns 1 and *root on Some. */
EXPORT int o_find_primitive_root64(u64 p, u64 degree, u64 *root_out)
{
    unsigned n = (unsigned)__builtin_ctzll(degree);
    u64 root = p - 1;
    /* get_q_s64: src/roots.rs:6-15 */
    u64 q = p - 1, s = 0;
    while (q % 2 == 0) { q /= 2; s += 1; }
    u64 z;
    if (!get_z64(p, &z)) return 0;
    for (unsigned i = 0; i + 1 < n; i++) {
        u64 r;
        if (!sqrt_mod_ex64(p, q, s, z, root, &r)) return 0;
        root = r;
    }
    *root_out = root;
    return 1;
}

EXPORT u64 o_exp_mod64(u64 p, u64 base, u64 pw) { return exp_mod64(p, base, pw); }
EXPORT u64 o_mul_mod64(u64 p, u64 a, u64 b) { return mul_mod64(p, a, b); }

/* src/lib.rs:118-121 */
static inline size_t bit_rev(unsigned nbits, size_t i)
{
    size_t r = 0;
    for (unsigned b = 0; b < nbits; b++) r |= ((i >> b) & 1) << (nbits - 1 - b);
    return r;
}

static inline unsigned ilog2_u64(u64 x) { return 63u - (unsigned)__builtin_clzll(x); }
static inline u32 min32(u32 a, u32 b) { return a < b ? a : b; }
static inline u64 min64(u64 a, u64 b) { return a < b ? a : b; }

/* ==========================================================================================
 * prime32::Plan   (src/prime32.rs)
 * ========================================================================================== */

#define RECURSION_THRESHOLD_32 2048 /* src/prime32.rs:12 */
#define RECURSION_THRESHOLD_64 1024 /* src/prime64.rs:7 */

typedef struct {
    size_t n;
    u32 p;
    u32 *twid, *twid_shoup, *inv_twid, *inv_twid_shoup; /* shoup tables NULL when p >= 2^31 */
    u32 p_barrett, big_q, n_inv_mod_p, n_inv_mod_p_shoup;
    u32 psi; /* the primitive 2n-th root the tables were built from (diagnostic) */
} o_plan32;

/* Status codes of o_plan*_new: 0 = Some(plan), 1 = None, 2 = the reference panics
 * (Div32::new / Div64::new assert divisor > 1: src/fastdiv.rs:49,99 run before validation at
 * src/prime32.rs:631, src/prime64.rs:705). */
EXPORT int o_plan32_new(size_t n, u32 p, o_plan32 **out)
{
    *out = NULL;
    if (p <= 1) return 2;
    /* src/prime32.rs:635-641 */
    u64 psi;
    if (n < 32 || (n & (n - 1)) != 0 || !o_is_prime64(p) || !o_find_primitive_root64(p, 2 * (u64)n, &psi))
        return 1;
    o_plan32 *pl = (o_plan32 *)calloc(1, sizeof(*pl));
    pl->n = n;
    pl->p = p;
    pl->psi = (u32)psi;
    pl->twid = (u32 *)malloc(n * sizeof(u32));
    pl->inv_twid = (u32 *)malloc(n * sizeof(u32));
    int shoup = p < ((u32)1 << 31);
    if (shoup) {
        pl->twid_shoup = (u32 *)malloc(n * sizeof(u32));
        pl->inv_twid_shoup = (u32 *)malloc(n * sizeof(u32));
    }
    /* init_negacyclic_twiddles[_shoup]: src/prime32.rs:223-282 */
    unsigned nbits = (unsigned)__builtin_ctzll(n);
    u32 w = (u32)psi, wk = 1;
    for (size_t k = 0; k < n; k++) {
        size_t fwd_idx = bit_rev(nbits, k);
        u32 wk_shoup = (u32)(((u64)wk << 32) / p);
        pl->twid[fwd_idx] = wk;
        if (shoup) pl->twid_shoup[fwd_idx] = wk_shoup;
        size_t inv_idx = bit_rev(nbits, (n - k) % n);
        if (k == 0) {
            pl->inv_twid[inv_idx] = wk;
            if (shoup) pl->inv_twid_shoup[inv_idx] = wk_shoup;
        } else {
            u32 x = p - wk;
            pl->inv_twid[inv_idx] = x;
            if (shoup) pl->inv_twid_shoup[inv_idx] = (u32)(((u64)x << 32) / p);
        }
        wk = (u32)(((u64)wk * w) % p);
    }
    /* src/prime32.rs:667-671 */
    pl->n_inv_mod_p = (u32)exp_mod64(p, (u64)n % p, (u64)p - 2); /* exp_mod32(p_div, n as u32, p-2) */
    pl->n_inv_mod_p_shoup = (u32)(((u64)pl->n_inv_mod_p << 32) / p);
    pl->big_q = ilog2_u64(p) + 1;
    u32 big_l = pl->big_q + 31;
    pl->p_barrett = (u32)(((u64)1 << big_l) / p);
    *out = pl;
    return 0;
}

EXPORT void o_plan32_free(o_plan32 *pl)
{
    if (!pl) return;
    free(pl->twid); free(pl->twid_shoup); free(pl->inv_twid); free(pl->inv_twid_shoup);
    free(pl);
}
EXPORT size_t o_plan32_ntt_size(const o_plan32 *pl) { return pl->n; }
EXPORT u32 o_plan32_modulus(const o_plan32 *pl) { return pl->p; }
EXPORT u32 o_plan32_psi(const o_plan32 *pl) { return pl->psi; }
EXPORT const u32 *o_plan32_twid(const o_plan32 *pl) { return pl->twid; }
EXPORT const u32 *o_plan32_inv_twid(const o_plan32 *pl) { return pl->inv_twid; }

/* ---- butterflies, p < 2^30 : src/prime32/less_than_30bit.rs:115-152, 265-302 ---- */
typedef struct { u32 a, b; } pair32;
typedef pair32 (*bf32_fn)(u32 z0, u32 z1, u32 w, u32 w_shoup, u32 p, u32 neg_p, u32 two_p);

static pair32 fwd_bf30(u32 z0, u32 z1, u32 w, u32 ws, u32 p, u32 neg_p, u32 two_p)
{
    (void)p;
    z0 = min32(z0, z0 - two_p);
    u32 shoup_q = (u32)(((u64)z1 * ws) >> 32);
    u32 t = z1 * w + shoup_q * neg_p;
    return (pair32){z0 + t, z0 - t + two_p};
}
static pair32 fwd_last_bf30(u32 z0, u32 z1, u32 w, u32 ws, u32 p, u32 neg_p, u32 two_p)
{
    z0 = min32(z0, z0 - two_p);
    z0 = min32(z0, z0 - p);
    u32 shoup_q = (u32)(((u64)z1 * ws) >> 32);
    u32 t = z1 * w + shoup_q * neg_p;
    t = min32(t, t - p);
    u32 r0 = z0 + t, r1 = z0 - t + p;
    return (pair32){min32(r0, r0 - p), min32(r1, r1 - p)};
}
static pair32 inv_bf30(u32 z0, u32 z1, u32 w, u32 ws, u32 p, u32 neg_p, u32 two_p)
{
    (void)p;
    u32 y0 = z0 + z1;
    y0 = min32(y0, y0 - two_p);
    u32 t = z0 - z1 + two_p;
    u32 shoup_q = (u32)(((u64)t * ws) >> 32);
    u32 y1 = t * w + shoup_q * neg_p;
    return (pair32){y0, y1};
}
static pair32 inv_last_bf30(u32 z0, u32 z1, u32 w, u32 ws, u32 p, u32 neg_p, u32 two_p)
{
    u32 y0 = z0 + z1;
    y0 = min32(y0, y0 - two_p);
    u32 t = z0 - z1 + two_p;
    u32 shoup_q = (u32)(((u64)t * ws) >> 32);
    u32 y1 = t * w + shoup_q * neg_p;
    return (pair32){min32(y0, y0 - p), min32(y1, y1 - p)};
}

/* ---- butterflies, 2^30 <= p < 2^31 : src/prime32/less_than_31bit.rs:117-157, 214-234 ---- */
static pair32 fwd_bf31(u32 z0, u32 z1, u32 w, u32 ws, u32 p, u32 neg_p, u32 two_p)
{
    (void)two_p;
    z0 = min32(z0, z0 - p);
    u32 shoup_q = (u32)(((u64)z1 * ws) >> 32);
    u32 t = z1 * w + shoup_q * neg_p;
    t = min32(t, t - p);
    return (pair32){z0 + t, z0 - t + p};
}
static pair32 fwd_last_bf31(u32 z0, u32 z1, u32 w, u32 ws, u32 p, u32 neg_p, u32 two_p)
{
    (void)two_p;
    z0 = min32(z0, z0 - p);
    u32 shoup_q = (u32)(((u64)z1 * ws) >> 32);
    u32 t = z1 * w + shoup_q * neg_p;
    t = min32(t, t - p);
    u32 r0 = z0 + t, r1 = z0 - t + p;
    return (pair32){min32(r0, r0 - p), min32(r1, r1 - p)};
}
static pair32 inv_bf31(u32 z0, u32 z1, u32 w, u32 ws, u32 p, u32 neg_p, u32 two_p)
{
    (void)two_p;
    u32 y0 = z0 + z1;
    y0 = min32(y0, y0 - p);
    u32 t = z0 - z1 + p;
    u32 shoup_q = (u32)(((u64)t * ws) >> 32);
    u32 y1 = t * w + shoup_q * neg_p;
    y1 = min32(y1, y1 - p);
    return (pair32){y0, y1};
}

/* ---- Shoup-form stage drivers: src/prime32/shoup.rs:582-708 (fwd), 1355-1481 (inv) ---- */
static void fwd_breadth_first32(u32 p, u32 *data, size_t n, const u32 *twid, const u32 *twid_shoup,
                                size_t depth, size_t half, bf32_fn butterfly, bf32_fn last_butterfly)
{
    size_t t = n, m = 1, w_idx = (m << depth) + half * m;
    u32 neg_p = (u32)0 - p, two_p = 2 * p;
    while (m < n) {
        t /= 2;
        const u32 *w = twid + w_idx, *ws = twid_shoup + w_idx;
        bf32_fn bf = (t == 1) ? last_butterfly : butterfly;
        for (size_t i = 0; i < m; i++) {
            u32 *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                pair32 r = bf(z0[j], z1[j], w[i], ws[i], p, neg_p, two_p);
                z0[j] = r.a; z1[j] = r.b;
            }
        }
        m *= 2;
        w_idx *= 2;
    }
}

static void fwd_depth_first32(u32 p, u32 *data, size_t n, const u32 *twid, const u32 *twid_shoup,
                              size_t depth, size_t half, bf32_fn butterfly, bf32_fn last_butterfly)
{
    if (n <= RECURSION_THRESHOLD_32) {
        fwd_breadth_first32(p, data, n, twid, twid_shoup, depth, half, butterfly, last_butterfly);
        return;
    }
    size_t t = n / 2, w_idx = ((size_t)1 << depth) + half;
    u32 neg_p = (u32)0 - p, two_p = 2 * p;
    u32 w = twid[w_idx], ws = twid_shoup[w_idx];
    for (size_t j = 0; j < t; j++) {
        pair32 r = butterfly(data[j], data[j + t], w, ws, p, neg_p, two_p);
        data[j] = r.a; data[j + t] = r.b;
    }
    fwd_depth_first32(p, data, n / 2, twid, twid_shoup, depth + 1, half * 2, butterfly, last_butterfly);
    fwd_depth_first32(p, data + n / 2, n / 2, twid, twid_shoup, depth + 1, half * 2 + 1, butterfly, last_butterfly);
}

static void inv_breadth_first32(u32 p, u32 *data, size_t n, const u32 *twid, const u32 *twid_shoup,
                                size_t depth, size_t half, bf32_fn butterfly, bf32_fn last_butterfly)
{
    size_t t = 1, m = n, w_idx = (m << depth) + half * m;
    u32 neg_p = (u32)0 - p, two_p = 2 * p;
    while (m > 1) {
        m /= 2;
        w_idx /= 2;
        const u32 *w = twid + w_idx, *ws = twid_shoup + w_idx;
        bf32_fn bf = (m == 1) ? last_butterfly : butterfly;
        for (size_t i = 0; i < m; i++) {
            u32 *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                pair32 r = bf(z0[j], z1[j], w[i], ws[i], p, neg_p, two_p);
                z0[j] = r.a; z1[j] = r.b;
            }
        }
        t *= 2;
    }
}

static void inv_depth_first32(u32 p, u32 *data, size_t n, const u32 *twid, const u32 *twid_shoup,
                              size_t depth, size_t half, bf32_fn butterfly, bf32_fn last_butterfly)
{
    if (n <= RECURSION_THRESHOLD_32) {
        inv_breadth_first32(p, data, n, twid, twid_shoup, depth, half, butterfly, last_butterfly);
        return;
    }
    /* inner calls get `butterfly` for both roles: src/prime32/shoup.rs:1436-1455 */
    inv_depth_first32(p, data, n / 2, twid, twid_shoup, depth + 1, half * 2, butterfly, butterfly);
    inv_depth_first32(p, data + n / 2, n / 2, twid, twid_shoup, depth + 1, half * 2 + 1, butterfly, butterfly);
    size_t t = n / 2, w_idx = ((size_t)1 << depth) + half;
    u32 neg_p = (u32)0 - p, two_p = 2 * p;
    u32 w = twid[w_idx], ws = twid_shoup[w_idx];
    for (size_t j = 0; j < t; j++) {
        pair32 r = last_butterfly(data[j], data[j + t], w, ws, p, neg_p, two_p);
        data[j] = r.a; data[j + t] = r.b;
    }
}

/* ---- generic p >= 2^31 : src/prime32/generic.rs:9-31 (add/sub/mul), 228-391 (drivers) ---- */
static inline u32 gadd32(u32 p, u32 a, u32 b) { u32 neg_b = p - b; return a >= neg_b ? a - neg_b : a + b; }
static inline u32 gsub32(u32 p, u32 a, u32 b) { u32 neg_b = p - b; return a >= b ? a - b : a + neg_b; }
static inline u32 gmul32(u32 p, u32 a, u32 b) { return (u32)(((u64)a * b) % p); }

static void gfwd_breadth_first32(u32 *data, size_t n, u32 p, const u32 *twid, size_t depth, size_t half)
{
    size_t t = n / 2, m = 1, w_idx = (m << depth) + half * m;
    while (m < n) {
        const u32 *w = twid + w_idx;
        for (size_t i = 0; i < m; i++) {
            u32 *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                u32 z1w = gmul32(p, z1[j], w[i]);
                u32 a = gadd32(p, z0[j], z1w), b = gsub32(p, z0[j], z1w);
                z0[j] = a; z1[j] = b;
            }
        }
        t /= 2; m *= 2; w_idx *= 2;
    }
}
static void gfwd_depth_first32(u32 *data, size_t n, u32 p, const u32 *twid, size_t depth, size_t half)
{
    if (n <= RECURSION_THRESHOLD_32) { gfwd_breadth_first32(data, n, p, twid, depth, half); return; }
    size_t t = n / 2;
    u32 w1 = twid[((size_t)1 << depth) + half];
    for (size_t j = 0; j < t; j++) {
        u32 z1w = gmul32(p, data[j + t], w1);
        u32 a = gadd32(p, data[j], z1w), b = gsub32(p, data[j], z1w);
        data[j] = a; data[j + t] = b;
    }
    gfwd_depth_first32(data, n / 2, p, twid, depth + 1, half * 2);
    gfwd_depth_first32(data + n / 2, n / 2, p, twid, depth + 1, half * 2 + 1);
}
static void ginv_breadth_first32(u32 *data, size_t n, u32 p, const u32 *inv_twid, size_t depth, size_t half)
{
    size_t t = 1, m = n, w_idx = (m << depth) + half * m;
    while (m > 1) {
        m /= 2; w_idx /= 2;
        const u32 *w = inv_twid + w_idx;
        for (size_t i = 0; i < m; i++) {
            u32 *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                u32 a = gadd32(p, z0[j], z1[j]), b = gmul32(p, gsub32(p, z0[j], z1[j]), w[i]);
                z0[j] = a; z1[j] = b;
            }
        }
        t *= 2;
    }
}
static void ginv_depth_first32(u32 *data, size_t n, u32 p, const u32 *inv_twid, size_t depth, size_t half)
{
    if (n <= RECURSION_THRESHOLD_32) { ginv_breadth_first32(data, n, p, inv_twid, depth, half); return; }
    ginv_depth_first32(data, n / 2, p, inv_twid, depth + 1, half * 2);
    ginv_depth_first32(data + n / 2, n / 2, p, inv_twid, depth + 1, half * 2 + 1);
    size_t t = n / 2;
    u32 w1 = inv_twid[((size_t)1 << depth) + half];
    for (size_t j = 0; j < t; j++) {
        u32 a = gadd32(p, data[j], data[j + t]), b = gmul32(p, gsub32(p, data[j], data[j + t]), w1);
        data[j] = a; data[j + t] = b;
    }
}

/* Plan::fwd / Plan::inv dispatch: src/prime32.rs:709-755, 762-808 (scalar arms) */
EXPORT void o_plan32_fwd(const o_plan32 *pl, u32 *buf)
{
    u32 p = pl->p;
    if (p < ((u32)1 << 30))
        fwd_depth_first32(p, buf, pl->n, pl->twid, pl->twid_shoup, 0, 0, fwd_bf30, fwd_last_bf30);
    else if (p < ((u32)1 << 31))
        fwd_depth_first32(p, buf, pl->n, pl->twid, pl->twid_shoup, 0, 0, fwd_bf31, fwd_last_bf31);
    else
        gfwd_depth_first32(buf, pl->n, p, pl->twid, 0, 0);
}
EXPORT void o_plan32_inv(const o_plan32 *pl, u32 *buf)
{
    u32 p = pl->p;
    if (p < ((u32)1 << 30))
        inv_depth_first32(p, buf, pl->n, pl->inv_twid, pl->inv_twid_shoup, 0, 0, inv_bf30, inv_last_bf30);
    else if (p < ((u32)1 << 31)) /* less_than_31bit.rs:362-380: inv_butterfly for both roles */
        inv_depth_first32(p, buf, pl->n, pl->inv_twid, pl->inv_twid_shoup, 0, 0, inv_bf31, inv_bf31);
    else
        ginv_depth_first32(buf, pl->n, p, pl->inv_twid, 0, 0);
}

/* Pointwise ops: src/prime32.rs:383-408 (mul_assign_normalize_scalar), 477-486 (normalize_scalar),
 * 575-598 (mul_accumulate_scalar), and the p >= 2^31 arms at 853-863, 893-901, 919-926.
 * `len` restates the zip-truncation: callers pass min(lhs.len(), rhs.len()). */
EXPORT void o_plan32_mul_assign_normalize(const o_plan32 *pl, u32 *lhs, const u32 *rhs, size_t len)
{
    u32 p = pl->p;
    if (p < ((u32)1 << 31)) {
        u32 big_q_m1 = pl->big_q - 1;
        for (size_t i = 0; i < len; i++) {
            u64 d = (u64)lhs[i] * rhs[i];
            u32 c1 = (u32)(d >> big_q_m1);
            u32 c3 = (u32)(((u64)c1 * pl->p_barrett) >> 32);
            u32 prod = (u32)d - p * c3;
            u32 shoup_q = (u32)(((u64)prod * pl->n_inv_mod_p_shoup) >> 32);
            u32 t = prod * pl->n_inv_mod_p - shoup_q * p;
            lhs[i] = min32(t, t - p);
        }
    } else {
        for (size_t i = 0; i < len; i++) {
            u32 prod = gmul32(p, lhs[i], rhs[i]);
            lhs[i] = gmul32(p, prod, pl->n_inv_mod_p);
        }
    }
}
EXPORT void o_plan32_normalize(const o_plan32 *pl, u32 *values, size_t len)
{
    u32 p = pl->p;
    if (p < ((u32)1 << 31)) {
        for (size_t i = 0; i < len; i++) {
            u32 val = values[i];
            u32 shoup_q = (u32)(((u64)val * pl->n_inv_mod_p_shoup) >> 32);
            u32 t = val * pl->n_inv_mod_p - shoup_q * p;
            values[i] = min32(t, t - p);
        }
    } else {
        for (size_t i = 0; i < len; i++) values[i] = gmul32(p, values[i], pl->n_inv_mod_p);
    }
}
EXPORT void o_plan32_mul_accumulate(const o_plan32 *pl, u32 *acc, const u32 *lhs, const u32 *rhs, size_t len)
{
    u32 p = pl->p;
    if (p < ((u32)1 << 31)) {
        u32 big_q_m1 = pl->big_q - 1;
        for (size_t i = 0; i < len; i++) {
            u64 d = (u64)lhs[i] * rhs[i];
            u32 c1 = (u32)(d >> big_q_m1);
            u32 c3 = (u32)(((u64)c1 * pl->p_barrett) >> 32);
            u32 prod = (u32)d - p * c3;
            prod = min32(prod, prod - p);
            u32 acc_ = prod + acc[i];
            acc[i] = min32(acc_, acc_ - p);
        }
    } else {
        for (size_t i = 0; i < len; i++) acc[i] = gadd32(p, acc[i], gmul32(p, lhs[i], rhs[i]));
    }
}

/* ==========================================================================================
 * prime64::Plan   (src/prime64.rs)
 * ========================================================================================== */

#define SOLINAS_P 0xFFFFFFFF00000001ull /* src/prime64/generic_solinas.rs:39 */

typedef struct {
    size_t n;
    u64 p;
    u64 *twid, *twid_shoup, *inv_twid, *inv_twid_shoup; /* shoup tables NULL when p >= 2^63 */
    u64 p_barrett, big_q, n_inv_mod_p, n_inv_mod_p_shoup;
    u64 psi;
} o_plan64;

EXPORT int o_plan64_new(size_t n, u64 p, o_plan64 **out)
{
    *out = NULL;
    if (p <= 1) return 2;
    /* src/prime64.rs:709-713 */
    u64 psi;
    if (n < 16 || (n & (n - 1)) != 0 || !o_is_prime64(p) || !o_find_primitive_root64(p, 2 * (u64)n, &psi))
        return 1;
    o_plan64 *pl = (o_plan64 *)calloc(1, sizeof(*pl));
    pl->n = n; pl->p = p; pl->psi = psi;
    pl->twid = (u64 *)malloc(n * sizeof(u64));
    pl->inv_twid = (u64 *)malloc(n * sizeof(u64));
    int shoup = p < ((u64)1 << 63);
    const unsigned bits = 64; /* has_ifma == false on the scalar path: src/prime64.rs:718-725 */
    if (shoup) {
        pl->twid_shoup = (u64 *)malloc(n * sizeof(u64));
        pl->inv_twid_shoup = (u64 *)malloc(n * sizeof(u64));
    }
    /* src/prime64.rs:158-218 */
    unsigned nbits = (unsigned)__builtin_ctzll(n);
    u64 w = psi, wk = 1;
    for (size_t k = 0; k < n; k++) {
        size_t fwd_idx = bit_rev(nbits, k);
        u64 wk_shoup = (u64)((((u128)wk) << bits) / p);
        pl->twid[fwd_idx] = wk;
        if (shoup) pl->twid_shoup[fwd_idx] = wk_shoup;
        size_t inv_idx = bit_rev(nbits, (n - k) % n);
        if (k == 0) {
            pl->inv_twid[inv_idx] = wk;
            if (shoup) pl->inv_twid_shoup[inv_idx] = wk_shoup;
        } else {
            u64 x = p - wk;
            pl->inv_twid[inv_idx] = x;
            if (shoup) pl->inv_twid_shoup[inv_idx] = (u64)((((u128)x) << bits) / p);
        }
        wk = mul_mod64(p, wk, w);
    }
    /* src/prime64.rs:752-756 */
    pl->n_inv_mod_p = exp_mod64(p, (u64)n % p, p - 2);
    pl->n_inv_mod_p_shoup = (u64)((((u128)pl->n_inv_mod_p) << bits) / p);
    pl->big_q = ilog2_u64(p) + 1;
    u64 big_l = pl->big_q + (bits - 1);
    /* for p >= 2^63 the shift is 127: 1u128 << 127 is fine; the field is unused on that path */
    pl->p_barrett = (u64)((((u128)1) << big_l) / p);
    *out = pl;
    return 0;
}
EXPORT void o_plan64_free(o_plan64 *pl)
{
    if (!pl) return;
    free(pl->twid); free(pl->twid_shoup); free(pl->inv_twid); free(pl->inv_twid_shoup);
    free(pl);
}
EXPORT size_t o_plan64_ntt_size(const o_plan64 *pl) { return pl->n; }
EXPORT u64 o_plan64_modulus(const o_plan64 *pl) { return pl->p; }
EXPORT u64 o_plan64_psi(const o_plan64 *pl) { return pl->psi; }
EXPORT const u64 *o_plan64_twid(const o_plan64 *pl) { return pl->twid; }
EXPORT const u64 *o_plan64_inv_twid(const o_plan64 *pl) { return pl->inv_twid; }

typedef struct { u64 a, b; } pair64;
typedef pair64 (*bf64_fn)(u64 z0, u64 z1, u64 w, u64 w_shoup, u64 p, u64 neg_p, u64 two_p);
static inline u64 mulhi64(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

/* p < 2^62: src/prime64/less_than_62bit.rs:117-157, 271-311 */
static pair64 fwd_bf62(u64 z0, u64 z1, u64 w, u64 ws, u64 p, u64 neg_p, u64 two_p)
{
    (void)p;
    z0 = min64(z0, z0 - two_p);
    u64 q = mulhi64(z1, ws);
    u64 t = z1 * w + q * neg_p;
    return (pair64){z0 + t, z0 - t + two_p};
}
static pair64 fwd_last_bf62(u64 z0, u64 z1, u64 w, u64 ws, u64 p, u64 neg_p, u64 two_p)
{
    z0 = min64(z0, z0 - two_p);
    z0 = min64(z0, z0 - p);
    u64 q = mulhi64(z1, ws);
    u64 t = z1 * w + q * neg_p;
    t = min64(t, t - p);
    u64 r0 = z0 + t, r1 = z0 - t + p;
    return (pair64){min64(r0, r0 - p), min64(r1, r1 - p)};
}
static pair64 inv_bf62(u64 z0, u64 z1, u64 w, u64 ws, u64 p, u64 neg_p, u64 two_p)
{
    (void)p;
    u64 y0 = z0 + z1;
    y0 = min64(y0, y0 - two_p);
    u64 t = z0 - z1 + two_p;
    u64 q = mulhi64(t, ws);
    u64 y1 = t * w + q * neg_p;
    return (pair64){y0, y1};
}
static pair64 inv_last_bf62(u64 z0, u64 z1, u64 w, u64 ws, u64 p, u64 neg_p, u64 two_p)
{
    u64 y0 = z0 + z1;
    y0 = min64(y0, y0 - two_p);
    y0 = min64(y0, y0 - p);
    u64 t = z0 - z1 + two_p;
    u64 q = mulhi64(t, ws);
    u64 y1 = t * w + q * neg_p;
    y1 = min64(y1, y1 - p);
    return (pair64){y0, y1};
}
/* 2^62 <= p < 2^63: src/prime64/less_than_63bit.rs:117-157, 214-234 */
static pair64 fwd_bf63(u64 z0, u64 z1, u64 w, u64 ws, u64 p, u64 neg_p, u64 two_p)
{
    (void)two_p;
    z0 = min64(z0, z0 - p);
    u64 q = mulhi64(z1, ws);
    u64 t = z1 * w + q * neg_p;
    t = min64(t, t - p);
    return (pair64){z0 + t, z0 - t + p};
}
static pair64 fwd_last_bf63(u64 z0, u64 z1, u64 w, u64 ws, u64 p, u64 neg_p, u64 two_p)
{
    (void)two_p;
    z0 = min64(z0, z0 - p);
    u64 q = mulhi64(z1, ws);
    u64 t = z1 * w + q * neg_p;
    t = min64(t, t - p);
    u64 r0 = z0 + t, r1 = z0 - t + p;
    return (pair64){min64(r0, r0 - p), min64(r1, r1 - p)};
}
static pair64 inv_bf63(u64 z0, u64 z1, u64 w, u64 ws, u64 p, u64 neg_p, u64 two_p)
{
    (void)two_p;
    u64 y0 = z0 + z1;
    y0 = min64(y0, y0 - p);
    u64 t = z0 - z1 + p;
    u64 q = mulhi64(t, ws);
    u64 y1 = t * w + q * neg_p;
    y1 = min64(y1, y1 - p);
    return (pair64){y0, y1};
}

/* Shoup-form drivers: src/prime64/shoup.rs:544-706 (fwd scalar), 1306-1468 (inv scalar) */
static void fwd_breadth_first64(u64 p, u64 *data, size_t n, const u64 *twid, const u64 *twid_shoup,
                                size_t depth, size_t half, bf64_fn butterfly, bf64_fn last_butterfly)
{
    size_t t = n, m = 1, w_idx = (m << depth) + half * m;
    u64 neg_p = (u64)0 - p, two_p = 2 * p;
    while (m < n) {
        t /= 2;
        const u64 *w = twid + w_idx, *ws = twid_shoup + w_idx;
        bf64_fn bf = (t == 1) ? last_butterfly : butterfly;
        for (size_t i = 0; i < m; i++) {
            u64 *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                pair64 r = bf(z0[j], z1[j], w[i], ws[i], p, neg_p, two_p);
                z0[j] = r.a; z1[j] = r.b;
            }
        }
        m *= 2; w_idx *= 2;
    }
}
static void fwd_depth_first64(u64 p, u64 *data, size_t n, const u64 *twid, const u64 *twid_shoup,
                              size_t depth, size_t half, bf64_fn butterfly, bf64_fn last_butterfly)
{
    if (n <= RECURSION_THRESHOLD_64) {
        fwd_breadth_first64(p, data, n, twid, twid_shoup, depth, half, butterfly, last_butterfly);
        return;
    }
    size_t t = n / 2, w_idx = ((size_t)1 << depth) + half;
    u64 neg_p = (u64)0 - p, two_p = 2 * p;
    u64 w = twid[w_idx], ws = twid_shoup[w_idx];
    for (size_t j = 0; j < t; j++) {
        pair64 r = butterfly(data[j], data[j + t], w, ws, p, neg_p, two_p);
        data[j] = r.a; data[j + t] = r.b;
    }
    fwd_depth_first64(p, data, n / 2, twid, twid_shoup, depth + 1, half * 2, butterfly, last_butterfly);
    fwd_depth_first64(p, data + n / 2, n / 2, twid, twid_shoup, depth + 1, half * 2 + 1, butterfly, last_butterfly);
}
static void inv_breadth_first64(u64 p, u64 *data, size_t n, const u64 *twid, const u64 *twid_shoup,
                                size_t depth, size_t half, bf64_fn butterfly, bf64_fn last_butterfly)
{
    size_t t = 1, m = n, w_idx = (m << depth) + half * m;
    u64 neg_p = (u64)0 - p, two_p = 2 * p;
    while (m > 1) {
        m /= 2; w_idx /= 2;
        const u64 *w = twid + w_idx, *ws = twid_shoup + w_idx;
        bf64_fn bf = (m == 1) ? last_butterfly : butterfly;
        for (size_t i = 0; i < m; i++) {
            u64 *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                pair64 r = bf(z0[j], z1[j], w[i], ws[i], p, neg_p, two_p);
                z0[j] = r.a; z1[j] = r.b;
            }
        }
        t *= 2;
    }
}
static void inv_depth_first64(u64 p, u64 *data, size_t n, const u64 *twid, const u64 *twid_shoup,
                              size_t depth, size_t half, bf64_fn butterfly, bf64_fn last_butterfly)
{
    if (n <= RECURSION_THRESHOLD_64) {
        inv_breadth_first64(p, data, n, twid, twid_shoup, depth, half, butterfly, last_butterfly);
        return;
    }
    inv_depth_first64(p, data, n / 2, twid, twid_shoup, depth + 1, half * 2, butterfly, butterfly);
    inv_depth_first64(p, data + n / 2, n / 2, twid, twid_shoup, depth + 1, half * 2 + 1, butterfly, butterfly);
    size_t t = n / 2, w_idx = ((size_t)1 << depth) + half;
    u64 neg_p = (u64)0 - p, two_p = 2 * p;
    u64 w = twid[w_idx], ws = twid_shoup[w_idx];
    for (size_t j = 0; j < t; j++) {
        pair64 r = last_butterfly(data[j], data[j + t], w, ws, p, neg_p, two_p);
        data[j] = r.a; data[j + t] = r.b;
    }
}

/* PrimeModulus for u64 and Solinas: src/prime64/generic_solinas.rs:42-129 */
static inline u64 gadd64(u64 p, u64 a, u64 b) { u64 neg_b = p - b; return a >= neg_b ? a - neg_b : a + b; }
static inline u64 gsub64(u64 p, u64 a, u64 b) { u64 neg_b = p - b; return a >= b ? a - b : a + neg_b; }
static inline u64 solinas_mul(u64 a, u64 b)
{
    const u64 p = SOLINAS_P;
    u128 wide = (u128)a * b;
    u64 lo = (u64)wide;
    u64 hi = (u64)(wide >> 64);
    u64 mid = hi & 0x00000000FFFFFFFFull;
    hi = (hi & 0xFFFFFFFF00000000ull) >> 32;
    u64 low2 = lo - hi;
    if (hi > lo) low2 += p;
    u64 product = mid << 32;
    product -= mid;
    u64 result = low2 + product;
    if (result < product || result >= p) result -= p;
    return result;
}
static inline u64 gmul64(u64 p, int solinas, u64 a, u64 b) { return solinas ? solinas_mul(a, b) : mul_mod64(p, a, b); }

/* generic drivers: src/prime64/generic_solinas.rs:449-562 (+ fwd_depth_first_scalar 1338-1388).
 * The reference monomorphises them per PrimeModulus impl (u64 / Solinas); the macro below does the
 * same so the Solinas multiply is inlined into the loops like in the Rust build. */
#define DEFINE_GENERIC64(SUF, MUL)                                                                              \
    static void gfwd_breadth_first64##SUF(u64 *data, size_t n, u64 p, const u64 *twid, size_t depth, size_t half) \
    {                                                                                                           \
        size_t t = n / 2, m = 1, w_idx = (m << depth) + half * m;                                               \
        while (m < n) {                                                                                         \
            const u64 *w = twid + w_idx;                                                                        \
            for (size_t i = 0; i < m; i++) {                                                                    \
                u64 *z0 = data + 2 * i * t, *z1 = z0 + t;                                                       \
                const u64 w1 = w[i];                                                                            \
                for (size_t j = 0; j < t; j++) {                                                                \
                    u64 z1w = MUL(p, z1[j], w1);                                                                \
                    u64 a = gadd64(p, z0[j], z1w), b = gsub64(p, z0[j], z1w);                                   \
                    z0[j] = a; z1[j] = b;                                                                       \
                }                                                                                               \
            }                                                                                                   \
            t /= 2; m *= 2; w_idx *= 2;                                                                         \
        }                                                                                                       \
    }                                                                                                           \
    static void gfwd_depth_first64##SUF(u64 *data, size_t n, u64 p, const u64 *twid, size_t depth, size_t half) \
    {                                                                                                           \
        if (n <= RECURSION_THRESHOLD_64) { gfwd_breadth_first64##SUF(data, n, p, twid, depth, half); return; }   \
        size_t t = n / 2;                                                                                       \
        u64 w1 = twid[((size_t)1 << depth) + half];                                                             \
        for (size_t j = 0; j < t; j++) {                                                                        \
            u64 z1w = MUL(p, data[j + t], w1);                                                                  \
            u64 a = gadd64(p, data[j], z1w), b = gsub64(p, data[j], z1w);                                       \
            data[j] = a; data[j + t] = b;                                                                       \
        }                                                                                                       \
        gfwd_depth_first64##SUF(data, n / 2, p, twid, depth + 1, half * 2);                                     \
        gfwd_depth_first64##SUF(data + n / 2, n / 2, p, twid, depth + 1, half * 2 + 1);                         \
    }                                                                                                           \
    static void ginv_breadth_first64##SUF(u64 *data, size_t n, u64 p, const u64 *inv_twid, size_t depth, size_t half) \
    {                                                                                                           \
        size_t t = 1, m = n, w_idx = (m << depth) + half * m;                                                   \
        while (m > 1) {                                                                                         \
            m /= 2; w_idx /= 2;                                                                                 \
            const u64 *w = inv_twid + w_idx;                                                                    \
            for (size_t i = 0; i < m; i++) {                                                                    \
                u64 *z0 = data + 2 * i * t, *z1 = z0 + t;                                                       \
                const u64 w1 = w[i];                                                                            \
                for (size_t j = 0; j < t; j++) {                                                                \
                    u64 a = gadd64(p, z0[j], z1[j]), b = MUL(p, gsub64(p, z0[j], z1[j]), w1);                   \
                    z0[j] = a; z1[j] = b;                                                                       \
                }                                                                                               \
            }                                                                                                   \
            t *= 2;                                                                                             \
        }                                                                                                       \
    }                                                                                                           \
    static void ginv_depth_first64##SUF(u64 *data, size_t n, u64 p, const u64 *inv_twid, size_t depth, size_t half) \
    {                                                                                                           \
        if (n <= RECURSION_THRESHOLD_64) { ginv_breadth_first64##SUF(data, n, p, inv_twid, depth, half); return; } \
        ginv_depth_first64##SUF(data, n / 2, p, inv_twid, depth + 1, half * 2);                                 \
        ginv_depth_first64##SUF(data + n / 2, n / 2, p, inv_twid, depth + 1, half * 2 + 1);                     \
        size_t t = n / 2;                                                                                       \
        u64 w1 = inv_twid[((size_t)1 << depth) + half];                                                         \
        for (size_t j = 0; j < t; j++) {                                                                        \
            u64 a = gadd64(p, data[j], data[j + t]), b = MUL(p, gsub64(p, data[j], data[j + t]), w1);           \
            data[j] = a; data[j + t] = b;                                                                       \
        }                                                                                                       \
    }
#define MUL_GENERIC(p, a, b) mul_mod64((p), (a), (b))
#define MUL_SOLINAS(p, a, b) solinas_mul((a), (b))
DEFINE_GENERIC64(_u64, MUL_GENERIC)
DEFINE_GENERIC64(_solinas, MUL_SOLINAS)

/* Plan::fwd / inv dispatch (scalar, non-nightly arms): src/prime64.rs:794-865, 872-943 */
EXPORT void o_plan64_fwd(const o_plan64 *pl, u64 *buf)
{
    u64 p = pl->p;
    if (p < ((u64)1 << 62))
        fwd_depth_first64(p, buf, pl->n, pl->twid, pl->twid_shoup, 0, 0, fwd_bf62, fwd_last_bf62);
    else if (p < ((u64)1 << 63))
        fwd_depth_first64(p, buf, pl->n, pl->twid, pl->twid_shoup, 0, 0, fwd_bf63, fwd_last_bf63);
    else if (p == SOLINAS_P)
        gfwd_depth_first64_solinas(buf, pl->n, p, pl->twid, 0, 0);
    else
        gfwd_depth_first64_u64(buf, pl->n, p, pl->twid, 0, 0);
}
EXPORT void o_plan64_inv(const o_plan64 *pl, u64 *buf)
{
    u64 p = pl->p;
    if (p < ((u64)1 << 62))
        inv_depth_first64(p, buf, pl->n, pl->inv_twid, pl->inv_twid_shoup, 0, 0, inv_bf62, inv_last_bf62);
    else if (p < ((u64)1 << 63))
        inv_depth_first64(p, buf, pl->n, pl->inv_twid, pl->inv_twid_shoup, 0, 0, inv_bf63, inv_bf63);
    else if (p == SOLINAS_P)
        ginv_depth_first64_solinas(buf, pl->n, p, pl->inv_twid, 0, 0);
    else
        ginv_depth_first64_u64(buf, pl->n, p, pl->inv_twid, 0, 0);
}

/* Pointwise: src/prime64.rs:534-584 (scalar Barrett+Shoup), 690-699 (normalize_scalar),
 * Solinas / generic arms 1013-1032, 1068-1082, 1116-1127. */
EXPORT void o_plan64_mul_assign_normalize(const o_plan64 *pl, u64 *lhs, const u64 *rhs, size_t len)
{
    u64 p = pl->p;
    if (p < ((u64)1 << 63)) {
        u64 big_q_m1 = pl->big_q - 1;
        for (size_t i = 0; i < len; i++) {
            u128 d = (u128)lhs[i] * rhs[i];
            u64 c1 = (u64)(d >> big_q_m1);
            u64 c3 = mulhi64(c1, pl->p_barrett);
            u64 prod = (u64)d - p * c3;
            u64 shoup_q = mulhi64(prod, pl->n_inv_mod_p_shoup);
            u64 t = prod * pl->n_inv_mod_p - shoup_q * p;
            lhs[i] = min64(t, t - p);
        }
    } else {
        int sol = p == SOLINAS_P;
        for (size_t i = 0; i < len; i++) {
            u64 prod = gmul64(p, sol, lhs[i], rhs[i]);
            lhs[i] = gmul64(p, sol, prod, pl->n_inv_mod_p);
        }
    }
}
EXPORT void o_plan64_normalize(const o_plan64 *pl, u64 *values, size_t len)
{
    u64 p = pl->p;
    if (p < ((u64)1 << 63)) {
        for (size_t i = 0; i < len; i++) {
            u64 val = values[i];
            u64 shoup_q = mulhi64(val, pl->n_inv_mod_p_shoup);
            u64 t = val * pl->n_inv_mod_p - shoup_q * p;
            values[i] = min64(t, t - p);
        }
    } else {
        int sol = p == SOLINAS_P;
        for (size_t i = 0; i < len; i++) values[i] = gmul64(p, sol, values[i], pl->n_inv_mod_p);
    }
}
EXPORT void o_plan64_mul_accumulate(const o_plan64 *pl, u64 *acc, const u64 *lhs, const u64 *rhs, size_t len)
{
    u64 p = pl->p;
    if (p < ((u64)1 << 63)) {
        u64 big_q_m1 = pl->big_q - 1;
        for (size_t i = 0; i < len; i++) {
            u128 d = (u128)lhs[i] * rhs[i];
            u64 c1 = (u64)(d >> big_q_m1);
            u64 c3 = mulhi64(c1, pl->p_barrett);
            u64 prod = (u64)d - p * c3;
            prod = min64(prod, prod - p);
            u64 acc_ = prod + acc[i];
            acc[i] = min64(acc_, acc_ - p);
        }
    } else {
        int sol = p == SOLINAS_P;
        for (size_t i = 0; i < len; i++) acc[i] = gadd64(p, acc[i], gmul64(p, sol, lhs[i], rhs[i]));
    }
}

/* ==========================================================================================
 * primes32 constants: src/lib.rs:453-462 and the derived Garner constants 512-594.
 * Derived at load time with the same formulas (const fn in the reference).
 * ========================================================================================== */
static const u32 P32[10] = {
    0x3F5A0001u, 0x3F5D0001u, 0x3F760001u, 0x3F820001u, 0x3FAC0001u,
    0x3FAF0001u, 0x3FB10001u, 0x3FBB0001u, 0x3FDE0001u, 0x3FFC0001u,
};
#define P0 (P32[0])
#define P1 (P32[1])
#define P2 (P32[2])
#define P3 (P32[3])
#define P4 (P32[4])
#define P5 (P32[5])
#define P6 (P32[6])
#define P7 (P32[7])
#define P8 (P32[8])
#define P9 (P32[9])

static struct {
    int ready;
    u32 P0_INV_MOD_P1, P01_INV_MOD_P2, P1_INV_MOD_P2, P3_INV_MOD_P4;
    u32 P2_INV_MOD_P3, P4_INV_MOD_P5, P6_INV_MOD_P7, P8_INV_MOD_P9;
    u64 P12, P34, P0_INV_MOD_P12, P0_INV_MOD_P12_SHOUP, P0_MOD_P34_SHOUP, P012_INV_MOD_P34, P012_INV_MOD_P34_SHOUP;
    u64 P01, P23, P45, P67, P89;
    u64 P01_MOD_P45_SHOUP, P01_MOD_P67_SHOUP, P01_MOD_P89_SHOUP, P23_MOD_P67_SHOUP, P23_MOD_P89_SHOUP, P45_MOD_P89_SHOUP;
    u64 P01_INV_MOD_P23, P01_INV_MOD_P23_SHOUP, P0123_INV_MOD_P45, P0123_INV_MOD_P45_SHOUP;
    u64 P012345_INV_MOD_P67, P012345_INV_MOD_P67_SHOUP, P01234567_INV_MOD_P89, P01234567_INV_MOD_P89_SHOUP;
    u128 P0123, P012345, P01234567, P0123456789;
} C;

static u32 inv_mod32(u32 m, u32 x) { return (u32)exp_mod64(m, x, (u64)m - 2); }        /* src/lib.rs:491-493 */
static u32 cmul_mod32(u32 m, u32 a, u32 b) { return (u32)(((u64)a * b) % m); }        /* src/lib.rs:486-489 */
static u64 shoup64c(u64 m, u64 w) { return (u64)((((u128)w) << 64) / m); }            /* src/lib.rs:508-510 */

static void consts_init(void)
{
    if (C.ready) return;
    C.P0_INV_MOD_P1 = inv_mod32(P1, P0);                                             /* lib.rs:512 */
    C.P01_INV_MOD_P2 = inv_mod32(P2, cmul_mod32(P2, P0, P1));                        /* lib.rs:514 */
    C.P1_INV_MOD_P2 = inv_mod32(P2, P1);                                             /* lib.rs:529 */
    C.P3_INV_MOD_P4 = inv_mod32(P4, P3);                                             /* lib.rs:531 */
    C.P12 = (u64)P1 * P2;                                                            /* lib.rs:533 */
    C.P34 = (u64)P3 * P4;
    C.P0_INV_MOD_P12 = exp_mod64(C.P12, P0, ((u64)P1 - 1) * ((u64)P2 - 1) - 1);      /* lib.rs:535-536 */
    C.P0_INV_MOD_P12_SHOUP = shoup64c(C.P12, C.P0_INV_MOD_P12);
    C.P0_MOD_P34_SHOUP = shoup64c(C.P34, P0);
    C.P012_INV_MOD_P34 = exp_mod64(C.P34, mul_mod64(C.P34, P0, C.P12), ((u64)P3 - 1) * ((u64)P4 - 1) - 1);
    C.P012_INV_MOD_P34_SHOUP = shoup64c(C.P34, C.P012_INV_MOD_P34);
    C.P2_INV_MOD_P3 = inv_mod32(P3, P2);                                             /* lib.rs:546-553 */
    C.P4_INV_MOD_P5 = inv_mod32(P5, P4);
    C.P6_INV_MOD_P7 = inv_mod32(P7, P6);
    C.P8_INV_MOD_P9 = inv_mod32(P9, P8);
    C.P01 = (u64)P0 * P1; C.P23 = (u64)P2 * P3; C.P45 = (u64)P4 * P5;                 /* lib.rs:555-559 */
    C.P67 = (u64)P6 * P7; C.P89 = (u64)P8 * P9;
    C.P01_MOD_P45_SHOUP = shoup64c(C.P45, C.P01);                                    /* lib.rs:561-568 */
    C.P01_MOD_P67_SHOUP = shoup64c(C.P67, C.P01);
    C.P01_MOD_P89_SHOUP = shoup64c(C.P89, C.P01);
    C.P23_MOD_P67_SHOUP = shoup64c(C.P67, C.P23);
    C.P23_MOD_P89_SHOUP = shoup64c(C.P89, C.P23);
    C.P45_MOD_P89_SHOUP = shoup64c(C.P89, C.P45);
    C.P01_INV_MOD_P23 = exp_mod64(C.P23, C.P01, ((u64)P2 - 1) * ((u64)P3 - 1) - 1);   /* lib.rs:570-589 */
    C.P01_INV_MOD_P23_SHOUP = shoup64c(C.P23, C.P01_INV_MOD_P23);
    C.P0123_INV_MOD_P45 = exp_mod64(C.P45, mul_mod64(C.P45, C.P01, C.P23), ((u64)P4 - 1) * ((u64)P5 - 1) - 1);
    C.P0123_INV_MOD_P45_SHOUP = shoup64c(C.P45, C.P0123_INV_MOD_P45);
    C.P012345_INV_MOD_P67 = exp_mod64(C.P67, mul_mod64(C.P67, mul_mod64(C.P67, C.P01, C.P23), C.P45),
                                      ((u64)P6 - 1) * ((u64)P7 - 1) - 1);
    C.P012345_INV_MOD_P67_SHOUP = shoup64c(C.P67, C.P012345_INV_MOD_P67);
    C.P01234567_INV_MOD_P89 = exp_mod64(
        C.P89, mul_mod64(C.P89, mul_mod64(C.P89, mul_mod64(C.P89, C.P01, C.P23), C.P45), C.P67),
        ((u64)P8 - 1) * ((u64)P9 - 1) - 1);
    C.P01234567_INV_MOD_P89_SHOUP = shoup64c(C.P89, C.P01234567_INV_MOD_P89);
    C.P0123 = (u128)C.P01 * (u128)C.P23;                                             /* lib.rs:591-594 (wrapping) */
    C.P012345 = C.P0123 * (u128)C.P45;
    C.P01234567 = C.P012345 * (u128)C.P67;
    C.P0123456789 = C.P01234567 * (u128)C.P89;
    C.ready = 1;
}

EXPORT u32 o_primes32(int i) { return P32[i]; }

/* src/native32.rs:21-25 */
static inline u32 n_mul_mod32(u32 p, u32 a, u32 b) { return (u32)(((u64)a * b) % p); }
/* src/native64.rs:36-41 */
static inline u64 n_mul_mod64(u64 p_neg, u64 a, u64 b, u64 b_shoup)
{
    u64 q = mulhi64(a, b_shoup);
    u64 r = a * b + p_neg * q;
    return min64(r, r + p_neg);
}

/* ---- Garner reconstructions ---- */
/* src/native_binary32.rs:21-41 */
EXPORT u32 o_reconstruct_32bit_01(u32 m0, u32 m1)
{
    consts_init();
    u32 v0 = m0;
    u32 v1 = n_mul_mod32(P1, C.P0_INV_MOD_P1, 2 * P1 + m1 - v0);
    int sign = v1 > (P1 / 2);
    u32 _0 = P0, _01 = _0 * P1;
    u32 pos = v0 + v1 * _0;
    u32 neg = pos - _01;
    return sign ? neg : pos;
}
/* src/native32.rs:27-56 */
EXPORT u32 o_reconstruct_32bit_012_u32(u32 m0, u32 m1, u32 m2)
{
    consts_init();
    u32 v0 = m0;
    u32 v1 = n_mul_mod32(P1, C.P0_INV_MOD_P1, 2 * P1 + m1 - v0);
    u32 v2 = n_mul_mod32(P2, C.P01_INV_MOD_P2, 2 * P2 + m2 - (v0 + n_mul_mod32(P2, P0, v1)));
    int sign = v2 > (P2 / 2);
    u32 _0 = P0, _01 = _0 * P1, _012 = _01 * P2;
    u32 pos = v0 + v1 * _0 + v2 * _01;
    u32 neg = pos - _012;
    return sign ? neg : pos;
}
/* src/native_binary64.rs:31-61 */
EXPORT u64 o_reconstruct_32bit_012_u64(u32 m0, u32 m1, u32 m2)
{
    consts_init();
    u32 v0 = m0;
    u32 v1 = n_mul_mod32(P1, C.P0_INV_MOD_P1, 2 * P1 + m1 - v0);
    u32 v2 = n_mul_mod32(P2, C.P01_INV_MOD_P2, 2 * P2 + m2 - (v0 + n_mul_mod32(P2, P0, v1)));
    int sign = v2 > (P2 / 2);
    u64 _0 = P0, _01 = _0 * (u64)P1, _012 = _01 * (u64)P2;
    u64 pos = (u64)v0 + (u64)v1 * _0 + (u64)v2 * _01;
    u64 neg = pos - _012;
    return sign ? neg : pos;
}
/* shared body of src/native64.rs:90-141 and src/native_binary128.rs:12-64 */
static inline void garner_01234(u32 m0, u32 m1, u32 m2, u32 m3, u32 m4, u64 *v0o, u64 *v12o, u64 *v34o)
{
    u64 mod_p12, mod_p34;
    {
        u32 v1 = m1;
        u32 v2 = n_mul_mod32(P2, C.P1_INV_MOD_P2, 2 * P2 + m2 - v1);
        mod_p12 = (u64)v1 + (u64)v2 * (u64)P1;
    }
    {
        u32 v3 = m3;
        u32 v4 = n_mul_mod32(P4, C.P3_INV_MOD_P4, 2 * P4 + m4 - v3);
        mod_p34 = (u64)v3 + (u64)v4 * (u64)P3;
    }
    u64 v0 = m0;
    u64 v12 = n_mul_mod64((u64)0 - C.P12, 2 * C.P12 + mod_p12 - v0, C.P0_INV_MOD_P12, C.P0_INV_MOD_P12_SHOUP);
    u64 v34 = n_mul_mod64((u64)0 - C.P34,
                          2 * C.P34 + mod_p34 - (v0 + n_mul_mod64((u64)0 - C.P34, v12, (u64)P0, C.P0_MOD_P34_SHOUP)),
                          C.P012_INV_MOD_P34, C.P012_INV_MOD_P34_SHOUP);
    *v0o = v0; *v12o = v12; *v34o = v34;
}
/* src/native64.rs:90-141 */
EXPORT u64 o_reconstruct_32bit_01234_u64(u32 m0, u32 m1, u32 m2, u32 m3, u32 m4)
{
    consts_init();
    u64 v0, v12, v34;
    garner_01234(m0, m1, m2, m3, m4, &v0, &v12, &v34);
    int sign = v34 > (C.P34 / 2);
    u64 _0 = P0, _012 = _0 * C.P12, _01234 = _012 * C.P34;
    u64 pos = v0 + v12 * _0 + v34 * _012;
    u64 neg = pos - _01234;
    return sign ? neg : pos;
}
/* src/native_binary128.rs:12-64; result written as two little-endian u64 limbs */
EXPORT void o_reconstruct_32bit_01234_u128(u32 m0, u32 m1, u32 m2, u32 m3, u32 m4, u64 out[2])
{
    consts_init();
    u64 v0, v12, v34;
    garner_01234(m0, m1, m2, m3, m4, &v0, &v12, &v34);
    int sign = v34 > (C.P34 / 2);
    u128 _0 = P0, _012 = _0 * (u128)C.P12, _01234 = _012 * (u128)C.P34;
    u128 pos = (u128)v0 + (u128)v12 * _0 + (u128)v34 * _012;
    u128 neg = pos - _01234;
    u128 r = sign ? neg : pos;
    out[0] = (u64)r; out[1] = (u64)(r >> 64);
}
/* src/native128.rs:19-118 */
EXPORT void o_reconstruct_32bit_0123456789_u128(const u32 m[10], u64 out[2])
{
    consts_init();
    u64 mod_p01, mod_p23, mod_p45, mod_p67, mod_p89;
    { u32 v0 = m[0]; u32 v1 = n_mul_mod32(P1, C.P0_INV_MOD_P1, 2 * P1 + m[1] - v0); mod_p01 = (u64)v0 + (u64)v1 * (u64)P0; }
    { u32 v2 = m[2]; u32 v3 = n_mul_mod32(P3, C.P2_INV_MOD_P3, 2 * P3 + m[3] - v2); mod_p23 = (u64)v2 + (u64)v3 * (u64)P2; }
    { u32 v4 = m[4]; u32 v5 = n_mul_mod32(P5, C.P4_INV_MOD_P5, 2 * P5 + m[5] - v4); mod_p45 = (u64)v4 + (u64)v5 * (u64)P4; }
    { u32 v6 = m[6]; u32 v7 = n_mul_mod32(P7, C.P6_INV_MOD_P7, 2 * P7 + m[7] - v6); mod_p67 = (u64)v6 + (u64)v7 * (u64)P6; }
    { u32 v8 = m[8]; u32 v9 = n_mul_mod32(P9, C.P8_INV_MOD_P9, 2 * P9 + m[9] - v8); mod_p89 = (u64)v8 + (u64)v9 * (u64)P8; }
    u64 n23 = (u64)0 - C.P23, n45 = (u64)0 - C.P45, n67 = (u64)0 - C.P67, n89 = (u64)0 - C.P89;
    u64 v01 = mod_p01;
    u64 v23 = n_mul_mod64(n23, 2 * C.P23 + mod_p23 - v01, C.P01_INV_MOD_P23, C.P01_INV_MOD_P23_SHOUP);
    u64 v45 = n_mul_mod64(n45, 2 * C.P45 + mod_p45 - (v01 + n_mul_mod64(n45, v23, C.P01, C.P01_MOD_P45_SHOUP)),
                          C.P0123_INV_MOD_P45, C.P0123_INV_MOD_P45_SHOUP);
    u64 v67 = n_mul_mod64(
        n67,
        2 * C.P67 + mod_p67 -
            (v01 + n_mul_mod64(n67, v23 + n_mul_mod64(n67, v45, C.P23, C.P23_MOD_P67_SHOUP), C.P01, C.P01_MOD_P67_SHOUP)),
        C.P012345_INV_MOD_P67, C.P012345_INV_MOD_P67_SHOUP);
    u64 v89 = n_mul_mod64(
        n89,
        2 * C.P89 + mod_p89 -
            (v01 + n_mul_mod64(n89,
                               v23 + n_mul_mod64(n89, v45 + n_mul_mod64(n89, v67, C.P45, C.P45_MOD_P89_SHOUP), C.P23,
                                                 C.P23_MOD_P89_SHOUP),
                               C.P01, C.P01_MOD_P89_SHOUP)),
        C.P01234567_INV_MOD_P89, C.P01234567_INV_MOD_P89_SHOUP);
    int sign = v89 > (C.P89 / 2);
    u128 pos = (u128)v01 + (u128)v23 * (u128)C.P01 + (u128)v45 * C.P0123 + (u128)v67 * C.P012345 + (u128)v89 * C.P01234567;
    u128 neg = pos - C.P0123456789;
    u128 r = sign ? neg : pos;
    out[0] = (u64)r; out[1] = (u64)(r >> 64);
}

/* ==========================================================================================
 * native / native_binary plans: tuples of prime32 plans on P0.. (try_new: native32.rs:338-345,
 * native64.rs:933-942, native128.rs:123-137, native_binary32.rs:190-193, native_binary64.rs:345-352,
 * native_binary128.rs:69-78)
 * ========================================================================================== */
typedef struct {
    size_t n;
    int nprimes;
    int word_bits; /* 32, 64, 128 */
    int binary;
    o_plan32 *pl[10];
} o_native;

/* kind: word_bits in {32,64,128}, binary in {0,1}.  status 0 = Some, 1 = None */
EXPORT int o_native_new(size_t n, int word_bits, int binary, o_native **out)
{
    consts_init();
    *out = NULL;
    int np;
    if (!binary) np = word_bits == 32 ? 3 : word_bits == 64 ? 5 : 10;
    else np = word_bits == 32 ? 2 : word_bits == 64 ? 3 : 5;
    o_native *nt = (o_native *)calloc(1, sizeof(*nt));
    nt->n = n; nt->nprimes = np; nt->word_bits = word_bits; nt->binary = binary;
    for (int i = 0; i < np; i++) {
        if (o_plan32_new(n, P32[i], &nt->pl[i]) != 0) {
            for (int j = 0; j < i; j++) o_plan32_free(nt->pl[j]);
            free(nt);
            return 1;
        }
    }
    *out = nt;
    return 0;
}
EXPORT void o_native_free(o_native *nt)
{
    if (!nt) return;
    for (int i = 0; i < nt->nprimes; i++) o_plan32_free(nt->pl[i]);
    free(nt);
}
EXPORT size_t o_native_ntt_size(const o_native *nt) { return nt->n; }
EXPORT int o_native_nprimes(const o_native *nt) { return nt->nprimes; }

static inline u128 load_word(const void *value, int word_bits, size_t i)
{
    if (word_bits == 32) return ((const u32 *)value)[i];
    if (word_bits == 64) return ((const u64 *)value)[i];
    const u64 *v = (const u64 *)value + 2 * i;
    return ((u128)v[1] << 64) | v[0];
}
static inline void store_word(void *value, int word_bits, size_t i, u128 x)
{
    if (word_bits == 32) ((u32 *)value)[i] = (u32)x;
    else if (word_bits == 64) ((u64 *)value)[i] = (u64)x;
    else { u64 *v = (u64 *)value + 2 * i; v[0] = (u64)x; v[1] = (u64)(x >> 64); }
}

/* Plan32::fwd: value % P_i then per-prime fwd.  native32.rs:366-377, native64.rs:971-999,
 * native128.rs:186-251, native_binary32.rs:201-208, native_binary64.rs:360-371, native_binary128.rs:85-112.
 * mod_p: nprimes planes of n u32 each, plane k at mod_p + k*n. */
EXPORT void o_native_fwd(const o_native *nt, const void *value, u32 *mod_p)
{
    size_t n = nt->n;
    for (size_t i = 0; i < n; i++) {
        u128 v = load_word(value, nt->word_bits, i);
        for (int k = 0; k < nt->nprimes; k++) mod_p[(size_t)k * n + i] = (u32)(v % P32[k]);
    }
    for (int k = 0; k < nt->nprimes; k++) o_plan32_fwd(nt->pl[k], mod_p + (size_t)k * n);
}
/* Plan32::fwd_binary: `*value as u32` with no reduction.  native_binary32.rs:210-217,
 * native_binary64.rs:372-389, native_binary128.rs:115-143 */
EXPORT void o_native_fwd_binary(const o_native *nt, const void *value, u32 *mod_p)
{
    size_t n = nt->n;
    for (size_t i = 0; i < n; i++) {
        u32 v = (u32)load_word(value, nt->word_bits, i);
        for (int k = 0; k < nt->nprimes; k++) mod_p[(size_t)k * n + i] = v;
    }
    for (int k = 0; k < nt->nprimes; k++) o_plan32_fwd(nt->pl[k], mod_p + (size_t)k * n);
}
/* Plan32::inv: per-prime inv (clobbers mod_p) then Garner.  native32.rs:379-407, native64.rs:1001-1038,
 * native128.rs:253-293, native_binary32.rs:219-241, native_binary64.rs:391-419, native_binary128.rs:145-167 */
EXPORT void o_native_inv(const o_native *nt, void *value, u32 *mod_p)
{
    size_t n = nt->n;
    for (int k = 0; k < nt->nprimes; k++) o_plan32_inv(nt->pl[k], mod_p + (size_t)k * n);
    for (size_t i = 0; i < n; i++) {
        u32 m[10];
        for (int k = 0; k < nt->nprimes; k++) m[k] = mod_p[(size_t)k * n + i];
        if (!nt->binary) {
            if (nt->word_bits == 32) ((u32 *)value)[i] = o_reconstruct_32bit_012_u32(m[0], m[1], m[2]);
            else if (nt->word_bits == 64) ((u64 *)value)[i] = o_reconstruct_32bit_01234_u64(m[0], m[1], m[2], m[3], m[4]);
            else o_reconstruct_32bit_0123456789_u128(m, (u64 *)value + 2 * i);
        } else {
            if (nt->word_bits == 32) ((u32 *)value)[i] = o_reconstruct_32bit_01(m[0], m[1]);
            else if (nt->word_bits == 64) ((u64 *)value)[i] = o_reconstruct_32bit_012_u64(m[0], m[1], m[2]);
            else o_reconstruct_32bit_01234_u128(m[0], m[1], m[2], m[3], m[4], (u64 *)value + 2 * i);
        }
    }
}
/* Plan32::negacyclic_polymul: native32.rs:411-432, native64.rs:1042-1069, native128.rs:297-348,
 * native_binary32.rs:245-262, native_binary64.rs:423-444, native_binary128.rs:169-196 */
EXPORT void o_native_polymul(const o_native *nt, void *prod, const void *lhs, const void *rhs)
{
    size_t n = nt->n;
    int np = nt->nprimes;
    u32 *l = (u32 *)malloc((size_t)np * n * sizeof(u32));
    u32 *r = (u32 *)malloc((size_t)np * n * sizeof(u32));
    o_native_fwd(nt, lhs, l);
    if (nt->binary) o_native_fwd_binary(nt, rhs, r);
    else o_native_fwd(nt, rhs, r);
    for (int k = 0; k < np; k++) o_plan32_mul_assign_normalize(nt->pl[k], l + (size_t)k * n, r + (size_t)k * n, n);
    o_native_inv(nt, prod, l);
    free(l); free(r);
}

/* ==========================================================================================
 * Schoolbook negacyclic convolutions -- the reference's own test oracles:
 * src/prime32.rs:957-978, src/prime64.rs:1170-1182, src/native128.rs:359-372.
 * p == 0 selects wrapping arithmetic.
 * ========================================================================================== */
EXPORT void o_schoolbook32(size_t n, u32 p, const u32 *lhs, const u32 *rhs, u32 *out)
{
    for (size_t i = 0; i < n; i++) out[i] = 0;
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {
            size_t k = (i + j) % n;
            int neg = (i + j) >= n;
            if (p == 0) {
                u32 pr = lhs[i] * rhs[j];
                out[k] = neg ? out[k] - pr : out[k] + pr;
            } else {
                u32 pr = (u32)(((u64)lhs[i] * rhs[j]) % p);
                out[k] = neg ? gsub32(p, out[k], pr) : gadd32(p, out[k], pr);
            }
        }
}
EXPORT void o_schoolbook64(size_t n, u64 p, const u64 *lhs, const u64 *rhs, u64 *out)
{
    for (size_t i = 0; i < n; i++) out[i] = 0;
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {
            size_t k = (i + j) % n;
            int neg = (i + j) >= n;
            if (p == 0) {
                u64 pr = lhs[i] * rhs[j];
                out[k] = neg ? out[k] - pr : out[k] + pr;
            } else {
                u64 pr = mul_mod64(p, lhs[i], rhs[j]);
                out[k] = neg ? gsub64(p, out[k], pr) : gadd64(p, out[k], pr);
            }
        }
}
EXPORT void o_schoolbook128(size_t n, const u64 *lhs, const u64 *rhs, u64 *out)
{
    u128 *acc = (u128 *)calloc(n, sizeof(u128));
    for (size_t i = 0; i < n; i++) {
        u128 a = load_word(lhs, 128, i);
        for (size_t j = 0; j < n; j++) {
            u128 b = load_word(rhs, 128, j);
            size_t k = (i + j) % n;
            if ((i + j) >= n) acc[k] -= a * b; else acc[k] += a * b;
        }
    }
    for (size_t i = 0; i < n; i++) store_word(out, 128, i, acc[i]);
    free(acc);
}

/* The same wrapping negacyclic convolution (the reference's test oracle, src/native64.rs:1176-1215) in a
 * form that is fast enough for n = 65536: word_bits in {32, 64, 128}, zero lhs coefficients skipped, rows
 * split at the wrap point, OpenMP over output halves.  Used to check the extended plans (n = 65536 has no
 * reference implementation at all; the wrapping product itself is the specification). */
EXPORT void o_negacyclic_wrapping(size_t n, int word_bits, const void *lhs, const void *rhs, void *out, int nthreads)
{
    (void)nthreads;
    if (word_bits == 32) {
        const u32 *a = (const u32 *)lhs, *b = (const u32 *)rhs; u32 *o = (u32 *)out;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
        for (size_t k = 0; k < n; k++) {
            u32 acc = 0;
            for (size_t i = 0; i <= k; i++) acc += a[i] * b[k - i];
            for (size_t i = k + 1; i < n; i++) acc -= a[i] * b[n + k - i];
            o[k] = acc;
        }
    } else if (word_bits == 64) {
        const u64 *a = (const u64 *)lhs, *b = (const u64 *)rhs; u64 *o = (u64 *)out;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
        for (size_t k = 0; k < n; k++) {
            u64 acc = 0;
            for (size_t i = 0; i <= k; i++) acc += a[i] * b[k - i];
            for (size_t i = k + 1; i < n; i++) acc -= a[i] * b[n + k - i];
            o[k] = acc;
        }
    } else {
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
        for (size_t k = 0; k < n; k++) {
            u128 acc = 0;
            for (size_t i = 0; i <= k; i++) { u128 x = load_word(lhs, 128, i); if (x) acc += x * load_word(rhs, 128, k - i); }
            for (size_t i = k + 1; i < n; i++) { u128 x = load_word(lhs, 128, i); if (x) acc -= x * load_word(rhs, 128, n + k - i); }
            store_word(out, 128, k, acc);
        }
    }
}

/* Direct evaluation of the transform definition, independent of the stage drivers:
 * out[j] = sum_i a[i] * psi^((2*brv(j)+1)*i) mod p   (SURVEY.md section 0). */
EXPORT void o_direct_fwd64(size_t n, u64 p, u64 psi, const u64 *a, u64 *out)
{
    unsigned nbits = (unsigned)__builtin_ctzll(n);
    for (size_t j = 0; j < n; j++) {
        u64 x = exp_mod64(p, psi, 2 * (u64)bit_rev(nbits, j) + 1);
        u64 acc = 0, xp = 1;
        for (size_t i = 0; i < n; i++) {
            acc = gadd64(p, acc, mul_mod64(p, a[i] % p, xp));
            xp = mul_mod64(p, xp, x);
        }
        out[j] = acc;
    }
}

/* ==========================================================================================
 * Batch drivers (OpenMP over independent polynomials) -- used only as the CPU baseline of
 * bench.py.  The reference has no batch API: callers loop over polynomials from many threads
 * (SURVEY.md section 2.2); this is that loop.
 * ========================================================================================== */
EXPORT int o_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
#ifdef _OPENMP
#define PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(nthreads)")
#else
#define PAR_FOR
#endif

/* CPU-baseline legs of bench.py: the SIMD ports of the reference's vector paths where the host has the ISA (cntt_simd.c: Solinas,
 * cntt_simd32.c: 32-bit Shoup-form primes) */
#include "cntt_simd.c"
#include "cntt_simd32.c"
static int g_batch_isa = -1; /* -1 best available, 0 scalar, 2 AVX2, 3 AVX-512 */
EXPORT void o_set_batch_isa(int isa) { g_batch_isa = isa; }
EXPORT void o_plan32_fwd_batch(const o_plan32 *pl, u32 *buf, size_t batch, int nthreads)
{
    (void)nthreads;
    const int isa = g_batch_isa;
    PAR_FOR
    for (long b = 0; b < (long)batch; b++) o_plan32_fwd_simd(pl, buf + (size_t)b * pl->n, isa);
}
EXPORT void o_plan32_inv_batch(const o_plan32 *pl, u32 *buf, size_t batch, int nthreads)
{
    (void)nthreads;
    const int isa = g_batch_isa;
    PAR_FOR
    for (long b = 0; b < (long)batch; b++) o_plan32_inv_simd(pl, buf + (size_t)b * pl->n, isa);
}
EXPORT void o_plan64_fwd_batch(const o_plan64 *pl, u64 *buf, size_t batch, int nthreads)
{
    (void)nthreads;
    const int isa = g_batch_isa;
    PAR_FOR
    for (long b = 0; b < (long)batch; b++) o_plan64_fwd_simd(pl, buf + (size_t)b * pl->n, isa);
}
EXPORT void o_plan64_inv_batch(const o_plan64 *pl, u64 *buf, size_t batch, int nthreads)
{
    (void)nthreads;
    const int isa = g_batch_isa;
    PAR_FOR
    for (long b = 0; b < (long)batch; b++) o_plan64_inv_simd(pl, buf + (size_t)b * pl->n, isa);
}
EXPORT void o_native_polymul_batch(const o_native *nt, void *prod, const void *lhs, const void *rhs, size_t batch,
                                   int nthreads)
{
    (void)nthreads;
    size_t stride = nt->n * (size_t)(nt->word_bits / 8);
    PAR_FOR
    for (long b = 0; b < (long)batch; b++)
        o_native_polymul(nt, (char *)prod + (size_t)b * stride, (const char *)lhs + (size_t)b * stride,
                         (const char *)rhs + (size_t)b * stride);
}
