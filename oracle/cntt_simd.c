/* cntt_simd.c -- AVX-512 and AVX2 ports of the reference's SIMD stage loops for p = 2^64 - 2^32 + 1 (Solinas).
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (included at the end of cntt_oracle.c).  The reference's CPU path for this prime
 * is the vectorised one (prime64/generic_solinas.rs:132-446: `PrimeModulusV3/V4 for Solinas`, the drivers
 * `fwd_breadth_first_avx2/avx512`, `inv_*` and the lane interleaves of the last three levels); the crate cannot be built
 * here (no cargo), so this file restates those loops with intrinsics so that bench.py's CPU arm times the algorithm the
 * crate would run on this host, not the scalar fallback.  Results are the canonical residues of the scalar path, checked bit
 * for bit in tests/test_oracle_simd.py.  Selected at run time (__builtin_cpu_supports); the scalar oracle stays the checker.
 *
 *   widening 64 x 64 -> 128 multiply from four 32 x 32 products     src/lib.rs:171-204 (V4), 283-316 (V3)
 *   Solinas reduction, add, sub with compare + select               generic_solinas.rs:400-446 (V4), 234-293 (V3)
 *   stage drivers, last levels through lane permutes                generic_solinas.rs:564-1336
 */
#if defined(__x86_64__)
#include <immintrin.h>

#define T512 __attribute__((target("avx512f,avx512dq,avx512vl")))
#define T256 __attribute__((target("avx2")))

/* ------------------------------------------------------------------ AVX-512: 8 lanes ---------------------- */
typedef struct { __m512i lo, hi; } wide512;
T512 static inline wide512 widening_mul512(__m512i x, __m512i y)
{
    const __m512i lo_mask = _mm512_set1_epi64(0x00000000FFFFFFFFll);
    __m512i x_hi = _mm512_shuffle_epi32(x, (_MM_PERM_ENUM)0xB1);
    __m512i y_hi = _mm512_shuffle_epi32(y, (_MM_PERM_ENUM)0xB1);
    __m512i z_lo_lo = _mm512_mul_epu32(x, y);
    __m512i z_lo_hi = _mm512_mul_epu32(x, y_hi);
    __m512i z_hi_lo = _mm512_mul_epu32(x_hi, y);
    __m512i z_hi_hi = _mm512_mul_epu32(x_hi, y_hi);
    __m512i z_lo_lo_shift = _mm512_srli_epi64(z_lo_lo, 32);
    __m512i sum_tmp = _mm512_add_epi64(z_lo_hi, z_lo_lo_shift);
    __m512i sum_lo = _mm512_and_si512(sum_tmp, lo_mask);
    __m512i sum_mid = _mm512_srli_epi64(sum_tmp, 32);
    __m512i sum_mid2 = _mm512_add_epi64(z_hi_lo, sum_lo);
    __m512i sum_mid2_hi = _mm512_srli_epi64(sum_mid2, 32);
    __m512i sum_hi = _mm512_add_epi64(z_hi_hi, sum_mid);
    wide512 r;
    r.hi = _mm512_add_epi64(sum_hi, sum_mid2_hi);
    r.lo = _mm512_add_epi64(_mm512_slli_epi64(_mm512_add_epi64(z_lo_hi, z_hi_lo), 32), z_lo_lo);
    return r;
}
T512 static inline __m512i solinas_mul512(__m512i a, __m512i b)
{
    const __m512i p = _mm512_set1_epi64((long long)SOLINAS_P);
    wide512 w = widening_mul512(a, b);
    __m512i mid = _mm512_and_si512(w.hi, _mm512_set1_epi64(0x00000000FFFFFFFFll));
    __m512i hi = _mm512_srli_epi64(w.hi, 32);
    __m512i low2 = _mm512_sub_epi64(w.lo, hi);
    low2 = _mm512_mask_add_epi64(low2, _mm512_cmpgt_epu64_mask(hi, w.lo), low2, p);
    __m512i product = _mm512_sub_epi64(_mm512_slli_epi64(mid, 32), mid);
    __m512i result = _mm512_add_epi64(low2, product);
    /* (result < product) || (result >= p) */
    __mmask8 cond = _mm512_cmpgt_epu64_mask(product, result) | _mm512_cmpge_epu64_mask(result, p);
    return _mm512_mask_sub_epi64(result, cond, result, p);
}
T512 static inline __m512i add512(__m512i a, __m512i b)
{
    const __m512i p = _mm512_set1_epi64((long long)SOLINAS_P);
    __m512i neg_b = _mm512_sub_epi64(p, b);
    __mmask8 ge = _mm512_cmpge_epu64_mask(a, neg_b);
    return _mm512_mask_sub_epi64(_mm512_add_epi64(a, b), ge, a, neg_b);
}
T512 static inline __m512i sub512(__m512i a, __m512i b)
{
    const __m512i p = _mm512_set1_epi64((long long)SOLINAS_P);
    __m512i neg_b = _mm512_sub_epi64(p, b);
    __mmask8 ge = _mm512_cmpge_epu64_mask(a, b);
    return _mm512_mask_sub_epi64(_mm512_add_epi64(a, neg_b), ge, a, b);
}
/* lane layouts of the last three levels: 16 consecutive words (A, B) <-> (Z0 = first halves of the blocks, Z1 = second halves) */
#define IDX8(a, b, c, d, e, f, g, h) _mm512_setr_epi64(a, b, c, d, e, f, g, h)
T512 static inline void split512(int t, __m512i A, __m512i B, __m512i *z0, __m512i *z1)
{
    if (t == 4) { *z0 = _mm512_shuffle_i64x2(A, B, 0x44); *z1 = _mm512_shuffle_i64x2(A, B, 0xEE); }
    else if (t == 2) { *z0 = _mm512_permutex2var_epi64(A, IDX8(0, 1, 4, 5, 8, 9, 12, 13), B); *z1 = _mm512_permutex2var_epi64(A, IDX8(2, 3, 6, 7, 10, 11, 14, 15), B); }
    else { *z0 = _mm512_permutex2var_epi64(A, IDX8(0, 2, 4, 6, 8, 10, 12, 14), B); *z1 = _mm512_permutex2var_epi64(A, IDX8(1, 3, 5, 7, 9, 11, 13, 15), B); }
}
T512 static inline void merge512(int t, __m512i z0, __m512i z1, __m512i *A, __m512i *B)
{
    if (t == 4) { *A = _mm512_shuffle_i64x2(z0, z1, 0x44); *B = _mm512_shuffle_i64x2(z0, z1, 0xEE); }
    else if (t == 2) { *A = _mm512_permutex2var_epi64(z0, IDX8(0, 1, 8, 9, 2, 3, 10, 11), z1); *B = _mm512_permutex2var_epi64(z0, IDX8(4, 5, 12, 13, 6, 7, 14, 15), z1); }
    else { *A = _mm512_permutex2var_epi64(z0, IDX8(0, 8, 1, 9, 2, 10, 3, 11), z1); *B = _mm512_permutex2var_epi64(z0, IDX8(4, 12, 5, 13, 6, 14, 7, 15), z1); }
}
/* the 8 / t twiddles of 16 consecutive words, each repeated t times */
T512 static inline __m512i twid512(int t, const u64 *w)
{
    if (t == 4) return _mm512_permutexvar_epi64(IDX8(0, 0, 0, 0, 1, 1, 1, 1), _mm512_maskz_loadu_epi64(0x03, w));
    if (t == 2) return _mm512_permutexvar_epi64(IDX8(0, 0, 1, 1, 2, 2, 3, 3), _mm512_maskz_loadu_epi64(0x0F, w));
    return _mm512_loadu_si512(w);
}
T512 static void fwd_breadth_first_avx512(u64 *data, size_t n, const u64 *twid, size_t depth, size_t half)
{
    size_t t = n / 2, m = 1, w_idx = (m << depth) + half * m;
    for (; t >= 8; t /= 2, m *= 2, w_idx *= 2) {
        const u64 *w = twid + w_idx;
        for (size_t i = 0; i < m; i++) {
            u64 *z0 = data + 2 * i * t, *z1 = z0 + t;
            const __m512i w1 = _mm512_set1_epi64((long long)w[i]);
            for (size_t j = 0; j < t; j += 8) {
                __m512i a = _mm512_loadu_si512(z0 + j), b = _mm512_loadu_si512(z1 + j);
                __m512i bw = solinas_mul512(b, w1);
                _mm512_storeu_si512(z0 + j, add512(a, bw));
                _mm512_storeu_si512(z1 + j, sub512(a, bw));
            }
        }
    }
    for (; t >= 1; t /= 2, m *= 2, w_idx *= 2) {
        const u64 *w = twid + w_idx;
        for (size_t k = 0; k < n; k += 16) {
            __m512i A = _mm512_loadu_si512(data + k), B = _mm512_loadu_si512(data + k + 8), a, b;
            split512((int)t, A, B, &a, &b);
            __m512i bw = solinas_mul512(b, twid512((int)t, w + k / (2 * t)));
            merge512((int)t, add512(a, bw), sub512(a, bw), &A, &B);
            _mm512_storeu_si512(data + k, A);
            _mm512_storeu_si512(data + k + 8, B);
        }
    }
}
T512 static void inv_breadth_first_avx512(u64 *data, size_t n, const u64 *inv_twid, size_t depth, size_t half)
{
    size_t t = 1, m = n, w_idx = (m << depth) + half * m;
    for (; t < 8 && m > 1; t *= 2) {
        m /= 2; w_idx /= 2;
        const u64 *w = inv_twid + w_idx;
        for (size_t k = 0; k < n; k += 16) {
            __m512i A = _mm512_loadu_si512(data + k), B = _mm512_loadu_si512(data + k + 8), a, b;
            split512((int)t, A, B, &a, &b);
            __m512i d = solinas_mul512(sub512(a, b), twid512((int)t, w + k / (2 * t)));
            merge512((int)t, add512(a, b), d, &A, &B);
            _mm512_storeu_si512(data + k, A);
            _mm512_storeu_si512(data + k + 8, B);
        }
    }
    for (; m > 1; t *= 2) {
        m /= 2; w_idx /= 2;
        const u64 *w = inv_twid + w_idx;
        for (size_t i = 0; i < m; i++) {
            u64 *z0 = data + 2 * i * t, *z1 = z0 + t;
            const __m512i w1 = _mm512_set1_epi64((long long)w[i]);
            for (size_t j = 0; j < t; j += 8) {
                __m512i a = _mm512_loadu_si512(z0 + j), b = _mm512_loadu_si512(z1 + j);
                _mm512_storeu_si512(z0 + j, add512(a, b));
                _mm512_storeu_si512(z1 + j, solinas_mul512(sub512(a, b), w1));
            }
        }
    }
}
T512 static void fwd_depth_first_avx512(u64 *data, size_t n, const u64 *twid, size_t depth, size_t half)
{
    if (n <= RECURSION_THRESHOLD_64) { fwd_breadth_first_avx512(data, n, twid, depth, half); return; }
    const size_t t = n / 2;
    const __m512i w1 = _mm512_set1_epi64((long long)twid[((size_t)1 << depth) + half]);
    for (size_t j = 0; j < t; j += 8) {
        __m512i a = _mm512_loadu_si512(data + j), b = _mm512_loadu_si512(data + j + t);
        __m512i bw = solinas_mul512(b, w1);
        _mm512_storeu_si512(data + j, add512(a, bw));
        _mm512_storeu_si512(data + j + t, sub512(a, bw));
    }
    fwd_depth_first_avx512(data, t, twid, depth + 1, half * 2);
    fwd_depth_first_avx512(data + t, t, twid, depth + 1, half * 2 + 1);
}
T512 static void inv_depth_first_avx512(u64 *data, size_t n, const u64 *inv_twid, size_t depth, size_t half)
{
    if (n <= RECURSION_THRESHOLD_64) { inv_breadth_first_avx512(data, n, inv_twid, depth, half); return; }
    const size_t t = n / 2;
    inv_depth_first_avx512(data, t, inv_twid, depth + 1, half * 2);
    inv_depth_first_avx512(data + t, t, inv_twid, depth + 1, half * 2 + 1);
    const __m512i w1 = _mm512_set1_epi64((long long)inv_twid[((size_t)1 << depth) + half]);
    for (size_t j = 0; j < t; j += 8) {
        __m512i a = _mm512_loadu_si512(data + j), b = _mm512_loadu_si512(data + j + t);
        _mm512_storeu_si512(data + j, add512(a, b));
        _mm512_storeu_si512(data + j + t, solinas_mul512(sub512(a, b), w1));
    }
}

/* ------------------------------------------------------------------ AVX2: 4 lanes -------------------------- */
typedef struct { __m256i lo, hi; } wide256;
T256 static inline __m256i cmpgt_epu64_256(__m256i a, __m256i b) /* unsigned a > b: flip the sign bits (pulp cmp_gt_u64x4) */
{
    const __m256i k = _mm256_set1_epi64x((long long)0x8000000000000000ull);
    return _mm256_cmpgt_epi64(_mm256_xor_si256(a, k), _mm256_xor_si256(b, k));
}
T256 static inline wide256 widening_mul256(__m256i x, __m256i y)
{
    const __m256i lo_mask = _mm256_set1_epi64x(0x00000000FFFFFFFFll);
    __m256i x_hi = _mm256_shuffle_epi32(x, 0xB1), y_hi = _mm256_shuffle_epi32(y, 0xB1);
    __m256i z_lo_lo = _mm256_mul_epu32(x, y), z_lo_hi = _mm256_mul_epu32(x, y_hi);
    __m256i z_hi_lo = _mm256_mul_epu32(x_hi, y), z_hi_hi = _mm256_mul_epu32(x_hi, y_hi);
    __m256i sum_tmp = _mm256_add_epi64(z_lo_hi, _mm256_srli_epi64(z_lo_lo, 32));
    __m256i sum_lo = _mm256_and_si256(sum_tmp, lo_mask), sum_mid = _mm256_srli_epi64(sum_tmp, 32);
    __m256i sum_mid2 = _mm256_add_epi64(z_hi_lo, sum_lo);
    wide256 r;
    r.hi = _mm256_add_epi64(_mm256_add_epi64(z_hi_hi, sum_mid), _mm256_srli_epi64(sum_mid2, 32));
    r.lo = _mm256_add_epi64(_mm256_slli_epi64(_mm256_add_epi64(z_lo_hi, z_hi_lo), 32), z_lo_lo);
    return r;
}
T256 static inline __m256i solinas_mul256(__m256i a, __m256i b)
{
    const __m256i p = _mm256_set1_epi64x((long long)SOLINAS_P);
    wide256 w = widening_mul256(a, b);
    __m256i mid = _mm256_and_si256(w.hi, _mm256_set1_epi64x(0x00000000FFFFFFFFll));
    __m256i hi = _mm256_srli_epi64(w.hi, 32);
    __m256i low2 = _mm256_sub_epi64(w.lo, hi);
    low2 = _mm256_blendv_epi8(low2, _mm256_add_epi64(low2, p), cmpgt_epu64_256(hi, w.lo));
    __m256i product = _mm256_sub_epi64(_mm256_slli_epi64(mid, 32), mid);
    __m256i result = _mm256_add_epi64(low2, product);
    __m256i keep = _mm256_andnot_si256(cmpgt_epu64_256(product, result), cmpgt_epu64_256(p, result)); /* !(r < product) && p > r */
    return _mm256_blendv_epi8(_mm256_sub_epi64(result, p), result, keep);
}
T256 static inline __m256i add256(__m256i a, __m256i b)
{
    const __m256i p = _mm256_set1_epi64x((long long)SOLINAS_P);
    __m256i neg_b = _mm256_sub_epi64(p, b);
    return _mm256_blendv_epi8(_mm256_sub_epi64(a, neg_b), _mm256_add_epi64(a, b), cmpgt_epu64_256(neg_b, a));
}
T256 static inline __m256i sub256(__m256i a, __m256i b)
{
    const __m256i p = _mm256_set1_epi64x((long long)SOLINAS_P);
    __m256i neg_b = _mm256_sub_epi64(p, b);
    return _mm256_blendv_epi8(_mm256_sub_epi64(a, b), _mm256_add_epi64(a, neg_b), cmpgt_epu64_256(b, a));
}
T256 static inline void split256(int t, __m256i A, __m256i B, __m256i *z0, __m256i *z1)
{
    if (t == 2) { *z0 = _mm256_permute2x128_si256(A, B, 0x20); *z1 = _mm256_permute2x128_si256(A, B, 0x31); }
    else { /* t == 1: even / odd words */
        __m256i lo = _mm256_unpacklo_epi64(A, B), hi = _mm256_unpackhi_epi64(A, B); /* (a0 b0 a2 b2), (a1 b1 a3 b3) */
        *z0 = _mm256_permute4x64_epi64(lo, 0xD8); *z1 = _mm256_permute4x64_epi64(hi, 0xD8); /* (a0 a2 b0 b2), (a1 a3 b1 b3) */
    }
}
T256 static inline void merge256(int t, __m256i z0, __m256i z1, __m256i *A, __m256i *B)
{
    if (t == 2) { *A = _mm256_permute2x128_si256(z0, z1, 0x20); *B = _mm256_permute2x128_si256(z0, z1, 0x31); }
    else {
        __m256i x = _mm256_permute4x64_epi64(z0, 0xD8), y = _mm256_permute4x64_epi64(z1, 0xD8); /* (e0 e4 e2 e6), (e1 e5 e3 e7) */
        *A = _mm256_unpacklo_epi64(x, y); *B = _mm256_unpackhi_epi64(x, y);                      /* (e0 e1 e2 e3), (e4 e5 e6 e7) */
    }
}
T256 static inline __m256i twid256(int t, const u64 *w)
{
    if (t == 2) return _mm256_setr_epi64x((long long)w[0], (long long)w[0], (long long)w[1], (long long)w[1]);
    return _mm256_loadu_si256((const __m256i *)w);
}
T256 static void fwd_breadth_first_avx2(u64 *data, size_t n, const u64 *twid, size_t depth, size_t half)
{
    size_t t = n / 2, m = 1, w_idx = (m << depth) + half * m;
    for (; t >= 4; t /= 2, m *= 2, w_idx *= 2) {
        const u64 *w = twid + w_idx;
        for (size_t i = 0; i < m; i++) {
            u64 *z0 = data + 2 * i * t, *z1 = z0 + t;
            const __m256i w1 = _mm256_set1_epi64x((long long)w[i]);
            for (size_t j = 0; j < t; j += 4) {
                __m256i a = _mm256_loadu_si256((const __m256i *)(z0 + j)), b = _mm256_loadu_si256((const __m256i *)(z1 + j));
                __m256i bw = solinas_mul256(b, w1);
                _mm256_storeu_si256((__m256i *)(z0 + j), add256(a, bw));
                _mm256_storeu_si256((__m256i *)(z1 + j), sub256(a, bw));
            }
        }
    }
    for (; t >= 1; t /= 2, m *= 2, w_idx *= 2) {
        const u64 *w = twid + w_idx;
        for (size_t k = 0; k < n; k += 8) {
            __m256i A = _mm256_loadu_si256((const __m256i *)(data + k)), B = _mm256_loadu_si256((const __m256i *)(data + k + 4)), a, b;
            split256((int)t, A, B, &a, &b);
            __m256i bw = solinas_mul256(b, twid256((int)t, w + k / (2 * t)));
            merge256((int)t, add256(a, bw), sub256(a, bw), &A, &B);
            _mm256_storeu_si256((__m256i *)(data + k), A);
            _mm256_storeu_si256((__m256i *)(data + k + 4), B);
        }
    }
}
T256 static void inv_breadth_first_avx2(u64 *data, size_t n, const u64 *inv_twid, size_t depth, size_t half)
{
    size_t t = 1, m = n, w_idx = (m << depth) + half * m;
    for (; t < 4 && m > 1; t *= 2) {
        m /= 2; w_idx /= 2;
        const u64 *w = inv_twid + w_idx;
        for (size_t k = 0; k < n; k += 8) {
            __m256i A = _mm256_loadu_si256((const __m256i *)(data + k)), B = _mm256_loadu_si256((const __m256i *)(data + k + 4)), a, b;
            split256((int)t, A, B, &a, &b);
            __m256i d = solinas_mul256(sub256(a, b), twid256((int)t, w + k / (2 * t)));
            merge256((int)t, add256(a, b), d, &A, &B);
            _mm256_storeu_si256((__m256i *)(data + k), A);
            _mm256_storeu_si256((__m256i *)(data + k + 4), B);
        }
    }
    for (; m > 1; t *= 2) {
        m /= 2; w_idx /= 2;
        const u64 *w = inv_twid + w_idx;
        for (size_t i = 0; i < m; i++) {
            u64 *z0 = data + 2 * i * t, *z1 = z0 + t;
            const __m256i w1 = _mm256_set1_epi64x((long long)w[i]);
            for (size_t j = 0; j < t; j += 4) {
                __m256i a = _mm256_loadu_si256((const __m256i *)(z0 + j)), b = _mm256_loadu_si256((const __m256i *)(z1 + j));
                _mm256_storeu_si256((__m256i *)(z0 + j), add256(a, b));
                _mm256_storeu_si256((__m256i *)(z1 + j), solinas_mul256(sub256(a, b), w1));
            }
        }
    }
}
T256 static void fwd_depth_first_avx2(u64 *data, size_t n, const u64 *twid, size_t depth, size_t half)
{
    if (n <= RECURSION_THRESHOLD_64) { fwd_breadth_first_avx2(data, n, twid, depth, half); return; }
    const size_t t = n / 2;
    const __m256i w1 = _mm256_set1_epi64x((long long)twid[((size_t)1 << depth) + half]);
    for (size_t j = 0; j < t; j += 4) {
        __m256i a = _mm256_loadu_si256((const __m256i *)(data + j)), b = _mm256_loadu_si256((const __m256i *)(data + j + t));
        __m256i bw = solinas_mul256(b, w1);
        _mm256_storeu_si256((__m256i *)(data + j), add256(a, bw));
        _mm256_storeu_si256((__m256i *)(data + j + t), sub256(a, bw));
    }
    fwd_depth_first_avx2(data, t, twid, depth + 1, half * 2);
    fwd_depth_first_avx2(data + t, t, twid, depth + 1, half * 2 + 1);
}
T256 static void inv_depth_first_avx2(u64 *data, size_t n, const u64 *inv_twid, size_t depth, size_t half)
{
    if (n <= RECURSION_THRESHOLD_64) { inv_breadth_first_avx2(data, n, inv_twid, depth, half); return; }
    const size_t t = n / 2;
    inv_depth_first_avx2(data, t, inv_twid, depth + 1, half * 2);
    inv_depth_first_avx2(data + t, t, inv_twid, depth + 1, half * 2 + 1);
    const __m256i w1 = _mm256_set1_epi64x((long long)inv_twid[((size_t)1 << depth) + half]);
    for (size_t j = 0; j < t; j += 4) {
        __m256i a = _mm256_loadu_si256((const __m256i *)(data + j)), b = _mm256_loadu_si256((const __m256i *)(data + j + t));
        _mm256_storeu_si256((__m256i *)(data + j), add256(a, b));
        _mm256_storeu_si256((__m256i *)(data + j + t), solinas_mul256(sub256(a, b), w1));
    }
}
#endif /* __x86_64__ */

/* 0 scalar, 2 AVX2, 3 AVX-512 (what the crate's runtime detection would pick: V4 needs the nightly feature, V3 is default) */
EXPORT int o_simd_isa(void)
{
#if defined(__x86_64__)
    static int isa = -1;
    if (isa < 0) {
        __builtin_cpu_init();
        isa = (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl")) ? 3
              : __builtin_cpu_supports("avx2") ? 2 : 0;
    }
    return isa;
#else
    return 0;
#endif
}
/* Plan::fwd / inv through the widest SIMD path the host has (isa: 0 scalar, 2, 3; -1 = best); only the Solinas class is
 * vectorised here, every other class runs the scalar path.  n >= 16 (prime64 plans) covers the 16-word tail loops. */
EXPORT void o_plan64_fwd_simd(const o_plan64 *pl, u64 *buf, int isa)
{
#if defined(__x86_64__)
    if (isa < 0) isa = o_simd_isa();
    if (pl->p == SOLINAS_P && isa == 3 && o_simd_isa() >= 3) { fwd_depth_first_avx512(buf, pl->n, pl->twid, 0, 0); return; }
    if (pl->p == SOLINAS_P && isa >= 2 && o_simd_isa() >= 2) { fwd_depth_first_avx2(buf, pl->n, pl->twid, 0, 0); return; }
#endif
    (void)isa;
    o_plan64_fwd(pl, buf);
}
EXPORT void o_plan64_inv_simd(const o_plan64 *pl, u64 *buf, int isa)
{
#if defined(__x86_64__)
    if (isa < 0) isa = o_simd_isa();
    if (pl->p == SOLINAS_P && isa == 3 && o_simd_isa() >= 3) { inv_depth_first_avx512(buf, pl->n, pl->inv_twid, 0, 0); return; }
    if (pl->p == SOLINAS_P && isa >= 2 && o_simd_isa() >= 2) { inv_depth_first_avx2(buf, pl->n, pl->inv_twid, 0, 0); return; }
#endif
    (void)isa;
    o_plan64_inv(pl, buf);
}
