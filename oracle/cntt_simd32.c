/* cntt_simd32.c -- AVX-512 and AVX2 ports of the reference's SIMD stage loops for 32-bit Shoup-form primes (p < 2^31).
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (included by cntt_oracle.c after cntt_simd.c).  For p < 2^31 the reference's CPU path
 * is vectorised: 16 (AVX-512, nightly feature) or 8 (AVX2) residues per instruction, Shoup multiplication from the high half of a
 * widening 32 x 32 product, and the last log2(lanes) levels through lane permutes.  This file restates those loops so that a
 * prime32 CPU figure is the algorithm the crate would run on this host.  The results are the scalar path's canonical residues,
 * bit for bit (tests/test_oracle_simd.py).
 *
 *   butterflies p < 2^30                     src/prime32/less_than_30bit.rs:7-112 (fwd), 157-262 (inv)
 *   butterflies 2^30 <= p < 2^31             src/prime32/less_than_31bit.rs:7-114, 159-211
 *   stage drivers and lane interleaves       src/prime32/shoup.rs:305-451 (AVX-512 fwd), 13-160 (AVX2 fwd), 709-1354 (inv)
 *   high half of the widening product        pulp widening_mul_u32x16 / u32x8: even lanes and odd lanes by vpmuludq, blended
 */
#if defined(__x86_64__)

/* ------------------------------------------------------------------ vector primitives ----------------------- */
T512 static inline __m512i mulhi512_u32(__m512i a, __m512i b)
{
    __m512i even = _mm512_mul_epu32(a, b);
    __m512i odd = _mm512_mul_epu32(_mm512_srli_epi64(a, 32), _mm512_srli_epi64(b, 32));
    return _mm512_mask_blend_epi32((__mmask16)0xAAAA, _mm512_srli_epi64(even, 32), odd);
}
T256 static inline __m256i mulhi256_u32(__m256i a, __m256i b)
{
    __m256i even = _mm256_mul_epu32(a, b);
    __m256i odd = _mm256_mul_epu32(_mm256_srli_epi64(a, 32), _mm256_srli_epi64(b, 32));
    return _mm256_blend_epi32(_mm256_srli_epi64(even, 32), odd, 0xAA);
}

/* one set of names per ISA, so that the butterflies and drivers below are written once (macro-expanded twice) */
#define V512 __m512i
#define V512_SET1(x) _mm512_set1_epi32((int)(x))
#define V512_ADD _mm512_add_epi32
#define V512_SUB _mm512_sub_epi32
#define V512_MIN _mm512_min_epu32
#define V512_MULLO _mm512_mullo_epi32
#define V512_MULHI mulhi512_u32
#define V512_LOAD(p) _mm512_loadu_si512((const void *)(p))
#define V512_STORE(p, v) _mm512_storeu_si512((void *)(p), (v))
#define V256 __m256i
#define V256_SET1(x) _mm256_set1_epi32((int)(x))
#define V256_ADD _mm256_add_epi32
#define V256_SUB _mm256_sub_epi32
#define V256_MIN _mm256_min_epu32
#define V256_MULLO _mm256_mullo_epi32
#define V256_MULHI mulhi256_u32
#define V256_LOAD(p) _mm256_loadu_si256((const __m256i *)(p))
#define V256_STORE(p, v) _mm256_storeu_si256((__m256i *)(p), (v))

/* lane interleaves of the last log2(LANES) levels: 2 * LANES consecutive words (A, B) <-> (Z0 = first halves of the blocks of 2t
 * words, Z1 = second halves), and the twiddle of lane i = entry i / t of the level's slice.  Index tables are built once. */
static int g_idx32_ready;
static int g_split0[5][16], g_split1[5][16], g_mergeA[5][16], g_mergeB[5][16], g_twid[5][16]; /* [log2 t], 16 lanes */
static int g_split0_8[4][8], g_split1_8[4][8], g_mergeA_8[4][8], g_mergeB_8[4][8], g_twid_8[4][8];
static void idx32_init(void)
{
    if (g_idx32_ready) return;
    for (int lt = 0; lt < 4; lt++) {
        const int t = 1 << lt;
        for (int i = 0; i < 16; i++) {
            const int b = i / t, o = i % t;
            g_split0[lt][i] = b * 2 * t + o;
            g_split1[lt][i] = b * 2 * t + t + o;
            g_twid[lt][i] = b;
        }
        for (int s = 0; s < 32; s++) {
            const int b = s / (2 * t), r = s % (2 * t);
            const int src = r < t ? b * t + r : 16 + b * t + (r - t);
            if (s < 16) g_mergeA[lt][s] = src; else g_mergeB[lt][s - 16] = src;
        }
    }
    for (int lt = 0; lt < 3; lt++) {
        const int t = 1 << lt;
        for (int i = 0; i < 8; i++) {
            const int b = i / t, o = i % t;
            g_split0_8[lt][i] = b * 2 * t + o;
            g_split1_8[lt][i] = b * 2 * t + t + o;
            g_twid_8[lt][i] = b;
        }
        for (int s = 0; s < 16; s++) {
            const int b = s / (2 * t), r = s % (2 * t);
            const int src = r < t ? b * t + r : 8 + b * t + (r - t);
            if (s < 8) g_mergeA_8[lt][s] = src; else g_mergeB_8[lt][s - 8] = src;
        }
    }
    __atomic_store_n(&g_idx32_ready, 1, __ATOMIC_RELEASE);
}
static inline int log2_small(size_t t) { return t == 1 ? 0 : t == 2 ? 1 : t == 4 ? 2 : 3; }

T512 static inline __m512i perm2_512(__m512i a, const int *idx, __m512i b) { return _mm512_permutex2var_epi32(a, _mm512_loadu_si512(idx), b); }
T512 static inline __m512i twid_512(const u32 *w, int lt) { return _mm512_permutexvar_epi32(_mm512_loadu_si512(g_twid[lt]), _mm512_loadu_si512(w)); }
/* AVX2 has no two-source 32-bit permute: permute each source by the low three index bits, pick by bit 3 */
T256 static inline __m256i perm2_256(__m256i a, const int *idx, __m256i b)
{
    const __m256i i = _mm256_loadu_si256((const __m256i *)idx);
    const __m256i pa = _mm256_permutevar8x32_epi32(a, i), pb = _mm256_permutevar8x32_epi32(b, i);
    return _mm256_castps_si256(_mm256_blendv_ps(_mm256_castsi256_ps(pa), _mm256_castsi256_ps(pb),
                                                 _mm256_castsi256_ps(_mm256_slli_epi32(i, 28)))); /* index bit 3 -> lane sign bit */
}
T256 static inline __m256i twid_256(const u32 *w, int lt)
{
    return _mm256_permutevar8x32_epi32(_mm256_loadu_si256((const __m256i *)w), _mm256_loadu_si256((const __m256i *)g_twid_8[lt]));
}

/* ------------------------------------------------------------------ butterflies + drivers, per ISA and class ----------------- */
#define DEFINE_SHOUP32(TGT, V, LANES, SUF, SPLIT0, SPLIT1, MERGEA, MERGEB, PERM2, TWID)                                          \
    typedef struct { V a, b; } pair##SUF;                                                                                        \
    TGT static inline pair##SUF fwd_bf30##SUF(V z0, V z1, V w, V ws, V p, V neg_p, V two_p, int last)                            \
    {                                                                                                                            \
        z0 = V##_MIN(z0, V##_SUB(z0, two_p));                                                                                    \
        if (last) z0 = V##_MIN(z0, V##_SUB(z0, p));                                                                              \
        V q = V##_MULHI(z1, ws);                                                                                                 \
        V t = V##_ADD(V##_MULLO(z1, w), V##_MULLO(q, neg_p));                                                                    \
        pair##SUF r;                                                                                                             \
        if (!last) { r.a = V##_ADD(z0, t); r.b = V##_ADD(V##_SUB(z0, t), two_p); return r; }                                     \
        t = V##_MIN(t, V##_SUB(t, p));                                                                                           \
        r.a = V##_ADD(z0, t); r.a = V##_MIN(r.a, V##_SUB(r.a, p));                                                               \
        r.b = V##_ADD(V##_SUB(z0, t), p); r.b = V##_MIN(r.b, V##_SUB(r.b, p));                                                   \
        return r;                                                                                                                \
    }                                                                                                                            \
    TGT static inline pair##SUF inv_bf30##SUF(V z0, V z1, V w, V ws, V p, V neg_p, V two_p, int last)                            \
    {                                                                                                                            \
        V y0 = V##_ADD(z0, z1);                                                                                                  \
        y0 = V##_MIN(y0, V##_SUB(y0, two_p));                                                                                    \
        V t = V##_ADD(V##_SUB(z0, z1), two_p);                                                                                   \
        V q = V##_MULHI(t, ws);                                                                                                  \
        V y1 = V##_ADD(V##_MULLO(t, w), V##_MULLO(q, neg_p));                                                                    \
        pair##SUF r;                                                                                                             \
        if (last) { y0 = V##_MIN(y0, V##_SUB(y0, p)); y1 = V##_MIN(y1, V##_SUB(y1, p)); }                                        \
        r.a = y0; r.b = y1;                                                                                                      \
        return r;                                                                                                                \
    }                                                                                                                            \
    TGT static inline pair##SUF fwd_bf31##SUF(V z0, V z1, V w, V ws, V p, V neg_p, V two_p, int last)                            \
    {                                                                                                                            \
        (void)two_p;                                                                                                             \
        z0 = V##_MIN(z0, V##_SUB(z0, p));                                                                                        \
        V q = V##_MULHI(z1, ws);                                                                                                 \
        V t = V##_ADD(V##_MULLO(z1, w), V##_MULLO(q, neg_p));                                                                    \
        t = V##_MIN(t, V##_SUB(t, p));                                                                                           \
        pair##SUF r;                                                                                                             \
        r.a = V##_ADD(z0, t); r.b = V##_ADD(V##_SUB(z0, t), p);                                                                  \
        if (last) { r.a = V##_MIN(r.a, V##_SUB(r.a, p)); r.b = V##_MIN(r.b, V##_SUB(r.b, p)); }                                  \
        return r;                                                                                                                \
    }                                                                                                                            \
    TGT static inline pair##SUF inv_bf31##SUF(V z0, V z1, V w, V ws, V p, V neg_p, V two_p, int last)                            \
    {                                                                                                                            \
        (void)two_p; (void)last; /* less_than_31bit.rs:362-380: the same butterfly in both roles */                              \
        V y0 = V##_ADD(z0, z1);                                                                                                  \
        y0 = V##_MIN(y0, V##_SUB(y0, p));                                                                                        \
        V t = V##_ADD(V##_SUB(z0, z1), p);                                                                                       \
        V q = V##_MULHI(t, ws);                                                                                                  \
        V y1 = V##_ADD(V##_MULLO(t, w), V##_MULLO(q, neg_p));                                                                    \
        pair##SUF r;                                                                                                             \
        r.a = y0; r.b = V##_MIN(y1, V##_SUB(y1, p));                                                                             \
        return r;                                                                                                                \
    }                                                                                                                            \
    /* cls: 30 or 31.  `last` marks the level that canonicalises: t == 1 forward, m == 1 inverse (breadth-first roles) */        \
    TGT static inline pair##SUF fwd_bf##SUF(int cls, V z0, V z1, V w, V ws, V p, V neg_p, V two_p, int last)                     \
    {                                                                                                                            \
        return cls == 30 ? fwd_bf30##SUF(z0, z1, w, ws, p, neg_p, two_p, last) : fwd_bf31##SUF(z0, z1, w, ws, p, neg_p, two_p, last); \
    }                                                                                                                            \
    TGT static inline pair##SUF inv_bf##SUF(int cls, V z0, V z1, V w, V ws, V p, V neg_p, V two_p, int last)                     \
    {                                                                                                                            \
        return cls == 30 ? inv_bf30##SUF(z0, z1, w, ws, p, neg_p, two_p, last) : inv_bf31##SUF(z0, z1, w, ws, p, neg_p, two_p, last); \
    }                                                                                                                            \
    TGT static void fwd_breadth_first32##SUF(int cls, u32 pp, u32 *data, size_t n, const u32 *twid, const u32 *twid_shoup,      \
                                             size_t depth, size_t half)                                                          \
    {                                                                                                                            \
        const V p = V##_SET1(pp), neg_p = V##_SET1((u32)0 - pp), two_p = V##_SET1(2 * pp);                                       \
        size_t t = n / 2, m = 1, w_idx = (m << depth) + half * m;                                                                \
        for (; t >= LANES; t /= 2, m *= 2, w_idx *= 2) {                                                                         \
            const u32 *w = twid + w_idx, *ws = twid_shoup + w_idx;                                                               \
            for (size_t i = 0; i < m; i++) {                                                                                     \
                u32 *z0 = data + 2 * i * t, *z1 = z0 + t;                                                                        \
                const V w1 = V##_SET1(w[i]), ws1 = V##_SET1(ws[i]);                                                              \
                for (size_t j = 0; j < t; j += LANES) {                                                                          \
                    pair##SUF r = fwd_bf##SUF(cls, V##_LOAD(z0 + j), V##_LOAD(z1 + j), w1, ws1, p, neg_p, two_p, 0);             \
                    V##_STORE(z0 + j, r.a); V##_STORE(z1 + j, r.b);                                                              \
                }                                                                                                                \
            }                                                                                                                    \
        }                                                                                                                        \
        for (; t >= 1; t /= 2, m *= 2, w_idx *= 2) {                                                                             \
            const u32 *w = twid + w_idx, *ws = twid_shoup + w_idx;                                                               \
            const int lt = log2_small(t);                                                                                        \
            for (size_t k = 0; k < n; k += 2 * LANES) {                                                                          \
                V A = V##_LOAD(data + k), B = V##_LOAD(data + k + LANES);                                                        \
                pair##SUF r = fwd_bf##SUF(cls, PERM2(A, SPLIT0[lt], B), PERM2(A, SPLIT1[lt], B), TWID(w + k / (2 * t), lt),      \
                                          TWID(ws + k / (2 * t), lt), p, neg_p, two_p, t == 1);                                  \
                V##_STORE(data + k, PERM2(r.a, MERGEA[lt], r.b));                                                                \
                V##_STORE(data + k + LANES, PERM2(r.a, MERGEB[lt], r.b));                                                        \
            }                                                                                                                    \
        }                                                                                                                        \
    }                                                                                                                            \
    TGT static void inv_breadth_first32##SUF(int cls, u32 pp, u32 *data, size_t n, const u32 *twid, const u32 *twid_shoup,      \
                                             size_t depth, size_t half, int top)                                                 \
    {                                                                                                                            \
        const V p = V##_SET1(pp), neg_p = V##_SET1((u32)0 - pp), two_p = V##_SET1(2 * pp);                                       \
        size_t t = 1, m = n, w_idx = (m << depth) + half * m;                                                                    \
        for (; t < LANES && m > 1; t *= 2) {                                                                                     \
            m /= 2; w_idx /= 2;                                                                                                  \
            const u32 *w = twid + w_idx, *ws = twid_shoup + w_idx;                                                               \
            const int lt = log2_small(t);                                                                                        \
            for (size_t k = 0; k < n; k += 2 * LANES) {                                                                          \
                V A = V##_LOAD(data + k), B = V##_LOAD(data + k + LANES);                                                        \
                pair##SUF r = inv_bf##SUF(cls, PERM2(A, SPLIT0[lt], B), PERM2(A, SPLIT1[lt], B), TWID(w + k / (2 * t), lt),      \
                                          TWID(ws + k / (2 * t), lt), p, neg_p, two_p, top && m == 1);                           \
                V##_STORE(data + k, PERM2(r.a, MERGEA[lt], r.b));                                                                \
                V##_STORE(data + k + LANES, PERM2(r.a, MERGEB[lt], r.b));                                                        \
            }                                                                                                                    \
        }                                                                                                                        \
        for (; m > 1; t *= 2) {                                                                                                  \
            m /= 2; w_idx /= 2;                                                                                                  \
            const u32 *w = twid + w_idx, *ws = twid_shoup + w_idx;                                                               \
            for (size_t i = 0; i < m; i++) {                                                                                     \
                u32 *z0 = data + 2 * i * t, *z1 = z0 + t;                                                                        \
                const V w1 = V##_SET1(w[i]), ws1 = V##_SET1(ws[i]);                                                              \
                for (size_t j = 0; j < t; j += LANES) {                                                                          \
                    pair##SUF r = inv_bf##SUF(cls, V##_LOAD(z0 + j), V##_LOAD(z1 + j), w1, ws1, p, neg_p, two_p, top && m == 1); \
                    V##_STORE(z0 + j, r.a); V##_STORE(z1 + j, r.b);                                                              \
                }                                                                                                                \
            }                                                                                                                    \
        }                                                                                                                        \
    }                                                                                                                            \
    TGT static void fwd_depth_first32##SUF(int cls, u32 pp, u32 *data, size_t n, const u32 *twid, const u32 *twid_shoup,        \
                                           size_t depth, size_t half)                                                            \
    {                                                                                                                            \
        if (n <= RECURSION_THRESHOLD_32) { fwd_breadth_first32##SUF(cls, pp, data, n, twid, twid_shoup, depth, half); return; }  \
        const V p = V##_SET1(pp), neg_p = V##_SET1((u32)0 - pp), two_p = V##_SET1(2 * pp);                                       \
        const size_t t = n / 2, w_idx = ((size_t)1 << depth) + half;                                                             \
        const V w1 = V##_SET1(twid[w_idx]), ws1 = V##_SET1(twid_shoup[w_idx]);                                                   \
        for (size_t j = 0; j < t; j += LANES) {                                                                                  \
            pair##SUF r = fwd_bf##SUF(cls, V##_LOAD(data + j), V##_LOAD(data + j + t), w1, ws1, p, neg_p, two_p, 0);             \
            V##_STORE(data + j, r.a); V##_STORE(data + j + t, r.b);                                                              \
        }                                                                                                                        \
        fwd_depth_first32##SUF(cls, pp, data, t, twid, twid_shoup, depth + 1, half * 2);                                         \
        fwd_depth_first32##SUF(cls, pp, data + t, t, twid, twid_shoup, depth + 1, half * 2 + 1);                                 \
    }                                                                                                                            \
    /* top: this call owns the level that canonicalises (the outermost one); inner calls run `butterfly` in both roles */        \
    TGT static void inv_depth_first32##SUF(int cls, u32 pp, u32 *data, size_t n, const u32 *twid, const u32 *twid_shoup,        \
                                           size_t depth, size_t half, int top)                                                   \
    {                                                                                                                            \
        if (n <= RECURSION_THRESHOLD_32) { inv_breadth_first32##SUF(cls, pp, data, n, twid, twid_shoup, depth, half, top); return; } \
        const size_t t = n / 2, w_idx = ((size_t)1 << depth) + half;                                                             \
        inv_depth_first32##SUF(cls, pp, data, t, twid, twid_shoup, depth + 1, half * 2, 0);                                      \
        inv_depth_first32##SUF(cls, pp, data + t, t, twid, twid_shoup, depth + 1, half * 2 + 1, 0);                              \
        const V p = V##_SET1(pp), neg_p = V##_SET1((u32)0 - pp), two_p = V##_SET1(2 * pp);                                       \
        const V w1 = V##_SET1(twid[w_idx]), ws1 = V##_SET1(twid_shoup[w_idx]);                                                   \
        for (size_t j = 0; j < t; j += LANES) {                                                                                  \
            pair##SUF r = inv_bf##SUF(cls, V##_LOAD(data + j), V##_LOAD(data + j + t), w1, ws1, p, neg_p, two_p, top);           \
            V##_STORE(data + j, r.a); V##_STORE(data + j + t, r.b);                                                              \
        }                                                                                                                        \
    }

DEFINE_SHOUP32(T512, V512, 16, _avx512, g_split0, g_split1, g_mergeA, g_mergeB, perm2_512, twid_512)
DEFINE_SHOUP32(T256, V256, 8, _avx2, g_split0_8, g_split1_8, g_mergeA_8, g_mergeB_8, perm2_256, twid_256)
#endif /* __x86_64__ */

/* Plan::fwd / inv (prime32) through the widest SIMD path the host has (isa: 0 scalar, 2, 3; -1 = best).  Vectorised classes:
 * p < 2^30 and p < 2^31; p >= 2^31 (prime32/generic.rs) and n < 64 run the scalar path. */
EXPORT void o_plan32_fwd_simd(const o_plan32 *pl, u32 *buf, int isa)
{
#if defined(__x86_64__)
    if (isa < 0) isa = o_simd_isa();
    if (pl->p < ((u32)1 << 31) && pl->n >= 64 && isa >= 2 && o_simd_isa() >= 2) {
        const int cls = pl->p < ((u32)1 << 30) ? 30 : 31;
        idx32_init();
        if (isa == 3 && o_simd_isa() >= 3) fwd_depth_first32_avx512(cls, pl->p, buf, pl->n, pl->twid, pl->twid_shoup, 0, 0);
        else fwd_depth_first32_avx2(cls, pl->p, buf, pl->n, pl->twid, pl->twid_shoup, 0, 0);
        return;
    }
#endif
    (void)isa;
    o_plan32_fwd(pl, buf);
}
EXPORT void o_plan32_inv_simd(const o_plan32 *pl, u32 *buf, int isa)
{
#if defined(__x86_64__)
    if (isa < 0) isa = o_simd_isa();
    if (pl->p < ((u32)1 << 31) && pl->n >= 64 && isa >= 2 && o_simd_isa() >= 2) {
        const int cls = pl->p < ((u32)1 << 30) ? 30 : 31;
        idx32_init();
        if (isa == 3 && o_simd_isa() >= 3) inv_depth_first32_avx512(cls, pl->p, buf, pl->n, pl->inv_twid, pl->inv_twid_shoup, 0, 0, 1);
        else inv_depth_first32_avx2(cls, pl->p, buf, pl->n, pl->inv_twid, pl->inv_twid_shoup, 0, 0, 1);
        return;
    }
#endif
    (void)isa;
    o_plan32_inv(pl, buf);
}
