// product_fused.cu -- see product_fused.hpp
#define CNTT_PRODUCT_HELPERS_ONLY
#include "product_fused.hpp"

namespace cntt {

struct ProductFusedDev {
    const uint2* tw[2];
    const uint2* last[2];
    Mod32 mod[2];
    TwHead<uint2> head[2]; // by value: the leading passes read their twiddles from the constant bank
};

template <class A, int LOGN>
struct ProductFusedCfg {
    // the engine of the prime plan's own CTA kernel, except at N = 1024: there prime32 runs 32 words per thread (one warp per
    // polynomial), which the fused kernels -- 64-bit words on top of the 32 residues -- pay for in registers (B200, batch 65536:
    // fwd 0.317 -> 0.346 ms, inv 0.328 -> 0.399 ms), so they stay on 16 words per thread with last-pass tables of their own
    static constexpr int LOGR = LOGN == 10 ? 4 : CtaCfg<A, LOGN>::LOGR;
    typedef Engine<A, LOGN, LOGR> E;
    static constexpr int T = E::T;
    static constexpr int GP = T >= 128 ? 1 : 128 / T;
    // forward output staged through a swizzled tile for coalesced stores, as in k_ntt_cta_pipe (ntt_kernels.cuh): pays
    // off up to N = 1024
    static constexpr bool STAGE = CNTT_STAGE_OUT != 0 && LOGN <= 10 && E::P >= 2 && E::R == 16;
    static constexpr size_t SMEM_XCHG = (size_t)GP * E::NBUF * E::SMEM_WORDS * sizeof(uint32_t);
    static constexpr size_t SMEM_FWD = SMEM_XCHG + (STAGE ? (size_t)GP * E::N * sizeof(uint32_t) : 0);
    // N = 4096 (256 threads): capped at 64 registers, four CTAs per SM (B200: fwd 0.398 -> 0.359 ms, inv 0.401 -> 0.377 ms
    // per 16384); smaller sizes lose 1-2 % with a cap and keep ptxas' default (0 = unspecified)
    static constexpr int MINBLK = LOGN == 12 ? 4 : 0;
};

// ---- fwd (product.rs:276-353 for count32 == 2, count64 == 0) ---------------------------------------------------------
template <class A, int LOGN>
__global__ void __launch_bounds__(ProductFusedCfg<A, LOGN>::GP * ProductFusedCfg<A, LOGN>::T, ProductFusedCfg<A, LOGN>::MINBLK)
k_product_fwd_fused(const ProductConsts c, const __grid_constant__ ProductFusedDev fp, uint64_t* __restrict__ ntt,
                    const uint64_t* __restrict__ standard, int mode, uint64_t bound, unsigned long long batch)
{
    typedef ProductFusedCfg<A, LOGN> Cfg;
    typedef typename Cfg::E E;
    constexpr int T = E::T, R = E::R, N = E::N, GP = Cfg::GP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* sm_all = reinterpret_cast<uint32_t*>(smem_raw);
    const int grp = (GP == 1) ? 0 : (int)(threadIdx.x / T);
    const int tid = (GP == 1) ? (int)threadIdx.x : (int)(threadIdx.x % T);
    unsigned long long b = (unsigned long long)blockIdx.x * GP + grp;
    const bool active = b < batch;
    if (!active) b = batch - 1;
    uint32_t* sm = sm_all + (size_t)grp * E::NBUF * E::SMEM_WORDS;
    const uint64_t* src = standard + b * N;
    uint32_t* dom32 = reinterpret_cast<uint32_t*>(ntt + b * c.domain_len);

    uint64_t s[R];
#pragma unroll
    for (int k = 0; k < R; k++) s[k] = ld_data(src + tid + k * T);
    // FwdMode::Bounded fast path (product.rs:305-322): values are centred representatives of magnitude <= bound < p0, p1
    const bool bounded = mode == PF_BOUNDED && bound < c.p[0] && bound < c.p[1];
    const uint64_t half = c.modulus / 2;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint64_t pj = c.p[j];
        uint32_t x[1][R];
#pragma unroll
        for (int k = 0; k < R; k++) {
            if (bounded) {
                const uint32_t s32 = (uint32_t)s[k];
                x[0][k] = s[k] < half ? s32 : (uint32_t)pj - ((uint32_t)c.modulus - s32);
            } else {
                // s mod p_j, limb by limb: Shoup products in [0, 2p) brought to [0, p) each, sum in [0, 2p) -- inside the
                // forward transform's input range for both classes; the output is canonicalised after the transform
                const uint32_t p32 = (uint32_t)pj;
                const uint32_t lo = (uint32_t)s[k], hi = (uint32_t)(s[k] >> 32);
                uint32_t a = lo * c.red32[j][0][0] - __umulhi(lo, c.red32[j][0][1]) * p32;
                uint32_t h = hi * c.red32[j][1][0] - __umulhi(hi, c.red32[j][1][1]) * p32;
                a = umin32(a, a - p32);
                h = umin32(h, h - p32);
                x[0][k] = a + h;
            }
        }
        E::template fwd<1>(x, sm, typename E::TwSrc{fp.tw[j], fp.last[j], &fp.head[j]}, 1u, tid, fp.mod[j]);
#pragma unroll
        for (int k = 0; k < R; k++) x[0][k] = A::canon_fwd(x[0][k], fp.mod[j]);
        if constexpr (Cfg::STAGE) {
            uint4* stage = reinterpret_cast<uint4*>(smem_raw + Cfg::SMEM_XCHG) + (size_t)grp * (N / 4);
            const int c0 = E::elem_last(tid, 0) / 4;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const int cc = c0 + v;
                stage[cc ^ ((cc >> 3) & 7)] = make_uint4(x[0][4 * v], x[0][4 * v + 1], x[0][4 * v + 2], x[0][4 * v + 3]);
            }
            __syncthreads(); // also orders prime 1's first scatter after prime 0's last gather
            if (active) {
                uint4* out = reinterpret_cast<uint4*>(dom32 + (size_t)j * N);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int cc = r * T + tid;
                    st_data(out + cc, stage[cc ^ ((cc >> 3) & 7)]);
                }
            }
        } else {
            if (active) store_contig<uint32_t, R>(dom32 + (size_t)j * N + E::elem_last(tid, 0), x[0]);
            if (j == 0 && E::P >= 2) __syncthreads(); // prime 1's first scatter vs prime 0's last gather
        }
    }
}

// ---- inv (product.rs:355-880 for count32 == 2, count64 == 0) ---------------------------------------------------------
template <class A, int LOGN>
__global__ void __launch_bounds__(ProductFusedCfg<A, LOGN>::GP * ProductFusedCfg<A, LOGN>::T, ProductFusedCfg<A, LOGN>::MINBLK)
k_product_inv_fused(const ProductConsts c, const __grid_constant__ ProductFusedDev fp, uint64_t* __restrict__ standard,
                    uint64_t* __restrict__ ntt, int mode, unsigned long long batch)
{
    typedef ProductFusedCfg<A, LOGN> Cfg;
    typedef typename Cfg::E E;
    constexpr int T = E::T, R = E::R, N = E::N, GP = Cfg::GP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* sm_all = reinterpret_cast<uint32_t*>(smem_raw);
    const int grp = (GP == 1) ? 0 : (int)(threadIdx.x / T);
    const int tid = (GP == 1) ? (int)threadIdx.x : (int)(threadIdx.x % T);
    unsigned long long b = (unsigned long long)blockIdx.x * GP + grp;
    const bool active = b < batch;
    if (!active) b = batch - 1;
    uint32_t* sm = sm_all + (size_t)grp * E::NBUF * E::SMEM_WORDS;
    uint64_t* dst = standard + b * N;
    uint32_t* dom32 = reinterpret_cast<uint32_t*>(ntt + b * c.domain_len);

    uint32_t r[2][R];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        uint32_t x[1][R];
        load_contig<uint32_t, R>(dom32 + (size_t)j * N + E::elem_last(tid, 0), x[0]);
        E::template inv<1>(x, sm, typename E::TwSrc{fp.tw[j], fp.last[j], &fp.head[j]}, 1u, tid, fp.mod[j]);
#pragma unroll
        for (int k = 0; k < R; k++) {
            r[j][k] = A::canon_inv(x[0][k], fp.mod[j]);
            // the reference's inv leaves the (un-normalised) inverse transforms in `ntt` (product.rs:388-399)
            if (active) st_data(dom32 + (size_t)j * N + tid + k * T, r[j][k]);
        }
        if (j == 0 && E::P >= 2) __syncthreads();
    }
    if (!active) return;
    const uint32_t p0 = (uint32_t)c.p[0], p1 = (uint32_t)c.p[1], i0 = c.inv10_32[0], i1 = c.inv10_32[1];
#pragma unroll
    for (int k = 0; k < R; k++) {
        // Knuth 4.3.2 mixed radix, as product.rs:826-869: v0 = r0, v1 = (r1 - v0) p0^-1 mod p1, lift = v1 p0 + v0;
        // all 32-bit (p0 < p1 < 2^31): the difference is brought to [0, p1), the Shoup product by p0^-1 to [0, p1)
        const uint32_t v0 = r[0][k];
        uint32_t d = r[1][k] - v0;
        d = r[1][k] >= v0 ? d : d + p1;
        uint32_t v1 = d * i0 - __umulhi(d, i1) * p1;
        v1 = umin32(v1, v1 - p1);
        const uint64_t lift = (uint64_t)v1 * p0 + v0;
        uint64_t* o = dst + tid + k * T;
        if (mode == PI_REPLACE) st_data(o, lift);
        else st_data(o, pdev::add_mod(c.modulus, ld_data(o), lift));
    }
}

// ---- launchers ---------------------------------------------------------------------------------------------------
template <class A, int LOGN, bool FWD>
static cudaError_t launch_one(const ProductConsts& c, const ProductFusedArgs& a, uint64_t* ntt, uint64_t* standard, int mode, uint64_t bound,
                              size_t batch, cudaStream_t st)
{
    typedef ProductFusedCfg<A, LOGN> Cfg;
    ProductFusedDev d;
    for (int j = 0; j < 2; j++) {
        d.tw[j] = FWD ? a.tw_fwd[j] : a.tw_inv[j];
        d.last[j] = FWD ? a.last_fwd[j] : a.last_inv[j];
        const TwHead<uint2>* h = FWD ? a.head_fwd[j] : a.head_inv[j];
        if (!d.tw[j] || !h || (Cfg::E::kLastXp && !d.last[j])) return cudaErrorInvalidValue;
        d.mod[j] = a.mod[j];
        d.head[j] = *h;
    }
    const unsigned long long nblk = (batch + Cfg::GP - 1) / Cfg::GP;
    if (nblk == 0) return cudaSuccess;
    if (nblk > 0x7fffffffull) return cudaErrorInvalidValue;
    if constexpr (FWD) {
        auto kern = k_product_fwd_fused<A, LOGN>;
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_FWD);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)nblk, Cfg::GP * Cfg::T, Cfg::SMEM_FWD, st>>>(c, d, ntt, standard, mode, bound, batch);
    } else {
        auto kern = k_product_inv_fused<A, LOGN>;
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_XCHG);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)nblk, Cfg::GP * Cfg::T, Cfg::SMEM_XCHG, st>>>(c, d, standard, ntt, mode, batch);
    }
    return cudaGetLastError();
}

template <class A, bool FWD>
static cudaError_t launch_class(const ProductConsts& c, const ProductFusedArgs& a, uint64_t* ntt, uint64_t* standard, int mode, uint64_t bound,
                                size_t batch, cudaStream_t st)
{
    switch (a.logn) {
    case 8: return launch_one<A, 8, FWD>(c, a, ntt, standard, mode, bound, batch, st);
    case 9: return launch_one<A, 9, FWD>(c, a, ntt, standard, mode, bound, batch, st);
    case 10: return launch_one<A, 10, FWD>(c, a, ntt, standard, mode, bound, batch, st);
    case 11: return launch_one<A, 11, FWD>(c, a, ntt, standard, mode, bound, batch, st);
    case 12: return launch_one<A, 12, FWD>(c, a, ntt, standard, mode, bound, batch, st);
    default: return cudaErrorNotSupported;
    }
}

bool product_fused_supported(int cls, int logn) { return (cls == 0 || cls == 1) && logn >= 8 && logn <= 12; }

cudaError_t product_fused_fwd(const ProductConsts& c, const ProductFusedArgs& a, uint64_t* ntt, const uint64_t* standard, int mode,
                              uint64_t bound, size_t batch, cudaStream_t st)
{
    if (!product_fused_supported(a.cls, a.logn)) return cudaErrorNotSupported;
    uint64_t* s = const_cast<uint64_t*>(standard);
    return a.cls == 0 ? launch_class<A32L4, true>(c, a, ntt, s, mode, bound, batch, st) : launch_class<A32L2, true>(c, a, ntt, s, mode, bound, batch, st);
}
cudaError_t product_fused_inv(const ProductConsts& c, const ProductFusedArgs& a, uint64_t* standard, uint64_t* ntt, int mode, size_t batch,
                              cudaStream_t st)
{
    if (!product_fused_supported(a.cls, a.logn)) return cudaErrorNotSupported;
    return a.cls == 0 ? launch_class<A32L4, false>(c, a, ntt, standard, mode, 0, batch, st) : launch_class<A32L2, false>(c, a, ntt, standard, mode, 0, batch, st);
}

bool product_fused_own_tables(int logn) { return logn == 10; }
template <class A>
static cudaError_t build_last_class(int logn, const uint2* heap, uint2* out, cudaStream_t st)
{
    switch (logn) {
    case 10: return launch_build_last_e<typename ProductFusedCfg<A, 10>::E>(heap, out, 0, st);
    default: return cudaErrorNotSupported;
    }
}
cudaError_t product_fused_build_last(int cls, int logn, const uint2* heap, uint2* out, cudaStream_t st)
{
    return cls == 0 ? build_last_class<A32L4>(logn, heap, out, st) : build_last_class<A32L2>(logn, heap, out, st);
}

} // namespace cntt
