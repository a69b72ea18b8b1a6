// native_split.cu -- fused forward kernel of the split-phase native API (Plan32::fwd / fwd_binary on residue planes,
// src/native64.rs:971-999): the workflow that keeps a key in the NTT domain and runs many mul_accumulate per inv.
//
//   k_native_fwd_fused   words read once; per prime: `% p` (or the low 32 bits, fwd_binary), forward NTT, canonical
//                        plane written once                       word + 4 np bytes per coefficient
//
// against reduce + np transform launches of the composition in capi.cu, which moves every plane three times
// (B200, native64 N=2048 batch 32768: 1.10 -> 0.78 ms).  Plan32::inv stays a composition: the same fusion (np inverse
// transforms + exact Garner lift + the write-back of the clobbered planes the reference leaves in mod_p) measured
// 0.96 -> 1.15 ms -- the np residue polynomials it must hold per CTA cost more occupancy than the saved traffic is worth.
#include "native.hpp"
#include "native_device.cuh"

namespace cntt {

constexpr int kSplitMinLogN = 8, kSplitMaxLogN = 12;
// the engine is the fused polymul's (native_fused_logr): its last-pass twiddle layout exists in every native plan

struct SplitParams {
    const uint2* tw[10];
    const uint2* tw_last[10];
    Mod32 mod[10];
};

template <int KIND, int LOGN>
struct SplitCfg {
    typedef Engine<A32L4, LOGN, native_fused_logr(KIND, LOGN)> E;
    static constexpr int NP = dev::KindInfo<KIND>::NP;
    static constexpr int T = E::T;
    static constexpr size_t SMEM_BYTES = (size_t)E::NBUF * E::SMEM_WORDS * sizeof(uint32_t);
};

template <int KIND, int LOGN, bool COPY_LOW32>
__global__ void __launch_bounds__(SplitCfg<KIND, LOGN>::T)
k_native_fwd_fused(const NativeConsts c, const SplitParams sp, const void* __restrict__ value, uint32_t* __restrict__ planes,
                   size_t plane_stride, unsigned long long batch)
{
    typedef SplitCfg<KIND, LOGN> Cfg;
    typedef typename Cfg::E E;
    typedef typename dev::KindInfo<KIND>::Word Word;
    constexpr int T = Cfg::T, R = E::R, N = E::N, NP = Cfg::NP, LIMBS = dev::KindInfo<KIND>::LIMBS;
    constexpr int WB = (int)sizeof(Word);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    const int tid = (int)threadIdx.x;
    const size_t base = (size_t)blockIdx.x * N;

    uint64_t lo[R], hi[WB == 16 ? R : 1];
#pragma unroll
    for (int k = 0; k < R; k++) {
        uint64_t h;
        dev::load_word<KIND>(value, base + tid + k * T, lo[k], h);
        if constexpr (WB == 16) hi[k] = h;
    }
#pragma unroll 1
    for (int pk = 0; pk < NP; pk++) {
        const Mod32 m = sp.mod[pk];
        uint32_t x[1][R];
#pragma unroll
        for (int k = 0; k < R; k++)
            x[0][k] = COPY_LOW32 ? (uint32_t)lo[k] : dev::residue<LIMBS, false>(lo[k], WB == 16 ? hi[k] : 0ull, c.red[pk], m.p);
        E::template fwd<1>(x, sm, typename E::TwSrc{sp.tw[pk], sp.tw_last[pk], nullptr}, 1u, tid, m);
#pragma unroll
        for (int k = 0; k < R; k++) x[0][k] = A32L4::canon_fwd(x[0][k], m);
        store_contig<uint32_t, R>(planes + (size_t)pk * plane_stride + base + E::elem_last(tid, 0), x[0]);
        if constexpr (E::P >= 2) __syncthreads(); // next prime's first scatter vs this prime's last gather
    }
}

template <int KIND, int LOGN>
static cudaError_t launch_split(const NativePlanDev& pl, void* value, uint32_t* planes, size_t plane_stride, size_t batch, int what,
                                cudaStream_t st)
{
    constexpr int NP = dev::KindInfo<KIND>::NP;
    if (what != 0 && what != 1) return cudaErrorNotSupported;
    SplitParams sp;
    for (int k = 0; k < NP; k++) {
        sp.tw[k] = pl.sub[k].tw_fwd;
        sp.tw_last[k] = pl.fused_fwd_last[k];
        if (!sp.tw[k] || !sp.tw_last[k]) return cudaErrorNotSupported;
        sp.mod[k] = pl.sub[k].mod;
    }
    if (batch > 0x7fffffffull) return cudaErrorInvalidValue;
    const NativeConsts& c = native_consts(pl.prime_set);
    auto go = [&](auto kern, size_t smem, int threads) -> cudaError_t {
        if (smem > (size_t)227 * 1024) return cudaErrorNotSupported;
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)batch, threads, smem, st>>>(c, sp, value, planes, plane_stride, (unsigned long long)batch);
        return cudaGetLastError();
    };
    if (what == 1) {
        if constexpr (KIND >= NK_BINARY32) return go(k_native_fwd_fused<KIND, LOGN, true>, SplitCfg<KIND, LOGN>::SMEM_BYTES, SplitCfg<KIND, LOGN>::T);
        else return cudaErrorInvalidValue;
    }
    return go(k_native_fwd_fused<KIND, LOGN, false>, SplitCfg<KIND, LOGN>::SMEM_BYTES, SplitCfg<KIND, LOGN>::T);
}

template <int KIND>
static cudaError_t launch_split_kind(const NativePlanDev& pl, void* value, uint32_t* planes, size_t plane_stride, size_t batch, int what,
                                     cudaStream_t st)
{
    switch (pl.logn) {
    case 8: return launch_split<KIND, 8>(pl, value, planes, plane_stride, batch, what, st);
    case 9: return launch_split<KIND, 9>(pl, value, planes, plane_stride, batch, what, st);
    case 10: return launch_split<KIND, 10>(pl, value, planes, plane_stride, batch, what, st);
    case 11: return launch_split<KIND, 11>(pl, value, planes, plane_stride, batch, what, st);
    case 12: return launch_split<KIND, 12>(pl, value, planes, plane_stride, batch, what, st);
    default: return cudaErrorNotSupported;
    }
}

// what: 0 = fwd, 1 = fwd_binary.  cudaErrorNotSupported: no fused variant (the caller composes the plan kernels).
cudaError_t native_split_fused(const NativePlanDev& pl, void* value, uint32_t* planes, size_t plane_stride, size_t batch, int what,
                               cudaStream_t st)
{
    if (pl.logn < kSplitMinLogN || pl.logn > kSplitMaxLogN) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    switch (pl.kind) {
    case NK_NATIVE32: return launch_split_kind<NK_NATIVE32>(pl, value, planes, plane_stride, batch, what, st);
    case NK_NATIVE64: return launch_split_kind<NK_NATIVE64>(pl, value, planes, plane_stride, batch, what, st);
    case NK_NATIVE128: return launch_split_kind<NK_NATIVE128>(pl, value, planes, plane_stride, batch, what, st);
    case NK_BINARY32: return launch_split_kind<NK_BINARY32>(pl, value, planes, plane_stride, batch, what, st);
    case NK_BINARY64: return launch_split_kind<NK_BINARY64>(pl, value, planes, plane_stride, batch, what, st);
    case NK_BINARY128: return launch_split_kind<NK_BINARY128>(pl, value, planes, plane_stride, batch, what, st);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace cntt
