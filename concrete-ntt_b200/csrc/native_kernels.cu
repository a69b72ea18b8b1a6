// native_kernels.cu -- kernels of the multi-prime plans: word -> residues, Garner reconstruction,
// and the fused negacyclic polymul.
#include "native.hpp"
#include "host_math.hpp"
#include "native_device.cuh"
#include <mutex>

namespace cntt {

// ---- host: constants -----------------------------------------------------------------------------
static NativeConsts g_consts[kNativePrimeSets];
static std::once_flag g_consts_once;

static void fill_consts_set(int set)
{
    using host::Fp;
    typedef unsigned __int128 u128;
    NativeConsts& c = g_consts[set];
    static const uint32_t PSETS[kNativePrimeSets][10] = {
        {0x3F5A0001u, 0x3F5D0001u, 0x3F760001u, 0x3F820001u, 0x3FAC0001u,
         0x3FAF0001u, 0x3FB10001u, 0x3FBB0001u, 0x3FDE0001u, 0x3FFC0001u}, // src/lib.rs:453-462
        {0x3F3A0001u, 0x3F540001u, 0x3F5A0001u, 0x3F760001u, 0x3F820001u,
         0x3FAC0001u, 0x3FD20001u, 0x3FDE0001u, 0x3FFC0001u, 0x3FFC0001u}, // every prime k 2^17 + 1 in (2^30 - 2^24, 2^30); slot 9 unused
    };
    const uint32_t* P = PSETS[set];
    auto shoup = [](uint64_t w, uint32_t p) { return make_uint2((uint32_t)w, (uint32_t)((w << 32) / p)); };
    for (int k = 0; k < 10; k++) {
        const uint32_t p = P[k];
        c.P[k] = p;
        uint32_t inv = 1; // Newton iteration for p^-1 mod 2^32 (p odd)
        for (int it = 0; it < 5; it++) inv *= 2u - p * inv;
        c.pinv[k] = inv;
        Fp f(p);
        uint64_t w = 1;
        for (int j = 0; j < 4; j++) {
            c.red[k][j] = shoup(w, p);
            w = f.mul(w, ((uint64_t)1 << 32) % p);
        }
        for (int j = 0; j < 10; j++) c.ginv[j][k] = j < k ? shoup(f.inv(P[j] % p), p) : make_uint2(0, 0);
        c.half_single[k] = p / 2;
        if (k >= 1) {
            const uint64_t h = ((uint64_t)P[k - 1] * p) / 2;
            c.half_pair_lo[k] = (uint32_t)(h % P[k - 1]);
            c.half_pair_hi[k] = (uint32_t)(h / P[k - 1]);
        } else {
            c.half_pair_lo[k] = c.half_pair_hi[k] = 0;
        }
    }
    static const int kNp[5] = {2, 3, 5, 10, 9};
    for (int cls = 0; cls < 5; cls++) {
        const int np = kNp[cls];
        u128 M = 1; // wrapping mod 2^128 is all the kernels need
        for (int j = 0; j < np; j++) M *= (u128)P[j];
        c.aM[cls][0] = (uint64_t)M; c.aM[cls][1] = (uint64_t)(M >> 64);
        for (int k = 0; k < 10; k++) {
            c.am[cls][k][0] = c.am[cls][k][1] = 0; c.acinv[cls][k] = 0;
            if (k >= np) continue;
            u128 mk = 1;        // M / P[k] mod 2^128
            uint64_t mk_modp = 1; // M / P[k] mod P[k]
            Fp f(P[k]);
            for (int j = 0; j < np; j++) {
                if (j == k) continue;
                mk *= (u128)P[j];
                mk_modp = f.mul(mk_modp, P[j] % P[k]);
            }
            c.am[cls][k][0] = (uint64_t)mk; c.am[cls][k][1] = (uint64_t)(mk >> 64);
            c.acinv[cls][k] = (uint32_t)f.inv(mk_modp);
        }
    }
    for (int k = 0; k < 10; k++) c.ainv[k] = 1.0f / (float)P[k];
    u128 m = 1;
    for (int j = 0; j <= 10; j++) {
        c.gm[j][0] = (uint64_t)m;
        c.gm[j][1] = (uint64_t)(m >> 64);
        if (j < 10) m *= (u128)P[j]; // wrapping, like the reference's u128::wrapping_mul (src/lib.rs:591-594)
    }
}

const NativeConsts& native_consts(int set)
{
    std::call_once(g_consts_once, [] { for (int s = 0; s < kNativePrimeSets; s++) fill_consts_set(s); });
    return g_consts[(set >= 0 && set < kNativePrimeSets) ? set : 0];
}

// lhs scale constants of the fused polymul: 2^(32 j) * 2^32 * N^-1 * acinv[cls][k] mod P[k] (Shoup pairs)
void native_lhs_scale(int logn, uint2 (*out)[4], int set, int np)
{
    const NativeConsts& c = native_consts(set);
    const int cls = native_np_class(np);
    for (int k = 0; k < 10; k++) {
        host::Fp f(c.P[k]);
        const uint64_t ninv = f.inv(((uint64_t)1 << logn) % c.P[k]);
        uint64_t w = f.mul(((uint64_t)1 << 32) % c.P[k], ninv);
        if (k < np) w = f.mul(w, c.acinv[cls][k]); // residues leave the inverse NTT pre-multiplied for reconstruct_bounded
        for (int j = 0; j < 4; j++) {
            out[k][j] = make_uint2((uint32_t)w, (uint32_t)((w << 32) / c.P[k]));
            w = f.mul(w, ((uint64_t)1 << 32) % c.P[k]);
        }
    }
}

// ---- kernels -------------------------------------------------------------------------------------
template <int WORD_BYTES, int NP, bool COPY_LOW32>
__global__ void __launch_bounds__(256)
k_native_reduce(const NativeConsts c, const void* __restrict__ value, uint32_t* __restrict__ planes, size_t plane_stride,
                unsigned long long nwords)
{
    constexpr int LIMBS = WORD_BYTES / 4;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
        uint64_t lo, hi = 0;
        if constexpr (WORD_BYTES == 4) lo = reinterpret_cast<const uint32_t*>(value)[i];
        else if constexpr (WORD_BYTES == 8) lo = reinterpret_cast<const uint64_t*>(value)[i];
        else {
            const uint4 q = reinterpret_cast<const uint4*>(value)[i];
            lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
            hi = (uint64_t)q.z | ((uint64_t)q.w << 32);
        }
#pragma unroll
        for (int k = 0; k < NP; k++) {
            uint32_t r;
            if constexpr (COPY_LOW32) r = (uint32_t)lo;
            else r = dev::residue<LIMBS, false>(lo, hi, c.red[k], c.P[k]); // lazy [0,4p): consumed by the forward NTT only
            planes[(size_t)k * plane_stride + i] = r;
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256)
k_native_crt(const NativeConsts c, void* __restrict__ value, const uint32_t* __restrict__ planes, size_t plane_stride,
             unsigned long long nwords)
{
    constexpr int NP = dev::KindInfo<KIND>::NP;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
        uint32_t r[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) r[k] = planes[(size_t)k * plane_stride + i];
        dev::store_word<KIND>(value, i, dev::reconstruct<KIND>(r, c));
    }
}

static unsigned grid_for(unsigned long long n)
{
    unsigned long long nblk = (n + 255) / 256;
    const unsigned long long cap = 148ull * 16ull;
    return (unsigned)(nblk > cap ? cap : nblk);
}

cudaError_t native_reduce(const NativePlanDev& pl, const void* value, uint32_t* planes, size_t plane_stride, size_t nwords,
                          bool copy_low32, cudaStream_t st)
{
    if (nwords == 0) return cudaSuccess;
    const NativeConsts& c = native_consts(pl.prime_set);
    const unsigned g = grid_for(nwords);
#define CNTT_RED(WB, NP)                                                                                           \
    do {                                                                                                           \
        if (copy_low32) k_native_reduce<WB, NP, true><<<g, 256, 0, st>>>(c, value, planes, plane_stride, nwords);  \
        else k_native_reduce<WB, NP, false><<<g, 256, 0, st>>>(c, value, planes, plane_stride, nwords);            \
    } while (0)
    switch (pl.kind) {
    case NK_NATIVE32: CNTT_RED(4, 3); break;
    case NK_NATIVE64: CNTT_RED(8, 5); break;
    case NK_NATIVE128: CNTT_RED(16, 10); break;
    case NK_BINARY32: CNTT_RED(4, 2); break;
    case NK_BINARY64: CNTT_RED(8, 3); break;
    case NK_BINARY128: CNTT_RED(16, 5); break;
    default: return cudaErrorInvalidValue;
    }
#undef CNTT_RED
    return cudaGetLastError();
}

cudaError_t native_crt(const NativePlanDev& pl, void* value, const uint32_t* planes, size_t plane_stride, size_t nwords,
                       cudaStream_t st)
{
    if (nwords == 0) return cudaSuccess;
    const NativeConsts& c = native_consts(pl.prime_set);
    const unsigned g = grid_for(nwords);
    switch (pl.kind) {
    case NK_NATIVE32: k_native_crt<NK_NATIVE32><<<g, 256, 0, st>>>(c, value, planes, plane_stride, nwords); break;
    case NK_NATIVE64: k_native_crt<NK_NATIVE64><<<g, 256, 0, st>>>(c, value, planes, plane_stride, nwords); break;
    case NK_NATIVE128: k_native_crt<NK_NATIVE128><<<g, 256, 0, st>>>(c, value, planes, plane_stride, nwords); break;
    case NK_BINARY32: k_native_crt<NK_BINARY32><<<g, 256, 0, st>>>(c, value, planes, plane_stride, nwords); break;
    case NK_BINARY64: k_native_crt<NK_BINARY64><<<g, 256, 0, st>>>(c, value, planes, plane_stride, nwords); break;
    case NK_BINARY128: k_native_crt<NK_BINARY128><<<g, 256, 0, st>>>(c, value, planes, plane_stride, nwords); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace cntt

#include "native_fused.cuh"

namespace cntt {
bool native_fused_supported(int logn) { return logn >= kFusedMinLogN && logn <= kFusedMaxLogN; }
template <int KIND>
static cudaError_t fused_build_last_kind(int logn, const uint2* heap, uint2* out, cudaStream_t st)
{
    switch (logn) {
    case 5: return fused_build_last_one<KIND, 5>(heap, out, st);
    case 6: return fused_build_last_one<KIND, 6>(heap, out, st);
    case 7: return fused_build_last_one<KIND, 7>(heap, out, st);
    case 8: return fused_build_last_one<KIND, 8>(heap, out, st);
    case 9: return fused_build_last_one<KIND, 9>(heap, out, st);
    case 10: return fused_build_last_one<KIND, 10>(heap, out, st);
    case 11: return fused_build_last_one<KIND, 11>(heap, out, st);
    case 12: return fused_build_last_one<KIND, 12>(heap, out, st);
    default: return cudaErrorNotSupported;
    }
}
cudaError_t native_fused_build_last(int kind, int logn, const uint2* heap, uint2* out, cudaStream_t st)
{
    // the layout depends on the kind only through native_fused_logr: one representative per class of that function
    const bool wide = kind == NK_NATIVE128 || kind == NK_BINARY128;
    if (wide) return fused_build_last_kind<NK_NATIVE128>(logn, heap, out, st);
    return kind >= NK_BINARY32 ? fused_build_last_kind<NK_BINARY64>(logn, heap, out, st) : fused_build_last_kind<NK_NATIVE64>(logn, heap, out, st);
}
cudaError_t native_polymul_fused_pre(const NativePlanDev& pl, void* prod, const void* lhs, const uint32_t* rhs_planes, size_t batch,
                                     size_t plane_stride, size_t poly_stride, cudaStream_t st)
{
    if (batch == 0) return cudaSuccess;
    switch (pl.kind) {
    case NK_NATIVE32: return launch_fused_pre_kind<NK_NATIVE32>(pl, prod, lhs, rhs_planes, batch, plane_stride, poly_stride, st);
    case NK_NATIVE64: return launch_fused_pre_kind<NK_NATIVE64>(pl, prod, lhs, rhs_planes, batch, plane_stride, poly_stride, st);
    case NK_NATIVE128: return launch_fused_pre_kind<NK_NATIVE128>(pl, prod, lhs, rhs_planes, batch, plane_stride, poly_stride, st);
    case NK_BINARY32: return launch_fused_pre_kind<NK_BINARY32>(pl, prod, lhs, rhs_planes, batch, plane_stride, poly_stride, st);
    case NK_BINARY64: return launch_fused_pre_kind<NK_BINARY64>(pl, prod, lhs, rhs_planes, batch, plane_stride, poly_stride, st);
    case NK_BINARY128: return launch_fused_pre_kind<NK_BINARY128>(pl, prod, lhs, rhs_planes, batch, plane_stride, poly_stride, st);
    default: return cudaErrorInvalidValue;
    }
}
cudaError_t native_polymul_fused(const NativePlanDev& pl, void* prod, const void* lhs, const void* rhs, size_t batch,
                                 cudaStream_t st)
{
    if (batch == 0) return cudaSuccess;
    switch (pl.kind) {
    case NK_NATIVE32: return launch_fused_kind<NK_NATIVE32>(pl, prod, lhs, rhs, batch, st);
    case NK_NATIVE64: return launch_fused_kind<NK_NATIVE64>(pl, prod, lhs, rhs, batch, st);
    case NK_NATIVE128: return launch_fused_kind<NK_NATIVE128>(pl, prod, lhs, rhs, batch, st);
    case NK_BINARY32: return launch_fused_kind<NK_BINARY32>(pl, prod, lhs, rhs, batch, st);
    case NK_BINARY64: return launch_fused_kind<NK_BINARY64>(pl, prod, lhs, rhs, batch, st);
    case NK_BINARY128: return launch_fused_kind<NK_BINARY128>(pl, prod, lhs, rhs, batch, st);
    default: return cudaErrorInvalidValue;
    }
}
} // namespace cntt

#include "native_large.cuh"

namespace cntt {
bool native_large_supported(int logn) { return logn >= kLargeMinLogN && logn <= kLargeMaxLogN; }
cudaError_t native_polymul_large(const NativePlanDev& pl, void* prod, const void* lhs, const void* rhs, size_t batch,
                                 uint32_t* planes_l, uint32_t* planes_r, cudaStream_t st)
{
    if (batch == 0) return cudaSuccess;
    if (!native_large_supported(pl.logn)) return cudaErrorNotSupported;
    switch (pl.kind) {
    case NK_NATIVE32: return launch_large_kind<NK_NATIVE32>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
    case NK_NATIVE64: return launch_large_kind<NK_NATIVE64>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
    case NK_NATIVE128: return launch_large_kind<NK_NATIVE128>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
    case NK_BINARY32: return launch_large_kind<NK_BINARY32>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
    case NK_BINARY64: return launch_large_kind<NK_BINARY64>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
    case NK_BINARY128: return launch_large_kind<NK_BINARY128>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
    default: return cudaErrorInvalidValue;
    }
}
} // namespace cntt
