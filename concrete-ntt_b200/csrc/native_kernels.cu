// native_kernels.cu -- kernels of the multi-prime plans: word -> residues, Garner reconstruction,
// and the fused negacyclic polymul.
#include "native.hpp"
#include "host_math.hpp"
#include "native_device.cuh"
#include <mutex>

namespace cntt {

// ---- host: constants (same formulas as src/lib.rs:512-594) ---------------------------------------
static NativeConsts g_consts;
static std::once_flag g_consts_once;

static void fill_consts()
{
    using host::Fp;
    typedef unsigned __int128 u128;
    NativeConsts& c = g_consts;
    static const uint32_t P[10] = {0x3F5A0001u, 0x3F5D0001u, 0x3F760001u, 0x3F820001u, 0x3FAC0001u,
                                   0x3FAF0001u, 0x3FB10001u, 0x3FBB0001u, 0x3FDE0001u, 0x3FFC0001u};
    for (int k = 0; k < 10; k++) {
        c.P[k] = P[k];
        c.barrett[k] = (uint64_t)((((u128)1) << 64) / P[k]);
        c.c64[k] = (uint32_t)((((u128)1) << 64) % P[k]);
    }
    auto inv32 = [](uint32_t m, uint64_t x) { return (uint32_t)Fp(m).inv(x % m); };
    auto shoup64 = [](uint64_t m, uint64_t w) { return (uint64_t)((((u128)w) << 64) / m); };
    // inverse modulo a product of two primes a*b via Euler: x^(phi - 1)
    auto inv_pair = [](uint64_t m, uint64_t x, uint32_t a, uint32_t b) { return Fp(m).pow(x, ((uint64_t)a - 1) * ((uint64_t)b - 1) - 1); };
    c.P0_INV_MOD_P1 = inv32(P[1], P[0]);
    c.P01_INV_MOD_P2 = inv32(P[2], (uint64_t)P[0] * P[1]);
    c.P1_INV_MOD_P2 = inv32(P[2], P[1]);
    c.P3_INV_MOD_P4 = inv32(P[4], P[3]);
    c.P2_INV_MOD_P3 = inv32(P[3], P[2]);
    c.P4_INV_MOD_P5 = inv32(P[5], P[4]);
    c.P6_INV_MOD_P7 = inv32(P[7], P[6]);
    c.P8_INV_MOD_P9 = inv32(P[9], P[8]);
    c.P12 = (uint64_t)P[1] * P[2];
    c.P34 = (uint64_t)P[3] * P[4];
    c.P0_INV_MOD_P12 = inv_pair(c.P12, P[0], P[1], P[2]);
    c.P0_INV_MOD_P12_SHOUP = shoup64(c.P12, c.P0_INV_MOD_P12);
    c.P0_MOD_P34_SHOUP = shoup64(c.P34, P[0]);
    c.P012_INV_MOD_P34 = inv_pair(c.P34, Fp(c.P34).mul(P[0], c.P12), P[3], P[4]);
    c.P012_INV_MOD_P34_SHOUP = shoup64(c.P34, c.P012_INV_MOD_P34);
    c.P01 = (uint64_t)P[0] * P[1];
    c.P23 = (uint64_t)P[2] * P[3];
    c.P45 = (uint64_t)P[4] * P[5];
    c.P67 = (uint64_t)P[6] * P[7];
    c.P89 = (uint64_t)P[8] * P[9];
    c.P01_MOD_P45_SHOUP = shoup64(c.P45, c.P01);
    c.P01_MOD_P67_SHOUP = shoup64(c.P67, c.P01);
    c.P01_MOD_P89_SHOUP = shoup64(c.P89, c.P01);
    c.P23_MOD_P67_SHOUP = shoup64(c.P67, c.P23);
    c.P23_MOD_P89_SHOUP = shoup64(c.P89, c.P23);
    c.P45_MOD_P89_SHOUP = shoup64(c.P89, c.P45);
    c.P01_INV_MOD_P23 = inv_pair(c.P23, c.P01, P[2], P[3]);
    c.P01_INV_MOD_P23_SHOUP = shoup64(c.P23, c.P01_INV_MOD_P23);
    {
        Fp f(c.P45);
        c.P0123_INV_MOD_P45 = inv_pair(c.P45, f.mul(c.P01 % c.P45, c.P23 % c.P45), P[4], P[5]);
        c.P0123_INV_MOD_P45_SHOUP = shoup64(c.P45, c.P0123_INV_MOD_P45);
    }
    {
        Fp f(c.P67);
        c.P012345_INV_MOD_P67 = inv_pair(c.P67, f.mul(f.mul(c.P01 % c.P67, c.P23 % c.P67), c.P45 % c.P67), P[6], P[7]);
        c.P012345_INV_MOD_P67_SHOUP = shoup64(c.P67, c.P012345_INV_MOD_P67);
    }
    {
        Fp f(c.P89);
        c.P01234567_INV_MOD_P89 =
            inv_pair(c.P89, f.mul(f.mul(f.mul(c.P01 % c.P89, c.P23 % c.P89), c.P45 % c.P89), c.P67 % c.P89), P[8], P[9]);
        c.P01234567_INV_MOD_P89_SHOUP = shoup64(c.P89, c.P01234567_INV_MOD_P89);
    }
    u128 p0123 = (u128)c.P01 * (u128)c.P23;
    u128 p012345 = p0123 * (u128)c.P45;
    u128 p01234567 = p012345 * (u128)c.P67;
    u128 p0123456789 = p01234567 * (u128)c.P89;
    auto split = [](u128 x, uint64_t out[2]) { out[0] = (uint64_t)x; out[1] = (uint64_t)(x >> 64); };
    split(p0123, c.P0123);
    split(p012345, c.P012345);
    split(p01234567, c.P01234567);
    split(p0123456789, c.P0123456789);
}

const NativeConsts& native_consts()
{
    std::call_once(g_consts_once, fill_consts);
    return g_consts;
}

// ---- kernels -------------------------------------------------------------------------------------
template <int WORD_BYTES, int NP, bool COPY_LOW32>
__global__ void __launch_bounds__(256)
k_native_reduce(const NativeConsts c, const void* __restrict__ value, uint32_t* __restrict__ planes, size_t plane_stride,
                unsigned long long nwords)
{
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
            else if constexpr (WORD_BYTES == 16) r = dev::mod_u128(lo, hi, c, k);
            else r = dev::mod_u64(lo, c.P[k], c.barrett[k]);
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
    const NativeConsts& c = native_consts();
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
    const NativeConsts& c = native_consts();
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
