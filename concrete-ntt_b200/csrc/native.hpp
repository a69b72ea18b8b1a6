// native.hpp -- constants and launch interface of the multi-prime ("native") plans.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ntt_kernels.cuh"

namespace cntt {

// Built-in primes P0..P9 (src/lib.rs:453-462) and everything Garner reconstruction needs
// (src/lib.rs:512-594).  Filled on the host by native_consts() with the reference's formulas and
// passed to kernels by value.
struct NativeConsts {
    uint32_t P[10];
    uint64_t barrett[10];  // floor(2^64 / P[k]) : value % P[k] without a divide
    uint32_t c64[10];      // 2^64 mod P[k]      : folds the high limb of a 128-bit word
    uint32_t P0_INV_MOD_P1, P01_INV_MOD_P2, P1_INV_MOD_P2, P3_INV_MOD_P4;
    uint32_t P2_INV_MOD_P3, P4_INV_MOD_P5, P6_INV_MOD_P7, P8_INV_MOD_P9;
    uint64_t P12, P34, P0_INV_MOD_P12, P0_INV_MOD_P12_SHOUP, P0_MOD_P34_SHOUP, P012_INV_MOD_P34, P012_INV_MOD_P34_SHOUP;
    uint64_t P01, P23, P45, P67, P89;
    uint64_t P01_MOD_P45_SHOUP, P01_MOD_P67_SHOUP, P01_MOD_P89_SHOUP, P23_MOD_P67_SHOUP, P23_MOD_P89_SHOUP, P45_MOD_P89_SHOUP;
    uint64_t P01_INV_MOD_P23, P01_INV_MOD_P23_SHOUP, P0123_INV_MOD_P45, P0123_INV_MOD_P45_SHOUP;
    uint64_t P012345_INV_MOD_P67, P012345_INV_MOD_P67_SHOUP, P01234567_INV_MOD_P89, P01234567_INV_MOD_P89_SHOUP;
    uint64_t P0123[2], P012345[2], P01234567[2], P0123456789[2]; // wrapping u128 products, {lo, hi}
};

const NativeConsts& native_consts();

enum NativeKind {            // word_bits / binary            reconstruction
    NK_NATIVE32 = 0,         // 32 / 0   3 primes              native32.rs:27-56
    NK_NATIVE64 = 1,         // 64 / 0   5 primes              native64.rs:90-141
    NK_NATIVE128 = 2,        // 128 / 0  10 primes             native128.rs:19-118
    NK_BINARY32 = 3,         // 32 / 1   2 primes              native_binary32.rs:21-41
    NK_BINARY64 = 4,         // 64 / 1   3 primes              native_binary64.rs:31-61
    NK_BINARY128 = 5,        // 128 / 1  5 primes              native_binary128.rs:12-64
};
inline int native_num_primes(int kind) { static const int np[6] = {3, 5, 10, 2, 3, 5}; return np[kind]; }
inline int native_word_bytes(int kind) { static const int wb[6] = {4, 8, 16, 4, 8, 16}; return wb[kind]; }

struct NativePlanDev {
    int kind;
    int logn;
    int nprimes;
    PlanDev<A32L4> sub[10]; // prime32 sub-plans on P0.. (all < 2^30)
};

// value (batch*n words) -> nprimes residue planes of batch*n u32, plane k at planes + k*plane_stride.
// copy_low32: fwd_binary's `*value as u32` (no reduction).
cudaError_t native_reduce(const NativePlanDev& pl, const void* value, uint32_t* planes, size_t plane_stride, size_t nwords,
                          bool copy_low32, cudaStream_t st);
// residue planes -> words (Garner, centred lift, wrapping)
cudaError_t native_crt(const NativePlanDev& pl, void* value, const uint32_t* planes, size_t plane_stride, size_t nwords,
                       cudaStream_t st);
// fused single-kernel polymul; returns cudaErrorNotSupported when no fused variant exists for (kind, logn)
cudaError_t native_polymul_fused(const NativePlanDev& pl, void* prod, const void* lhs, const void* rhs, size_t batch,
                                 cudaStream_t st);

} // namespace cntt
