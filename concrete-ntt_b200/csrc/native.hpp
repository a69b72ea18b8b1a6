// native.hpp -- constants and launch interface of the multi-prime ("native") plans.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ntt_kernels.cuh"

namespace cntt {

// Everything the native kernels need about the built-in primes P0..P9 (src/lib.rs:453-462), derived on
// the host by native_consts() and passed to kernels by value (constant bank, compile-time indices).
//
// Garner reconstruction works on 32-bit mixed-radix digits d_0..d_{np-1} of the unique x in [0, prod P_k)
// with x = r_k (mod P_k):   x = d_0 + d_1 P_0 + d_2 P_0 P_1 + ...
//   d_k = (..((r_k - d_0) P_0^-1 - d_1) P_1^-1 - ...) P_{k-1}^-1  mod P_k
// These are the digits the reference's reconstruct_* functions compute (their v_i, or pairs of them:
// v12 = d_1 + d_2 P_1, v34 = d_3 + d_4 P_3 in src/native64.rs:90-141), so the lifted word and the sign
// rule (top digit / top digit pair > half) are identical; only the evaluation order differs.
struct NativeConsts {
    uint32_t P[10];
    uint32_t pinv[10];      // P[k]^-1 mod 2^32 (Montgomery)
    uint2 red[10][4];       // red[k][j] = {2^(32 j) mod P[k], its Shoup companion}, j = 1..3  (j = 0 unused)
    uint2 ginv[10][10];     // ginv[j][k] = {P[j]^-1 mod P[k], Shoup companion}, j < k
    uint64_t gm[11][2];     // gm[j] = P_0 ... P_{j-1} mod 2^128 as {lo, hi}; gm[0] = 1
    uint32_t half_single[10]; // floor(P[k] / 2): sign threshold when the top digit decides (np = 2, 3)
    uint32_t half_pair_lo[10], half_pair_hi[10]; // floor(P[k-1] P[k] / 2) = lo + hi * P[k-1] (np = 5, 10)
    // Quotient-estimate CRT of the fused polymul kernels (native_device.cuh, reconstruct_bounded), per prime
    // count class cls = 0..4 for np = 2, 3, 5, 10, 9 with M = P_0 ... P_{np-1}:
    uint64_t am[5][10][2];  // (M / P[k]) mod 2^128 as {lo, hi}
    uint64_t aM[5][2];      // M mod 2^128
    uint32_t acinv[5][10];  // ((M / P[k]) mod P[k])^-1 mod P[k]: folded into the lhs scale constants
    float ainv[10];         // 1 / P[k]
};
__host__ __device__ constexpr int native_np_class(int np) { return np == 2 ? 0 : np == 3 ? 1 : np == 5 ? 2 : np == 9 ? 4 : 3; }

// Prime sets.  Set 0 = the reference's P0..P9.  Set 1 = the "extended" set: the nine primes k 2^17 + 1 in
// (2^30 - 2^24, 2^30), ascending -- six of them are the reference's P0, P2, P3, P4, P8, P9 -- which admit
// negacyclic transforms up to N = 65536, one size beyond the reference's limit (P1 - 1 = 2^16 * odd makes
// native*::Plan32::try_new(65536) return None, src/native64.rs:933-942).  A polymul result does not depend on
// which primes carried it (it is the exact integer product, wrapped), so extended plans are bit-compatible
// with the reference wherever both exist; they are an opt-in extension (cntt_native_plan_new_ext), see
// DESIGN.md section 8.  Unused table slots of set 1 repeat its last prime.
constexpr int kNativePrimeSets = 2;
constexpr int kExtPrimes = 9;
const NativeConsts& native_consts(int set = 0);

enum NativeKind {            // word_bits / binary            reference reconstruction
    NK_NATIVE32 = 0,         // 32 / 0   3 primes              native32.rs:27-56
    NK_NATIVE64 = 1,         // 64 / 0   5 primes              native64.rs:90-141
    NK_NATIVE128 = 2,        // 128 / 0  10 primes             native128.rs:19-118
    NK_BINARY32 = 3,         // 32 / 1   2 primes              native_binary32.rs:21-41
    NK_BINARY64 = 4,         // 64 / 1   3 primes              native_binary64.rs:31-61
    NK_BINARY128 = 5,        // 128 / 1  5 primes              native_binary128.rs:12-64
};
inline int native_num_primes(int kind) { static const int np[6] = {3, 5, 10, 2, 3, 5}; return np[kind]; }
// Primes the FUSED polymul (N <= 4096) carries the product on.  A product coefficient is the exact integer
// sum_i +-a_i b_j, |x| < n 2^(2w), and only x mod 2^w is returned, so any prime set with prod P_k > 2 |x| gives the
// reference's result.  native128 at n <= 4096: |x| < 2^268 and P_0 ... P_8 = 2^269.92, so nine primes carry it
// (|x / M| < 0.27, far inside the rounding margin of reconstruct_bounded) and the tenth prime's three transforms,
// two reductions and CRT term are not computed.  The split fwd / inv API and N > 4096 keep the reference's ten.
constexpr int native_fused_np(int kind, int np_ref) { return kind == 2 /* NK_NATIVE128 */ ? 9 : np_ref; }
inline int native_word_bytes(int kind) { static const int wb[6] = {4, 8, 16, 4, 8, 16}; return wb[kind]; }

struct NativePlanDev {
    int kind;
    int logn;
    int nprimes;
    int prime_set;          // 0: reference primes, 1: extended set (N up to 65536)
    PlanDev<A32L4> sub[10]; // prime32 sub-plans on P0.. (all < 2^30)
    uint2 lscale[10][4];    // 2^(32 j) * (2^32 / N) * acinv[cls][k] mod P[k] (Shoup pairs): lhs scaling of the fused polymul
    const uint2* fused_fwd_last[10]; // last-pass twiddle layouts of the fused kernel's engine (Engine::TwSrc::last); N > 4096: of the
                                     // large path's 4096-word row engine (native_large_build_last)
    const uint2* fused_inv_last[10];
    // binary plans, fused polymul: table of the first register pass of the {0,1} operand's forward transform, per prime
    // (native_fused.cuh, binary_pass0); nullptr = not built
    const uint32_t* bin0[10];
};
// entries of one prime's bin0 table: [nibble q][4-bit pattern v][output slot j of a 2^r1-slot set]
constexpr int kBin0Entries = 4 * 16 * 16;

void native_lhs_scale(int logn, uint2 (*out)[4], int set, int np);

// value (batch*n words) -> nprimes residue planes of batch*n u32, plane k at planes + k*plane_stride.
// copy_low32: fwd_binary's `*value as u32` (no reduction).  Residues are written in the lazy range
// [0, 4p) accepted by the forward NTT kernels (the planes are only ever consumed by them).
cudaError_t native_reduce(const NativePlanDev& pl, const void* value, uint32_t* planes, size_t plane_stride, size_t nwords,
                          bool copy_low32, cudaStream_t st);
// residue planes -> words (Garner, centred lift, wrapping)
cudaError_t native_crt(const NativePlanDev& pl, void* value, const uint32_t* planes, size_t plane_stride, size_t nwords,
                       cudaStream_t st);
// fused kernel: is there a variant for this size, and its engine's last-pass table builder (heap -> out, n entries)
bool native_fused_supported(int logn);
cudaError_t native_fused_build_last(int kind, int logn, const uint2* heap, uint2* out, cudaStream_t st);
// Words per thread (log2) of the engine behind the fused polymul and the fused split forward kernel, and therefore of
// the plan's last-pass twiddle layout.  B200, batch 65536: 16 words per thread beat 8 for the 32/64-bit kinds from
// N = 2048 on (native64 13.9 -> 14.9, native32 25.9 -> 27.3, binary64 24.8 -> 26.2 M polymul/s; N = 1024 -1 %), and lose
// for 128-bit words (native128 N = 4096 2.47 -> 2.35).
constexpr int native_fused_logr(int kind, int logn)
{
    const bool wide = kind == NK_NATIVE128 || kind == NK_BINARY128;
    const int r = (!wide && logn >= 11) ? 4 : 3;
    return logn < r ? logn : r;
}
// fused single-kernel polymul; returns cudaErrorNotSupported when no fused variant exists for (kind, logn)
cudaError_t native_polymul_fused(const NativePlanDev& pl, void* prod, const void* lhs, const void* rhs, size_t batch,
                                 cudaStream_t st);

// the same with the rhs operand already in the NTT domain (residue planes as written by cntt_native_fwd): plane k of key b at
// rhs_planes + k * plane_stride + b * poly_stride words; poly_stride = 0 shares one key.  cudaErrorNotSupported outside 256 <= N <= 4096
cudaError_t native_polymul_fused_pre(const NativePlanDev& pl, void* prod, const void* lhs, const uint32_t* rhs_planes, size_t batch,
                                     size_t plane_stride, size_t poly_stride, cudaStream_t st);

// fused split-phase forward kernel (native_split.cu): what = 0 fwd, 1 fwd_binary; cudaErrorNotSupported when no fused
// variant exists for the plan's size (the caller then composes reduce / transforms / CRT)
cudaError_t native_split_fused(const NativePlanDev& pl, void* value, uint32_t* planes, size_t plane_stride, size_t batch, int what,
                               cudaStream_t st);

// three-kernel polymul for 4096 < N <= 65536 (native_large.cuh; 65536 only exists for extended plans).  planes_l / planes_r: nprimes planes of batch * n
// u32 each (scratch).  Returns cudaErrorNotSupported for other sizes.
bool native_large_supported(int logn);
cudaError_t native_large_build_last(int logn, const uint2* heap, uint2* out, cudaStream_t st);
cudaError_t native_polymul_large(const NativePlanDev& pl, void* prod, const void* lhs, const void* rhs, size_t batch,
                                 uint32_t* planes_l, uint32_t* planes_r, cudaStream_t st);

} // namespace cntt
