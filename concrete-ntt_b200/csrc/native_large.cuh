// native_large.cuh -- negacyclic polymul of the native / native_binary plans for 4096 < N <= 32768 (65536: extended plans).
//
// A residue polynomial of N = C * 4096 u32 words (C = 2, 4, 8, 16) no longer fits one CTA, so the product is
// computed by three kernels around the residue planes, each touching every byte once:
//
//   k_large_lead_fwd   per coefficient column j < 4096: the C words value[j + 4096 r] of an operand are read once,
//                      and for EVERY prime reduced (lhs: times 2^32 / N, as in the fused kernel), run through the
//                      log2 C leading Cooley-Tukey levels in registers and written to the prime's plane
//   k_large_mid        per (prime, polynomial, 4096-word row): both operands' rows are loaded, transformed
//                      together (shared twiddle loads) with the sub-tree rooted at heap node C + row -- exactly
//                      the reference's (recursion_depth, recursion_half) call, prime32/shoup.rs:686-706 --,
//                      multiplied pointwise (Montgomery, x N^-1), inverse-transformed and written over the lhs row
//   k_large_lead_inv   per column: for every prime the C row values are read, run through the log2 C trailing
//                      Gentleman-Sande levels, canonicalised; the Garner lift (native_device.cuh) produces the
//                      C product words, written once
//
// HBM bytes per polymul (binary64, N = 32768: 3 primes): 1.25 + 1.125 + 0.625 MiB = 3.0 MiB, against 7.4 MiB of the
// plan-API composition (reduce, 2 x [strided + CTA] per NTT, pointwise, CRT) it replaces.
#pragma once
#include "native_device.cuh"

namespace cntt {

constexpr int kLargeRowLog = 12;                 // rows of 4096 words: the CTA engine's largest transform
constexpr int kLargeMinLogN = 13, kLargeMaxLogN = 16; // 16 (C = 16 rows): extended prime set only

struct LargeParams {
    const uint2* tw_fwd[10];
    const uint2* tw_inv[10];
    const uint2* tw_fwd_last[10]; // Engine<A32L4, 12, 4> last-pass layouts, one slice per row (sub-block)
    const uint2* tw_inv_last[10];
    Mod32 mod[10];
    uint2 lscale[10][4];
};

// ---- leading levels, forward: words -> planes -------------------------------------------------------------
// grid.x covers batch * 4096 columns; grid.y = 2 (0: lhs -> planes_l, 1: rhs -> planes_r)
template <int KIND, int LOGC>
__global__ void __launch_bounds__(256)
k_large_lead_fwd(const NativeConsts c, const LargeParams lp, const void* __restrict__ lhs, const void* __restrict__ rhs,
                 uint32_t* __restrict__ planes_l, uint32_t* __restrict__ planes_r, size_t plane_stride,
                 unsigned long long ncols)
{
    typedef Engine<A32L4, LOGC, LOGC> E; // a single register pass of LOGC levels, nu = 1
    constexpr int C = 1 << LOGC, NP = dev::KindInfo<KIND>::NP, LIMBS = dev::KindInfo<KIND>::LIMBS;
    constexpr bool BINARY = KIND >= NK_BINARY32;
    constexpr int WB = (int)sizeof(typename dev::KindInfo<KIND>::Word);
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncols) return;
    const bool is_rhs = blockIdx.y != 0;
    const void* src = is_rhs ? rhs : lhs;
    uint32_t* dst = is_rhs ? planes_r : planes_l;
    const unsigned long long b = idx >> kLargeRowLog;
    const unsigned j = (unsigned)(idx & ((1u << kLargeRowLog) - 1u));
    const size_t base = ((size_t)b << (kLargeRowLog + LOGC)) + j;

    uint64_t lo[C], hi[WB == 16 ? C : 1];
#pragma unroll
    for (int r = 0; r < C; r++) {
        uint64_t h;
        dev::load_word<KIND>(src, base + ((size_t)r << kLargeRowLog), lo[r], h);
        if constexpr (WB == 16) hi[r] = h;
    }
#pragma unroll 1
    for (int pk = 0; pk < NP; pk++) {
        const Mod32 m = lp.mod[pk];
        uint32_t x[1][C];
#pragma unroll
        for (int r = 0; r < C; r++) {
            if (!is_rhs) x[0][r] = dev::residue<LIMBS, true>(lo[r], WB == 16 ? hi[r] : 0ull, lp.lscale[pk], m.p);
            else x[0][r] = BINARY ? (uint32_t)lo[r] : dev::residue<LIMBS, false>(lo[r], WB == 16 ? hi[r] : 0ull, c.red[pk], m.p);
        }
        const typename E::TwSrc tws = {lp.tw_fwd[pk], nullptr, nullptr};
        E::template fwd_pass<0, 1>(x, tws, 1u, 0, m);
        uint32_t* plane = dst + (size_t)pk * plane_stride + base;
#pragma unroll
        for (int r = 0; r < C; r++) plane[(size_t)r << kLargeRowLog] = x[0][r]; // lazy [0,4p): consumed by k_large_mid
    }
}

// ---- rows: fwd x 2, pointwise, inv --------------------------------------------------------------------------
// one CTA of 256 threads per (row of a polynomial, prime): blockIdx.x = polynomial * C + row, blockIdx.y = prime
#ifndef CNTT_LARGE_MID_MINBLK
#define CNTT_LARGE_MID_MINBLK 0 // minimum resident CTAs the row kernel is compiled for (0: unspecified; 3 and 4 measured within 2 %)
#endif
template <int LOGC>
__global__ void __launch_bounds__(Geo<kLargeRowLog, 4>::T, CNTT_LARGE_MID_MINBLK)
k_large_mid(const NativeConsts c, const LargeParams lp, uint32_t* __restrict__ planes_l, const uint32_t* __restrict__ planes_r,
            size_t plane_stride)
{
    typedef Engine<A32L4, kLargeRowLog, 4> E;
    constexpr int T = E::T, R = E::R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    const int tid = (int)threadIdx.x;
    const int pk = (int)blockIdx.y;
    const unsigned row = blockIdx.x & ((1u << LOGC) - 1u);
    const unsigned nu0 = (1u << LOGC) + row;
    const Mod32 m = lp.mod[pk];
    uint32_t* L = planes_l + (size_t)pk * plane_stride + ((size_t)blockIdx.x << kLargeRowLog);
    const uint32_t* Rr = planes_r + (size_t)pk * plane_stride + ((size_t)blockIdx.x << kLargeRowLog);

    uint32_t x[2][R];
#pragma unroll
    for (int k = 0; k < R; k++) { x[0][k] = L[tid + k * T]; x[1][k] = Rr[tid + k * T]; }
    E::template fwd<2>(x, sm, typename E::TwSrc{lp.tw_fwd[pk], lp.tw_fwd_last[pk] + (size_t)row * E::LAST_WORDS, nullptr}, nu0, tid, m);
    uint32_t y[1][R];
    const uint32_t p = m.p, pinv = c.pinv[pk];
#pragma unroll
    for (int k = 0; k < R; k++) y[0][k] = dev::mont(dev::red2p(x[0][k], p), dev::red2p(x[1][k], p), p, pinv);
    if constexpr (E::NBUF == 1 && E::P >= 2) __syncthreads();
    E::template inv<1>(y, sm, typename E::TwSrc{lp.tw_inv[pk], lp.tw_inv_last[pk] + (size_t)row * E::LAST_WORDS, nullptr}, nu0, tid, m);
#pragma unroll
    for (int k = 0; k < R; k++) L[tid + k * T] = y[0][k]; // lazy [0,2p): consumed by k_large_lead_inv
}

// ---- trailing levels, inverse, and the Garner lift: planes -> words ------------------------------------------
template <int KIND, int LOGC>
__global__ void __launch_bounds__(128)
k_large_lead_inv(const NativeConsts c, const LargeParams lp, void* __restrict__ prod, const uint32_t* __restrict__ planes,
                 size_t plane_stride, unsigned long long ncols)
{
    typedef Engine<A32L4, LOGC, LOGC> E;
    constexpr int C = 1 << LOGC, NP = dev::KindInfo<KIND>::NP;
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncols) return;
    const unsigned long long b = idx >> kLargeRowLog;
    const unsigned j = (unsigned)(idx & ((1u << kLargeRowLog) - 1u));
    const size_t base = ((size_t)b << (kLargeRowLog + LOGC)) + j;

    uint32_t res[NP][C];
#pragma unroll
    for (int pk = 0; pk < NP; pk++) { // unrolled: res[][] must stay in registers
        const Mod32 m = lp.mod[pk];
        uint32_t x[1][C];
        const uint32_t* plane = planes + (size_t)pk * plane_stride + base;
#pragma unroll
        for (int r = 0; r < C; r++) x[0][r] = plane[(size_t)r << kLargeRowLog];
        const typename E::TwSrc tws = {lp.tw_inv[pk], nullptr, nullptr};
        E::template inv_pass<0, 1>(x, tws, 1u, 0, m);
#pragma unroll
        for (int r = 0; r < C; r++) res[pk][r] = A32L4::canon_inv(x[0][r], m);
    }
#pragma unroll
    for (int r = 0; r < C; r++) {
        uint32_t rr[NP];
#pragma unroll
        for (int pk = 0; pk < NP; pk++) rr[pk] = res[pk][r];
        dev::store_word<KIND>(prod, base + ((size_t)r << kLargeRowLog), dev::reconstruct_bounded<KIND>(rr, c));
    }
}

// ---- launcher ------------------------------------------------------------------------------------------------
template <int KIND, int LOGC>
static cudaError_t launch_large_kc(const NativePlanDev& pl, void* prod, const void* lhs, const void* rhs, size_t batch,
                                   uint32_t* planes_l, uint32_t* planes_r, cudaStream_t st)
{
    typedef Engine<A32L4, kLargeRowLog, 4> E;
    constexpr int NP = dev::KindInfo<KIND>::NP;
    LargeParams lp;
    for (int k = 0; k < NP; k++) {
        lp.tw_fwd[k] = pl.sub[k].tw_fwd; lp.tw_inv[k] = pl.sub[k].tw_inv;
        lp.tw_fwd_last[k] = pl.fused_fwd_last[k]; lp.tw_inv_last[k] = pl.fused_inv_last[k]; // built for 4096-word rows by native_large_build_last
        if (!lp.tw_fwd_last[k] || !lp.tw_inv_last[k]) return cudaErrorInvalidValue;
        lp.mod[k] = pl.sub[k].mod;
        for (int j = 0; j < 4; j++) lp.lscale[k][j] = pl.lscale[k][j];
    }
    const NativeConsts& c = native_consts(pl.prime_set);
    const size_t plane_stride = batch << (kLargeRowLog + LOGC);
    const unsigned long long ncols = (unsigned long long)batch << kLargeRowLog;
    if (((ncols + 127) / 128) > 0x7fffffffull || (batch << LOGC) > 0x7fffffffull) return cudaErrorInvalidValue;
    k_large_lead_fwd<KIND, LOGC><<<dim3((unsigned)((ncols + 255) / 256), 2), 256, 0, st>>>(c, lp, lhs, rhs, planes_l, planes_r, plane_stride, ncols);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)2 * E::NBUF * E::SMEM_WORDS * sizeof(uint32_t);
    auto mid = k_large_mid<LOGC>;
    if ((e = ensure_dyn_smem(reinterpret_cast<const void*>(mid), smem)) != cudaSuccess) return e;
    mid<<<dim3((unsigned)(batch << LOGC), NP), E::T, smem, st>>>(c, lp, planes_l, planes_r, plane_stride);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    k_large_lead_inv<KIND, LOGC><<<(unsigned)((ncols + 127) / 128), 128, 0, st>>>(c, lp, prod, planes_l, plane_stride, ncols);
    return cudaGetLastError();
}

template <int KIND>
static cudaError_t launch_large_kind(const NativePlanDev& pl, void* prod, const void* lhs, const void* rhs, size_t batch,
                                     uint32_t* planes_l, uint32_t* planes_r, cudaStream_t st)
{
    switch (pl.logn - kLargeRowLog) {
    case 1: return launch_large_kc<KIND, 1>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
    case 2: return launch_large_kc<KIND, 2>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
    case 3: return launch_large_kc<KIND, 3>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
    case 4: // N = 65536: extended plans; the 10-prime kind has no extended set (nine primes only)
        if constexpr (KIND != NK_NATIVE128) return launch_large_kc<KIND, 4>(pl, prod, lhs, rhs, batch, planes_l, planes_r, st);
        else return cudaErrorNotSupported;
    default: return cudaErrorNotSupported;
    }
}

// heap (2^logn entries) -> last-pass layouts of Engine<A32L4, 12, 4>, one slice per 4096-word row (n entries in all).
// The large path owns these tables: the prime32 sub-plans lay theirs out for whatever block size launch_ntt uses.
cudaError_t native_large_build_last(int logn, const uint2* heap, uint2* out, cudaStream_t st)
{
    if (!native_large_supported(logn)) return cudaErrorInvalidValue;
    return launch_build_last_e<Engine<A32L4, kLargeRowLog, 4>>(heap, out, logn - kLargeRowLog, st);
}

} // namespace cntt
