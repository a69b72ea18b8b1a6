// ntt_kernels.cuh -- batched single-prime kernels built on the engine, and their launchers.
//
//   k_ntt_cta      one polynomial (or contiguous sub-block of a large polynomial) per thread group, all of its stages in
//                  registers + shared memory: N' <= 4096 with 16 words per thread; 32-bit words also N' = 8192 / 16384 with
//                  32 words per thread and one exchange buffer; 64-bit words N' = 8192 (forward)
//   k_ntt_cta_pipe persistent, software-pipelined flavour of the same (kept for A/B runs; the one-shot kernel is what ships)
//   k_ntt_strided  the leading 1..5 stages of a larger transform, strided through global memory (shift butterflies for Solinas)
//   k_ntt_cluster  N = 32768 on two-CTA clusters with a DSMEM exchange (experiment, off: measured slower than strided + CTA)
//   k_pointwise    mul_assign_normalize / normalize / mul_accumulate streams
//
// Batch layout everywhere: polynomial-major contiguous, buf[b*N + i] (the reference's slices
// concatenated).
#pragma once
#include "ntt_engine.cuh"
#include <atomic>

namespace cntt {

constexpr int kMaxCtaLogN = 12; // largest block a CTA keeps on chip when a transform is cut into blocks
// 32-bit words: N = 8192 still fits one CTA (512 threads x 16 words, 66 KB of exchange buffers), which saves the
// strided pass through HBM that the block scheme would spend on its single leading level.
#ifndef CNTT_CTA13
#define CNTT_CTA13 1
#endif
#ifndef CNTT_CTA13_MINTHREADS_FWD
#define CNTT_CTA13_MINTHREADS_FWD 0 // 0: the 32-bit defaults below
#endif
#ifndef CNTT_CTA13_MINTHREADS_INV
#define CNTT_CTA13_MINTHREADS_INV 0
#endif
#ifndef CNTT_CTA14
#define CNTT_CTA14 2 // N = 16384 (32-bit words) in one CTA of 1024 threads x 16 words (135 KB of exchange buffers, one CTA per
                     // SM): 0 off, 1 both directions, 2 forward only.  B200, batch 16384: forward 0.971 -> 0.840 ms; the inverse
                     // loses 6 % (its phases no longer overlap with a second resident CTA), so it keeps the block scheme
#endif
#ifndef CNTT_CTA13_64
#define CNTT_CTA13_64 2 // the same for 64-bit words (512 threads x 16 words, one 67 KB exchange buffer): 0 off, 1 both directions,
                        // 2 forward only.  B200, Solinas, batch 16384, kernel capped to 64 registers (two CTAs per SM): forward
                        // 1.713 -> 1.549 ms (Shoup-64 1.543 -> 1.495), inverse 1.63 -> 1.83 ms, so only the forward uses it
#endif
#ifndef CNTT_CTA13_64_MINTHREADS
#define CNTT_CTA13_64_MINTHREADS 1024 // resident threads the 64-bit N = 8192 kernel is compiled for (0: the 64-bit default below)
#endif
// 32-bit words, 32 words per thread (LOGR = 5) for N = 2^CNTT_R32_MINLOGN .. 2^CNTT_R32_MAXBLK (0 = off, the r01 scheme): three
// register passes (two exchanges) instead of four, 256 / 512 threads per polynomial, ONE exchange buffer (two barriers per
// exchange) so that four / two CTAs stay resident, and no persistent software-pipelined forward variant (its second register
// set halves the resident CTAs).  B200, r02 (profiles/r02_experiments.txt), M NTT/s fwd / inv:
//   N = 8192   45.8 / 46.1 -> 47.7 / 54.8      N = 16384  19.5 / 17.4 -> 20.5 / 21.6
// With two exchange buffers the same kernels lose (32.2 / 43.8 at N = 8192).  N = 32768 also fits one CTA this way (1024 threads,
// one 135 KB buffer: a single launch that reads and writes every word once) but runs at 8.1 / 8.5 against 8.1 / 8.9 for the
// two-pass scheme -- one resident CTA per SM cannot overlap its load, compute and store phases (ncu: issue 40 %, lg_throttle the
// top stall) -- and as the 32768-word block of N = 65536 it is slower (2.7 against 4.0), so blocks stop at 16384 words.
#ifndef CNTT_R32_MINLOGN
#define CNTT_R32_MINLOGN 13
#endif
#ifndef CNTT_R32_MAXBLK
#define CNTT_R32_MAXBLK 14
#endif
#ifndef CNTT_CLUSTER32
#define CNTT_CLUSTER32 0 // 1: N = 32768 (32-bit words) as one launch of two-CTA clusters (k_ntt_cluster, experiment)
#endif
constexpr bool kR32 = CNTT_R32_MINLOGN != 0;
template <class A, int LOGN> constexpr bool r32_size() { return kR32 && sizeof(typename A::W) == 4 && LOGN >= CNTT_R32_MINLOGN && LOGN <= CNTT_R32_MAXBLK; }
// 32-bit words beyond the single-CTA sizes: levels one strided launch may run (words per thread = 2^k) and the block size the CTA
// kernel then transforms.  r01: k <= 4 and 4096-word blocks in both directions.  r02 (profiles/r02_experiments.txt, "strided depth"):
// the FORWARD transform gains from five leading levels in one strided launch and 1024 / 2048-word blocks (N = 32768 8.0 -> 9.6,
// N = 65536 4.0 -> 4.6 M NTT/s); six levels lose, and the inverse loses or stays flat with any of it, so it keeps the r01 scheme.
#ifndef CNTT_STRIDED_MAXK32_FWD
#define CNTT_STRIDED_MAXK32_FWD 5
#endif
#ifndef CNTT_STRIDED_MAXK32_INV
#define CNTT_STRIDED_MAXK32_INV 4
#endif
#define CNTT_STRIDED_MAXK32 (CNTT_STRIDED_MAXK32_FWD > CNTT_STRIDED_MAXK32_INV ? CNTT_STRIDED_MAXK32_FWD : CNTT_STRIDED_MAXK32_INV)
#ifndef CNTT_LARGE_MINBLK32_FWD
#define CNTT_LARGE_MINBLK32_FWD 10
#endif
// Solinas: all five power-of-two levels (heap nodes < 32) in ONE strided launch of shift butterflies (32 words per thread), then
// small blocks (which run faster per word than 4096-word ones); 4 = the r01 scheme.  Measured (profiles/r02_experiments.txt,
// "k64s"): within +-4 % of the r01 scheme up to N = 65536 (the 32-word strided kernel needs 174-188 registers), so 4 stays.
#ifndef CNTT_STRIDED_MAXK64S
#define CNTT_STRIDED_MAXK64S 4
#endif
#ifndef CNTT_LARGE_MINBLK64S
#define CNTT_LARGE_MINBLK64S 9
#endif
#ifndef CNTT_LARGE_MINLOGN64S
#define CNTT_LARGE_MINLOGN64S 14
#endif
constexpr int large_blk32(int logn, bool fwd);
// size of the contiguous blocks the CTA kernel transforms for a plan of 2^logn words
template <class A> constexpr int cta_block_logn(int logn, bool fwd)
{
    if (logn <= kMaxCtaLogN) return logn;
    if (sizeof(typename A::W) == 4 && kR32 && CNTT_CLUSTER32 && logn == 15) return 14; // cluster of two 16384-word halves (k_ntt_cluster)
    if (sizeof(typename A::W) == 4 && kR32) return (logn >= CNTT_R32_MINLOGN && logn <= CNTT_R32_MAXBLK) ? logn : large_blk32(logn, fwd);
    if (ShiftHead<A>::value && CNTT_STRIDED_MAXK64S > 4 && logn >= CNTT_LARGE_MINLOGN64S && logn - CNTT_STRIDED_MAXK64S <= kMaxCtaLogN)
        return logn - CNTT_STRIDED_MAXK64S < CNTT_LARGE_MINBLK64S ? CNTT_LARGE_MINBLK64S : logn - CNTT_STRIDED_MAXK64S;
    if (logn == 13 && (sizeof(typename A::W) == 4 ? CNTT_CTA13 != 0 : (CNTT_CTA13_64 == 1 || (CNTT_CTA13_64 == 2 && fwd)))) return 13;
#if CNTT_CTA14
    if (logn == 14 && sizeof(typename A::W) == 4 && (CNTT_CTA14 == 1 || fwd)) return 14;
#endif
    return kMaxCtaLogN;
}
constexpr int large_blk32(int logn, bool fwd)
{
    // forward: the smallest block >= CNTT_LARGE_MINBLK32_FWD that ONE strided launch reaches; inverse (and everything one launch
    // cannot reach): 4096-word blocks
    if (fwd && CNTT_STRIDED_MAXK32_FWD > 4) {
        const int blk = logn - CNTT_STRIDED_MAXK32_FWD;
        if (blk <= kMaxCtaLogN) return blk < CNTT_LARGE_MINBLK32_FWD ? CNTT_LARGE_MINBLK32_FWD : blk;
    }
    return kMaxCtaLogN;
}
constexpr int kMaxLogN = 26;    // two-level + repeated strided passes; table memory is the limit

template <class A>
struct PlanDev {
    int logn;
    typename A::Mod mod;
    const typename A::Tw* tw_fwd; // heap order, N entries, entry 0 unused (= 1)
    const typename A::Tw* tw_inv;
    // last-pass tables of the CTA kernel (Engine::TwSrc::last), N entries each, built by launch_build_last;
    // nullptr when the class / size does not use them
    const typename A::Tw* tw_fwd_last;
    const typename A::Tw* tw_inv_last;
    // host copies of heap[0 .. kHeadEntries) (entries beyond N are zero), passed by value to the CTA kernel
    const TwHead<typename A::Tw>* head_fwd;
    const TwHead<typename A::Tw>* head_inv;
};

// LOGR of the CTA kernel per word size, and whether the last-pass table exists for a transform size
#ifndef CNTT_LOGR64
#define CNTT_LOGR64 4
#endif
#ifndef CNTT_LOGR32
#define CNTT_LOGR32 4
#endif
// N = 1024 x 32-bit with 32 words per thread: ONE warp per polynomial and ONE exchange (5 + 5 levels), warp-level barrier only
#ifndef CNTT_R32_LOGN10
#define CNTT_R32_LOGN10 1
#endif
#ifndef CNTT_R32_WHOLE_MASK
#define CNTT_R32_WHOLE_MASK 0 // A/B runs: bit LOGN = whole transforms of 2^LOGN x 32-bit words at 32 words per thread (N = 512: rejected, r02 "warpsync")
#endif
// WHOLE: the kernel transforms whole polynomials (log_sub == 0).  It only matters at N = 1024 x 32-bit: as the 1024-word block of a
// larger transform (N = 32768 forward) the 32-word flavour loses (0.844 -> 0.887 ms per 8192 polynomials), so blocks keep 16 words
template <class A, int LOGN, bool WHOLE = true> struct CtaCfg {
    static constexpr int LOGR_MAX = sizeof(typename A::W) == 8 ? CNTT_LOGR64 : (r32_size<A, LOGN>() || (WHOLE && CNTT_R32_LOGN10 != 0 && LOGN == 10) || (WHOLE && ((CNTT_R32_WHOLE_MASK >> LOGN) & 1) != 0)) ? 5 : CNTT_LOGR32;
    static constexpr int LOGR = LOGN < LOGR_MAX ? LOGN : LOGR_MAX;
    typedef Engine<A, LOGN, LOGR> E;
};

// ---- batch data accesses ----------------------------------------------------------------------
// Batch words are touched exactly once per launch; with CNTT_STREAM_LDST they bypass L1 allocation (ld.global.cs /
// st.global.cs) so that the few KB of L1 left beside the shared-memory carve-out keep the twiddle tables, which
// every polynomial re-reads.
#ifndef CNTT_STREAM_LDST
#define CNTT_STREAM_LDST 1
#endif
template <class T> __device__ __forceinline__ T ld_data(const T* p)
{
#if CNTT_STREAM_LDST
    return __ldcs(p);
#else
    return *p;
#endif
}
template <class T> __device__ __forceinline__ void st_data(T* p, T v)
{
#if CNTT_STREAM_LDST
    __stcs(p, v);
#else
    *p = v;
#endif
}

// ---- vector helpers -------------------------------------------------------------------------
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  A thread's R consecutive words leave in 32-byte pieces, so
// every store instruction fills whole 32-byte sectors; with 128-bit stores each sector of a thread's chunk was written
// by two instructions (ncu: 2x excessive L2 sectors on the forward kernels' stores).  Needs 32-byte alignment, which
// the helpers below test at run time (the C ABI only promises the 16 bytes of cudaMalloc'd sub-buffers).
#ifndef CNTT_WIDE256
#define CNTT_WIDE256 1
#endif
__device__ __forceinline__ void st256(void* p, const uint32_t (&w)[8])
{
#if CNTT_STREAM_LDST
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
#else
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
#endif
                 :: "l"(__cvta_generic_to_global(p)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void ld256(const void* p, uint32_t (&w)[8])
{
#if CNTT_STREAM_LDST
    asm volatile("ld.global.cs.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(__cvta_generic_to_global(p)) : "memory");
}
template <class W, int R>
__device__ __forceinline__ void store_contig(W* __restrict__ dst, const W (&x)[R])
{
    constexpr int BYTES = R * (int)sizeof(W);
    if constexpr (CNTT_WIDE256 && BYTES % 32 == 0) {
        if ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
            constexpr int PER8 = 32 / (int)sizeof(W);
#pragma unroll
            for (int v = 0; v < R / PER8; v++) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if constexpr (sizeof(W) == 4) w[i] = (uint32_t)x[8 * v + i];
                    else w[i] = (uint32_t)((uint64_t)x[4 * v + i / 2] >> (32 * (i & 1)));
                }
                st256(reinterpret_cast<unsigned char*>(dst) + 32 * v, w);
            }
            return;
        }
    }
    if constexpr (BYTES % 16 == 0) {
        constexpr int PER = 16 / (int)sizeof(W);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int v = 0; v < R / PER; v++) {
            uint4 q;
            if constexpr (sizeof(W) == 4) {
                q = make_uint4((uint32_t)x[4 * v], (uint32_t)x[4 * v + 1], (uint32_t)x[4 * v + 2], (uint32_t)x[4 * v + 3]);
            } else {
                uint64_t a = (uint64_t)x[2 * v], b = (uint64_t)x[2 * v + 1];
                q = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
            }
            st_data(d4 + v, q);
        }
    } else {
#pragma unroll
        for (int k = 0; k < R; k++) st_data(dst + k, x[k]);
    }
}
template <class W, int R>
__device__ __forceinline__ void load_contig(const W* __restrict__ src, W (&x)[R])
{
    constexpr int BYTES = R * (int)sizeof(W);
    if constexpr (CNTT_WIDE256 && BYTES % 32 == 0) {
        if ((reinterpret_cast<uintptr_t>(src) & 31u) == 0) {
            constexpr int PER8 = 32 / (int)sizeof(W);
#pragma unroll
            for (int v = 0; v < R / PER8; v++) {
                uint32_t w[8];
                ld256(reinterpret_cast<const unsigned char*>(src) + 32 * v, w);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if constexpr (sizeof(W) == 4) x[8 * v + i] = w[i];
                }
                if constexpr (sizeof(W) == 8) {
#pragma unroll
                    for (int i = 0; i < 4; i++) x[4 * v + i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
                }
            }
            return;
        }
    }
    if constexpr (BYTES % 16 == 0) {
        constexpr int PER = 16 / (int)sizeof(W);
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
        for (int v = 0; v < R / PER; v++) {
            uint4 q = ld_data(s4 + v);
            if constexpr (sizeof(W) == 4) {
                x[4 * v] = q.x; x[4 * v + 1] = q.y; x[4 * v + 2] = q.z; x[4 * v + 3] = q.w;
            } else {
                x[2 * v] = (uint64_t)q.x | ((uint64_t)q.y << 32);
                x[2 * v + 1] = (uint64_t)q.z | ((uint64_t)q.w << 32);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < R; k++) x[k] = ld_data(src + k);
    }
}

// ---- CTA kernel -------------------------------------------------------------------------------
// nvpoly "virtual polynomials" of N = 2^LOGN words each, contiguous.  With log_sub > 0 they are the
// 2^log_sub contiguous sub-blocks of larger polynomials and virtual polynomial v uses the twiddle
// sub-tree rooted at heap node (1 << log_sub) + (v mod 2^log_sub).
// HEAD: whole transforms (log_sub == 0); the leading passes read their twiddles from the by-value `head`.
// NP: polynomials per thread group.  Every polynomial of a batch uses the same twiddles, so a thread that
// carries NP polynomials loads each twiddle once for NP butterflies: the last pass reads R-1 table entries
// per thread (2x the bytes of the data itself for u32 Shoup pairs), and L1TEX wavefronts -- 73 % of them
// global loads, mostly twiddles -- were the limiter of the u32 kernels at NP = 1 (ncu r01).
// Output staging of the forward kernel.  After the last pass a thread owns 16 consecutive words, and the conflict-free
// thread -> block map interleaves the blocks of neighbouring warps: written straight from registers, one 128-bit store
// instruction of a warp touches 32 different 128-byte lines (16 bytes each).  With CNTT_STAGE_OUT the words go through a
// swizzled shared-memory tile (chunk c of 16 bytes at c ^ ((c >> 3) & 7): conflict-free on both sides) and leave as fully
// coalesced 512-byte warp stores, for one more barrier per polynomial.  B200, prime32 batch sweep: N=256 1892 -> 2322,
// N=512 897 -> 937, N=1024 439 -> 461 M NTT/s; N=2048 and 4096 lose 3-4 % (the barrier spans 128 / 256 threads), so
// the staging was used up to N = 1024 -- until the 256-bit stores above: with whole-sector stores straight from
// registers the tile is a loss at every size (N=256 2291 -> 2331, N=512 924 -> 982, N=1024 469 -> 481 M NTT/s without it),
// so it is off; the code stays for 16-byte-aligned experiments.
#ifndef CNTT_STAGE_OUT
#define CNTT_STAGE_OUT 0
#endif
template <class A, int LOGN, int LOGR, int NP, bool FWD>
__host__ __device__ constexpr bool cta_stages_out()
{
    return CNTT_STAGE_OUT != 0 && FWD && NP == 1 && sizeof(typename A::W) == 4 && LOGR == 4 && Geo<LOGN, LOGR>::P >= 2 && LOGN <= 10;
}
// (The mirror image for the inverse kernel's strided 128-bit loads -- coalesced loads into a tile, barrier, pick up --
// measured slower: prime32 N=1024 inverse 572 -> 509 M NTT/s; loads, unlike stores, are already merged by L1.)
// resident threads per SM the 64-bit kernels are compiled for (register cap 65536 / this; 0: none).  B200, r01:
// 768 (<= 85 registers) vs uncapped: Solinas N=4096 inverse 1.52 -> 1.27 ms per 32768, N=2048 +1 %, Shoup-64 +3 %;
// 896 and 1024 lose on the N=2048 inverse.
#ifndef CNTT_CTA_MINTHREADS64
#define CNTT_CTA_MINTHREADS64 768
#endif
// 32-bit kernels (B200, r01, prime32 batch sweep): a 1024-thread cap (<= 64 registers) speeds the inverse up (N=2048
// 267 -> 276, N=4096 111 -> 122, N=65536 3.6 -> 4.1 M NTT/s) and slows the whole-transform forward down (N=2048 226 -> 213),
// whose sub-block flavour (the second level of N > 4096) gains from 1280 (3.5 -> 3.7 M NTT/s at N=65536)
#ifndef CNTT_CTA_MINTHREADS32_INV
#define CNTT_CTA_MINTHREADS32_INV 1024
#endif
#ifndef CNTT_CTA_MINTHREADS32_FWD_SUB
#define CNTT_CTA_MINTHREADS32_FWD_SUB 1280
#endif
template <class A, int LOGN, int LOGR, int GP, bool FWD, bool HEAD>
constexpr int cta_min_blocks()
{
    constexpr int want = sizeof(typename A::W) == 8 ? ((LOGN == 13 && CNTT_CTA13_64_MINTHREADS) ? CNTT_CTA13_64_MINTHREADS : CNTT_CTA_MINTHREADS64) :
                         (LOGN == 13 && (FWD ? CNTT_CTA13_MINTHREADS_FWD : CNTT_CTA13_MINTHREADS_INV)) ? (FWD ? CNTT_CTA13_MINTHREADS_FWD : CNTT_CTA13_MINTHREADS_INV) : !FWD ? CNTT_CTA_MINTHREADS32_INV : !HEAD ? CNTT_CTA_MINTHREADS32_FWD_SUB : 0;
    return want > GP * Geo<LOGN, LOGR>::T ? want / (GP * Geo<LOGN, LOGR>::T) : 0; // 0 = unspecified (not the same as 1: ptxas then keeps its default register heuristic)
}
template <class A, int LOGN, int LOGR, int GP, bool FWD, bool HEAD, int NP>
__global__ void __launch_bounds__(GP * Geo<LOGN, LOGR>::T, cta_min_blocks<A, LOGN, LOGR, GP, FWD, HEAD>())
k_ntt_cta(const typename A::Tw* __restrict__ tw, const typename A::Tw* __restrict__ tw_last, const typename A::Mod m,
          typename A::W* __restrict__ data, unsigned long long nvpoly, int log_sub, unsigned long long poly_stride,
          const __grid_constant__ TwHead<typename A::Tw> head)
{
    typedef Engine<A, LOGN, LOGR> E;
    typedef typename A::W W;
    constexpr int T = E::T, R = E::R, N = E::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* sm_all = reinterpret_cast<W*>(smem_raw);

    const int grp = (GP == 1) ? 0 : (int)(threadIdx.x / T);
    const int tid = (GP == 1) ? (int)threadIdx.x : (int)(threadIdx.x % T);
    // group g owns polynomials [g NP, g NP + NP); with log_sub > 0 they are sub-blocks and NP == 1
    const unsigned long long vp0 = ((unsigned long long)blockIdx.x * GP + grp) * NP;
    W* base[NP];
    bool active[NP];
#pragma unroll
    for (int np = 0; np < NP; np++) {
        unsigned long long vp = vp0 + np;
        active[np] = vp < nvpoly;
        if (!active[np]) vp = nvpoly - 1; // keep the group in lock-step (barriers), discard its result
        // polynomial (vp >> log_sub) starts at poly_stride words per polynomial (product plans interleave planes;
        // everywhere else poly_stride == N << log_sub and this is vp * N), sub-block (vp mod 2^log_sub) inside it
        base[np] = data + (vp >> log_sub) * poly_stride + (vp & ((1ull << log_sub) - 1ull)) * (unsigned long long)N;
    }
    const unsigned sub = HEAD ? 0u : (unsigned)((vp0 < nvpoly ? vp0 : nvpoly - 1) & ((1ull << log_sub) - 1ull));
    const unsigned nu0 = HEAD ? 1u : (1u << log_sub) + sub;
    W* sm = sm_all + (size_t)grp * NP * E::SMEM_WORDS * E::NBUF;
    const typename E::TwSrc tws = {tw, tw_last + (size_t)sub * E::LAST_WORDS, HEAD ? &head : nullptr};

    W x[NP][R];
    if constexpr (FWD) {
#pragma unroll
        for (int np = 0; np < NP; np++)
#pragma unroll
            for (int k = 0; k < R; k++) x[np][k] = ld_data(base[np] + tid + k * T);
        E::template fwd<NP>(x, sm, tws, nu0, tid, m);
#pragma unroll
        for (int np = 0; np < NP; np++) {
#pragma unroll
            for (int k = 0; k < R; k++) x[np][k] = A::canon_fwd(x[np][k], m);
            if (active[np]) store_contig<W, R>(base[np] + E::elem_last(tid, 0), x[np]);
        }
    } else {
#pragma unroll
        for (int np = 0; np < NP; np++) load_contig<W, R>(base[np] + E::elem_last(tid, 0), x[np]);
        E::template inv<NP>(x, sm, tws, nu0, tid, m);
#pragma unroll
        for (int np = 0; np < NP; np++) {
            if (active[np]) {
#pragma unroll
                for (int k = 0; k < R; k++) st_data(base[np] + tid + k * T, A::canon_inv(x[np][k], m));
            }
        }
    }
}

// ---- persistent, software-pipelined CTA kernel (whole transforms only) ------------------------------------
// The one-shot kernel above gives every thread group a single polynomial: load, wait ~1 us for HBM, transform,
// store, exit.  With CTAs this short-lived the 32-bit kernels spend their time in load latency and CTA
// turnover, not in the integer pipes (ncu r01, prime32 N=1024 fwd: fmaheavy 53 %, long_scoreboard + drain among
// the top stalls; the inverse, whose loads are 128-bit, ran 1.35x faster than the forward).  Here the grid is
// sized to the resident CTAs of the device and every group walks the batch with stride gridDim.x * GP; the
// loads of the group's NEXT polynomial are issued into a second register set before the current one is
// transformed, so HBM latency is hidden behind a whole transform.
#ifndef CNTT_PIPE32
#define CNTT_PIPE32 1
#endif
#ifndef CNTT_PIPE32_MINLOGN
#define CNTT_PIPE32_MINLOGN 13
#endif
#ifndef CNTT_PIPE32_MAXLOGN
#define CNTT_PIPE32_MAXLOGN 14 // the second register set does not fit the 64 registers of a 1024-thread CTA (N = 32768)
#endif
#ifndef CNTT_PIPE64
#define CNTT_PIPE64 0
#endif
template <class A, int LOGN, int LOGR, int GP, bool FWD, int NP>
__global__ void __launch_bounds__(GP * Geo<LOGN, LOGR>::T, cta_min_blocks<A, LOGN, LOGR, GP, FWD, true>())
k_ntt_cta_pipe(const typename A::Tw* __restrict__ tw, const typename A::Tw* __restrict__ tw_last, const typename A::Mod m,
               typename A::W* __restrict__ data, unsigned long long nvpoly, unsigned long long poly_stride,
               const __grid_constant__ TwHead<typename A::Tw> head)
{
    typedef Engine<A, LOGN, LOGR> E;
    typedef typename A::W W;
    constexpr int T = E::T, R = E::R;
    // does iteration i+1's first scatter need a barrier against iteration i's last gather?  (not when the two
    // use different ping-pong buffers, i.e. exactly two exchanges on two buffers, or when every exchange of the
    // run-time pass loop already ends with a barrier)
    constexpr bool kTailSync = E::P >= 2 && !(E::NBUF == 2 && E::P == 3) && !(E::kLoopPasses && E::NBUF == 1 && E::P >= 3);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* sm_all = reinterpret_cast<W*>(smem_raw);
    const int grp = (GP == 1) ? 0 : (int)(threadIdx.x / T);
    const int tid = (GP == 1) ? (int)threadIdx.x : (int)(threadIdx.x % T);
    W* sm = sm_all + (size_t)grp * NP * E::SMEM_WORDS * E::NBUF;
    constexpr bool kStage = cta_stages_out<A, LOGN, LOGR, NP, FWD>();
    uint4* stage = reinterpret_cast<uint4*>(sm_all + (size_t)GP * NP * E::SMEM_WORDS * E::NBUF) + (size_t)grp * (E::N / 4); // kStage only
    const typename E::TwSrc tws = {tw, tw_last, &head};
    const unsigned long long ngroups = (unsigned long long)gridDim.x * GP;
    const unsigned long long ngrp_total = (nvpoly + NP - 1) / NP;
    const unsigned long long iters = (ngrp_total + ngroups - 1) / ngroups; // uniform across the CTA (barriers)
    unsigned long long g = (unsigned long long)blockIdx.x * GP + grp;
    const int off_in = FWD ? tid : E::elem_last(tid, 0);

    auto fetch = [&](W (&dst)[NP][R], unsigned long long gi) {
#pragma unroll
        for (int np = 0; np < NP; np++) {
            unsigned long long vp = gi * NP + np;
            if (vp >= nvpoly) vp = nvpoly - 1; // keep the group in lock-step, result discarded
            const W* src = data + vp * poly_stride + off_in;
            if constexpr (FWD) {
#pragma unroll
                for (int k = 0; k < R; k++) dst[np][k] = ld_data(src + k * T);
            } else {
                load_contig<W, R>(src, dst[np]);
            }
        }
    };

    W nx[NP][R];
    fetch(nx, g);
#pragma unroll 1
    for (unsigned long long it = 0; it < iters; ++it, g += ngroups) {
        W x[NP][R];
#pragma unroll
        for (int np = 0; np < NP; np++)
#pragma unroll
            for (int k = 0; k < R; k++) x[np][k] = nx[np][k];
        if (it + 1 < iters) fetch(nx, g + ngroups);
        if constexpr (FWD) {
            E::template fwd<NP>(x, sm, tws, 1u, tid, m);
#pragma unroll
            for (int np = 0; np < NP; np++) {
#pragma unroll
                for (int k = 0; k < R; k++) x[np][k] = A::canon_fwd(x[np][k], m);
                const unsigned long long vp = g * NP + np;
                if constexpr (kStage) {
                    const int c0 = E::elem_last(tid, 0) / 4; // first of this thread's four 16-byte chunks
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        const int c = c0 + v;
                        stage[c ^ ((c >> 3) & 7)] = make_uint4((uint32_t)x[np][4 * v], (uint32_t)x[np][4 * v + 1], (uint32_t)x[np][4 * v + 2], (uint32_t)x[np][4 * v + 3]);
                    }
                    __syncthreads();
                    if (vp < nvpoly) {
                        uint4* out = reinterpret_cast<uint4*>(data + vp * poly_stride);
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const int c = r * T + tid;
                            st_data(out + c, stage[c ^ ((c >> 3) & 7)]);
                        }
                    }
                } else {
                    if (vp < nvpoly) store_contig<W, R>(data + vp * poly_stride + E::elem_last(tid, 0), x[np]);
                }
            }
        } else {
            E::template inv<NP>(x, sm, tws, 1u, tid, m);
#pragma unroll
            for (int np = 0; np < NP; np++) {
                const unsigned long long vp = g * NP + np;
                if (vp < nvpoly) {
                    W* dst = data + vp * poly_stride;
#pragma unroll
                    for (int k = 0; k < R; k++) st_data(dst + tid + k * T, A::canon_inv(x[np][k], m));
                }
            }
        }
        if constexpr (kTailSync) __syncthreads();
    }
}

// ---- persistent forward kernel with bulk-asynchronous (TMA engine) input staging: EXPERIMENT, off by default --------------------
// north_star: "TMA-staged tiles where they measurably help".  Same walk over the batch as k_ntt_cta_pipe, but the NEXT polynomial is
// fetched by one cp.async.bulk (global -> shared, completion on an mbarrier) issued by one thread per group while the current one is
// transformed, instead of R register loads per thread: no second register set, no LSU instructions for the global side.  The tile
// costs N words of shared memory per group on top of the exchange buffers, and the R loads per thread come back as R LDS.
// Measured on B200 (profiles/r02_experiments.txt, "bulk"): see the log; not shipped.
#ifndef CNTT_BULK32
#define CNTT_BULK32 0
#endif
#ifndef CNTT_BULK_MINLOGN
#define CNTT_BULK_MINLOGN 10
#endif
#ifndef CNTT_BULK_MAXLOGN
#define CNTT_BULK_MAXLOGN 13
#endif
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <class A, int LOGN, int LOGR, int GP>
__global__ void __launch_bounds__(GP * Geo<LOGN, LOGR>::T)
k_ntt_cta_bulk(const typename A::Tw* __restrict__ tw, const typename A::Tw* __restrict__ tw_last, const typename A::Mod m,
               typename A::W* __restrict__ data, unsigned long long nvpoly, unsigned long long poly_stride,
               const __grid_constant__ TwHead<typename A::Tw> head)
{
    typedef Engine<A, LOGN, LOGR> E;
    typedef typename A::W W;
    constexpr int T = E::T, R = E::R, N = E::N;
    constexpr unsigned kBytes = N * sizeof(W);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [GP stage tiles of N words][GP mbarriers][exchange buffers]
    W* stage_all = reinterpret_cast<W*>(smem_raw);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)GP * kBytes);
    W* sm_all = reinterpret_cast<W*>(smem_raw + (size_t)GP * kBytes + 128);
    const int grp = (GP == 1) ? 0 : (int)(threadIdx.x / T);
    const int tid = (GP == 1) ? (int)threadIdx.x : (int)(threadIdx.x % T);
    W* sm = sm_all + (size_t)grp * E::SMEM_WORDS * E::NBUF;
    W* stage = stage_all + (size_t)grp * N;
    const uint32_t bar = smem_u32(bars + grp), dst = smem_u32(stage);
    const typename E::TwSrc tws = {tw, tw_last, &head};
    const unsigned long long ngroups = (unsigned long long)gridDim.x * GP;
    const unsigned long long iters = (nvpoly + ngroups - 1) / ngroups; // uniform across the CTA (barriers)
    unsigned long long g = (unsigned long long)blockIdx.x * GP + grp;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](unsigned long long vp) {
        if (vp >= nvpoly) vp = nvpoly - 1;
        const W* src = data + vp * poly_stride;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(kBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(__cvta_generic_to_global(src)), "r"(kBytes), "r"(bar) : "memory");
    };
    if (tid == 0) issue(g);
    unsigned phase = 0;
#pragma unroll 1
    for (unsigned long long it = 0; it < iters; ++it, g += ngroups) {
        // wait for this group's tile
        asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}" :: "r"(bar), "r"(phase) : "memory");
        phase ^= 1u;
        W x[1][R];
#pragma unroll
        for (int k = 0; k < R; k++) x[0][k] = stage[tid + k * T];
        __syncthreads(); // every thread of the CTA has read its tile: the tile may be overwritten (and the previous iteration's exchanges are done)
        if (tid == 0 && it + 1 < iters) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy reads of the tile before its async-proxy overwrite
            issue(g + ngroups);
        }
        E::template fwd<1>(x, sm, tws, 1u, tid, m);
#pragma unroll
        for (int k = 0; k < R; k++) x[0][k] = A::canon_fwd(x[0][k], m);
        if (g < nvpoly) store_contig<W, R>(data + g * poly_stride + E::elem_last(tid, 0), x[0]);
    }
}

// ---- last-pass twiddle table builder (plan time) ---------------------------------------------------
// One thread per (sub-block, engine thread): copies the R - 1 heap entries of its last-pass sub-tree into
// the thread-innermost layout read by Engine::tw_at.  Uses the engine's own thread -> node map, so the
// table cannot disagree with the kernel that consumes it.
template <class E>
__global__ void k_build_last(const typename E::Tw* __restrict__ heap, typename E::Tw* __restrict__ out, int log_sub)
{
    constexpr int T = E::T, R = E::R, LOGR = E::G::levels(E::P - 1);
    const unsigned sub = blockIdx.x;
    for (int tid = threadIdx.x; tid < T; tid += blockDim.x) {
        const unsigned nu = E::template node<E::P - 1>(tid, (1u << log_sub) + sub);
        typename E::Tw* o = out + (size_t)sub * E::LAST_WORDS;
        for (int j = 0; j < LOGR; j++)
            for (int g = 0; g < (1 << j); g++) o[((1 << j) - 1 + g) * T + tid] = heap[(nu << j) + g];
    }
    (void)R;
}
template <class E>
cudaError_t launch_build_last_e(const typename E::Tw* heap, typename E::Tw* out, int log_sub, cudaStream_t st)
{
    if constexpr (!E::kLastXp) { (void)heap; (void)out; (void)log_sub; (void)st; return cudaSuccess; }
    else {
        k_build_last<E><<<1u << log_sub, E::T < 256 ? E::T : 256, 0, st>>>(heap, out, log_sub);
        return cudaGetLastError();
    }
}
// Sizes whose whole-transform kernel has two geometries (CtaCfg WHOLE / !WHOLE): the FORWARD transform of N = 1024 x 32-bit runs 32 words per
// thread (one warp per polynomial, one exchange) from CNTT_R32_LOGN10_MINBATCH polynomials on, and the 16-word kernel (64 threads per
// polynomial) below that and for the inverse.  B200, M NTT/s fwd / inv (profiles/r02_experiments.txt, "warpsync"):
//   batch     16 words      32 words            a single transform under a CUDA graph: 2.5 us against 3.5 us
//   9472      410 / 424     396 / 391
//   18944     475 / 512     488 / 490
//   65536     515 / 572     562 / 571
// Such plans carry both last-pass layouts, the second behind the first.
#ifndef CNTT_R32_LOGN10_MINBATCH
#define CNTT_R32_LOGN10_MINBATCH 16384
#endif
template <class A, int LOGN> constexpr bool cta_has_alt() { return CtaCfg<A, LOGN, true>::LOGR != CtaCfg<A, LOGN, false>::LOGR; }
// entries of one last-pass table (per direction) of a plan of 2^logn words
template <class A> constexpr bool cta_has_alt_rt(int logn)
{
    return logn == 9 ? cta_has_alt<A, 9>() : logn == 10 ? cta_has_alt<A, 10>() : logn == 11 ? cta_has_alt<A, 11>() : logn == 12 ? cta_has_alt<A, 12>() : false;
}
template <class A> constexpr size_t last_table_entries(int logn) { return ((size_t)1 << logn) * (cta_has_alt_rt<A>(logn) ? 2 : 1); }
// does the CTA kernel of a transform of 2^logn words (class A) read a last-pass table?
template <class A, int LOGN> constexpr bool cta_uses_last() { return CtaCfg<A, LOGN>::E::kLastXp; }
template <class A>
bool cta_uses_last_rt(int l)
{
    if (l == 13) return cta_uses_last<A, 13>();
    if constexpr (sizeof(typename A::W) == 4) {
        if (l == 14) return cta_uses_last<A, 14>();
        if constexpr (r32_size<A, 15>()) { if (l == 15) return cta_uses_last<A, 15>(); }
    }
    switch (l) {
    case 4: return cta_uses_last<A, 4>();
    case 5: return cta_uses_last<A, 5>();
    case 6: return cta_uses_last<A, 6>();
    case 7: return cta_uses_last<A, 7>();
    case 8: return cta_uses_last<A, 8>();
    case 9: return cta_uses_last<A, 9>();
    case 10: return cta_uses_last<A, 10>();
    case 11: return cta_uses_last<A, 11>();
    case 12: return cta_uses_last<A, 12>();
    default: return false;
    }
}
template <class A>
bool plan_uses_last(int logn) { return cta_uses_last_rt<A>(cta_block_logn<A>(logn, true)) || cta_uses_last_rt<A>(cta_block_logn<A>(logn, false)); }
// one size with (possibly) two geometries: blocks of a larger transform take the block layout; whole transforms the whole-transform
// layout and, where the two differ, the block layout behind it (last_table_entries)
template <class A, int L>
cudaError_t launch_build_last_size(const typename A::Tw* heap, typename A::Tw* out, int log_sub, cudaStream_t st)
{
    if (log_sub != 0) return launch_build_last_e<typename CtaCfg<A, L, false>::E>(heap, out, log_sub, st);
    if constexpr (cta_has_alt<A, L>()) {
        if (cudaError_t e = launch_build_last_e<typename CtaCfg<A, L, false>::E>(heap, out + ((size_t)1 << L), 0, st); e != cudaSuccess) return e;
    }
    return launch_build_last_e<typename CtaCfg<A, L, true>::E>(heap, out, 0, st);
}
// heap (2^logn entries) -> last-pass table (last_table_entries) for the CTA kernel this plan size launches
template <class A>
cudaError_t launch_build_last(int logn, bool fwd, const typename A::Tw* heap, typename A::Tw* out, cudaStream_t st)
{
    const int l = cta_block_logn<A>(logn, fwd);
    const int log_sub = logn - l;
    if (l == 13) return launch_build_last_e<typename CtaCfg<A, 13>::E>(heap, out, log_sub, st);
    if constexpr (sizeof(typename A::W) == 4) {
        if (l == 14) return launch_build_last_e<typename CtaCfg<A, 14>::E>(heap, out, log_sub, st);
        if constexpr (r32_size<A, 15>()) { if (l == 15) return launch_build_last_e<typename CtaCfg<A, 15>::E>(heap, out, log_sub, st); }
    }
    switch (l) {
    case 4: return launch_build_last_e<typename CtaCfg<A, 4>::E>(heap, out, log_sub, st);
    case 5: return launch_build_last_e<typename CtaCfg<A, 5>::E>(heap, out, log_sub, st);
    case 6: return launch_build_last_e<typename CtaCfg<A, 6>::E>(heap, out, log_sub, st);
    case 7: return launch_build_last_e<typename CtaCfg<A, 7>::E>(heap, out, log_sub, st);
    case 8: return launch_build_last_e<typename CtaCfg<A, 8>::E>(heap, out, log_sub, st);
    case 9: return launch_build_last_size<A, 9>(heap, out, log_sub, st);
    case 10: return launch_build_last_size<A, 10>(heap, out, log_sub, st);
    case 11: return launch_build_last_size<A, 11>(heap, out, log_sub, st);
    case 12: return launch_build_last_size<A, 12>(heap, out, log_sub, st);
    default: return cudaErrorInvalidValue;
    }
}

// ---- strided kernel ---------------------------------------------------------------------------
// Stages s0 .. s0+LOGK-1 of polynomials of 2^logn words.  One thread owns 2^LOGK words spaced
// (N >> s0) >> LOGK apart; adjacent threads own adjacent columns (coalesced).
template <class A, int LOGK, bool FWD>
__global__ void __launch_bounds__(256)
k_ntt_strided(const typename A::Tw* __restrict__ tw, const typename A::Mod m, typename A::W* __restrict__ data,
              unsigned long long nitems, int logn, int s0, unsigned long long poly_stride)
{
    typedef Engine<A, LOGK, LOGK> E; // a single register pass of LOGK levels
    typedef typename A::W W;
    constexpr int K = 1 << LOGK;
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nitems) return;
    const int log_items = logn - LOGK;            // items per polynomial
    const int log_bsub = logn - s0 - LOGK;        // column count of one block
    const unsigned long long b = idx >> log_items;
    const unsigned it = (unsigned)(idx & ((1ull << log_items) - 1ull));
    const unsigned blk = it >> log_bsub;
    const unsigned o = it & ((1u << log_bsub) - 1u);
    const size_t bsub = (size_t)1 << log_bsub;
    W* base = data + b * poly_stride + ((size_t)blk << (logn - s0)) + o;
    const unsigned nu = (1u << s0) + blk;
    W x[1][K];
    const typename E::TwSrc tws = {tw, nullptr, nullptr};
#pragma unroll
    for (int k = 0; k < K; k++) x[0][k] = base[(size_t)k * bsub];
    // Solinas, leading levels of the whole transform (s0 == 0: heap nodes < 2^LOGK <= 32): multiplier-free shift butterflies
    bool shifted = false;
    if constexpr (ShiftHead<A>::value && LOGK <= 5) shifted = s0 == 0 && m.shift_head != 0;
    if constexpr (FWD) {
        if constexpr (ShiftHead<A>::value && LOGK <= 5) {
            if (shifted) E::template shift_levels_fwd<1>(x, std::make_integer_sequence<int, LOGK>{});
        }
        if (!shifted) E::template fwd_pass<0, 1>(x, tws, nu, 0, m);
#pragma unroll
        for (int k = 0; k < K; k++) base[(size_t)k * bsub] = x[0][k]; // lazy range, consumed by the next level
    } else {
        if constexpr (ShiftHead<A>::value && LOGK <= 5) {
            if (shifted) E::template shift_levels_inv<1>(x, std::make_integer_sequence<int, LOGK>{});
        }
        if (!shifted) E::template inv_pass<0, 1>(x, tws, nu, 0, m);
#pragma unroll
        for (int k = 0; k < K; k++) base[(size_t)k * bsub] = A::canon_inv(x[0][k], m);
    }
}

// ---- cluster kernel: EXPERIMENT (CNTT_CLUSTER32, off by default) ---------------------------------------------------
// One polynomial of 2^LOGN 32-bit words per thread-block cluster of C = 2^LOGC CTAs: a single launch that reads and writes every
// word once, for a size whose exchange buffer does not fit one CTA at useful occupancy.  The cluster as a whole is the engine
// Engine<A, LOGN, 5> (thread u = rank * T + tid): its pass 0 pairs words N/2 .. N/32 apart, i.e. words of different CTAs' halves,
// and runs in registers on words read straight from global memory.  The exchange after pass 0 is the only one that crosses CTAs:
// register slot k belongs to CTA k / (32 / C) afterwards -- a compile-time property of the slot -- so every thread PUSHES its words
// into the owner's shared memory (st.shared::cluster; remote stores are fire-and-forget, remote loads would wait ~215 cycles each),
// one cluster barrier, and from there each CTA is the sub-block engine Engine<A, LOGN - LOGC, 5> with nu0 = C + rank on local
// shared memory only.  The inverse mirrors it (push in the pass-0 layout after the local passes).
__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
__device__ __forceinline__ uint32_t cluster_map(const void* smem_ptr, unsigned rank) // this CTA's shared address -> CTA `rank`'s
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

template <class A, int LOGN, int LOGC, bool FWD>
__global__ void __launch_bounds__(Geo<LOGN - LOGC, 5>::T, 2)
k_ntt_cluster(const typename A::Tw* __restrict__ tw, const typename A::Tw* __restrict__ tw_last, const typename A::Mod m,
              typename A::W* __restrict__ data, unsigned long long poly_stride, const __grid_constant__ TwHead<typename A::Tw> head)
{
    typedef Engine<A, LOGN, 5> EC;          // the cluster: only its pass 0 is used
    typedef Engine<A, LOGN - LOGC, 5> ES;   // one CTA's sub-block: passes 1 .. P-1
    typedef typename A::W W;
    static_assert(sizeof(W) == 4, "32-bit words");
    static_assert(EC::P == ES::P && EC::G::R1 == ES::G::R1 + LOGC && EC::G::R1 <= 5, "pass 0 of the cluster = LOGC levels + pass 0 of the sub-block");
    static_assert(!ES::kXor && !ES::kLoopPasses, "padded layout, compile-time pass chain");
    constexpr int C = 1 << LOGC, T = ES::T, TC = EC::T, R = 32, NS = ES::N, PER = R / C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* sm = reinterpret_cast<W*>(smem_raw);
    const unsigned rank = cluster_ctarank();
    const int tid = (int)threadIdx.x, u = (int)rank * T + tid;
    W* base = data + (unsigned long long)(blockIdx.x >> LOGC) * poly_stride;
    const typename EC::TwSrc twc = {tw, nullptr, &head};
    const typename ES::TwSrc tws = {tw, tw_last + (size_t)rank * ES::LAST_WORDS, nullptr};
    const unsigned nu0 = (unsigned)C + rank;
    W x[1][R];
    if constexpr (FWD) {
        cluster_arrive(); // a CTA's shared memory may only be written once that CTA runs: everybody says so first ...
#pragma unroll
        for (int k = 0; k < R; k++) x[0][k] = ld_data(base + u + k * TC);
        EC::template fwd_pass<0, 1>(x, twc, 1u, u, m);
        cluster_wait();   // ... and is heard after the loads and pass 0
        // slot k = owner (k / PER), local element (k % PER) * TC + u, in the padded layout gather<1> of the sub-block engine reads
#pragma unroll
        for (int d = 0; d < C; d++) {
            if ((unsigned)d == rank) {
#pragma unroll
                for (int kk = 0; kk < PER; kk++) sm[pad_idx<W>(kk * TC + u)] = x[0][d * PER + kk];
            } else {
                const uint32_t rb = cluster_map(sm, (unsigned)d);
#pragma unroll
                for (int kk = 0; kk < PER; kk++) st_cluster(rb + 4u * (uint32_t)pad_idx<W>(kk * TC + u), x[0][d * PER + kk]);
            }
        }
        cluster_sync_all();
        ES::template gather<1, 1>(x, sm, tid);
        if constexpr (ES::NBUF == 1 && 2 < ES::P) __syncthreads();
        ES::template fwd_from<1, 1>(x, sm, tws, nu0, tid, m);
#pragma unroll
        for (int k = 0; k < R; k++) x[0][k] = A::canon_fwd(x[0][k], m);
        store_contig<W, R>(base + (size_t)rank * NS + ES::elem_last(tid, 0), x[0]);
    } else {
        load_contig<W, R>(base + (size_t)rank * NS + ES::elem_last(tid, 0), x[0]);
        ES::template inv_down<ES::P - 1, 1, 1>(x, sm, tws, nu0, tid, m);
        // pass-1 layout of the sub-block: element blk * B1 + o + k * S1 (blk = tid / S1, o = tid % S1), global element rank * NS + that
        // = slot (rank * PER + blk) of cluster thread (o + k * S1) in the pass-0 layout: owner CTA (o + k S1) / T, its thread (o + k S1) % T
        constexpr int S1 = ES::template stride<1>();
        static_assert(ES::template blk_words<1>() == TC && S1 * R == TC, "pass-1 blocks of the sub-block = pass-0 stride of the cluster");
        const int blk = tid / S1, o = tid % S1;
        const int slot = (int)rank * PER + blk;
        cluster_sync_all(); // every CTA of the cluster is done reading its exchange buffer
#pragma unroll
        for (int d = 0; d < C; d++) {
            constexpr int KPER = T / S1; // register slots per owner
            if ((unsigned)d == rank) {
#pragma unroll
                for (int kk = 0; kk < KPER; kk++) sm[slot * T + o + kk * S1] = x[0][d * KPER + kk];
            } else {
                const uint32_t rb = cluster_map(sm, (unsigned)d);
#pragma unroll
                for (int kk = 0; kk < KPER; kk++) st_cluster(rb + 4u * (uint32_t)(slot * T + o + kk * S1), x[0][d * KPER + kk]);
            }
        }
        cluster_sync_all();
#pragma unroll
        for (int k = 0; k < R; k++) x[0][k] = sm[k * T + tid];
        EC::template inv_pass<0, 1>(x, twc, 1u, u, m);
#pragma unroll
        for (int k = 0; k < R; k++) st_data(base + u + k * TC, A::canon_inv(x[0][k], m));
    }
}

// ---- pointwise --------------------------------------------------------------------------------
enum PointwiseOp { OP_MUL_ASSIGN_NORMALIZE = 0, OP_NORMALIZE = 1, OP_MUL_ACCUMULATE = 2 };

template <class A, int OP>
__global__ void __launch_bounds__(256)
k_pointwise(const typename A::Mod m, typename A::W* __restrict__ dst, const typename A::W* __restrict__ a,
            const typename A::W* __restrict__ b, unsigned long long nvec)
{
    typedef typename A::W W;
    constexpr int PER = 16 / (int)sizeof(W);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        W d[PER], xa[PER], xb[PER];
        load_contig<W, PER>(dst + v * PER, d);
        if constexpr (OP != OP_NORMALIZE) load_contig<W, PER>(a + v * PER, xa);
        if constexpr (OP == OP_MUL_ACCUMULATE) load_contig<W, PER>(b + v * PER, xb);
#pragma unroll
        for (int i = 0; i < PER; i++) {
            if constexpr (OP == OP_MUL_ASSIGN_NORMALIZE) d[i] = A::mul_norm(d[i], xa[i], m);
            else if constexpr (OP == OP_NORMALIZE) d[i] = A::norm(d[i], m);
            else d[i] = A::mul_acc(d[i], xa[i], xb[i], m);
        }
        store_contig<W, PER>(dst + v * PER, d);
    }
}

// Same streams when the polynomials of a batch are poly_stride words apart (product plans: the planes of one
// polynomial are interleaved with the other primes' planes).  blockIdx.y = polynomial.
template <class A, int OP>
__global__ void __launch_bounds__(256)
k_pointwise_strided(const typename A::Mod m, typename A::W* __restrict__ dst, const typename A::W* __restrict__ a,
                    const typename A::W* __restrict__ b, unsigned nvec_per_poly, unsigned long long poly_stride)
{
    typedef typename A::W W;
    constexpr int PER = 16 / (int)sizeof(W);
    const unsigned long long off = (unsigned long long)blockIdx.y * poly_stride;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec_per_poly; v += gridDim.x * blockDim.x) {
        W d[PER], xa[PER], xb[PER];
        load_contig<W, PER>(dst + off + (size_t)v * PER, d);
        if constexpr (OP != OP_NORMALIZE) load_contig<W, PER>(a + off + (size_t)v * PER, xa);
        if constexpr (OP == OP_MUL_ACCUMULATE) load_contig<W, PER>(b + off + (size_t)v * PER, xb);
#pragma unroll
        for (int i = 0; i < PER; i++) {
            if constexpr (OP == OP_MUL_ASSIGN_NORMALIZE) d[i] = A::mul_norm(d[i], xa[i], m);
            else if constexpr (OP == OP_NORMALIZE) d[i] = A::norm(d[i], m);
            else d[i] = A::mul_acc(d[i], xa[i], xb[i], m);
        }
        store_contig<W, PER>(dst + off + (size_t)v * PER, d);
    }
}

// ---- launchers --------------------------------------------------------------------------------
#ifndef CNTT_NP32
#define CNTT_NP32 1
#endif
template <class A, int LOGN, bool FWD, int NP, bool WHOLE = true>
cudaError_t launch_cta_np(const PlanDev<A>& pl, typename A::W* data, unsigned long long nvpoly, int log_sub, size_t poly_stride, cudaStream_t st)
{
    constexpr int LOGR = CtaCfg<A, LOGN, WHOLE>::LOGR;
    typedef typename CtaCfg<A, LOGN, WHOLE>::E E;
    constexpr int T = E::T;
    constexpr int GP = T >= 128 ? 1 : 128 / T;
    const size_t smem_xchg = (size_t)GP * NP * E::NBUF * E::SMEM_WORDS * sizeof(typename A::W);
    const size_t smem = smem_xchg;
    const unsigned long long ngrp = (nvpoly + NP - 1) / NP;
    const unsigned long long nblk = (ngrp + GP - 1) / GP;
    if (nblk == 0) return cudaSuccess;
    if (nblk > 0x7fffffffull) return cudaErrorInvalidValue;
    const typename A::Tw* last = FWD ? pl.tw_fwd_last : pl.tw_inv_last;
    if (E::kLastXp && last == nullptr) return cudaErrorInvalidValue; // plan built without its last-pass table
    if constexpr (!WHOLE) { if (log_sub == 0 && last != nullptr) last += (size_t)1 << LOGN; } // whole transform on the block geometry: second layout
    const TwHead<typename A::Tw>* head = FWD ? pl.head_fwd : pl.head_inv;
    auto launch = [&](auto kern, const TwHead<typename A::Tw>& h) -> cudaError_t {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)nblk, GP * T, smem, st>>>(FWD ? pl.tw_fwd : pl.tw_inv, last, pl.mod, data, nvpoly, log_sub, poly_stride, h);
        return cudaGetLastError();
    };
    // measured on B200 (prime32, batch 65536): forward +6 % at N = 1024..4096; the inverse loses 14 % (its 128-bit
    // loads were never latency-bound and the second register set costs occupancy), so only the forward pipelines
    // (re-measured after the 256-bit stores: the one-shot kernel now wins up to N = 4096 -- N=256 2311 -> 2631, N=1024 478 -> 517
    // M NTT/s, N=2048 / 4096 within 1 % -- and the persistent one keeps N = 8192, 42.3 -> 45.9, where only two CTAs are resident)
    constexpr bool kBulk = CNTT_BULK32 != 0 && FWD && NP == 1 && sizeof(typename A::W) == 4 && LOGN >= CNTT_BULK_MINLOGN && LOGN <= CNTT_BULK_MAXLOGN;
    if constexpr (kBulk) {
        if (log_sub == 0 && head != nullptr && poly_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(data) & 15u) == 0) {
            auto kern = k_ntt_cta_bulk<A, LOGN, LOGR, GP>;
            const size_t smem_b = smem_xchg + (size_t)GP * E::N * sizeof(typename A::W) + 128;
            static std::atomic<int> resident_b[64];
            int dev = 0;
            cudaError_t e = cudaGetDevice(&dev);
            if (e != cudaSuccess) return e;
            if (dev < 0 || dev >= 64) dev = 0;
            int res = resident_b[dev].load(std::memory_order_relaxed);
            if (res == 0) {
                if ((e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem_b)) != cudaSuccess) return e;
                int per_sm = 0, sms = 0;
                if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GP * T, smem_b)) != cudaSuccess) return e;
                if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
                res = per_sm * sms > 0 ? per_sm * sms : 1;
                resident_b[dev].store(res, std::memory_order_relaxed);
            }
            if (nblk >= 2ull * (unsigned long long)res) {
                kern<<<(unsigned)res, GP * T, smem_b, st>>>(pl.tw_fwd, last, pl.mod, data, nvpoly, poly_stride, *head);
                return cudaGetLastError();
            }
        }
    }
    constexpr bool kPipe = FWD && (sizeof(typename A::W) == 4 ? (CNTT_PIPE32 != 0 && LOGN >= CNTT_PIPE32_MINLOGN && LOGN <= CNTT_PIPE32_MAXLOGN && !r32_size<A, LOGN>()) : CNTT_PIPE64 != 0);
    if constexpr (kPipe) {
        // persistent variant: needs whole transforms and at least two polynomials per resident group to pipeline
        if (log_sub == 0 && head != nullptr) {
            auto kern = k_ntt_cta_pipe<A, LOGN, LOGR, GP, FWD, NP>;
            const size_t smem = smem_xchg + (cta_stages_out<A, LOGN, LOGR, NP, FWD>() ? (size_t)GP * E::N * sizeof(typename A::W) : 0);
            static std::atomic<int> resident[64]; // CTAs the device holds at once, per device ordinal (0: not queried yet)
            int dev = 0;
            cudaError_t e = cudaGetDevice(&dev);
            if (e != cudaSuccess) return e;
            if (dev < 0 || dev >= 64) dev = 0;
            int res = resident[dev].load(std::memory_order_relaxed);
            if (res == 0) {
                if ((e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem)) != cudaSuccess) return e;
                int per_sm = 0, sms = 0;
                if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GP * T, smem)) != cudaSuccess) return e;
                if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
                res = per_sm * sms > 0 ? per_sm * sms : 1;
                resident[dev].store(res, std::memory_order_relaxed); // racing threads compute the same value
            }
            if (nblk >= 2ull * (unsigned long long)res) {
                kern<<<(unsigned)res, GP * T, smem, st>>>(FWD ? pl.tw_fwd : pl.tw_inv, last, pl.mod, data, nvpoly, poly_stride, *head);
                return cudaGetLastError();
            }
        }
    }
    if (log_sub == 0 && head != nullptr) return launch(k_ntt_cta<A, LOGN, LOGR, GP, FWD, true, NP>, *head);
    static const TwHead<typename A::Tw> none = {};
    if constexpr (NP == 1) return launch(k_ntt_cta<A, LOGN, LOGR, GP, FWD, false, 1>, none);
    else return cudaErrorInvalidValue; // sub-block launches carry one block per group
}
// u32 whole transforms carry CNTT_NP32 polynomials per thread group (twiddle reuse); sub-blocks of a large
// transform, 64-bit words and tiny batches carry one.
template <class A, int LOGN, bool FWD>
cudaError_t launch_cta_one(const PlanDev<A>& pl, typename A::W* data, unsigned long long nvpoly, int log_sub, size_t poly_stride, cudaStream_t st)
{
    constexpr int NPW = (sizeof(typename A::W) == 4 && CtaCfg<A, LOGN>::E::P >= 2) ? CNTT_NP32 : 1;
    if constexpr (NPW > 1) {
        const TwHead<typename A::Tw>* head = FWD ? pl.head_fwd : pl.head_inv;
        if (log_sub == 0 && head != nullptr && nvpoly >= 2ull * NPW * 148ull) return launch_cta_np<A, LOGN, FWD, NPW>(pl, data, nvpoly, log_sub, poly_stride, st);
    }
    if constexpr (cta_has_alt<A, LOGN>()) {
        if (log_sub > 0 || !FWD || nvpoly < (unsigned long long)CNTT_R32_LOGN10_MINBATCH) return launch_cta_np<A, LOGN, FWD, 1, false>(pl, data, nvpoly, log_sub, poly_stride, st);
    }
    return launch_cta_np<A, LOGN, FWD, 1>(pl, data, nvpoly, log_sub, poly_stride, st);
}

template <class A, bool FWD>
cudaError_t launch_cta(const PlanDev<A>& pl, int logn_sub, typename A::W* data, unsigned long long nvpoly, int log_sub, size_t poly_stride, cudaStream_t st)
{
    if (logn_sub == 13) return launch_cta_one<A, 13, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    if constexpr (sizeof(typename A::W) == 4) {
        if (logn_sub == 14) return launch_cta_one<A, 14, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
        if constexpr (r32_size<A, 15>()) { if (logn_sub == 15) return launch_cta_one<A, 15, FWD>(pl, data, nvpoly, log_sub, poly_stride, st); }
    }
    switch (logn_sub) {
    case 4: return launch_cta_one<A, 4, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    case 5: return launch_cta_one<A, 5, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    case 6: return launch_cta_one<A, 6, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    case 7: return launch_cta_one<A, 7, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    case 8: return launch_cta_one<A, 8, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    case 9: return launch_cta_one<A, 9, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    case 10: return launch_cta_one<A, 10, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    case 11: return launch_cta_one<A, 11, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    case 12: return launch_cta_one<A, 12, FWD>(pl, data, nvpoly, log_sub, poly_stride, st);
    default: return cudaErrorInvalidValue;
    }
}

template <class A, int LOGK, bool FWD>
cudaError_t launch_strided_one(const PlanDev<A>& pl, typename A::W* data, size_t batch, int s0, size_t poly_stride, cudaStream_t st)
{
    const unsigned long long nitems = (unsigned long long)batch << (pl.logn - LOGK);
    const unsigned long long nblk = (nitems + 255) / 256;
    if (nblk == 0) return cudaSuccess;
    if (nblk > 0x7fffffffull) return cudaErrorInvalidValue;
    k_ntt_strided<A, LOGK, FWD><<<(unsigned)nblk, 256, 0, st>>>(FWD ? pl.tw_fwd : pl.tw_inv, pl.mod, data, nitems, pl.logn, s0, poly_stride);
    return cudaGetLastError();
}
template <class A, bool FWD>
cudaError_t launch_strided(const PlanDev<A>& pl, int logk, typename A::W* data, size_t batch, int s0, size_t poly_stride, cudaStream_t st)
{
    switch (logk) {
    case 1: return launch_strided_one<A, 1, FWD>(pl, data, batch, s0, poly_stride, st);
    case 2: return launch_strided_one<A, 2, FWD>(pl, data, batch, s0, poly_stride, st);
    case 3: return launch_strided_one<A, 3, FWD>(pl, data, batch, s0, poly_stride, st);
    case 4: return launch_strided_one<A, 4, FWD>(pl, data, batch, s0, poly_stride, st);
    default: break;
    }
    if constexpr (ShiftHead<A>::value && CNTT_STRIDED_MAXK64S >= 5) { if (logk == 5) return launch_strided_one<A, 5, FWD>(pl, data, batch, s0, poly_stride, st); }
    if constexpr (sizeof(typename A::W) == 4 && CNTT_STRIDED_MAXK32 >= 5) { if (logk == 5) return launch_strided_one<A, 5, FWD>(pl, data, batch, s0, poly_stride, st); }
    if constexpr (sizeof(typename A::W) == 4 && CNTT_STRIDED_MAXK32 >= 6) { if (logk == 6) return launch_strided_one<A, 6, FWD>(pl, data, batch, s0, poly_stride, st); }
    return cudaErrorInvalidValue;
}

template <class A, int LOGN, int LOGC, bool FWD>
cudaError_t launch_cluster(const PlanDev<A>& pl, typename A::W* data, size_t batch, size_t poly_stride, cudaStream_t st)
{
    typedef Engine<A, LOGN - LOGC, 5> ES;
    const TwHead<typename A::Tw>* head = FWD ? pl.head_fwd : pl.head_inv;
    if (head == nullptr) return cudaErrorInvalidValue;
    if (((unsigned long long)batch << LOGC) > 0x7fffffffull) return cudaErrorInvalidValue;
    auto kern = k_ntt_cluster<A, LOGN, LOGC, FWD>;
    const size_t smem = (size_t)ES::SMEM_WORDS * ES::NBUF * sizeof(typename A::W);
    if (cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem); e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(batch << LOGC));
    cfg.blockDim = dim3(ES::T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1 << LOGC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, FWD ? pl.tw_fwd : pl.tw_inv, FWD ? pl.tw_fwd_last : pl.tw_inv_last, pl.mod, data,
                              (unsigned long long)poly_stride, *head);
}

// Full transform of `batch` polynomials.  Sizes one CTA holds (cta_block_logn): one CTA-kernel launch.  Larger: leading stages
// strided (kmax per launch) until the remaining contiguous blocks have the size cta_block_logn names, then the CTA
// kernel on all batch * 2^s blocks; the inverse runs the same schedule backwards.
// poly_stride: distance in words between consecutive polynomials of the batch (0: contiguous, = N)
template <class A, bool FWD>
cudaError_t launch_ntt(const PlanDev<A>& pl, typename A::W* data, size_t batch, cudaStream_t st, size_t poly_stride = 0)
{
    if (batch == 0) return cudaSuccess;
    if (poly_stride == 0) poly_stride = (size_t)1 << pl.logn;
    const int blk = cta_block_logn<A>(pl.logn, FWD);
    if (pl.logn == blk) return launch_cta<A, FWD>(pl, pl.logn, data, batch, 0, poly_stride, st);
    const int lead = pl.logn - blk;
#if CNTT_CLUSTER32
    if constexpr (sizeof(typename A::W) == 4) {
        if (pl.logn == 15 && blk == 14 && (FWD ? pl.head_fwd : pl.head_inv) != nullptr) return launch_cluster<A, 15, 1, FWD>(pl, data, batch, poly_stride, st);
    }
#endif
    constexpr int kmax = sizeof(typename A::W) == 4 ? (FWD ? CNTT_STRIDED_MAXK32_FWD : CNTT_STRIDED_MAXK32_INV) : ShiftHead<A>::value ? CNTT_STRIDED_MAXK64S : 4;
    cudaError_t e;
    if constexpr (FWD) {
        int s0 = 0;
        while (s0 < lead) {
            const int k = (lead - s0) < kmax ? (lead - s0) : kmax;
            if ((e = launch_strided<A, true>(pl, k, data, batch, s0, poly_stride, st)) != cudaSuccess) return e;
            s0 += k;
        }
        return launch_cta<A, true>(pl, blk, data, (unsigned long long)batch << lead, lead, poly_stride, st);
    } else {
        if ((e = launch_cta<A, false>(pl, blk, data, (unsigned long long)batch << lead, lead, poly_stride, st)) != cudaSuccess) return e;
        // mirror of the forward schedule: last forward chunk first
        int chunks[8], nc = 0, s = 0;
        while (s < lead) { const int k = (lead - s) < kmax ? (lead - s) : kmax; chunks[nc++] = k; s += k; }
        for (int c = nc - 1; c >= 0; c--) {
            s -= chunks[c];
            if ((e = launch_strided<A, false>(pl, chunks[c], data, batch, s, poly_stride, st)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
}

template <class A, int OP>
cudaError_t launch_pointwise(const PlanDev<A>& pl, typename A::W* dst, const typename A::W* a, const typename A::W* b,
                             size_t nwords, cudaStream_t st)
{
    constexpr int PER = 16 / (int)sizeof(typename A::W);
    const unsigned long long nvec = nwords / PER; // n >= 16 words, so nwords is a multiple of PER
    if (nvec == 0) return cudaSuccess;
    unsigned long long nblk = (nvec + 255) / 256;
    const unsigned long long cap = 148ull * 32ull;
    if (nblk > cap) nblk = cap;
    k_pointwise<A, OP><<<(unsigned)nblk, 256, 0, st>>>(pl.mod, dst, a, b, nvec);
    return cudaGetLastError();
}

// pointwise op on `batch` polynomials of 2^logn words, poly_stride words apart (all three streams share the layout)
template <class A, int OP>
cudaError_t launch_pointwise_strided(const PlanDev<A>& pl, typename A::W* dst, const typename A::W* a, const typename A::W* b,
                                     size_t batch, size_t poly_stride, cudaStream_t st)
{
    constexpr int PER = 16 / (int)sizeof(typename A::W);
    const unsigned nvec = (unsigned)(((size_t)1 << pl.logn) / PER);
    const unsigned gx = (nvec + 255) / 256;
    for (size_t b0 = 0; b0 < batch; b0 += 65535) { // grid.y limit
        const size_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
        k_pointwise_strided<A, OP><<<dim3(gx, (unsigned)nb), 256, 0, st>>>(pl.mod, dst + b0 * poly_stride, a ? a + b0 * poly_stride : a,
                                                                         b ? b + b0 * poly_stride : b, nvec, poly_stride);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

} // namespace cntt
