// ntt_engine.cuh -- the register/shared-memory NTT engine shared by every kernel in the library.
//
// Transform (reference semantics, SURVEY.md section 0 / prime32.rs:704-708):
//   fwd: natural-order in -> bit-reversed out, merged-psi Cooley-Tukey; stage with m blocks uses
//        twid[m + block] (heap order: twid[brv(k)] = psi^k, prime32.rs:248-282)
//   inv: bit-reversed in -> natural out, Gentleman-Sande, un-normalised (x N)
//
// Decomposition (B200-first, not the reference's loop nest):
//   A polynomial of N = 2^LOGN words is owned by T = N/R threads, R = 2^LOGR words per thread in
//   registers.  The LOGN stages are cut into P = ceil(LOGN/LOGR) "passes".  Inside a pass every
//   thread runs up to LOGR butterfly levels entirely in registers; between passes the words are
//   re-distributed through (padded, conflict-free) shared memory.  A pass owned by a thread is a
//   sub-tree of the twiddle heap rooted at node `nu`; level j of the pass uses heap entries
//   (nu << j) + g, g < 2^j.
//   Pass 0 covers the first R1 = LOGN - (P-1) LOGR stages with layout   i = tid + k T
//   Pass q>=1 covers LOGR stages at s0 = R1 + (q-1) LOGR with layout    i = blk B + o + k S,
//        B = N >> s0, S = B / R, blk = tid / S, o = tid % S,  nu = (nu0 << s0) + blk
//   so the last pass leaves R consecutive words per thread (vector stores), and the first pass
//   reads words strided by T (coalesced).  `nu0` = (1 << depth) + half lets the same code run as
//   the contiguous second level of a two-level large-N transform -- exactly the reference's
//   (recursion_depth, recursion_half) call (prime32/shoup.rs:597,686-706).
#pragma once
#include "arith.cuh"
#include <atomic>
#include <utility>

namespace cntt {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-function, per-device property: set it the first time a kernel that
// needs more than 48 KB is launched on a device, not on every launch (r01: 27.6 us per fwd + inv round trip at batch 1
// through two calls against 4.6 us under a graph -- most of the gap was this call).  Lock-free open-addressing table
// keyed by the kernel's address; a lost race just sets the attribute twice.
inline cudaError_t ensure_dyn_smem(const void* kern, size_t smem)
{
    if (smem <= 48 * 1024) return cudaSuccess;
    constexpr unsigned kSlots = 1024;
    static std::atomic<const void*> key[kSlots];
    static std::atomic<unsigned long long> devmask[kSlots];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = (dev >= 0 && dev < 64) ? 1ull << dev : 0ull;
    unsigned h = (unsigned)((reinterpret_cast<uintptr_t>(kern) >> 4) * 2654435761u) % kSlots;
    for (unsigned probe = 0; probe < kSlots; probe++, h = (h + 1) % kSlots) {
        const void* k = key[h].load(std::memory_order_acquire);
        if (k == nullptr) {
            const void* expect = nullptr;
            if (!key[h].compare_exchange_strong(expect, kern, std::memory_order_acq_rel) && expect != kern) continue;
            k = kern;
        }
        if (k != kern) continue;
        if (bit && (devmask[h].load(std::memory_order_acquire) & bit)) return cudaSuccess;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess && bit) devmask[h].fetch_or(bit, std::memory_order_acq_rel);
        return e;
    }
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); // table full: uncached
}

// FF ("first full"): when LOGN is not a multiple of LOGR one pass is short.  By default it is pass 0 (R1 < LOGR levels); with FF
// pass 0 runs LOGR levels and the LAST pass is the short one: it keeps the last-pass layout (R consecutive words per thread) and
// simply skips the first LOGR - RL levels of its LOGR-level frame, i.e. it handles 2^(LOGR-RL) adjacent blocks of 2^RL words.
// Used for the Solinas class, whose pass-0 levels are multiplier-free (ShiftHead): N = 2048 then has four of them instead of three.
template <int LOGN, int LOGR, bool FF = false>
struct Geo {
    static_assert(LOGN >= LOGR, "need at least R words per polynomial");
    static constexpr int N = 1 << LOGN;
    static constexpr int R = 1 << LOGR;
    static constexpr int T = N / R;                           // threads per polynomial
    static constexpr int P = (LOGN + LOGR - 1) / LOGR;        // passes
    static constexpr int RX = LOGN - (P - 1) * LOGR;          // levels of the short pass
    static constexpr bool kFF = FF && P >= 2 && RX < LOGR;
    static constexpr int R1 = kFF ? LOGR : RX;                // levels in pass 0
    static constexpr int RL = P == 1 ? RX : kFF ? RX : LOGR;  // levels in the last pass
    static __host__ __device__ constexpr int s0(int q) { return q == 0 ? 0 : kFF ? (q == P - 1 ? LOGN - LOGR : q * LOGR) : R1 + (q - 1) * LOGR; }
    static __host__ __device__ constexpr int levels(int q) { return q == 0 ? R1 : q == P - 1 ? RL : LOGR; }
    // first level of pass q inside its LOGR-level frame (level j of the frame pairs slots R >> (j + 1) apart)
    static __host__ __device__ constexpr int jlo(int q) { return (kFF && q == P - 1) ? LOGR - RL : 0; }
};
template <class A> struct FirstFull { static constexpr bool value = false; };
template <> struct FirstFull<A64S> { static constexpr bool value = CNTT_FIRST_FULL_64S != 0; };

// shared-memory padding: one extra word per 128-byte row makes every power-of-two stride
// (<= one row) conflict-free for both the scatter and the gather side of an exchange.
template <class W> struct PadCfg { static constexpr int LOGROW = (sizeof(W) == 4) ? 5 : 4; };
template <class W> __host__ __device__ constexpr int padded_words(int n) { return n + (n >> PadCfg<W>::LOGROW); }
template <class W> __device__ __forceinline__ int pad_idx(int i) { return i + (i >> PadCfg<W>::LOGROW); }

// The first kHeadEntries entries of a heap table, passed BY VALUE as a __grid_constant__ kernel parameter: the
// leading passes of a transform use the same few twiddles in every thread (pass 0) or in every thread of a
// warp (middle passes whose blocks span >= 32 threads), so they are read from the constant bank -- as
// immediate c[0x0][..] operands when the index is a compile-time value -- instead of through L1TEX, which was
// the limiter of the u32 kernels (ncu r01: l1tex 82-86 % busy, a third of its requests were twiddle loads).
// exchange buffers per polynomial for 64-bit words: 1 = one buffer and a second barrier per exchange.  The 64-bit
// kernels are occupancy-limited by shared memory (2 x 17 KB per polynomial at N = 2048), and the extra warps
// are worth more than the saved barrier (B200, Solinas N = 2048: fwd 1.172 -> 1.132 ms, inv 1.348 -> 1.161 ms).
#ifndef CNTT_NBUF64
#define CNTT_NBUF64 1
#endif
#ifndef CNTT_WARPSYNC
#define CNTT_WARPSYNC 1
#endif
#ifndef CNTT_NBUF32
#define CNTT_NBUF32 2
#endif
#ifndef CNTT_HEAD_MINS
#define CNTT_HEAD_MINS 16
#endif
// ping-pong exchange buffers only while both fit in this many bytes per polynomial: beyond N = 4096 x u32 (2 x 16.9 KB) the
// second buffer costs resident CTAs, which are worth more than the saved barrier (ntt_kernels.cuh, CNTT_R32_MINLOGN)
#ifndef CNTT_NBUF2_MAXBYTES
#define CNTT_NBUF2_MAXBYTES 40000
#endif
constexpr int kHeadLog = 8;
constexpr int kHeadEntries = 1 << kHeadLog;
template <class Tw> struct TwHead { Tw e[kHeadEntries * 8 / (sizeof(Tw) < 8 ? 8 : sizeof(Tw))]; }; // 2 KB of parameter space
template <class Tw> __host__ __device__ constexpr int head_log() { return sizeof(Tw) <= 8 ? kHeadLog : kHeadLog - 1; }

template <class A, int LOGN, int LOGR>
struct Engine {
    typedef Geo<LOGN, LOGR, FirstFull<A>::value> G;
    typedef typename A::W W;
    typedef typename A::Tw Tw;
    typedef typename A::Mod Mod;
    static constexpr int N = G::N, R = G::R, T = G::T, P = G::P;
    static constexpr int LOGROW = PadCfg<W>::LOGROW;
    static constexpr int PH = 1 << LOGROW;                      // lanes served by one shared-memory phase

    // Exchange barrier of ONE polynomial group.  A group of at most 32 threads is (part of) one warp -- groups are aligned to their
    // size inside the CTA -- so a warp-level barrier orders its shared-memory traffic, and the other groups of the CTA are not
    // stalled on it (CNTT_WARPSYNC; profiles/r02_experiments.txt, "warpsync")
    static __device__ __forceinline__ void gsync()
    {
        if constexpr (CNTT_WARPSYNC != 0 && T <= 32) __syncwarp();
        else __syncthreads();
    }

    // ---- layouts ------------------------------------------------------------------------
    // Pass q >= 1 works on blocks of B = N >> s0(q) words, S = B / R apart inside a thread.  When S is
    // smaller than a phase, one phase touches PH / S different blocks; with the padded layout those
    // blocks must be D = PH / R apart to land on distinct banks, so the thread -> block map is permuted
    // (blocks {b0, b0 + D, ...} share a phase).  Verified conflict-free by enumeration for every
    // (LOGN, LOGR, word size) the library instantiates.  When a polynomial has too
    // few phases for that permutation (small N) the engine switches to an XOR swizzle
    // i ^ ((i >> LOGR) & (PH - 1)) of the unpadded index, which is conflict-free for every pass.
    template <int Q> static __host__ __device__ constexpr int blk_words() { return Q == 0 ? N : (N >> G::s0(Q)); }
    template <int Q> static __host__ __device__ constexpr int stride() { return Q == 0 ? T : (blk_words<Q>() >> LOGR); }
    template <int Q> static __host__ __device__ constexpr bool wants_perm()
    {
        return Q >= 1 && stride<Q>() < PH && blk_words<Q>() >= PH && R < PH;
    }
    static constexpr int D = (R < PH) ? PH / R : 1;
    static constexpr bool kPermFeasible = (T / PH) >= D && ((T / PH) % D) == 0;
    template <int Q> static __host__ __device__ constexpr bool any_wants_perm()
    {
        if constexpr (Q >= P) return false;
        else return wants_perm<Q>() || any_wants_perm<Q + 1>();
    }
    static constexpr bool kXor = any_wants_perm<0>() && !kPermFeasible;
    static constexpr int SMEM_WORDS = kXor ? N : padded_words<W>(N); // per polynomial, per buffer
    static constexpr int NBUF = (P >= 3 && !(sizeof(W) == 8 && CNTT_NBUF64 == 1) && !(sizeof(W) == 4 && CNTT_NBUF32 == 1) &&
                                 2 * SMEM_WORDS * (int)sizeof(W) <= CNTT_NBUF2_MAXBYTES) ? 2 : 1; // ping-pong when >1 exchange

    template <int Q> static __device__ __forceinline__ void decomp(int tid, int& blk, int& o)
    {
        constexpr int S = stride<Q>();
        if constexpr (Q == 0) {
            blk = 0; o = tid;
        } else if constexpr (wants_perm<Q>() && !kXor) {
            const int w = tid / PH, l = tid % PH;
            blk = (w / D) * (D * (PH / S)) + (w % D) + D * (l / S);
            o = l % S;
        } else {
            blk = tid / S; o = tid % S;
        }
    }
    // element index of register slot k of thread `tid` in pass q
    template <int Q> static __device__ __forceinline__ int elem(int tid, int k)
    {
        int blk, o;
        decomp<Q>(tid, blk, o);
        return blk * blk_words<Q>() + o + k * stride<Q>();
    }
    template <int Q> static __device__ __forceinline__ unsigned node(int tid, unsigned nu0)
    {
        if constexpr (Q == 0) {
            return nu0;
        } else {
            int blk, o;
            decomp<Q>(tid, blk, o);
            return (nu0 << G::s0(Q)) + (unsigned)blk;
        }
    }
    // shared-memory index of slot k.  Padded mode: base computed once, per-k offsets are compile-time
    // whenever the split (base + kS) >> LOGROW == (base >> LOGROW) + (kS >> LOGROW) is exact (always for
    // Q >= 1; for Q == 0 iff T is a multiple of the row).  XOR mode: the fields (block | slot | offset) of
    // the element index are bit-disjoint, so swz(base + kS) = swz(base) ^ swz(kS).
    static __host__ __device__ constexpr int swz(int i) { return i ^ ((i >> LOGR) & (PH - 1)); }
    template <int Q> static __host__ __device__ constexpr bool split_ok()
    {
        if constexpr (Q == 0) return (T % PH) == 0;
        else return true;
    }
    template <int Q> static __device__ __forceinline__ int sbase(int base_elem)
    {
        if constexpr (kXor) return swz(base_elem);
        else return pad_idx<W>(base_elem);
    }
    template <int Q> static __device__ __forceinline__ int sidx(int base_elem, int base_s, int k)
    {
        constexpr int S = stride<Q>();
        if constexpr (kXor) return base_s ^ swz(k * S);
        else if constexpr (split_ok<Q>()) return base_s + k * S + ((k * S) >> LOGROW);
        else return pad_idx<W>(base_elem + k * S);
    }

    template <int Q, int NP>
    static __device__ __forceinline__ void scatter(W (&x)[NP][R], W* sm, int tid)
    {
        const int be = elem<Q>(tid, 0);
        const int bs = sbase<Q>(be);
#pragma unroll
        for (int np = 0; np < NP; np++)
#pragma unroll
            for (int k = 0; k < R; k++) sm[np * SMEM_WORDS * NBUF + sidx<Q>(be, bs, k)] = x[np][k];
    }
    template <int Q, int NP>
    static __device__ __forceinline__ void gather(W (&x)[NP][R], const W* sm, int tid)
    {
        const int be = elem<Q>(tid, 0);
        const int bs = sbase<Q>(be);
#pragma unroll
        for (int np = 0; np < NP; np++)
#pragma unroll
            for (int k = 0; k < R; k++) x[np][k] = sm[np * SMEM_WORDS * NBUF + sidx<Q>(be, bs, k)];
    }

    // ---- twiddle sources ------------------------------------------------------------------
    // `heap` is the plan's heap-ordered table.  In the last pass every thread owns a different sub-tree, so
    // heap reads are strided by 2^j entries across a warp (ncu: the L1 tag stage was the limiter of the
    // u32 kernels, 4.4x more global-load wavefronts than shared-memory wavefronts).  `last` is the same
    // data re-laid per (level, g) with the thread index innermost, last[((1 << j) - 1 + g) * T + tid],
    // built on the device by k_build_last with this very engine's thread map: one coalesced request per load.
    struct TwSrc {
        const Tw* __restrict__ heap;
        const Tw* __restrict__ last; // this polynomial's (sub-block's) slice; unused when !kLastXp
        const TwHead<Tw>* head;      // kernel-parameter copy of heap[0 .. kHeadEntries); nullptr: not available
    };
    // may pass Q take its twiddles from the constant bank?  (only for whole transforms, nu0 == 1: the caller
    // says so by passing a non-null `head`; the decision per pass is static)
    template <int Q> static __host__ __device__ constexpr bool head_ok()
    {
        if constexpr (Q == P - 1 && P >= 2) return false;                                 // per-thread sub-trees
        else if constexpr (Q == 0) return G::R1 <= head_log<Tw>();                        // indices < 2^R1
        else return stride<Q>() >= CNTT_HEAD_MINS && (G::s0(Q) + LOGR) <= head_log<Tw>(); // (nearly) warp-uniform node
    }
    static constexpr int LAST_WORDS = (R - 1) * T;             // entries of one sub-block's last-pass table
    template <int Q> static __device__ __forceinline__ Tw tw_at(const TwSrc& s, unsigned nu, int tid, int j, int g)
    {
        if constexpr (kLastXp && Q == P - 1) return __ldg(s.last + ((1 << j) - 1 + g) * T + tid);
        else if constexpr (head_ok<Q>()) {
            if (s.head != nullptr) return s.head->e[(nu << j) + g];
            return __ldg(s.heap + (nu << j) + g);
        } else return __ldg(s.heap + (nu << j) + g);
    }
    static __device__ __forceinline__ Tw ldtw(const Tw* __restrict__ tw, unsigned idx) { return __ldg(tw + idx); }

    // ---- register passes ------------------------------------------------------------------
    template <int Q, int NP>
    static __device__ __forceinline__ void fwd_pass(W (&x)[NP][R], const TwSrc& tw, unsigned nu, int tid, const Mod& m)
    {
        constexpr int L = G::levels(Q), J0 = G::jlo(Q);
#pragma unroll
        for (int j = J0; j < J0 + L; j++) {
            const int half = R >> (j + 1);
#pragma unroll
            for (int g = 0; g < (1 << j); g++) {
                const Tw t = tw_at<Q>(tw, nu, tid, j, g);
#pragma unroll
                for (int u = 0; u < half; u++) {
#pragma unroll
                    for (int np = 0; np < NP; np++) A::fwd_bf(x[np][2 * half * g + u], x[np][2 * half * g + u + half], t, m);
                }
            }
        }
    }
    template <int Q, int NP>
    static __device__ __forceinline__ void inv_pass(W (&x)[NP][R], const TwSrc& tw, unsigned nu, int tid, const Mod& m)
    {
        constexpr int L = G::levels(Q), J0 = G::jlo(Q);
#pragma unroll
        for (int j = J0 + L - 1; j >= J0; j--) {
            const int half = R >> (j + 1);
#pragma unroll
            for (int g = 0; g < (1 << j); g++) {
                const Tw t = tw_at<Q>(tw, nu, tid, j, g);
#pragma unroll
                for (int u = 0; u < half; u++) {
#pragma unroll
                    for (int np = 0; np < NP; np++) A::inv_bf(x[np][2 * half * g + u], x[np][2 * half * g + u + half], t, m);
                }
            }
        }
    }

    // ---- whole transforms -------------------------------------------------------------------
    // fwd: x enters in pass-0 layout (slot k <-> element tid + kT), leaves in pass-(P-1) layout
    //      (slot k <-> element tid*R + k when P >= 2), lazy range of the policy (not canonical).
    // `sm` points at this polynomial group's shared memory (NP * NBUF * SMEM_WORDS words).
    // All threads of the group must call (uses gsync()).
    template <int Q, int NP>
    static __device__ __forceinline__ void fwd_from(W (&x)[NP][R], W* sm, const TwSrc& tw, unsigned nu0, int tid, const Mod& m)
    {
        fwd_pass<Q, NP>(x, tw, node<Q>(tid, nu0), tid, m);
        if constexpr (Q + 1 < P) {
            W* buf = sm + ((NBUF == 2 && (Q & 1)) ? SMEM_WORDS : 0);
            scatter<Q, NP>(x, buf, tid);
            gsync();
            gather<Q + 1, NP>(x, buf, tid);
            if constexpr (NBUF == 1 && Q + 2 < P) gsync();
            fwd_from<Q + 1, NP>(x, sm, tw, nu0, tid, m);
        }
    }
    // the rest of a forward transform whose pass 0 the caller has already done (x in pass-0 layout): exchange, passes 1 .. P-1
    template <int NP>
    static __device__ __forceinline__ void fwd_after_pass0(W (&x)[NP][R], W* sm, const TwSrc& tw, unsigned nu0, int tid, const Mod& m)
    {
        static_assert(!kLoopPasses && P >= 2, "compile-time pass chain only");
        scatter<0, NP>(x, sm, tid);
        gsync();
        gather<1, NP>(x, sm, tid);
        if constexpr (NBUF == 1 && 2 < P) gsync();
        fwd_from<1, NP>(x, sm, tw, nu0, tid, m);
    }
    template <int NP>
    static __device__ __forceinline__ void fwd(W (&x)[NP][R], W* sm, const TwSrc& tw, unsigned nu0, int tid, const Mod& m)
    {
        if constexpr (kLoopPasses) fwd_loop<NP>(x, sm, tw.heap, nu0, tid, m);
        else fwd_from<0, NP>(x, sm, tw, nu0, tid, m);
    }

    // inv: x enters in pass-(P-1) layout, leaves in pass-0 layout, lazy range of the policy.
    template <int Q, int NP>
    static __device__ __forceinline__ void inv_from(W (&x)[NP][R], W* sm, const TwSrc& tw, unsigned nu0, int tid, const Mod& m)
    {
        inv_pass<Q, NP>(x, tw, node<Q>(tid, nu0), tid, m);
        if constexpr (Q > 0) {
            W* buf = sm + ((NBUF == 2 && (Q & 1)) ? SMEM_WORDS : 0);
            scatter<Q, NP>(x, buf, tid);
            gsync();
            gather<Q - 1, NP>(x, buf, tid);
            if constexpr (NBUF == 1 && Q >= 2) gsync();
            inv_from<Q - 1, NP>(x, sm, tw, nu0, tid, m);
        }
    }
    // the inverse passes P-1 .. QSTOP only (x leaves in pass-QSTOP layout): the cluster kernel runs pass 0 across CTAs itself
    template <int Q, int QSTOP, int NP>
    static __device__ __forceinline__ void inv_down(W (&x)[NP][R], W* sm, const TwSrc& tw, unsigned nu0, int tid, const Mod& m)
    {
        static_assert(!kLoopPasses, "compile-time pass chain only");
        inv_pass<Q, NP>(x, tw, node<Q>(tid, nu0), tid, m);
        if constexpr (Q > QSTOP) {
            W* buf = sm + ((NBUF == 2 && (Q & 1)) ? SMEM_WORDS : 0);
            scatter<Q, NP>(x, buf, tid);
            gsync();
            gather<Q - 1, NP>(x, buf, tid);
            if constexpr (NBUF == 1) gsync();
            inv_down<Q - 1, QSTOP, NP>(x, sm, tw, nu0, tid, m);
        }
    }
    template <int NP>
    static __device__ __forceinline__ void inv(W (&x)[NP][R], W* sm, const TwSrc& tw, unsigned nu0, int tid, const Mod& m)
    {
        if constexpr (kLoopPasses) inv_loop<NP>(x, sm, tw.heap, nu0, tid, m);
        else inv_from<P - 1, NP>(x, sm, tw, nu0, tid, m);
    }

    // ---- pass loop (64-bit words) ---------------------------------------------------------------
    // A 64-bit butterfly is ~40 SASS instructions, so the fully unrolled transform (R/2 * LOGN
    // butterflies) is ~55 KB of code for N = 2048 and thrashes the 32 KB instruction cache (ncu:
    // stall_no_instruction was the top stall).  The butterfly levels are therefore emitted once and the
    // passes iterate at run time; only the (tiny) per-pass exchanges stay specialised at compile time.
    static constexpr bool kLoopPasses = (sizeof(W) == 8) && (P >= 2);
    static constexpr bool kLastXp = (P >= 2) && !kLoopPasses;  // transposed last-pass twiddle table in use

    template <int Q> static __device__ __forceinline__ unsigned node_rt(int q, int tid, unsigned nu0)
    {
        if constexpr (Q + 1 >= P) return node<Q>(tid, nu0);
        else return q == Q ? node<Q>(tid, nu0) : node_rt<Q + 1>(q, tid, nu0);
    }
    // exchange between pass Q and pass Q+1 (either direction), selected at run time
    template <int Q, int NP, bool FWD>
    static __device__ __forceinline__ void xchg_rt(int q, W (&x)[NP][R], W* sm, int tid)
    {
        if constexpr (Q + 1 < P) {
            if (q == Q) {
                W* buf = sm + ((NBUF == 2 && (Q & 1)) ? SMEM_WORDS : 0);
                if constexpr (FWD) {
                    scatter<Q, NP>(x, buf, tid);
                    gsync();
                    gather<Q + 1, NP>(x, buf, tid);
                } else {
                    scatter<Q + 1, NP>(x, buf, tid);
                    gsync();
                    gather<Q, NP>(x, buf, tid);
                }
                if constexpr (NBUF == 1 && P >= 3) gsync();
            } else {
                xchg_rt<Q + 1, NP, FWD>(q, x, sm, tid);
            }
        }
    }
    template <int NP>
    static __device__ __forceinline__ void levels_fwd(W (&x)[NP][R], const Tw* __restrict__ tw, unsigned nu, int jlo, int jhi, const Mod& m)
    {
#pragma unroll
        for (int j = 0; j < LOGR; j++) {
            if (G::RX == LOGR || (j >= jlo && j < jhi)) {
                const int half = R >> (j + 1);
#pragma unroll
                for (int g = 0; g < (1 << j); g++) {
                    const Tw t = ldtw(tw, (nu << j) + g);
#pragma unroll
                    for (int u = 0; u < half; u++)
#pragma unroll
                        for (int np = 0; np < NP; np++) A::fwd_bf(x[np][2 * half * g + u], x[np][2 * half * g + u + half], t, m);
                }
            }
        }
    }
    template <int NP>
    static __device__ __forceinline__ void levels_inv(W (&x)[NP][R], const Tw* __restrict__ tw, unsigned nu, int jlo, int jhi, const Mod& m)
    {
#pragma unroll
        for (int j = LOGR - 1; j >= 0; j--) {
            if (G::RX == LOGR || (j >= jlo && j < jhi)) {
                const int half = R >> (j + 1);
#pragma unroll
                for (int g = 0; g < (1 << j); g++) {
                    const Tw t = ldtw(tw, (nu << j) + g);
#pragma unroll
                    for (int u = 0; u < half; u++)
#pragma unroll
                        for (int np = 0; np < NP; np++) A::inv_bf(x[np][2 * half * g + u], x[np][2 * half * g + u + half], t, m);
                }
            }
        }
    }
    // ---- pass 0 of a whole Solinas transform: the twiddles of levels 0 .. R1-1 (heap nodes < 2^R1 <= 16) are compile-time powers
    //      of two, so the butterflies are shift butterflies (arith.cuh, A64S::fwd_bf_shift) emitted once, outside the pass loop
    static constexpr bool kShiftHead = ShiftHead<A>::value && kLoopPasses && G::R1 <= 4;
    template <int J, int GI, bool FWD, int NP>
    static __device__ __forceinline__ void shift_group(W (&x)[NP][R])
    {
        constexpr int half = R >> (J + 1);
        constexpr int K = FWD ? shift_exp((1 << J) + GI) : (192 - shift_exp((1 << J) + GI)) % 192;
#pragma unroll
        for (int u = 0; u < half; u++)
#pragma unroll
            for (int np = 0; np < NP; np++) {
                if constexpr (FWD) A::template fwd_bf_shift<K>(x[np][2 * half * GI + u], x[np][2 * half * GI + u + half]);
                else A::template inv_bf_shift<K>(x[np][2 * half * GI + u], x[np][2 * half * GI + u + half]);
            }
    }
    template <int J, bool FWD, int NP, int... GI>
    static __device__ __forceinline__ void shift_level(W (&x)[NP][R], std::integer_sequence<int, GI...>)
    {
        (shift_group<J, GI, FWD, NP>(x), ...);
    }
    template <int NP, int... J>
    static __device__ __forceinline__ void shift_levels_fwd(W (&x)[NP][R], std::integer_sequence<int, J...>)
    {
        (shift_level<J, true, NP>(x, std::make_integer_sequence<int, (1 << J)>{}), ...);
    }
    template <int NP, int... J>
    static __device__ __forceinline__ void shift_levels_inv(W (&x)[NP][R], std::integer_sequence<int, J...>)
    {
        (shift_level<G::R1 - 1 - J, false, NP>(x, std::make_integer_sequence<int, (1 << (G::R1 - 1 - J))>{}), ...);
    }
    template <int NP>
    static __device__ __forceinline__ void fwd_loop(W (&x)[NP][R], W* sm, const Tw* __restrict__ tw, unsigned nu0, int tid, const Mod& m)
    {
        int q0 = 0;
        if constexpr (kShiftHead) {
            if (nu0 == 1u && m.shift_head) { // whole transform on the verified table: uniform branch
                shift_levels_fwd<NP>(x, std::make_integer_sequence<int, G::R1>{});
                xchg_rt<0, NP, true>(0, x, sm, tid);
                q0 = 1;
            }
        }
#pragma unroll 1
        for (int q = q0; q < P; q++) {
            levels_fwd<NP>(x, tw, node_rt<0>(q, tid, nu0), G::jlo(q), G::jlo(q) + G::levels(q), m);
            xchg_rt<0, NP, true>(q, x, sm, tid);
        }
    }
    template <int NP>
    static __device__ __forceinline__ void inv_loop(W (&x)[NP][R], W* sm, const Tw* __restrict__ tw, unsigned nu0, int tid, const Mod& m)
    {
        int qend = 0;
        if constexpr (kShiftHead) {
            if (nu0 == 1u && m.shift_head) qend = 1;
        }
#pragma unroll 1
        for (int q = P - 1; q >= qend; q--) {
            levels_inv<NP>(x, tw, node_rt<0>(q, tid, nu0), G::jlo(q), G::jlo(q) + G::levels(q), m);
            xchg_rt<0, NP, false>(q - 1, x, sm, tid);
        }
        if constexpr (kShiftHead) {
            if (qend == 1) shift_levels_inv<NP>(x, std::make_integer_sequence<int, G::R1>{});
        }
    }

    // element index held in slot k after fwd / expected before inv
    static __device__ __forceinline__ int elem_last(int tid, int k) { return elem<P - 1>(tid, k); }
    static __device__ __forceinline__ int elem_first(int tid, int k) { return elem<0>(tid, k); }
};

} // namespace cntt
