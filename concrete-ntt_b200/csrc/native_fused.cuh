// native_fused.cuh -- fused negacyclic polymul of the native / native_binary plans (N <= 4096).
//
// One thread group owns one polymul.  lhs and rhs are read from HBM exactly once into registers;
// for every prime the group reduces both operands, runs the two forward NTTs together (shared
// twiddle loads), multiplies pointwise (x N^-1), runs the inverse NTT, and parks the residue
// polynomial in shared memory; after the last prime each thread lifts its own coefficients with the
// Garner reconstruction and writes the product to HBM exactly once.  Residues never touch HBM
// (the reference round-trips 2 x nprimes scratch vectors per call, src/native64.rs:1047-1068).
#pragma once
#include "native_device.cuh"

namespace cntt {

// Per-kind configuration:
//   RELOAD: the operands are re-read from global memory (L2 hits after the first prime) for every prime instead of
//           living in registers across the prime loop (128-bit words: 64 registers per thread)
//   ACC:    the reconstruction sum of reconstruct_bounded is accumulated per prime in registers (word + float per
//           coefficient) instead of parking np residue polynomials in shared memory -- for the 128-bit kinds that
//           is 160 KB at N = 4096, which pinned the kernel at one CTA per SM
// The two go together (fused_acc_reload): without the stash a CTA needs 34 KB of shared memory instead of 75 KB, without the
// operand registers it fits 128 registers, and the kernel is then compiled for 512 resident threads (four CTAs of 128 threads
// per SM instead of three).  Measured per kind on B200 (profiles/r02_experiments.txt, "acc occupancy"), M polymul/s at N = 2048:
//   128-bit kinds (N <= 2048)   native128 5.43 -> 5.64 (N = 1024: 12.1 -> 13.2), binary128 11.1 -> 11.9              -> on
//   native32                    27.3 -> 28.6 (N = 1024: 58.2 -> 56.5)                                                  -> on at N = 2048
//   pre-transformed rhs         native32 35.0 -> 38.6, native64 18.5 -> 20.2, binary32 47.3 -> 55.4, binary64 32.7 -> 32.3
//                               (N = 1024 loses 2 %)                                                                   -> on at N = 2048 but binary64
//   native64, binary32/64       14.9 -> 14.3, 43.0 -> 41.7, 26.9 -> 25.9 (the accumulators spill)                      -> off
// CNTT_FUSED_RELOAD_MASK / CNTT_FUSED_ACC_MASK (bit k = NativeKind k) override the table for A/B runs (tools/build_variant.sh).
#ifndef CNTT_FUSED_RELOAD_MASK
#define CNTT_FUSED_RELOAD_MASK (-1)
#endif
#ifndef CNTT_FUSED_ACC_MASK
#define CNTT_FUSED_ACC_MASK (-1)
#endif
#ifndef CNTT_FUSED_SEQ_MASK
#define CNTT_FUSED_SEQ_MASK 0
#endif
#ifndef CNTT_FUSED_MINTHREADS
#define CNTT_FUSED_MINTHREADS 0 // resident threads per SM the kernel is compiled for (register cap 65536 / this); 0: 512 with ACC + RELOAD, else 768
#endif
constexpr bool fused_acc_reload(int kind, int logn, bool pre)
{
    if (kind == NK_NATIVE128 || kind == NK_BINARY128) return logn <= 11;
    if (logn != 11) return false;
    return kind == NK_NATIVE32 || (pre && (kind == NK_NATIVE64 || kind == NK_BINARY32));
}
constexpr bool fused_reload(int kind, int logn, bool pre) { return CNTT_FUSED_RELOAD_MASK >= 0 ? (((CNTT_FUSED_RELOAD_MASK >> kind) & 1) != 0 && logn <= 11) : fused_acc_reload(kind, logn, pre); }
constexpr bool fused_acc(int kind, int logn, bool pre) { return CNTT_FUSED_ACC_MASK >= 0 ? (((CNTT_FUSED_ACC_MASK >> kind) & 1) != 0 && logn <= 11) : fused_acc_reload(kind, logn, pre); }
constexpr int fused_minthreads(int kind, int logn, bool pre)
{
    return CNTT_FUSED_MINTHREADS != 0 ? CNTT_FUSED_MINTHREADS : (fused_reload(kind, logn, pre) && fused_acc(kind, logn, pre)) ? 512 : 768;
}
constexpr int kFusedMinLogN = 5, kFusedMaxLogN = 12;

// Binary plans: rhs in {0,1}^N (src/native_binary64.rs:423-444; anything else is unspecified there).  The first register pass of
// its forward transform is then a LINEAR map of bits with plan-time coefficients -- the same for every thread, because pass 0
// uses the root sub-tree -- so it is evaluated from a table instead of with butterflies: the 2^R1 slots of a set are cut into
// nibbles, tab[q][v][j] = sum over the bits of v of (coefficient of input 4q+bit in output j), and output j is the sum of one row
// per nibble (at most four canonical values: inside the lazy range [0,4p)).  N = 2048: 24 butterflies per thread become 8 vector
// loads and 16 additions.  Only bit 0 of an rhs word is read.  B200 (profiles/r02_experiments.txt): binary64 N=2048 26.2 -> 26.8,
// binary32 40.8 -> 43.0, binary64 N=256 262 -> 275 M polymul/s; with four nibbles (R1 = 4, N = 4096) it loses 8 %, so R1 <= 3 only.
#ifndef CNTT_BINARY_LUT
#define CNTT_BINARY_LUT 1
#endif
template <class E>
__device__ __forceinline__ void binary_pass0(uint32_t (&x)[E::R], const uint32_t* __restrict__ tab)
{
    constexpr int R1 = E::G::R1, S = 1 << R1, NSETS = E::R >> R1, QB = S < 4 ? S : 4, NQ = S / QB;
#pragma unroll
    for (int s = 0; s < NSETS; s++) {
        uint32_t y[S];
#pragma unroll
        for (int j = 0; j < S; j++) y[j] = 0;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            uint32_t v = 0;
#pragma unroll
            for (int i = 0; i < QB; i++) v |= (x[s + (QB * q + i) * NSETS] & 1u) << i;
            const uint32_t* row = tab + (q * 16 + v) * S;
            if constexpr (S >= 4) {
#pragma unroll
                for (int j = 0; j < S; j += 4) {
                    const uint4 t = __ldg(reinterpret_cast<const uint4*>(row + j));
                    y[j] += t.x; y[j + 1] += t.y; y[j + 2] += t.z; y[j + 3] += t.w;
                }
            } else {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(row));
                y[0] += t.x; y[1] += t.y;
            }
        }
#pragma unroll
        for (int j = 0; j < S; j++) x[s + j * NSETS] = y[j];
    }
}

struct FusedParams {
    // PRE (pre-transformed rhs): `rhs` points at residue planes in the NTT domain, as written by cntt_native_fwd / fwd_binary; plane k of
    // key b at rhs + k * pre_plane_stride + b * pre_poly_stride words (pre_poly_stride = 0: one key for the whole batch)
    unsigned long long pre_plane_stride, pre_poly_stride;
    const uint32_t* bin0[10];
    const uint2* tw_fwd[10];
    const uint2* tw_inv[10];
    const uint2* tw_fwd_last[10];
    const uint2* tw_inv_last[10];
    Mod32 mod[10];
    uint2 lscale[10][4];
};

template <int KIND, int LOGN, int LOGR, bool PRE = false>
struct FusedCfg {
    typedef Engine<A32L4, LOGN, LOGR> E;
    static constexpr int NP = native_fused_np(KIND, dev::KindInfo<KIND>::NP); // native128: nine of the ten primes (native.hpp)
    static constexpr int T = E::T;
    static constexpr int GP = T >= 128 ? 1 : 128 / T;
    static constexpr bool ACC = fused_acc(KIND, LOGN, PRE);
    // SEQ (experiment, CNTT_FUSED_SEQ_MASK): the two forward transforms of a prime run one after the other through ONE set of exchange
    // buffers, the last prime's residues stay in registers instead of the stash, the operands are re-read per prime, and the kernel is
    // compiled for 512 resident threads: 50 KB instead of 75 KB of shared memory at N = 2048 with five primes, four CTAs per SM
    static constexpr bool SEQ = ((CNTT_FUSED_SEQ_MASK >> KIND) & 1) != 0 && LOGN == 11 && !PRE && !ACC && KIND < NK_BINARY32;
    static constexpr int XCHG_WORDS = ((PRE || SEQ) ? 1 : 2) * E::NBUF * E::SMEM_WORDS;   // polynomials in flight: lhs and rhs, or one at a time
    static constexpr bool RELOAD = fused_reload(KIND, LOGN, PRE) || SEQ;
    static constexpr int MINTHREADS = SEQ ? 512 : fused_minthreads(KIND, LOGN, PRE);
    static constexpr int STASH_WORDS = ACC ? 0 : (SEQ ? NP - 1 : NP) * E::N;
    static constexpr size_t SMEM_BYTES = (size_t)GP * (XCHG_WORDS + STASH_WORDS) * sizeof(uint32_t);
    static constexpr int BLK_BY_THREADS = MINTHREADS > GP * T ? MINTHREADS / (GP * T) : 1;
    static constexpr int BLK_BY_SMEM = (int)((size_t)227 * 1024 / (SMEM_BYTES + 1024)) > 0 ? (int)((size_t)227 * 1024 / (SMEM_BYTES + 1024)) : 1;
    static constexpr int MINBLK = BLK_BY_THREADS < BLK_BY_SMEM ? BLK_BY_THREADS : BLK_BY_SMEM; // no point capping registers below what shared memory admits
};

// Per prime p_k (all arithmetic 32-bit, lazy ranges of policy A32L4):
//   lhs residue * (2^32 / N)   [0,4p)    Shoup multiplies by per-limb constants
//   rhs residue                [0,4p)    (binary plans: the low 32 bits, src/native_binary64.rs:372-389)
//   two forward NTTs, twiddle loads shared
//   pointwise Montgomery product  A B 2^-32 = a b / N   in (0,2p)    (replaces mul_assign_normalize)
//   inverse NTT, canonical residue parked in shared memory
//
// PRE = the rhs operand arrives already transformed (the TFHE shape: the key stays in the NTT domain, src/prime32.rs:905-927 is what a
// caller of the reference does per prime): only the lhs is reduced and transformed, the rhs residues are read in the last-pass layout
// (R consecutive words per thread), and two of the three transforms per prime remain.
template <int KIND, int LOGN, int LOGR, bool PRE = false>
__global__ void __launch_bounds__(FusedCfg<KIND, LOGN, LOGR, PRE>::GP * FusedCfg<KIND, LOGN, LOGR, PRE>::T, FusedCfg<KIND, LOGN, LOGR, PRE>::MINBLK)
k_polymul_fused(const NativeConsts c, const FusedParams fp, void* __restrict__ prod, const void* __restrict__ lhs,
                const void* __restrict__ rhs, unsigned long long batch)
{
    typedef FusedCfg<KIND, LOGN, LOGR, PRE> Cfg;
    typedef typename Cfg::E E;
    typedef typename dev::KindInfo<KIND>::Word Word;
    constexpr int T = Cfg::T, R = E::R, N = E::N, NP = Cfg::NP, GP = Cfg::GP;
    constexpr int LIMBS = dev::KindInfo<KIND>::LIMBS;
    constexpr bool BINARY = KIND >= NK_BINARY32;
    constexpr int WB = (int)sizeof(Word);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* sm_all = reinterpret_cast<uint32_t*>(smem_raw);

    const int grp = (GP == 1) ? 0 : (int)(threadIdx.x / T);
    const int tid = (GP == 1) ? (int)threadIdx.x : (int)(threadIdx.x % T);
    unsigned long long b = (unsigned long long)blockIdx.x * GP + grp;
    const bool active = b < batch;
    if (!active) b = batch - 1;
    uint32_t* sm = sm_all + (size_t)grp * (Cfg::XCHG_WORDS + Cfg::STASH_WORDS);
    uint32_t* stash = sm + Cfg::XCHG_WORDS;
    const size_t base = (size_t)b * N;

    // operands: read once and kept in registers for all primes, or (RELOAD) re-read per prime
    constexpr bool RELOAD = Cfg::RELOAD, ACC = Cfg::ACC;
    constexpr int CLS = native_np_class(NP);
    constexpr int RK = RELOAD ? 1 : R;
    uint64_t llo[RK], rlo[PRE ? 1 : RK];
    uint64_t lhi[(WB == 16 && !RELOAD) ? R : 1], rhi[(WB == 16 && !RELOAD && !PRE) ? R : 1];
    if constexpr (!RELOAD) {
#pragma unroll
        for (int k = 0; k < R; k++) {
            uint64_t h0, h1;
            dev::load_word<KIND>(lhs, base + tid + k * T, llo[k], h0);
            if constexpr (WB == 16) lhi[k] = h0;
            if constexpr (!PRE) {
                dev::load_word<KIND>(rhs, base + tid + k * T, rlo[k], h1);
                if constexpr (WB == 16) rhi[k] = h1;
            }
        }
    }
    Word acc[ACC ? R : 1];
    float accf[ACC ? R : 1];
    if constexpr (ACC) {
#pragma unroll
        for (int k = 0; k < R; k++) { acc[k] = 0; accf[k] = 0.0f; }
    }

    uint32_t y[1][R]; // one prime's product residues; after the loop: the last prime's (SEQ reads them from here)
#pragma unroll 1
    for (int pk = 0; pk < NP; pk++) {
        const Mod32 m = fp.mod[pk];
        const uint32_t p = m.p;
        uint32_t x[2][R];
        if constexpr (RELOAD) {
#pragma unroll
            for (int k = 0; k < R; k++) {
                uint64_t alo, ahi, blo, bhi;
                dev::load_word<KIND>(lhs, base + tid + k * T, alo, ahi);
                x[0][k] = dev::residue<LIMBS, true>(alo, ahi, fp.lscale[pk], p);
                if constexpr (!PRE) {
                    dev::load_word<KIND>(rhs, base + tid + k * T, blo, bhi);
                    x[1][k] = BINARY ? (uint32_t)blo : dev::residue<LIMBS, false>(blo, bhi, c.red[pk], p);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; k++) {
                x[0][k] = dev::residue<LIMBS, true>(llo[k], WB == 16 ? lhi[k] : 0ull, fp.lscale[pk], p);
                if constexpr (!PRE) x[1][k] = BINARY ? (uint32_t)rlo[k] : dev::residue<LIMBS, false>(rlo[k], WB == 16 ? rhi[k] : 0ull, c.red[pk], p);
            }
        }
        if constexpr (PRE) {
            uint32_t xl[1][R];
#pragma unroll
            for (int k = 0; k < R; k++) xl[0][k] = x[0][k];
            E::template fwd<1>(xl, sm, typename E::TwSrc{fp.tw_fwd[pk], fp.tw_fwd_last[pk]}, 1u, tid, m);
            const uint32_t* rp = reinterpret_cast<const uint32_t*>(rhs) + (size_t)pk * fp.pre_plane_stride + (size_t)b * fp.pre_poly_stride + E::elem_last(tid, 0);
            load_contig<uint32_t, R>(rp, x[1]);   // canonical residues, the lane's R consecutive words of the last-pass layout
#pragma unroll
            for (int k = 0; k < R; k++) x[0][k] = xl[0][k];
        } else if constexpr (BINARY && CNTT_BINARY_LUT != 0 && E::P >= 2 && !E::kLoopPasses && E::G::R1 <= 3) {
            const typename E::TwSrc tws = {fp.tw_fwd[pk], fp.tw_fwd_last[pk]};
            uint32_t xl[1][R];
#pragma unroll
            for (int k = 0; k < R; k++) xl[0][k] = x[0][k];
            E::template fwd_pass<0, 1>(xl, tws, 1u, tid, m);     // lhs: pass 0 with butterflies
            binary_pass0<E>(x[1], fp.bin0[pk]);                   // rhs: pass 0 from the table
#pragma unroll
            for (int k = 0; k < R; k++) x[0][k] = xl[0][k];
            E::template fwd_after_pass0<2>(x, sm, tws, 1u, tid, m);
        } else if constexpr (Cfg::SEQ) {
            static_assert(!Cfg::SEQ || (E::P == 3 && E::NBUF == 2), "back-to-back transforms without a barrier in between: two exchanges on two buffers");
            const typename E::TwSrc tws = {fp.tw_fwd[pk], fp.tw_fwd_last[pk]};
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t xl[1][R];
#pragma unroll
                for (int k = 0; k < R; k++) xl[0][k] = x[h][k];
                E::template fwd<1>(xl, sm, tws, 1u, tid, m);
#pragma unroll
                for (int k = 0; k < R; k++) x[h][k] = xl[0][k];
            }
        } else {
            E::template fwd<2>(x, sm, typename E::TwSrc{fp.tw_fwd[pk], fp.tw_fwd_last[pk]}, 1u, tid, m);
        }
        const uint32_t pinv = c.pinv[pk];
#pragma unroll
        for (int k = 0; k < R; k++) y[0][k] = dev::mont(dev::red2p(x[0][k], p), dev::red2p(x[1][k], p), p, pinv);
        if constexpr (E::NBUF == 1 && E::P >= 2) __syncthreads(); // single exchange buffer: fwd gather vs inv scatter
        E::template inv<1>(y, sm, typename E::TwSrc{fp.tw_inv[pk], fp.tw_inv_last[pk]}, 1u, tid, m);
        if constexpr (ACC) {
            // reconstruct_bounded, one prime at a time: acc += y (M/P_k) mod 2^w,  accf += y / P_k
            const float ip = c.ainv[pk];
            Word mk;
            if constexpr (WB == 16) mk = ((dev::u128)c.am[CLS][pk][1] << 64) | c.am[CLS][pk][0];
            else mk = (Word)c.am[CLS][pk][0];
#pragma unroll
            for (int k = 0; k < R; k++) {
                const uint32_t yk = A32L4::canon_inv(y[0][k], m);
                acc[k] += (Word)yk * mk;
                accf[k] = fmaf(__uint2float_rn(yk), ip, accf[k]);
            }
        } else if (!Cfg::SEQ || pk < NP - 1) {
#pragma unroll
            for (int k = 0; k < R; k++) stash[pk * N + tid + k * T] = A32L4::canon_inv(y[0][k], m);
        } else {
#pragma unroll
            for (int k = 0; k < R; k++) y[0][k] = A32L4::canon_inv(y[0][k], m);
        }
        if constexpr (E::NBUF == 1 && E::P >= 2) __syncthreads(); // inv gather vs next prime's fwd scatter
    }

    if (active) {
        if constexpr (ACC) {
            Word bigm;
            if constexpr (WB == 16) bigm = ((dev::u128)c.aM[CLS][1] << 64) | c.aM[CLS][0];
            else bigm = (Word)c.aM[CLS][0];
#pragma unroll
            for (int k = 0; k < R; k++) {
                const uint32_t q = (uint32_t)__float2int_rn(accf[k]);
                dev::store_word<KIND>(prod, base + tid + k * T, (Word)(acc[k] - (Word)q * bigm));
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; k++) {
                uint32_t r[NP];
#pragma unroll
                for (int pk = 0; pk < NP; pk++) r[pk] = (Cfg::SEQ && pk == NP - 1) ? y[0][k] : stash[pk * N + tid + k * T];
                dev::store_word<KIND>(prod, base + tid + k * T, dev::reconstruct_bounded<KIND, NP>(r, c));
            }
        }
    }
}

template <int KIND, int LOGN, bool PRE = false>
static cudaError_t launch_fused_one(const NativePlanDev& pl, void* prod, const void* lhs, const void* rhs, size_t batch, cudaStream_t st,
                                    size_t pre_plane_stride = 0, size_t pre_poly_stride = 0)
{
    constexpr int LOGR = native_fused_logr(KIND, LOGN);
    typedef FusedCfg<KIND, LOGN, LOGR, PRE> Cfg;
    FusedParams fp;
    fp.pre_plane_stride = pre_plane_stride;
    fp.pre_poly_stride = pre_poly_stride;
    for (int k = 0; k < Cfg::NP; k++) {
        fp.bin0[k] = pl.bin0[k];
        if (KIND >= NK_BINARY32 && CNTT_BINARY_LUT != 0 && Cfg::E::P >= 2 && Cfg::E::G::R1 <= 3 && !fp.bin0[k]) return cudaErrorInvalidValue;
        fp.tw_fwd[k] = pl.sub[k].tw_fwd;
        fp.tw_inv[k] = pl.sub[k].tw_inv;
        fp.tw_fwd_last[k] = pl.fused_fwd_last[k];
        fp.tw_inv_last[k] = pl.fused_inv_last[k];
        if (Cfg::E::kLastXp && (!fp.tw_fwd_last[k] || !fp.tw_inv_last[k])) return cudaErrorInvalidValue;
        fp.mod[k] = pl.sub[k].mod;
        for (int j = 0; j < 4; j++) fp.lscale[k][j] = pl.lscale[k][j];
    }
    auto kern = k_polymul_fused<KIND, LOGN, LOGR, PRE>;
    if (Cfg::SMEM_BYTES > 227 * 1024) return cudaErrorNotSupported;
    if (cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES); e != cudaSuccess) return e;
    const unsigned long long nblk = (batch + Cfg::GP - 1) / Cfg::GP;
    if (nblk > 0x7fffffffull) return cudaErrorInvalidValue;
    kern<<<(unsigned)nblk, Cfg::GP * Cfg::T, Cfg::SMEM_BYTES, st>>>(native_consts(pl.prime_set), fp, prod, lhs, rhs, batch);
    return cudaGetLastError();
}

template <int KIND>
static cudaError_t launch_fused_kind(const NativePlanDev& pl, void* prod, const void* lhs, const void* rhs, size_t batch, cudaStream_t st)
{
    switch (pl.logn) {
    case 5: return launch_fused_one<KIND, 5>(pl, prod, lhs, rhs, batch, st);
    case 6: return launch_fused_one<KIND, 6>(pl, prod, lhs, rhs, batch, st);
    case 7: return launch_fused_one<KIND, 7>(pl, prod, lhs, rhs, batch, st);
    case 8: return launch_fused_one<KIND, 8>(pl, prod, lhs, rhs, batch, st);
    case 9: return launch_fused_one<KIND, 9>(pl, prod, lhs, rhs, batch, st);
    case 10: return launch_fused_one<KIND, 10>(pl, prod, lhs, rhs, batch, st);
    case 11: return launch_fused_one<KIND, 11>(pl, prod, lhs, rhs, batch, st);
    case 12: return launch_fused_one<KIND, 12>(pl, prod, lhs, rhs, batch, st);
    default: return cudaErrorNotSupported; // larger N: unfused two-level pipeline (capi.cu)
    }
}
// pre-transformed rhs (k_polymul_fused<.., PRE = true>): 256 <= N <= 4096
template <int KIND>
static cudaError_t launch_fused_pre_kind(const NativePlanDev& pl, void* prod, const void* lhs, const uint32_t* rhs_planes, size_t batch, size_t plane_stride,
                                         size_t poly_stride, cudaStream_t st)
{
    switch (pl.logn) {
    case 8: return launch_fused_one<KIND, 8, true>(pl, prod, lhs, rhs_planes, batch, st, plane_stride, poly_stride);
    case 9: return launch_fused_one<KIND, 9, true>(pl, prod, lhs, rhs_planes, batch, st, plane_stride, poly_stride);
    case 10: return launch_fused_one<KIND, 10, true>(pl, prod, lhs, rhs_planes, batch, st, plane_stride, poly_stride);
    case 11: return launch_fused_one<KIND, 11, true>(pl, prod, lhs, rhs_planes, batch, st, plane_stride, poly_stride);
    case 12: return launch_fused_one<KIND, 12, true>(pl, prod, lhs, rhs_planes, batch, st, plane_stride, poly_stride);
    default: return cudaErrorNotSupported;
    }
}

template <int KIND, int LOGN>
static cudaError_t fused_build_last_one(const uint2* heap, uint2* out, cudaStream_t st)
{
    return launch_build_last_e<Engine<A32L4, LOGN, native_fused_logr(KIND, LOGN)>>(heap, out, 0, st);
}

} // namespace cntt
