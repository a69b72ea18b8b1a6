// native_device.cuh -- per-coefficient device code of the native plans: word -> residue (lazy) and the
// Garner reconstruction on 32-bit mixed-radix digits (see native.hpp for why this equals the reference's
// reconstruct_* functions, including the sign rule: src/native32.rs:39, native64.rs:125,
// native128.rs:105, native_binary32.rs:27, native_binary64.rs:42, native_binary128.rs:47).
//
// The native kernels are bound by the FMA-heavy pipe (IMAD 1 slot, IMAD.HI 2, IMAD.WIDE 2.6 -- measured,
// profiles/r01_ubench_int_pipes.txt), so everything here is 32-bit Shoup arithmetic: no 64-bit multiplies.
#pragma once
#include "native.hpp"

namespace cntt {
namespace dev {

typedef unsigned __int128 u128;

// a * c mod p in [0, 2p) for any 32-bit a; cc = {c, floor(c 2^32 / p)}
__device__ __forceinline__ uint32_t mulc(uint32_t a, uint2 cc, uint32_t p)
{
    const uint32_t q = __umulhi(a, cc.y);
    return a * cc.x - q * p;
}
__device__ __forceinline__ uint32_t red2p(uint32_t x, uint32_t p) { return umin32(x, x - 2u * p); } // [0,4p) -> [0,2p)
__device__ __forceinline__ uint32_t red1p(uint32_t x, uint32_t p) { return umin32(x, x - p); }      // [0,2p) -> [0,p)
// any 32-bit w -> [0, 2p):  w - (w >> 30) p   (2^30 - p < 2^24 for every built-in prime)
__device__ __forceinline__ uint32_t fold32(uint32_t w, uint32_t p) { return w - (w >> 30) * p; }

// word -> residue mod P[k] in [0, 4p).  Limb j (weight 2^(32 j)) is multiplied by scale[j] (Shoup pair);
// limb 0 is multiplied by scale[0] when SCALE0, else folded.  With scale = red[k] this is plain reduction;
// the fused polymul passes scale = red * (2^32 / N) for the lhs operand (see native_fused.cuh).
template <int NLIMBS, bool SCALE0>
__device__ __forceinline__ uint32_t residue(uint64_t lo, uint64_t hi, const uint2* scale, uint32_t p)
{
    uint32_t x = SCALE0 ? mulc((uint32_t)lo, scale[0], p) : fold32((uint32_t)lo, p); // [0,2p)
    if constexpr (NLIMBS >= 2) x += mulc((uint32_t)(lo >> 32), scale[1], p);           // [0,4p)
    if constexpr (NLIMBS >= 4) {
        x = red2p(x, p) + mulc((uint32_t)hi, scale[2], p);
        x = red2p(x, p) + mulc((uint32_t)(hi >> 32), scale[3], p);
    }
    return x;
}

// Montgomery product a b 2^-32 mod p in (0, 2p) for a b < 2^32 p
__device__ __forceinline__ uint32_t mont(uint32_t a, uint32_t b, uint32_t p, uint32_t pinv)
{
    const uint64_t t = (uint64_t)a * b;
    const uint32_t m = (uint32_t)t * pinv;
    return (uint32_t)(t >> 32) - __umulhi(m, p) + p;
}

template <int KIND> struct KindInfo;
template <> struct KindInfo<NK_NATIVE32> { static constexpr int NP = 3, LIMBS = 1; typedef uint32_t Word; };
template <> struct KindInfo<NK_NATIVE64> { static constexpr int NP = 5, LIMBS = 2; typedef uint64_t Word; };
template <> struct KindInfo<NK_NATIVE128> { static constexpr int NP = 10, LIMBS = 4; typedef u128 Word; };
template <> struct KindInfo<NK_BINARY32> { static constexpr int NP = 2, LIMBS = 1; typedef uint32_t Word; };
template <> struct KindInfo<NK_BINARY64> { static constexpr int NP = 3, LIMBS = 2; typedef uint64_t Word; };
template <> struct KindInfo<NK_BINARY128> { static constexpr int NP = 5, LIMBS = 4; typedef u128 Word; };

template <int KIND>
__device__ __forceinline__ void load_word(const void* p, size_t i, uint64_t& lo, uint64_t& hi)
{
    typedef typename KindInfo<KIND>::Word Word;
    if constexpr (sizeof(Word) == 4) { lo = reinterpret_cast<const uint32_t*>(p)[i]; hi = 0; }
    else if constexpr (sizeof(Word) == 8) { lo = reinterpret_cast<const uint64_t*>(p)[i]; hi = 0; }
    else {
        const uint4 q = reinterpret_cast<const uint4*>(p)[i];
        lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
        hi = (uint64_t)q.z | ((uint64_t)q.w << 32);
    }
}
template <int KIND>
__device__ __forceinline__ void store_word(void* value, unsigned long long i, typename KindInfo<KIND>::Word w)
{
    typedef typename KindInfo<KIND>::Word Word;
    if constexpr (sizeof(Word) == 16) {
        const uint64_t lo = (uint64_t)w, hi = (uint64_t)(w >> 64);
        reinterpret_cast<uint4*>(value)[i] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    } else {
        reinterpret_cast<Word*>(value)[i] = w;
    }
}

// canonical residues r[0..NP) -> centred lift wrapped to the word
template <int KIND>
__device__ __forceinline__ typename KindInfo<KIND>::Word reconstruct(const uint32_t* r, const NativeConsts& c)
{
    typedef typename KindInfo<KIND>::Word Word;
    constexpr int NP = KindInfo<KIND>::NP;
    uint32_t d[NP];
    d[0] = r[0];
#pragma unroll
    for (int k = 1; k < NP; k++) {
        const uint32_t p = c.P[k];
        uint32_t x = r[k];
#pragma unroll
        for (int j = 0; j < k; j++) x = mulc(x - d[j] + p, c.ginv[j][k], p); // d[j] < P[j] < P[k]; x < 2p
        d[k] = red1p(x, p);
    }
    bool neg;
    if constexpr (NP <= 3) {
        neg = d[NP - 1] > c.half_single[NP - 1];
    } else {
        neg = d[NP - 1] > c.half_pair_hi[NP - 1] || (d[NP - 1] == c.half_pair_hi[NP - 1] && d[NP - 2] > c.half_pair_lo[NP - 1]);
    }
    if constexpr (sizeof(Word) == 4) {
        uint32_t pos = d[0];
#pragma unroll
        for (int j = 1; j < NP; j++) pos += d[j] * (uint32_t)c.gm[j][0];
        return neg ? pos - (uint32_t)c.gm[NP][0] : pos;
    } else if constexpr (sizeof(Word) == 8) {
        uint64_t pos = d[0];
#pragma unroll
        for (int j = 1; j < NP; j++) pos += (uint64_t)d[j] * c.gm[j][0];
        return neg ? pos - c.gm[NP][0] : pos;
    } else {
        u128 pos = d[0];
#pragma unroll
        for (int j = 1; j < NP; j++) pos += (u128)d[j] * (((u128)c.gm[j][1] << 64) | c.gm[j][0]);
        return neg ? pos - (((u128)c.gm[NP][1] << 64) | c.gm[NP][0]) : pos;
    }
}

// Reconstruction for the fused polymul kernels.  There the integer behind the residues is a negacyclic product
// coefficient, |x| < n 2^(2w) (n 2^w for binary plans), at most 2^-5.9 of M = prod P_k for every plan the library
// builds -- so the quotient of the classical CRT sum is known from a float estimate and no mixed-radix chain is
// needed.  With y_k = x (M/P_k)^-1 mod P_k (the factor is folded into the lhs scale constants, so y_k is what
// the inverse NTT already delivers):   x = sum_k y_k (M/P_k) - q M,   q = round(sum_k y_k / P_k),
// because sum_k y_k / P_k = x / M + q exactly and |x / M| <= 2^-5.9 while the float sum is off by < 2^-19.
// Only the low word of x is wanted, so the sum runs modulo 2^w.  The result is the centred lift wrapped to the
// word -- the value the reference's reconstruct_* returns (src/native64.rs:90-141 etc.) whenever the bound holds,
// i.e. for every negacyclic_polymul input.  (Plan32::inv on caller-supplied residues keeps the exact Garner
// chain above: there the bound is the caller's business and the reference's sign rule must be reproduced.)
template <int KIND, int NP = KindInfo<KIND>::NP>
__device__ __forceinline__ typename KindInfo<KIND>::Word reconstruct_bounded(const uint32_t* y, const NativeConsts& c)
{
    typedef typename KindInfo<KIND>::Word Word;
    constexpr int CLS = native_np_class(NP);
    float f = 0.0f;
#pragma unroll
    for (int k = 0; k < NP; k++) f = fmaf(__uint2float_rn(y[k]), c.ainv[k], f);
    const uint32_t q = (uint32_t)__float2int_rn(f);
    if constexpr (sizeof(Word) == 4) {
        uint32_t x = 0u - q * (uint32_t)c.aM[CLS][0];
#pragma unroll
        for (int k = 0; k < NP; k++) x += y[k] * (uint32_t)c.am[CLS][k][0];
        return x;
    } else if constexpr (sizeof(Word) == 8) {
        uint64_t x = 0ull - (uint64_t)q * c.aM[CLS][0];
#pragma unroll
        for (int k = 0; k < NP; k++) x += (uint64_t)y[k] * c.am[CLS][k][0];
        return x;
    } else {
        u128 x = (u128)0 - (u128)q * (((u128)c.aM[CLS][1] << 64) | c.aM[CLS][0]);
#pragma unroll
        for (int k = 0; k < NP; k++) x += (u128)y[k] * (((u128)c.am[CLS][k][1] << 64) | c.am[CLS][k][0]);
        return x;
    }
}

} // namespace dev
} // namespace cntt
