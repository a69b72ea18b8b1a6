// native_device.cuh -- per-coefficient device code of the native plans: word % P_k and the six Garner
// reconstructions.  The reconstructions evaluate the reference's formulas term by term so that even
// arbitrary (not polymul-generated) residues lift to the same word, including the sign rule
// (src/native32.rs:39, native64.rs:125, native128.rs:105, native_binary32.rs:27,
// native_binary64.rs:42, native_binary128.rs:47).
#pragma once
#include "native.hpp"

namespace cntt {
namespace dev {

typedef unsigned __int128 u128;

// v mod p with M = floor(2^64 / p): the quotient estimate is off by at most one.
__device__ __forceinline__ uint32_t mod_u64(uint64_t v, uint32_t p, uint64_t M)
{
    const uint64_t q = __umul64hi(v, M);
    uint64_t r = v - q * p;
    if (r >= p) r -= p;
    return (uint32_t)r;
}
__device__ __forceinline__ uint32_t mod_u128(uint64_t lo, uint64_t hi, const NativeConsts& c, int k)
{
    const uint32_t h = mod_u64(hi, c.P[k], c.barrett[k]);
    const uint32_t l = mod_u64(lo, c.P[k], c.barrett[k]);
    return mod_u64((uint64_t)h * c.c64[k] + l, c.P[k], c.barrett[k]);
}
// native32::mul_mod32 (src/native32.rs:21-25): (a * b) % P[k]
__device__ __forceinline__ uint32_t mul_mod32(const NativeConsts& c, int k, uint32_t a, uint32_t b)
{
    return mod_u64((uint64_t)a * b, c.P[k], c.barrett[k]);
}
// native64::mul_mod64 (src/native64.rs:36-41)
__device__ __forceinline__ uint64_t mul_mod64(uint64_t p_neg, uint64_t a, uint64_t b, uint64_t b_shoup)
{
    const uint64_t q = __umul64hi(a, b_shoup);
    const uint64_t r = a * b + p_neg * q;
    const uint64_t r2 = r + p_neg;
    return r < r2 ? r : r2;
}
__device__ __forceinline__ u128 mk128(const uint64_t w[2]) { return ((u128)w[1] << 64) | w[0]; }

template <int KIND> struct KindInfo;
template <> struct KindInfo<NK_NATIVE32> { static constexpr int NP = 3; typedef uint32_t Word; };
template <> struct KindInfo<NK_NATIVE64> { static constexpr int NP = 5; typedef uint64_t Word; };
template <> struct KindInfo<NK_NATIVE128> { static constexpr int NP = 10; typedef u128 Word; };
template <> struct KindInfo<NK_BINARY32> { static constexpr int NP = 2; typedef uint32_t Word; };
template <> struct KindInfo<NK_BINARY64> { static constexpr int NP = 3; typedef uint64_t Word; };
template <> struct KindInfo<NK_BINARY128> { static constexpr int NP = 5; typedef u128 Word; };

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

// mixed-radix digits shared by native64 / native_binary128: (v0, v12, v34)
__device__ __forceinline__ void garner_01234(const uint32_t* r, const NativeConsts& c, uint64_t& v0, uint64_t& v12, uint64_t& v34)
{
    const uint32_t v2 = mul_mod32(c, 2, c.P1_INV_MOD_P2, 2 * c.P[2] + r[2] - r[1]);
    const uint64_t mod_p12 = (uint64_t)r[1] + (uint64_t)v2 * c.P[1];
    const uint32_t v4 = mul_mod32(c, 4, c.P3_INV_MOD_P4, 2 * c.P[4] + r[4] - r[3]);
    const uint64_t mod_p34 = (uint64_t)r[3] + (uint64_t)v4 * c.P[3];
    v0 = r[0];
    v12 = mul_mod64(0 - c.P12, 2 * c.P12 + mod_p12 - v0, c.P0_INV_MOD_P12, c.P0_INV_MOD_P12_SHOUP);
    v34 = mul_mod64(0 - c.P34, 2 * c.P34 + mod_p34 - (v0 + mul_mod64(0 - c.P34, v12, (uint64_t)c.P[0], c.P0_MOD_P34_SHOUP)),
                    c.P012_INV_MOD_P34, c.P012_INV_MOD_P34_SHOUP);
}

template <int KIND>
__device__ __forceinline__ typename KindInfo<KIND>::Word reconstruct(const uint32_t* r, const NativeConsts& c)
{
    if constexpr (KIND == NK_BINARY32) {
        const uint32_t v0 = r[0];
        const uint32_t v1 = mul_mod32(c, 1, c.P0_INV_MOD_P1, 2 * c.P[1] + r[1] - v0);
        const uint32_t pos = v0 + v1 * c.P[0];
        return v1 > (c.P[1] / 2) ? pos - c.P[0] * c.P[1] : pos;
    } else if constexpr (KIND == NK_NATIVE32 || KIND == NK_BINARY64) {
        const uint32_t v0 = r[0];
        const uint32_t v1 = mul_mod32(c, 1, c.P0_INV_MOD_P1, 2 * c.P[1] + r[1] - v0);
        const uint32_t v2 = mul_mod32(c, 2, c.P01_INV_MOD_P2, 2 * c.P[2] + r[2] - (v0 + mul_mod32(c, 2, c.P[0], v1)));
        const bool sign = v2 > (c.P[2] / 2);
        if constexpr (KIND == NK_NATIVE32) {
            const uint32_t _01 = c.P[0] * c.P[1];
            const uint32_t pos = v0 + v1 * c.P[0] + v2 * _01;
            return sign ? pos - _01 * c.P[2] : pos;
        } else {
            const uint64_t _01 = (uint64_t)c.P[0] * c.P[1];
            const uint64_t pos = (uint64_t)v0 + (uint64_t)v1 * c.P[0] + (uint64_t)v2 * _01;
            return sign ? pos - _01 * c.P[2] : pos;
        }
    } else if constexpr (KIND == NK_NATIVE64) {
        uint64_t v0, v12, v34;
        garner_01234(r, c, v0, v12, v34);
        const uint64_t _0 = c.P[0], _012 = _0 * c.P12;
        const uint64_t pos = v0 + v12 * _0 + v34 * _012;
        return v34 > (c.P34 / 2) ? pos - _012 * c.P34 : pos;
    } else if constexpr (KIND == NK_BINARY128) {
        uint64_t v0, v12, v34;
        garner_01234(r, c, v0, v12, v34);
        const u128 _0 = c.P[0], _012 = _0 * (u128)c.P12;
        const u128 pos = (u128)v0 + (u128)v12 * _0 + (u128)v34 * _012;
        return v34 > (c.P34 / 2) ? pos - _012 * (u128)c.P34 : pos;
    } else { // NK_NATIVE128
        uint64_t mp[5];
#pragma unroll
        for (int t = 0; t < 5; t++) {
            const uint32_t inv = t == 0 ? c.P0_INV_MOD_P1 : t == 1 ? c.P2_INV_MOD_P3 : t == 2 ? c.P4_INV_MOD_P5
                                 : t == 3 ? c.P6_INV_MOD_P7 : c.P8_INV_MOD_P9;
            const uint32_t a = r[2 * t];
            const uint32_t b = mul_mod32(c, 2 * t + 1, inv, 2 * c.P[2 * t + 1] + r[2 * t + 1] - a);
            mp[t] = (uint64_t)a + (uint64_t)b * c.P[2 * t];
        }
        const uint64_t n23 = 0 - c.P23, n45 = 0 - c.P45, n67 = 0 - c.P67, n89 = 0 - c.P89;
        const uint64_t v01 = mp[0];
        const uint64_t v23 = mul_mod64(n23, 2 * c.P23 + mp[1] - v01, c.P01_INV_MOD_P23, c.P01_INV_MOD_P23_SHOUP);
        const uint64_t v45 = mul_mod64(n45, 2 * c.P45 + mp[2] - (v01 + mul_mod64(n45, v23, c.P01, c.P01_MOD_P45_SHOUP)),
                                       c.P0123_INV_MOD_P45, c.P0123_INV_MOD_P45_SHOUP);
        const uint64_t v67 = mul_mod64(
            n67,
            2 * c.P67 + mp[3] -
                (v01 + mul_mod64(n67, v23 + mul_mod64(n67, v45, c.P23, c.P23_MOD_P67_SHOUP), c.P01, c.P01_MOD_P67_SHOUP)),
            c.P012345_INV_MOD_P67, c.P012345_INV_MOD_P67_SHOUP);
        const uint64_t v89 = mul_mod64(
            n89,
            2 * c.P89 + mp[4] -
                (v01 + mul_mod64(n89,
                                 v23 + mul_mod64(n89, v45 + mul_mod64(n89, v67, c.P45, c.P45_MOD_P89_SHOUP), c.P23,
                                                 c.P23_MOD_P89_SHOUP),
                                 c.P01, c.P01_MOD_P89_SHOUP)),
            c.P01234567_INV_MOD_P89, c.P01234567_INV_MOD_P89_SHOUP);
        const u128 pos = (u128)v01 + (u128)v23 * (u128)c.P01 + (u128)v45 * mk128(c.P0123) + (u128)v67 * mk128(c.P012345) +
                         (u128)v89 * mk128(c.P01234567);
        return v89 > (c.P89 / 2) ? pos - mk128(c.P0123456789) : pos;
    }
}

} // namespace dev
} // namespace cntt
