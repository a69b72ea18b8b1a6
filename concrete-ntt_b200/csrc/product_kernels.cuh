// product_kernels.cuh -- per-coefficient kernels of product::Plan (src/product.rs:139-967): the composite-modulus
// plan (modulus = product of distinct primes < 2^64) that tfhe-rs drives for the NTT-based PBS.
//
// NTT-domain layout of ONE polynomial (product.rs:261-278): all u32 planes first, bit-cast into the front of the
// &mut [u64], then the u64 planes: domain_len = (n/2) * count32 + n * count64 u64 words.  A batch is the
// reference's slices concatenated: polynomial b at ntt + b * domain_len.  The per-prime transforms and pointwise
// ops run the prime32 / prime64 kernels with poly_stride = 2 * domain_len (u32 words) / domain_len (u64 words).
//
//   k_product_reduce   standard -> residues (fwd's loops, product.rs:283-353): Generic `% p`, the Bounded fast
//                      path of the two-u32-prime case, and the truncating / copying single-prime special cases
//   k_product_crt      planes -> standard (inv's Knuth 4.3.2 mixed-radix lift, product.rs:360-880), Replace or
//                      Accumulate (add_mod_u64), with the same special cases
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cntt {

constexpr int kProductMaxPrimes = 8; // 7 primes = 1 (mod 64) already exceed 2^64; see DESIGN.md

struct ProductConsts {
    int count32, count64;          // planes, primes sorted ascending (u32 ones first)
    uint64_t p[kProductMaxPrimes]; // the primes
    uint64_t recip[kProductMaxPrimes]; // floor((2^64 - 1) / p): Barrett quotient estimate for x mod p, x < 2^64
    // u32 primes below 2^31: Shoup pairs {c, floor(c 2^32 / p)} of c = 1 and c = 2^32 mod p, so that a 64-bit word
    // reduces limb by limb on the 32-bit multiplier (the fused kernels; 64-bit mul.hi is 5x slower on this GPU)
    uint32_t red32[kProductMaxPrimes][2][2];
    uint32_t inv10_32[2];          // Shoup pair of inv[1][0] = p_0^-1 mod p_1 when both are u32 primes below 2^31
    uint64_t inv[kProductMaxPrimes][kProductMaxPrimes]; // inv[j][i] = p_i^-1 mod p_j, i < j (product.rs:205-226)
    uint64_t modulus;
    uint64_t n;                    // polynomial size
    uint64_t domain_len;           // u64 words per polynomial in the NTT domain
};

enum ProductFwdMode { PF_GENERIC = 0, PF_BOUNDED = 1 };
enum ProductInvMode { PI_REPLACE = 0, PI_ACCUMULATE = 1 };

namespace pdev {
// x mod p for any x < 2^64 (2 <= p): q = mulhi(x, recip) is floor(x / p) or one less
__device__ __forceinline__ uint64_t rem64(uint64_t x, uint64_t p, uint64_t recip)
{
    const uint64_t q = __umul64hi(x, recip);
    uint64_t r = x - q * p;
    if (r >= p) r -= p;
    if (r >= p) r -= p; // recip is floor((2^64-1)/p): one more unit of slack than floor(2^64/p)
    return r;
}
__device__ __forceinline__ uint64_t sub_mod(uint64_t p, uint64_t a, uint64_t b) { return a >= b ? a - b : a - b + p; } // product.rs:65-81
__device__ __forceinline__ uint64_t add_mod(uint64_t p, uint64_t a, uint64_t b) // product.rs:83-91
{
    const uint64_t s = a + b;
    return (s >= p || s < a) ? s - p : s;
}
__device__ __forceinline__ uint64_t mul_mod(uint64_t a, uint64_t b, uint64_t p, uint64_t recip)
{
    const uint64_t hi = __umul64hi(a, b);
    if (hi == 0) return rem64(a * b, p, recip);
    return (uint64_t)(((unsigned __int128)a * b) % p); // u64 factors only
}
} // namespace pdev

#ifndef CNTT_PRODUCT_HELPERS_ONLY // product_fused.cu needs the constants and helpers, not a second copy of the kernels
// one thread per coefficient (b, i)
__global__ void __launch_bounds__(256)
k_product_reduce(const ProductConsts c, uint64_t* __restrict__ ntt, const uint64_t* __restrict__ standard, int mode, uint64_t bound,
                 unsigned long long ncoef)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncoef) return;
    const unsigned long long b = idx / c.n;
    const uint64_t i = idx - b * c.n;
    const uint64_t s = standard[idx];
    uint64_t* dom = ntt + b * c.domain_len;
    uint32_t* dom32 = reinterpret_cast<uint32_t*>(dom);
    uint64_t* dom64 = dom + (c.n / 2) * c.count32;
    if (c.count32 == 0 && c.count64 == 1) { dom64[i] = s; return; }                  // product.rs:283-287 (copy, no reduction)
    if (c.count32 == 1 && c.count64 == 0) { dom32[i] = (uint32_t)s; return; }        // product.rs:288-294 (truncation)
    if (c.count32 == 2 && c.count64 == 0 && mode == PF_BOUNDED && bound < c.p[0] && bound < c.p[1]) { // product.rs:305-322
        const bool positive = s < c.modulus / 2;
        const uint32_t s32 = (uint32_t)s;
        const uint32_t complement = (uint32_t)c.modulus - s32;
        dom32[i] = positive ? s32 : (uint32_t)c.p[0] - complement;
        dom32[c.n + i] = positive ? s32 : (uint32_t)c.p[1] - complement;
        return;
    }
    for (int k = 0; k < c.count32; k++) dom32[(uint64_t)k * c.n + i] = (uint32_t)pdev::rem64(s, c.p[k], c.recip[k]);
    for (int k = 0; k < c.count64; k++) dom64[(uint64_t)k * c.n + i] = pdev::rem64(s, c.p[c.count32 + k], c.recip[c.count32 + k]);
}

__global__ void __launch_bounds__(256)
k_product_crt(const ProductConsts c, uint64_t* __restrict__ standard, const uint64_t* __restrict__ ntt, int mode,
              unsigned long long ncoef)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncoef) return;
    const unsigned long long b = idx / c.n;
    const uint64_t i = idx - b * c.n;
    const uint64_t* dom = ntt + b * c.domain_len;
    const uint32_t* dom32 = reinterpret_cast<const uint32_t*>(dom);
    const uint64_t* dom64 = dom + (c.n / 2) * c.count32;
    const int np = c.count32 + c.count64;
    if (np == 0) { if (mode == PI_REPLACE) standard[idx] = 0; return; }              // product.rs:380-386
    if (c.count32 == 1 && c.count64 == 0) {                                          // product.rs:401-417
        const uint32_t v = dom32[i];
        if (mode == PI_REPLACE) standard[idx] = v;
        else standard[idx] = (uint64_t)(uint32_t)pdev::add_mod(c.p[0], (uint64_t)(uint32_t)standard[idx], v); // add_mod_u32 on the truncated word
        return;
    }
    uint64_t v[kProductMaxPrimes];
#pragma unroll
    for (int j = 0; j < kProductMaxPrimes; j++) {
        if (j < np) {
            uint64_t x = j < c.count32 ? (uint64_t)dom32[(uint64_t)j * c.n + i] : dom64[(uint64_t)(j - c.count32) * c.n + i];
            const uint64_t pj = c.p[j];
#pragma unroll
            for (int t = 0; t < kProductMaxPrimes; t++)
                if (t < j) x = pdev::mul_mod(pdev::sub_mod(pj, x, v[t]), c.inv[j][t], pj, c.recip[j]); // product.rs:826-857
            v[j] = x;
        }
    }
    uint64_t acc = 0;                                                                // product.rs:859-869 (Horner, wrapping)
#pragma unroll
    for (int j = kProductMaxPrimes - 1; j >= 0; j--)
        if (j < np) acc = acc * c.p[j] + v[j];
    standard[idx] = mode == PI_REPLACE ? acc : pdev::add_mod(c.modulus, standard[idx], acc);
}
#endif // CNTT_PRODUCT_HELPERS_ONLY

} // namespace cntt
