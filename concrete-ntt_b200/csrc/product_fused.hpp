// product_fused.hpp -- fused kernels of product::Plan for its hot shape: a modulus that is the product of TWO primes
// below 2^32 of the same arithmetic class (the tfhe-rs NTT-PBS modulus, src/product.rs:444-445) and N <= 4096.
//
//   fwd:  standard -> (% p0, forward NTT) -> plane 0,  (% p1, forward NTT) -> plane 1          8 + 8 bytes / coefficient
//   inv:  plane j -> inverse NTT -> plane j (the reference leaves the inverse transforms in `ntt`) -> mixed-radix
//         lift -> standard (Replace) or standard += lift mod modulus (Accumulate)              8 + 8 + 8 (+ 8) bytes
//
// against 32 / 32-40 bytes per coefficient of the three-launch composition (k_product_reduce, two prime32 transforms,
// k_product_crt) they replace; every other shape keeps that composition (capi.cu).
#pragma once
#include "ntt_kernels.cuh"
#include "product_kernels.cuh"

namespace cntt {

struct ProductFusedArgs {
    int cls;   // 0: A32L4 (p < 2^30), 1: A32L2 (p < 2^31); both primes
    int logn;
    const uint2* tw_fwd[2];
    const uint2* tw_inv[2];
    const uint2* last_fwd[2];
    const uint2* last_inv[2];
    Mod32 mod[2];
    const TwHead<uint2>* head_fwd[2]; // host copies
    const TwHead<uint2>* head_inv[2];
};

bool product_fused_supported(int cls, int logn);
// true: the fused kernels of this size use an engine geometry of their own, and last_fwd / last_inv must be tables built by
// product_fused_build_last (heap: the prime plan's heap-ordered table, out: 2^logn entries), not the prime plans' own
bool product_fused_own_tables(int logn);
cudaError_t product_fused_build_last(int cls, int logn, const uint2* heap, uint2* out, cudaStream_t st);
// cudaErrorNotSupported when no fused variant exists for (cls, logn)
cudaError_t product_fused_fwd(const ProductConsts& c, const ProductFusedArgs& a, uint64_t* ntt, const uint64_t* standard, int mode,
                              uint64_t bound, size_t batch, cudaStream_t st);
cudaError_t product_fused_inv(const ProductConsts& c, const ProductFusedArgs& a, uint64_t* standard, uint64_t* ntt, int mode,
                              size_t batch, cudaStream_t st);

} // namespace cntt
