// native52_kernels.cuh -- per-coefficient kernels of the Plan52 twins (native32 / native64 / native_binary32 /
// native_binary64 ::Plan52, e.g. src/native64.rs:1072-1165): the same plans on one to three ~50-bit primes
// (primes52, src/lib.rs:598-652) with u64 residue planes, transformed by prime64 plans.
//
// In the reference these types exist only under feature = "nightly" on AVX-512-IFMA hosts, and their reconstruction
// has no scalar twin; what it computes is the mixed-radix (Garner) form of the centred CRT lift,
//   v_0 = m_0,  v_k = (m_k - (v_0 + v_1 P_0 + ... + v_{k-1} P_0..P_{k-2})) (P_0..P_{k-1})^-1 mod P_k,
//   value = v_0 + v_1 P_0 + v_2 P_0 P_1 - (v_top > P_top / 2 ? P_0..P_top : 0)      (wrapping in the word)
// (src/native64.rs:770-828, native32.rs:222-253, native_binary32.rs:111-125) -- every v_k canonical, so the function
// of the residues is fixed by arithmetic, not by the instruction sequence.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cntt {

struct Native52Consts {
    int np;            // 1, 2 or 3 primes
    int word_bytes;    // 4 or 8
    int reduce;        // fwd takes `value % P_k` (64-bit words) or `value as u64` (32-bit words: value < P_k)
    uint64_t p[3];
    uint64_t inv[3];   // inv[k] = (P_0 .. P_{k-1})^-1 mod P_k   (k >= 1)
    uint64_t pre[3];   // pre[k] = P_0 .. P_{k-1} mod 2^64 (wrapping), pre[0] = 1
    uint64_t full;     // P_0 .. P_{np-1} mod 2^64
};

__device__ __forceinline__ uint64_t mulmod52(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)(((unsigned __int128)a * b) % p); }

template <bool COPY>
__global__ void __launch_bounds__(256)
k_native52_reduce(const Native52Consts c, const void* __restrict__ value, uint64_t* __restrict__ planes, size_t plane_stride,
                  unsigned long long nwords)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
        const uint64_t v = c.word_bytes == 4 ? (uint64_t)reinterpret_cast<const uint32_t*>(value)[i] : reinterpret_cast<const uint64_t*>(value)[i];
        for (int k = 0; k < c.np; k++) planes[(size_t)k * plane_stride + i] = (COPY || !c.reduce) ? v : v % c.p[k];
    }
}

__global__ void __launch_bounds__(256)
k_native52_crt(const Native52Consts c, void* __restrict__ value, const uint64_t* __restrict__ planes, size_t plane_stride,
               unsigned long long nwords)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
        uint64_t v[3];
        v[0] = planes[i];
        uint64_t acc = v[0];
        for (int k = 1; k < c.np; k++) {
            const uint64_t pk = c.p[k];
            // partial = v_0 + v_1 P_0 + ... mod P_k, then v_k = (m_k - partial) inv_k mod P_k
            uint64_t partial = v[0] % pk, w = 1;
            for (int j = 1; j < k; j++) {
                w = mulmod52(w, c.p[j - 1] % pk, pk);
                partial = (partial + mulmod52(v[j], w, pk)) % pk;
            }
            const uint64_t mk = planes[(size_t)k * plane_stride + i];
            const uint64_t d = mk >= partial ? mk - partial : mk + pk - partial;
            v[k] = mulmod52(d, c.inv[k], pk);
            acc += v[k] * c.pre[k];
        }
        if (v[c.np - 1] > c.p[c.np - 1] / 2) acc -= c.full;
        if (c.word_bytes == 4) reinterpret_cast<uint32_t*>(value)[i] = (uint32_t)acc;
        else reinterpret_cast<uint64_t*>(value)[i] = acc;
    }
}

} // namespace cntt
