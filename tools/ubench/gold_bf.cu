// gold_bf.cu -- throughput of Goldilocks (p = 2^64 - 2^32 + 1) butterfly formulations in isolation (registers only).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iconcrete-ntt_b200/csrc -o build/gold_bf tools/ubench/gold_bf.cu
// Each thread owns 16 words and runs 4 butterfly levels (32 butterflies) per iteration, like one register pass of the
// CTA kernel.  Variants: 0 = A64S::fwd_bf with table twiddles (the shipped butterfly), 1 = shift butterflies with the
// compile-time exponents of transform levels 0..3, 2 = A64S::inv_bf, 3 = inverse shift butterflies.
#include "arith.cuh"
#include <cstdio>
#include <cstdlib>
using namespace cntt;
typedef uint64_t u64; typedef uint32_t u32;
static constexpr u64 P = 0xFFFFFFFF00000001ull;

// ---- x * 2^K mod p for any 64-bit x, result <= p - 1 ... see DESIGN notes; K in [0,192)
template <int B> __device__ __forceinline__ u64 shl_a0(u64 x)   // x * 2^B, 1 <= B <= 31
{
    const u32 y2 = (u32)(x >> (64 - B));
    const u64 A = x << B;
    const u64 D = ((u64)(~y2) << 32) | (u64)(y2 + 1u);           // p - y2 * EPS
    return A64S::sub_lazy(A, D);
}
template <int B> __device__ __forceinline__ u64 shl_a2(u64 x)   // x * 2^(64+B)
{
    const u32 y0 = (u32)x << B;
    const u64 A = (u64)y0 * 0xFFFFFFFFull;
    const u64 D = x >> (32 - B);
    return A64S::sub_lazy(A, D);
}
template <int B> __device__ __forceinline__ u64 shl_a1(u64 x)   // x * 2^(32+B): generic 4-limb reduce of (y2,y1,y0,0)
{
    const u32 y0 = (u32)x << B, y1 = (u32)(x >> (32 - B)), y2 = (u32)(x >> (64 - B));
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        ".reg .pred q;\n\t"
        "sub.cc.u32      %0, 0, %4;\n\t"          // (y0:0) - y2
        "subc.cc.u32     %1, %2, 0;\n\t"
        "subc.u32        m, 0, 0;\n\t"
        "sub.cc.u32      %0, %0, m;\n\t"
        "subc.u32        %1, %1, 0;\n\t"
        "mad.lo.cc.u32   %0, %3, 0xFFFFFFFF, %0;\n\t"
        "madc.hi.cc.u32  %1, %3, 0xFFFFFFFF, %1;\n\t"
        "addc.u32        m, 0, 0;\n\t"
        "setp.eq.u32     q, %1, 0xFFFFFFFF;\n\t"
        "setp.ne.and.u32 q, %0, 0, q;\n\t"
        "setp.ne.or.u32  q, m, 0, q;\n\t"
        "@q add.cc.u32   %0, %0, 0xFFFFFFFF;\n\t"
        "@q addc.u32     %1, %1, 0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1) : "r"(y0), "r"(y1), "r"(y2));
    return A64S::pack(r0, r1);
}
template <int K> __device__ __forceinline__ u64 shl_mod(u64 x)  // |x * 2^K| with the sign dropped (K mod 96)
{
    constexpr int k = K % 96, a = k / 32, b = k % 32;
    static_assert(b != 0 || k == 0, "limb-aligned shifts do not occur");
    if constexpr (k == 0) return x;
    else if constexpr (a == 0) return shl_a0<b>(x);
    else if constexpr (a == 1) return shl_a1<b>(x);
    else return shl_a2<b>(x);
}
template <int K> __device__ __forceinline__ void fwd_bf_shift(u64& z0, u64& z1)
{
    const u64 t = shl_mod<K>(z1);
    const u64 a = A64S::add_lazy(z0, t), b = A64S::sub_lazy(z0, t);
    if constexpr ((K % 192) >= 96) { z0 = b; z1 = a; } else { z0 = a; z1 = b; }
}
template <int K> __device__ __forceinline__ void inv_bf_shift(u64& z0, u64& z1)   // canonical in / out
{
    const u64 a = A64S::add(z0, z1);
    const u64 d = (K % 192) >= 96 ? A64S::sub_lazy(z1, z0) : A64S::sub_lazy(z0, z1);
    z0 = a; z1 = shl_mod<K>(d);
}
// add_lazy with the end-around fix on the multiply pipe: r + carry * EPS as ONE IMAD.WIDE.  EPS must be opaque to the compiler (a
// kernel parameter), or ptxas strength-reduces the multiply back into ALU adds.
__device__ __forceinline__ u64 add_lazy_w(u64 z, u64 x, u32 eps)
{
    u64 r;
    asm("{\n\t"
        ".reg .u32 r0, r1, m;\n\t"
        ".reg .u64 rr;\n\t"
        "add.cc.u32   r0, %1, %3;\n\t"
        "addc.cc.u32  r1, %2, %4;\n\t"
        "addc.u32     m, 0, 0;\n\t"
        "mov.b64      rr, {r0, r1};\n\t"
        "mad.wide.u32 %0, m, %5, rr;\n\t"
        "}"
        : "=l"(r) : "r"((u32)z), "r"((u32)(z >> 32)), "r"((u32)x), "r"((u32)(x >> 32)), "r"(eps));
    return r;
}
template <int K> __device__ __forceinline__ void fwd_bf_shift_w(u64& z0, u64& z1, u32 eps)
{
    const u64 t = shl_mod<K>(z1);
    const u64 a = add_lazy_w(z0, t, eps), b = A64S::sub_lazy(z0, t);
    if constexpr ((K % 192) >= 96) { z0 = b; z1 = a; } else { z0 = a; z1 = b; }
}
__device__ __forceinline__ void fwd_bf_w(u64& z0, u64& z1, u64 t, u32 eps)
{
    const u64 x = A64S::mul(z1, t);
    const u64 a = add_lazy_w(z0, x, eps), b = A64S::sub_lazy(z0, x);
    z0 = a; z1 = b;
}
// Karatsuba: three 32 x 32 products instead of four.  The 33-bit sums and the 66-bit middle term are
// left to the compiler (unsigned __int128); the 128 -> 64-bit reduction is A64S::mul's.
__device__ __forceinline__ u64 mul_kara(u64 a, u64 b)
{
    typedef unsigned __int128 u128;
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    const u64 z0 = (u64)a0 * b0, z2 = (u64)a1 * b1;
    const u64 sa = (u64)a0 + a1, sb = (u64)b0 + b1;
    const u32 sal = (u32)sa, sbl = (u32)sb;
    const bool ca = (sa >> 32) != 0, cb = (sb >> 32) != 0;
    u128 z1 = (u128)((u64)sal * sbl) + ((u128)(ca ? sbl : 0u) << 32) + ((u128)(cb ? sal : 0u) << 32) + ((u128)((ca && cb) ? 1u : 0u) << 64);
    z1 -= z0; z1 -= z2;
    const u128 w = (u128)z0 + (z1 << 32) + ((u128)z2 << 64);
    const u32 c0 = (u32)w, c1 = (u32)(w >> 32), c2 = (u32)(w >> 64), c3 = (u32)(w >> 96);
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        ".reg .pred q;\n\t"
        "sub.cc.u32      %0, %2, %5;\n\t"
        "subc.cc.u32     %1, %3, 0;\n\t"
        "subc.u32        m, 0, 0;\n\t"
        "sub.cc.u32      %0, %0, m;\n\t"
        "subc.u32        %1, %1, 0;\n\t"
        "mad.lo.cc.u32   %0, %4, 0xFFFFFFFF, %0;\n\t"
        "madc.hi.cc.u32  %1, %4, 0xFFFFFFFF, %1;\n\t"
        "addc.u32        m, 0, 0;\n\t"
        "setp.eq.u32     q, %1, 0xFFFFFFFF;\n\t"
        "setp.ne.and.u32 q, %0, 0, q;\n\t"
        "setp.ne.or.u32  q, m, 0, q;\n\t"
        "@q add.cc.u32   %0, %0, 0xFFFFFFFF;\n\t"
        "@q addc.u32     %1, %1, 0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1) : "r"(c0), "r"(c1), "r"(c2), "r"(c3));
    return (u64)r0 | ((u64)r1 << 32);
}
__device__ __forceinline__ void fwd_bf_kara(u64& z0, u64& z1, u64 t)
{
    const u64 x = mul_kara(z1, t);
    const u64 a = A64S::add_lazy(z0, x), b = A64S::sub_lazy(z0, x);
    z0 = a; z1 = b;
}
// exponents of the first four transform levels (heap order, entry h = 2^level + block): tw[h] = 2^E[h]
__device__ constexpr int kExp[16] = {0, 48, 120, 168, 156, 12, 84, 132, 78, 126, 6, 54, 42, 90, 162, 18};

template <int J, int G, int U> struct BfIdx { static constexpr int half = 16 >> (J + 1); static constexpr int i0 = 2 * half * G + U; };

template <int V> __device__ __forceinline__ void pass(u64 (&x)[16], const u64* tw, const Mod64& m, u32 eps)
{
    if constexpr (V == 6) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int half = 16 >> (j + 1);
#pragma unroll
            for (int g = 0; g < (1 << j); g++) {
                const u64 t = tw[(1 << j) + g];
#pragma unroll
                for (int u = 0; u < half; u++) fwd_bf_kara(x[2 * half * g + u], x[2 * half * g + u + half], t);
            }
        }
    } else if constexpr (V == 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int half = 16 >> (j + 1);
#pragma unroll
            for (int g = 0; g < (1 << j); g++) {
                const u64 t = tw[(1 << j) + g];
#pragma unroll
                for (int u = 0; u < half; u++) fwd_bf_w(x[2 * half * g + u], x[2 * half * g + u + half], t, eps);
            }
        }
    } else if constexpr (V == 0 || V == 2) {
#pragma unroll
        for (int j0 = 0; j0 < 4; j0++) {
            const int j = V == 0 ? j0 : 3 - j0;
            const int half = 16 >> (j + 1);
#pragma unroll
            for (int g = 0; g < (1 << j); g++) {
                const u64 t = tw[(1 << j) + g];
#pragma unroll
                for (int u = 0; u < half; u++) {
                    if constexpr (V == 0) A64S::fwd_bf(x[2 * half * g + u], x[2 * half * g + u + half], t, m);
                    else A64S::inv_bf(x[2 * half * g + u], x[2 * half * g + u + half], t, m);
                }
            }
        }
    } else {
#define BF(J, G) { constexpr int half = 16 >> ((J) + 1); _Pragma("unroll") for (int u = 0; u < half; u++) { \
        if constexpr (V == 1) fwd_bf_shift<kExp[(1 << (J)) + (G)]>(x[2 * half * (G) + u], x[2 * half * (G) + u + half]); \
        else if constexpr (V == 5) fwd_bf_shift_w<kExp[(1 << (J)) + (G)]>(x[2 * half * (G) + u], x[2 * half * (G) + u + half], eps); \
        else inv_bf_shift<(192 - kExp[(1 << (J)) + (G)]) % 192>(x[2 * half * (G) + u], x[2 * half * (G) + u + half]); } }
        if constexpr (V == 1 || V == 5) {
            BF(0, 0) BF(1, 0) BF(1, 1) BF(2, 0) BF(2, 1) BF(2, 2) BF(2, 3)
            BF(3, 0) BF(3, 1) BF(3, 2) BF(3, 3) BF(3, 4) BF(3, 5) BF(3, 6) BF(3, 7)
        } else {
            BF(3, 0) BF(3, 1) BF(3, 2) BF(3, 3) BF(3, 4) BF(3, 5) BF(3, 6) BF(3, 7)
            BF(2, 0) BF(2, 1) BF(2, 2) BF(2, 3) BF(1, 0) BF(1, 1) BF(0, 0)
        }
#undef BF
    }
}

struct TwArr { u64 e[16]; };
template <int V>
__global__ void __launch_bounds__(128) k_bench(u64* out, const u64* in, const __grid_constant__ TwArr tw, int iters, u32 eps)
{
    u64 x[16];
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = in[t * 16 + k];
    Mod64 m = {};
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        pass<V>(x, tw.e, m, eps);
        if constexpr (V == 2 || V == 3) { /* inverse keeps canonical values */ }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) out[t * 16 + k] = A64S::canon(x[k]);
}

// ---- host reference -------------------------------------------------------------------------------
static u64 mulmod(u64 a, u64 b) { return (u64)((unsigned __int128)a * b % P); }
static u64 powmod(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = mulmod(r, a); a = mulmod(a, a); e >>= 1; } return r; }
static void ref_pass(u64* x, const u64* tw, bool fwd)
{
    for (int j0 = 0; j0 < 4; j0++) {
        int j = fwd ? j0 : 3 - j0, half = 16 >> (j + 1);
        for (int g = 0; g < (1 << j); g++)
            for (int u = 0; u < half; u++) {
                u64 &a = x[2 * half * g + u], &b = x[2 * half * g + u + half];
                typedef unsigned __int128 u128;
                if (fwd) { u64 t = mulmod(b % P, tw[(1 << j) + g]); u64 s = (u64)(((u128)(a % P) + t) % P), d = (u64)(((u128)(a % P) + P - t) % P); a = s; b = d; }
                else { u64 s = (u64)(((u128)a + b) % P), d = mulmod((u64)(((u128)a + P - b) % P), tw[(1 << j) + g]); a = s; b = d; }
            }
    }
}
template <int V> static int run(const char* name, int sms, double ghz, const TwArr& tw, u64* d_in, u64* d_out, const u64* h_in, size_t nthreads)
{
    const int iters = 64, blocks = (int)(nthreads / 128);
    // correctness: one iteration against the host reference
    k_bench<V><<<blocks, 128>>>(d_out, d_in, tw, 1, 0xFFFFFFFFu);
    u64* h_out = (u64*)malloc(nthreads * 16 * 8);
    cudaMemcpy(h_out, d_out, nthreads * 16 * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (size_t t = 0; t < nthreads && t < 4096; t++) {
        u64 x[16];
        for (int k = 0; k < 16; k++) x[k] = h_in[t * 16 + k];
        ref_pass(x, tw.e, V == 0 || V == 1 || V == 4 || V == 5 || V == 6);
        for (int k = 0; k < 16; k++) if (x[k] % P != h_out[t * 16 + k]) bad++;
    }
    free(h_out);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_bench<V><<<blocks, 128>>>(d_out, d_in, tw, iters, 0xFFFFFFFFu);
    cudaEventRecord(a);
    k_bench<V><<<blocks, 128>>>(d_out, d_in, tw, iters, 0xFFFFFFFFu);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double bf = (double)nthreads * iters * 32;
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_bench<V>, 128, 0);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_bench<V>);
    printf("%-44s %8.3f ms  %6.2f butterflies/clk/SM  (regs %d, %d CTAs/SM, mismatches %d) %s\n", name, ms, bf / (ms * 1e-3) / sms / (ghz * 1e9), fa.numRegs, nb, bad, cudaGetErrorString(cudaGetLastError()));
    return bad;
}
int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount, khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    printf("device %s  SMs %d  clock attr %.3f GHz\n", prop.name, sms, ghz);
    TwArr tw, twi; int E[16] = {0, 48, 120, 168, 156, 12, 84, 132, 78, 126, 6, 54, 42, 90, 162, 18};
    for (int i = 0; i < 16; i++) { tw.e[i] = powmod(2, E[i]); twi.e[i] = powmod(2, (192 - E[i]) % 192); }
    const size_t nthreads = (size_t)sms * 128 * 16;
    u64* h_in = (u64*)malloc(nthreads * 16 * 8);
    srand(7);
    for (size_t i = 0; i < nthreads * 16; i++) {
        u64 v = ((u64)rand() << 62) ^ ((u64)rand() << 31) ^ (u64)rand();
        if (i % 97 == 0) v = P - 1 - (i % 3); if (i % 101 == 0) v = (u64)(i % 5) << 63; if (i % 103 == 0) v = 0xFFFFFFFF00000000ull;
        h_in[i] = v % P;
    }
    u64 *d_in, *d_out; cudaMalloc(&d_in, nthreads * 16 * 8); cudaMalloc(&d_out, nthreads * 16 * 8);
    cudaMemcpy(d_in, h_in, nthreads * 16 * 8, cudaMemcpyHostToDevice);
    int bad = 0;
    bad += run<0>("fwd: A64S::fwd_bf, table twiddles (shipped)", sms, ghz, tw, d_in, d_out, h_in, nthreads);
    bad += run<1>("fwd: shift butterflies (levels 0..3)", sms, ghz, tw, d_in, d_out, h_in, nthreads);
    bad += run<4>("fwd: fwd_bf, add fix as IMAD.WIDE", sms, ghz, tw, d_in, d_out, h_in, nthreads);
    bad += run<5>("fwd: shift butterflies, add fix as IMAD.WIDE", sms, ghz, tw, d_in, d_out, h_in, nthreads);
    bad += run<6>("fwd: fwd_bf with a Karatsuba (3-product) multiply", sms, ghz, tw, d_in, d_out, h_in, nthreads);
    bad += run<2>("inv: A64S::inv_bf, table twiddles (shipped)", sms, ghz, twi, d_in, d_out, h_in, nthreads);
    bad += run<3>("inv: shift butterflies (levels 3..0)", sms, ghz, twi, d_in, d_out, h_in, nthreads);
    return bad ? 1 : 0;
}
