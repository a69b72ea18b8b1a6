// xchg.cu -- cost of the inter-pass exchange (a 16 x 16 word transpose among 16 threads) done through shared memory, as the engine
// does it, against a __shfl_xor_sync butterfly network (north_star: "warp-shuffle radix stages").  Registers: 16 u32 per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/xchg tools/ubench/xchg.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int M> __device__ __forceinline__ void transpose16(uint32_t (&x)[16], uint32_t* sm, int lane)
{
    const int o = lane & 15, half = lane >> 4;
    if constexpr (M == 0) { // shared memory: thread o writes word k to element o + 16 k, reads elements 16 o .. 16 o + 15 (one pad word per 32)
        uint32_t* s = sm + half * (256 + 8);
#pragma unroll
        for (int k = 0; k < 16; k++) { const int e = o + 16 * k; s[e + (e >> 5)] = x[k]; }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++) { const int e = 16 * o + k; x[k] = s[e + (e >> 5)]; }
        __syncwarp();
    } else {                // four shfl_xor stages, half of the registers travel in each
#pragma unroll
        for (int s = 8; s >= 1; s >>= 1) {
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if ((k & s) == 0) {
                    const uint32_t a = x[k], b = x[k | s];
                    const uint32_t recv = __shfl_xor_sync(0xffffffffu, up ? a : b, s);
                    x[k] = up ? recv : a;
                    x[k | s] = up ? b : recv;
                }
            }
        }
    }
}
template <int M> __global__ void __launch_bounds__(128) k(uint32_t* out, const uint32_t* in, int iters)
{
    __shared__ uint32_t smem[4 * 2 * (256 + 8)];
    uint32_t x[16];
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t* sm = smem + (threadIdx.x >> 5) * 2 * (256 + 8);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = in[t * 16 + i];
    for (int it = 0; it < iters; it++) {
        transpose16<M>(x, sm, lane);
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = x[i] * 2654435761u + 12345u; // one IMAD per word keeps the values live and distinct
    }
#pragma unroll
    for (int i = 0; i < 16; i++) out[t * 16 + i] = x[i];
}
template <int M> static void run(const char* name, int sms, double ghz, uint32_t* d_in, uint32_t* d_out, uint32_t* h_ref, size_t nthreads)
{
    const int iters = 256, blocks = (int)(nthreads / 128);
    k<M><<<blocks, 128>>>(d_out, d_in, iters);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<M><<<blocks, 128>>>(d_out, d_in, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    uint32_t* h = (uint32_t*)malloc(nthreads * 16 * 4);
    cudaMemcpy(h, d_out, nthreads * 16 * 4, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    if (M == 0) for (size_t i = 0; i < nthreads * 16; i++) h_ref[i] = h[i];
    else for (size_t i = 0; i < nthreads * 16; i++) bad += h_ref[i] != h[i];
    free(h);
    const double words = (double)nthreads * 16 * iters;
    printf("%-34s %8.3f ms  %7.2f words/clk/SM exchanged (+1 IMAD each)  mismatches vs shared-memory version %zu  %s\n", name, ms,
           words / (ms * 1e-3) / sms / (ghz * 1e9), bad, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount, khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    const size_t nthreads = (size_t)sms * 128 * 16;
    uint32_t *d_in, *d_out; cudaMalloc(&d_in, nthreads * 64); cudaMalloc(&d_out, nthreads * 64);
    uint32_t* h = (uint32_t*)malloc(nthreads * 64);
    for (size_t i = 0; i < nthreads * 16; i++) h[i] = (uint32_t)(i * 2246822519u + 7);
    cudaMemcpy(d_in, h, nthreads * 64, cudaMemcpyHostToDevice);
    printf("device %s  SMs %d  clock %.3f GHz; 16 x 16 word transpose among 16 threads, 256 per thread\n", prop.name, sms, ghz);
    run<0>("shared memory (16 STS + 16 LDS)", sms, ghz, d_in, d_out, h, nthreads);
    run<1>("shfl_xor network (32 SHFL + SEL)", sms, ghz, d_in, d_out, h, nthreads);
    return 0;
}
