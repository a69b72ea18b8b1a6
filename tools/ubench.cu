// ubench.cu -- integer-pipe and copy micro-benchmarks that fix the roofline denominators used in DESIGN.md.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ILP 8
#define ITERS 4096

template <int OP>
__global__ void __launch_bounds__(256) k_int(uint32_t* out, uint32_t seed, uint32_t mulc)
{
    uint32_t x[ILP];
    uint64_t y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = seed + threadIdx.x * 7 + i; y[i] = ((uint64_t)x[i] << 32) | (x[i] * 3u); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) x[i] = x[i] * mulc + seed;                                     // IMAD
            else if (OP == 1) x[i] = __umulhi(x[i], mulc) + seed;                       // IMAD.HI (+IADD)
            else if (OP == 2) y[i] = (uint64_t)(uint32_t)y[i] * mulc + y[i];            // IMAD.WIDE.U32
            else if (OP == 3) y[i] = y[i] * (y[i] | 1);                                 // 64-bit mul.lo
            else if (OP == 4) y[i] = __umul64hi(y[i], y[i] | 0x8000000000000001ull) + 1; // 64-bit mul.hi
            else if (OP == 5) x[i] = x[i] + seed + mulc;                                // IADD3
            else if (OP == 6) x[i] = min(x[i] ^ seed, x[i] - mulc);                     // LOP3 + IADD + MNMX
            else if (OP == 7) {                                                         // Shoup butterfly core (3 mul + 4 alu)
                uint32_t q = __umulhi(x[i], mulc);
                uint32_t t = x[i] * seed - q * 1062862849u;
                uint32_t c = min(x[(i + 1) % ILP], x[(i + 1) % ILP] - 2125725698u);
                x[i] = c + t;
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= x[i] ^ (uint32_t)y[i] ^ (uint32_t)(y[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

template <int OP> static void run(const char* name, double ops_per_iter, uint32_t* out, int sms, double clk_ghz)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int blocks = sms * 8;
    k_int<OP><<<blocks, 256>>>(out, 12345u, 2654435761u);
    cudaEventRecord(a);
    k_int<OP><<<blocks, 256>>>(out, 12345u, 2654435761u);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * 256 * ITERS * ILP * ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Gop/s  %7.2f ops/clk/SM (at %.3f GHz)\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / (clk_ghz * 1e9), clk_ghz);
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    printf("device %s  SMs %d  clock attr %.3f GHz\n", prop.name, sms, ghz);
    uint32_t* out; cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    run<0>("IMAD (lo)", 1, out, sms, ghz);
    run<1>("IMAD.HI + IADD", 1, out, sms, ghz);
    run<2>("IMAD.WIDE.U32", 1, out, sms, ghz);
    run<3>("mul.lo.u64", 1, out, sms, ghz);
    run<4>("mul.hi.u64 (+add)", 1, out, sms, ghz);
    run<5>("IADD3", 1, out, sms, ghz);
    run<6>("LOP3+IADD+MNMX", 1, out, sms, ghz);
    run<7>("Shoup bfly core (3mul+4alu)", 1, out, sms, ghz);
    size_t bytes = (size_t)2 << 30;
    uint4 *src, *dst; cudaMalloc(&src, bytes); cudaMalloc(&dst, bytes);
    cudaMemset(src, 1, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        k_copy<<<sms * 16, 256>>>(src, dst, bytes / 16);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("copy 2 GiB (r+w 4 GiB)       %8.3f ms  %8.1f GB/s\n", ms, 2.0 * bytes / ms * 1e-6);
    }
    return 0;
}
