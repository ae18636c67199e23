// Throughput of the legacy warp-level tensor-core instructions on sm_100a that an exact integer formulation of the tone
// block sums could use (int16 samples split into two 8-bit slices, phasors into 8-bit signed digits):
//   mma.sync.m16n8k32 u8/s8 -> s32   and   mma.sync.m16n8k16 bf16 -> f32,
// next to DMMA m8n8k4 (tools/ubench.cu: 16.5 SMSP-cycles each).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_imma tools/ubench_imma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma16832(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void bmma16816(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ILP>
__global__ void k_imma(int* out, int iters, unsigned seed) {
    unsigned a[4] = {seed, seed * 3u, seed * 5u, seed * 7u}, b[2] = {seed * 11u, seed * 13u};
    int c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) imma16832(c[i], a, b);
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 0x12345678) out[threadIdx.x] = s;
}
template <int ILP>
__global__ void k_bmma(float* out, int iters, unsigned seed) {
    unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u ^ (seed & 1u), 0x3f803f80u};
    float c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = (float)i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) bmma16816(c[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 1.2345f) out[threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double clk = p.clockRate * 1e3;
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int* d; cudaMalloc(&d, 4096);
    const int iters = 20000;
    for (int wps = 1; wps <= 4; wps *= 2) {
        float ms = time_ms([&] { k_imma<8><<<p.multiProcessorCount, 128 * wps>>>(d, iters, 12345u); });
        double n = (double)p.multiProcessorCount * 4 * wps * iters * 8;
        printf("IMMA m16n8k32 u8*s8, ILP 8, %d warps/SMSP: %.1f SMSP-cycles per instruction, %.1f TMAC/s\n", wps,
               ms * 1e-3 * clk * p.multiProcessorCount * 4 / n, n * 4096 / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_bmma<8><<<p.multiProcessorCount, 128 * wps>>>((float*)d, iters, 12345u); });
        printf("BF16 m16n8k16,       ILP 8, %d warps/SMSP: %.1f SMSP-cycles per instruction, %.1f TMAC/s\n", wps,
               ms * 1e-3 * clk * p.multiProcessorCount * 4 / n, n * 2048 / (ms * 1e-3) / 1e12);
    }
    {
        float ms = time_ms([&] { k_imma<1><<<p.multiProcessorCount, 32>>>(d, iters, 12345u); });
        printf("IMMA latency (dependent chain, one warp): %.1f cycles\n", ms * 1e-3 * clk / iters);
    }
    return 0;
}
