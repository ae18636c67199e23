// tools/ubench.cu -- pipe throughput / latency microbenchmark used to calibrate the kernel design
// (FP32 FMA, FP64 FMA, I2F, shared-memory loads) on the box's GPU.  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void k_fma(T* out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == (T)12345.678) out[0] = s;
}
__global__ void k_i2f(float* out, int iters, int seed) {
    int v[8]; float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = seed + threadIdx.x + i; acc[i] = 0.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i] += (float)(short)v[i]; v[i] = v[i] * 3 + 1; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}
template <typename T, int ILP>
static void run(const char* name, int blocks, int threads) {
    T* d; cudaMalloc(&d, 64);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_fma<T, ILP><<<blocks, threads>>>(d, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e0);
    k_fma<T, ILP><<<blocks, threads>>>(d, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)blocks * threads * iters * ILP;
    printf("%-10s ILP=%d blocks=%d threads=%d : %.3f ms  %.2f TFMA/s  (%.1f lanes/clk/SM @1.965GHz,148SM)\n", name, ILP, blocks, threads, ms,
           fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    run<float, 8>("fp32", 148 * 8, 256);
    run<float, 1>("fp32", 148 * 8, 256);
    run<float, 1>("fp32lat", 148, 32);
    run<double, 8>("fp64", 148 * 8, 256);
    run<double, 2>("fp64", 148 * 8, 256);
    run<double, 1>("fp64", 148 * 8, 256);
    run<double, 1>("fp64lat", 148, 32);
    run<double, 8>("fp64 1w", 148 * 4, 32);
    {
        float* d; cudaMalloc(&d, 64);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_i2f<<<148 * 8, 256>>>(d, 4096, 1);
        cudaEventRecord(e0);
        k_i2f<<<148 * 8, 256>>>(d, 4096, 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double n = 148.0 * 8 * 256 * 4096 * 8;
        printf("i2f+fadd+imad: %.3f ms, %.1f conv lanes/clk/SM\n", ms, n / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
