// Micro-benchmark: FP64 DFMA vs DMMA (mma.sync m8n8k4 f64) rate on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k_dfma(double *out, double a0, double b0, int iters) {
    double acc[CH], a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { acc[i] = threadIdx.x * 0.001 + i; a[i] = a0 + i * 0.01; b[i] = b0 + i * 0.02; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) acc[i] = fma(a[i], b[(i + r) % CH], acc[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dmma(double *out, double a0, double b0, int iters) {
    double c0[CH], c1[CH], a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x * 0.001 + i; c1[i] = i; a[i] = a0 + i * 0.01; b[i] = b0 + i * 0.02; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a[i]), "d"(b[(i + r) % CH]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double *out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    const int iters = 1024, blocks = 148 * 4, threads = 512;
    constexpr int CH = 8;
    float ms1 = timeit([&] { k_dfma<CH><<<blocks, threads>>>(out, 1.0001, 0.9999, iters); });
    float ms2 = timeit([&] { k_dmma<CH><<<blocks, threads>>>(out, 1.0001, 0.9999, iters); });
    double fma1 = (double)blocks * threads * iters * 8 * CH;
    double fma2 = (double)blocks * (threads / 32) * iters * 8 * CH * (8.0 * 8 * 4);   // MACs per warp-mma
    printf("DFMA : %.3f ms  %.2f TFLOP/s\n", ms1, 2 * fma1 / ms1 / 1e9);
    printf("DMMA : %.3f ms  %.2f TFLOP/s\n", ms2, 2 * fma2 / ms2 / 1e9);
    return 0;
}
