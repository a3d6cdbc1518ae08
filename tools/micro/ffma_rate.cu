// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int CH>
__global__ void k_ffma(float *out, float a0, float b0, int iters) {
    float acc[CH], a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { acc[i] = threadIdx.x * 0.001f + i; a[i] = a0 + i * 0.01f; b[i] = b0 + i * 0.02f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) acc[i] = fmaf(a[i], b[(i + r) % CH], acc[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_ffma2(float *out, float a0, float b0, int iters) {
    unsigned long long acc[CH], a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        float2 t = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
        acc[i] = *reinterpret_cast<unsigned long long *>(&t);
        float2 ta = make_float2(a0 + i * 0.01f, a0 - i * 0.01f);
        float2 tb = make_float2(b0 + i * 0.02f, b0 - i * 0.02f);
        a[i] = *reinterpret_cast<unsigned long long *>(&ta);
        b[i] = *reinterpret_cast<unsigned long long *>(&tb);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) acc[i] = ffma2(a[i], b[(i + r) % CH], acc[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { float2 t = *reinterpret_cast<float2 *>(&acc[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FFMA2 whose first operand is a scalar broadcast (pack2(g, g) -> SASS "R.F32" operand)
template <int CH>
__global__ void k_ffma2_bcast(float *out, float a0, float b0, int iters) {
    unsigned long long acc[CH], b[CH];
    float a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        float2 t = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
        acc[i] = *reinterpret_cast<unsigned long long *>(&t);
        a[i] = a0 + i * 0.01f;
        float2 tb = make_float2(b0 + i * 0.02f, b0 - i * 0.02f);
        b[i] = *reinterpret_cast<unsigned long long *>(&tb);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                unsigned long long aa;
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a[(i + r) % CH]));
                acc[i] = ffma2(aa, b[i], acc[i]);
            }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { float2 t = *reinterpret_cast<float2 *>(&acc[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    const int iters = 4096, blocks = 148 * 4, threads = 512;
    constexpr int CH = 8;
    float ms1 = timeit([&] { k_ffma<CH><<<blocks, threads>>>(out, 1.0001f, 0.9999f, iters); });
    float ms2 = timeit([&] { k_ffma2<CH><<<blocks, threads>>>(out, 1.0001f, 0.9999f, iters); });
    double fma1 = (double)blocks * threads * iters * 8 * CH;      // FMAs
    double fma2 = fma1 * 2;
    printf("FFMA : %.3f ms  %.2f TFLOP/s\n", ms1, 2 * fma1 / ms1 / 1e9);
    printf("FFMA2: %.3f ms  %.2f TFLOP/s\n", ms2, 2 * fma2 / ms2 / 1e9);
    float ms3 = timeit([&] { k_ffma2_bcast<CH><<<blocks, threads>>>(out, 1.0001f, 0.9999f, iters); });
    printf("FFMA2 (scalar-broadcast operand): %.3f ms  %.2f TFLOP/s\n", ms3, 2 * fma2 / ms3 / 1e9);
    return 0;
}
