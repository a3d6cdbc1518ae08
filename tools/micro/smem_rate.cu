// Micro-benchmark: shared-memory round trip (LDS.128 x4 -> STS.128 x4 per group) on a 64 KiB tile,
// CTAs of NT threads, several CTAs per SM, with and without the 128 FMAs of a dense 2-qubit gate.
#include <cstdio>
#include <cuda_runtime.h>
template <int NT, int FMA>
__global__ void __launch_bounds__(NT) k(float4 *out, int gates, int stride_log) {
    extern __shared__ float4 tile[];
    for (int i = threadIdx.x; i < 4096; i += NT) tile[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const unsigned off0 = 1u << stride_log, off1 = 2u << stride_log;
    for (int g = 0; g < gates; ++g) {
        for (unsigned grp = threadIdx.x; grp < 1024; grp += NT) {
            // insert two zero bits at stride_log, stride_log+1
            const unsigned lo = grp & (off0 - 1), hi = grp >> stride_log;
            const unsigned b = (hi << (stride_log + 2)) | lo;
            float4 x0 = tile[b], x1 = tile[b | off0], x2 = tile[b | off1], x3 = tile[b | off0 | off1];
            float4 y0 = x1, y1 = x2, y2 = x3, y3 = x0;      // rotate so the round trip is not a no-op
            if (FMA == 2) {
                // 64 FFMA2 (same FLOPs as the 128 scalar FMAs)
                unsigned long long *X = reinterpret_cast<unsigned long long *>(&x0);
                unsigned long long a0 = X[0], a1 = X[1], b0, b1, c0, c1, d0, d1;
                b0 = reinterpret_cast<unsigned long long *>(&x1)[0]; b1 = reinterpret_cast<unsigned long long *>(&x1)[1];
                c0 = reinterpret_cast<unsigned long long *>(&x2)[0]; c1 = reinterpret_cast<unsigned long long *>(&x2)[1];
                d0 = reinterpret_cast<unsigned long long *>(&x3)[0]; d1 = reinterpret_cast<unsigned long long *>(&x3)[1];
                unsigned long long p0 = a0, p1 = a1, p2 = b0, p3 = b1, p4 = c0, p5 = c1, p6 = d0, p7 = d1;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p0) : "l"(b0), "l"(c1));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p1) : "l"(b1), "l"(c0));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p2) : "l"(a0), "l"(d1));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p3) : "l"(a1), "l"(d0));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p4) : "l"(a0), "l"(b1));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p5) : "l"(a1), "l"(b0));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p6) : "l"(c0), "l"(b1));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p7) : "l"(c1), "l"(a0));
                }
                reinterpret_cast<unsigned long long *>(&y0)[0] = p0; reinterpret_cast<unsigned long long *>(&y0)[1] = p1;
                reinterpret_cast<unsigned long long *>(&y1)[0] = p2; reinterpret_cast<unsigned long long *>(&y1)[1] = p3;
                reinterpret_cast<unsigned long long *>(&y2)[0] = p4; reinterpret_cast<unsigned long long *>(&y2)[1] = p5;
                reinterpret_cast<unsigned long long *>(&y3)[0] = p6; reinterpret_cast<unsigned long long *>(&y3)[1] = p7;
            } else if (FMA == 1) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    y0.x = fmaf(x1.x, 0.5f, y0.x); y0.y = fmaf(x2.y, 0.25f, y0.y); y0.z = fmaf(x3.z, 0.5f, y0.z); y0.w = fmaf(x1.w, 0.5f, y0.w);
                    y1.x = fmaf(x0.x, 0.5f, y1.x); y1.y = fmaf(x2.y, 0.25f, y1.y); y1.z = fmaf(x3.z, 0.5f, y1.z); y1.w = fmaf(x0.w, 0.5f, y1.w);
                    y2.x = fmaf(x1.x, 0.5f, y2.x); y2.y = fmaf(x0.y, 0.25f, y2.y); y2.z = fmaf(x3.z, 0.5f, y2.z); y2.w = fmaf(x1.w, 0.5f, y2.w);
                    y3.x = fmaf(x1.x, 0.5f, y3.x); y3.y = fmaf(x2.y, 0.25f, y3.y); y3.z = fmaf(x0.z, 0.5f, y3.z); y3.w = fmaf(x1.w, 0.5f, y3.w);
                }
            }
            tile[b] = y0; tile[b | off0] = y1; tile[b | off1] = y2; tile[b | off0 | off1] = y3;
        }
        __syncthreads();
    }
    out[blockIdx.x * NT + threadIdx.x] = tile[threadIdx.x];
}
template <int NT, int FMA> void run(float4 *out, int ctas_per_sm, const char *name) {
    cudaFuncSetAttribute(k<NT, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int gates = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NT, FMA><<<148 * ctas_per_sm, NT, 65536>>>(out, 10, 8); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<NT, FMA><<<148 * ctas_per_sm, NT, 65536>>>(out, gates, 8); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per SM: ctas_per_sm tiles x gates gate-tiles
    double cyc = ms * 1e-3 * 1.965e9 / (gates * (double)ctas_per_sm);
    printf("%-28s NT=%d ctas/SM=%d: %.0f cycles per gate-tile (64 KiB read + 64 KiB write)\n", name, NT, ctas_per_sm, cyc);
}
int main() {
    float4 *out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float4));
    run<128, 0>(out, 3, "smem only"); run<128, 1>(out, 3, "smem + 128 FFMA"); run<128, 2>(out, 3, "smem + 64 FFMA2");
    run<256, 0>(out, 2, "smem only"); run<256, 1>(out, 2, "smem + 128 FFMA"); run<256, 2>(out, 2, "smem + 64 FFMA2");
    run<512, 0>(out, 1, "smem only"); run<512, 1>(out, 1, "smem + 128 FFMA"); run<512, 2>(out, 1, "smem + 64 FFMA2");
    return 0;
}
