// Micro-benchmark: 2-qubit gate applied to 16 register-resident amplitudes per thread, matrix
// operands from the constant bank through UNIFORM registers (no vector registers for the
// matrix).  Variant A: scalar FFMA with a UR operand.  Variant B: packed FFMA2 with a
// UR-broadcast operand + 2 FADD per amplitude to combine.  Prints ms per 2-qubit gate
// extrapolated to a 2^30-amplitude state (the fused pass's gate-phase floor).
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float2 c_m[4096];
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 bc(float g) { u64 d; asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(g)); return d; }

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_ffma_ur(float2 *data, const int *offs, int ng, int iters) {
    float2 a[16];
    for (int i = 0; i < 16; ++i) a[i] = data[threadIdx.x * 16 + i + blockIdx.x * 4096];
    for (int it = 0; it < iters; ++it)
        for (int g = 0; g < ng; ++g) {
            const float2 *M = c_m + offs[g];
#pragma unroll
            for (int grp = 0; grp < 4; ++grp) {
                float2 x[4], y[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) x[c] = a[grp * 4 + c];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float2 m = M[r * 4 + 0];
                    y[r].x = m.x * x[0].x - m.y * x[0].y;
                    y[r].y = m.x * x[0].y + m.y * x[0].x;
#pragma unroll
                    for (int c = 1; c < 4; ++c) {
                        m = M[r * 4 + c];
                        y[r].x = fmaf(m.x, x[c].x, y[r].x);
                        y[r].x = fmaf(-m.y, x[c].y, y[r].x);
                        y[r].y = fmaf(m.x, x[c].y, y[r].y);
                        y[r].y = fmaf(m.y, x[c].x, y[r].y);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) a[grp * 4 + c] = y[c];
            }
        }
    for (int i = 0; i < 16; ++i) data[threadIdx.x * 16 + i + blockIdx.x * 4096] = a[i];
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_ffma2_ur(u64 *data, const int *offs, int ng, int iters) {
    u64 a[16];
    for (int i = 0; i < 16; ++i) a[i] = data[threadIdx.x * 16 + i + blockIdx.x * 4096];
    for (int it = 0; it < iters; ++it)
        for (int g = 0; g < ng; ++g) {
            const float2 *M = c_m + offs[g];
#pragma unroll
            for (int grp = 0; grp < 4; ++grp) {
                u64 x[4], P[4], Q[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) x[c] = a[grp * 4 + c];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float2 m = M[r * 4 + 0];
                    P[r] = fmul2(bc(m.x), x[0]);
                    Q[r] = fmul2(bc(m.y), x[0]);
#pragma unroll
                    for (int c = 1; c < 4; ++c) {
                        m = M[r * 4 + c];
                        P[r] = ffma2(bc(m.x), x[c], P[r]);
                        Q[r] = ffma2(bc(m.y), x[c], Q[r]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float2 p = *reinterpret_cast<float2 *>(&P[c]), q = *reinterpret_cast<float2 *>(&Q[c]);
                    float2 y = make_float2(p.x - q.y, p.y + q.x);
                    a[grp * 4 + c] = *reinterpret_cast<u64 *>(&y);
                }
            }
        }
    for (int i = 0; i < 16; ++i) data[threadIdx.x * 16 + i + blockIdx.x * 4096] = a[i];
}

// Variant C: the matrix in VECTOR registers (loaded from shared memory), scalar 3-register FFMA:
// what the round-1 gate phase issued
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_ffma_vr(float2 *data, const float2 *gm, int ng, int iters) {
    __shared__ float2 sm[16 * 8];
    if (threadIdx.x < 128) sm[threadIdx.x] = gm[threadIdx.x];
    __syncthreads();
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = data[threadIdx.x * 16 + i + blockIdx.x * 4096];
    for (int it = 0; it < iters; ++it)
        for (int g = 0; g < ng; ++g) {
            float2 M[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) M[e] = sm[(g & 7) * 16 + e];
#pragma unroll
            for (int grp = 0; grp < 2; ++grp) {
                float2 x[4], y[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) x[c] = a[grp * 4 + c];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float2 m = M[r * 4 + 0];
                    y[r].x = m.x * x[0].x - m.y * x[0].y;
                    y[r].y = m.x * x[0].y + m.y * x[0].x;
#pragma unroll
                    for (int c = 1; c < 4; ++c) {
                        m = M[r * 4 + c];
                        y[r].x = fmaf(m.x, x[c].x, y[r].x);
                        y[r].x = fmaf(-m.y, x[c].y, y[r].x);
                        y[r].y = fmaf(m.x, x[c].y, y[r].y);
                        y[r].y = fmaf(m.y, x[c].x, y[r].y);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) a[grp * 4 + c] = y[c];
            }
        }
    for (int i = 0; i < 8; ++i) data[threadIdx.x * 16 + i + blockIdx.x * 4096] = a[i];
}

template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float2 hm[4096];
    for (int i = 0; i < 4096; ++i) hm[i] = make_float2(0.25f * cosf(i * 0.37f), 0.25f * sinf(i * 0.37f));
    cudaMemcpyToSymbol(c_m, hm, sizeof(hm));
    float2 *gm; cudaMalloc(&gm, sizeof(hm)); cudaMemcpy(gm, hm, sizeof(hm), cudaMemcpyHostToDevice);
    int hoffs[8] = {0, 16, 32, 48, 64, 80, 96, 112}, *offs;
    cudaMalloc(&offs, sizeof(hoffs)); cudaMemcpy(offs, hoffs, sizeof(hoffs), cudaMemcpyHostToDevice);
    const int ng = 8, iters = 512;
    for (int per_sm = 1; per_sm <= 4; ++per_sm) {
        const int blocks = sms * per_sm;
        float2 *d; cudaMalloc(&d, (size_t)blocks * 4096 * sizeof(float2)); cudaMemset(d, 0, (size_t)blocks * 4096 * sizeof(float2));
        const double gate_amps = (double)blocks * 256 * 16 * ng * iters;     // amplitude updates by 2-qubit gates
        float ms;
        ms = timeit([&] { if (per_sm <= 3) k_ffma_ur<3><<<blocks, 256>>>(d, offs, ng, iters); else k_ffma_ur<4><<<blocks, 256>>>(d, offs, ng, iters); });
        printf("CTAs/SM=%d  FFMA.UR  : %.3f ms  %.2f TFLOP/s  -> %.3f ms per 2q gate at 2^30 amps\n", per_sm, ms, gate_amps * 32 / ms / 1e9, ms / gate_amps * 1073741824.0);
        ms = timeit([&] { if (per_sm <= 3) k_ffma2_ur<3><<<blocks, 256>>>((u64 *)d, offs, ng, iters); else k_ffma2_ur<4><<<blocks, 256>>>((u64 *)d, offs, ng, iters); });
        printf("CTAs/SM=%d  FFMA2.UR : %.3f ms  %.2f TFLOP/s  -> %.3f ms per 2q gate at 2^30 amps\n", per_sm, ms, gate_amps * 32 / ms / 1e9, ms / gate_amps * 1073741824.0);
        const double gate_amps_v = gate_amps / 2;
        ms = timeit([&] { if (per_sm <= 3) k_ffma_vr<3><<<blocks, 256>>>(d, gm, ng, iters); else k_ffma_vr<4><<<blocks, 256>>>(d, gm, ng, iters); });
        printf("CTAs/SM=%d  FFMA.VR  : %.3f ms  %.2f TFLOP/s  -> %.3f ms per 2q gate at 2^30 amps\n", per_sm, ms, gate_amps_v * 32 / ms / 1e9, ms / gate_amps_v * 1073741824.0);
        cudaFree(d);
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
