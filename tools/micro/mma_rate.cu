// Legacy tensor path (mma.sync) rates on sm_100a: TF32 m16n8k8 and BF16 m16n8k16.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void tf32_kernel(float *out, int iters) {
    float d[ILP][4];
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800001u, 0x3f800002u, 0x3f800003u};
    unsigned b[2] = {0x3f000000u + threadIdx.x, 0x3f000001u};
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void bf16_kernel(float *out, int iters) {
    float d[ILP][4];
    unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    unsigned b[2] = {0x3f003f00u + threadIdx.x, 0x3f003f00u};
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 4 * 1024);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int threads = warps * 32;
        {
            float ms = time_ms([&] { tf32_kernel<8><<<sms, threads>>>(out, iters); });
            double flop = 2.0 * 16 * 8 * 8 * 8 * (double)iters * warps * sms;
            printf("tf32 m16n8k8  warps/SM=%2d ILP=8: %.3f ms  %.1f TFLOP/s\n", warps, ms, flop / ms / 1e9);
        }
        {
            float ms = time_ms([&] { bf16_kernel<8><<<sms, threads>>>(out, iters); });
            double flop = 2.0 * 16 * 8 * 16 * 8 * (double)iters * warps * sms;
            printf("bf16 m16n8k16 warps/SM=%2d ILP=8: %.3f ms  %.1f TFLOP/s\n", warps, ms, flop / ms / 1e9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
