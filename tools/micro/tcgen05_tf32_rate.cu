// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (5th-generation tensor cores, fp32
// accumulator in TMEM) for the tile shapes a dense 5-qubit complex64 gate block would use
// (M = 128 state columns, N = 64 = real form of a 32 x 32 complex matrix, K = 8 per instruction),
// next to larger N.  One CTA per SM, one elected thread issues; operands are K-major,
// 128-byte-swizzled tiles in shared memory (contents irrelevant for a rate test).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_tf32_rate tcgen05_tf32_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}\n"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // K-major, SWIZZLE_128B: start address (>>4), LBO ignored (1), SBO = 8 rows x 128 B = 1024 B, version 1
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    // A: 128 rows x 128 B (K = 32 tf32), B: N rows x 128 B
    float *fa = reinterpret_cast<float *>(smem + (base - smem_u32(smem)));
    for (int i = threadIdx.x; i < (128 + N) * 32; i += 128) fa[i] = 1e-3f * (float)(i % 97);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(N < 32 ? 32 : N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_d = tmem_base;
    // instruction descriptor: D = f32, A = B = tf32, K-major both, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint64_t ad = make_desc(base), bd = make_desc(base + 128 * 128);
        uint32_t parity = 0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k)      // K = 32 in four K = 8 steps: +32 B inside the swizzle atom
                mma_tf32(tmem_d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it | k) ? 1u : 0u);
            if ((it & 63) == 63 || it == iters - 1) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                mbar_wait(smem_u32(&bar), parity);
                parity ^= 1;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (threadIdx.x < 32) {
        uint32_t r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(tmem_d));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        if (sink) sink[blockIdx.x * 32 + threadIdx.x] = __uint_as_float(r0) + __uint_as_float(r1) + __uint_as_float(r2) + __uint_as_float(r3);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(N < 32 ? 32 : N));
    }
}

template <int N> void run(int sms, float *sink) {
    const int iters = 4096;
    const size_t smem = (128 + N) * 128 + 1024;
    cudaFuncSetAttribute(k_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_rate<N><<<sms, 128, smem>>>(64, sink);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(err)); return; }
    cudaEventRecord(e0);
    k_rate<N><<<sms, 128, smem>>>(iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = (double)sms * iters * 4 * 128.0 * N * 8 * 2;
    printf("tcgen05.mma kind::tf32 M=128 N=%3d K=8: %.3f ms  %.1f TFLOP/s dense  (3xTF32 split: %.1f effective)  status %s\n",
           N, ms, flop / ms / 1e9, flop / ms / 1e9 / 3, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *sink; cudaMalloc(&sink, sms * 32 * sizeof(float));
    run<64>(sms, sink);
    run<128>(sms, sink);
    run<256>(sms, sink);
    float h[4]; cudaMemcpy(h, sink, sizeof(h), cudaMemcpyDeviceToHost);
    printf("sample accumulator values: %g %g %g %g\n", h[0], h[1], h[2], h[3]);
    return 0;
}
