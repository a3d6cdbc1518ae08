// Shared device/host helpers for the unitair_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/unitair_b200.h"

namespace ua {

// ------------------------------------------------------------------ error plumbing
void set_error(const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;
int check_launch(const char *what);   // counts the launch, maps cudaGetLastError()

// ------------------------------------------------------------------ complex helpers
template <typename R> struct CplxOf;
template <> struct CplxOf<float>  { using type = float2;  };
template <> struct CplxOf<double> { using type = double2; };

// 16-byte vector that the streaming kernels move per load: 2 complex64 or 1 complex128
template <typename R> struct VecOf;
template <> struct VecOf<float>  { using type = float4;  static constexpr int APV = 2; };
template <> struct VecOf<double> { using type = double2; static constexpr int APV = 1; };

__device__ __forceinline__ float2  mk(float x, float y)   { return make_float2(x, y); }
__device__ __forceinline__ double2 mk(double x, double y) { return make_double2(x, y); }

template <typename C> __device__ __forceinline__ C cconj(C a) { a.y = -a.y; return a; }

// acc += a*b (complex), 4 FMAs
template <typename C> __device__ __forceinline__ void cfma(C &acc, const C a, const C b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
template <typename C> __device__ __forceinline__ C cmul(const C a, const C b) {
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}

// ------------------------------------------------------------------ packed fp32 (FFMA2)
// Blackwell issues fma.rn.f32x2 (SASS FFMA2) at ~1.5x the scalar FFMA FLOP rate (measured
// 69.7 vs 45.7 TFLOP/s, tools/micro/ffma_rate.cu).  A complex MAC acc += g*x is two FFMA2:
//   acc = (g.re,g.re)*(x.re,x.im) + acc ;  acc = (-g.im,g.im)*(x.im,x.re) + acc
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(f32x2_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// ------------------------------------------------------------------ index helpers
// insert a zero bit at position p (bits >= p move up by one)
__host__ __device__ __forceinline__ uint64_t insert_zero(uint64_t x, int p) {
    const uint64_t lo = x & ((1ull << p) - 1ull);
    return ((x >> p) << (p + 1)) | lo;
}

// ------------------------------------------------------------------ streaming ld/st
// Streaming (evict-first) 16-byte accesses.  Plain coherent path (no .nc) so the
// kernels stay correct when out aliases in.
template <bool STREAM> __device__ __forceinline__ float4 ld16(const float4 *p) {
    if (STREAM) return __ldcs(p);
    return *p;
}
template <bool STREAM> __device__ __forceinline__ double2 ld16(const double2 *p) {
    if (STREAM) return __ldcs(p);
    return *p;
}
template <bool STREAM> __device__ __forceinline__ void st16(float4 *p, float4 v) {
    if (STREAM) __stcs(p, v); else *p = v;
}
template <bool STREAM> __device__ __forceinline__ void st16(double2 *p, double2 v) {
    if (STREAM) __stcs(p, v); else *p = v;
}

// 32-byte (256-bit, sm_100+) accesses: two adjacent 16-byte vectors in one instruction.
// `p` must be 32-byte aligned.
template <bool STREAM> __device__ __forceinline__ void ld32(const float4 *p, float4 &a, float4 &b) {
    if (STREAM)
        asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
    else
        asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
template <bool STREAM> __device__ __forceinline__ void ld32(const double2 *p, double2 &a, double2 &b) {
    if (STREAM)
        asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
    else
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
template <bool STREAM> __device__ __forceinline__ void st32(float4 *p, const float4 a, const float4 b) {
    if (STREAM)
        asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
    else
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
template <bool STREAM> __device__ __forceinline__ void st32(double2 *p, const double2 a, const double2 b) {
    if (STREAM)
        asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
    else
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

// SMs of the current device (grid caps are multiples of this, never a literal 148)
static inline int sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

static inline bool is_pow2(long long x) { return x > 0 && (x & (x - 1)) == 0; }
static inline int ilog2(long long x) { int r = 0; while ((1ll << (r + 1)) <= x) ++r; return r; }

}  // namespace ua
