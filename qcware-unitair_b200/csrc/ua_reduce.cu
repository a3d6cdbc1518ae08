// Fused map-reduce over states: |psi|^2, sum |psi|^2, sum d_k |psi_k|^2, sum conj(a) b.
//
// Replaces (s.conj()*s).real and torch.sum(...) chains of
// src/unitair/states/innerprod.py:26,46,59,65 (2-4 elementwise/reduce passes + temporaries)
// with one read of the state.  Row sums are accumulated in fp64 across threads and blocks
// (deterministic two-stage reduction, no atomics) and rounded once to the state's precision.
//
// Algorithmic traffic per amplitude: abs_squared 8+4 = 12 B (c64) / 24 B (c128);
// norm 8/16 B; diag expectation 8+4 = 12 B / 24 B; inner product 16 B / 32 B.
#include "ua_common.cuh"

namespace ua {

enum ReduceOp { OP_NORM = 0, OP_DIAG = 1, OP_INNER = 2 };

struct ReduceArgs {
    const void *a;       // state (NORM, DIAG) or left state (INNER)
    const void *b;       // diag (DIAG, real) or right state (INNER)
    void *out;           // final output (R or complex R), one per row
    double2 *partial;    // [batch * chunks]
    long long elems, a_bstride, b_bstride;
    int chunks;
};

static inline int chunks_for(long long batch, long long elems) {
    // enough blocks to fill the machine, at most 1024 partials per row
    long long per_row = (elems + 8191) / 8192;
    if (per_row > 1024) per_row = 1024;
    if (per_row < 1) per_row = 1;
    return (int)per_row;
}

__device__ __forceinline__ double2 block_reduce_256(double2 v) {
    __shared__ double2 sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (l < 8) ? sh[l] : make_double2(0.0, 0.0);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
            v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        }
    }
    return v;  // valid in thread 0
}

template <typename R, int OP>
__global__ void __launch_bounds__(256) reduce_rows_kernel(const ReduceArgs r) {
    using C = typename CplxOf<R>::type;
    const long long row = blockIdx.x / r.chunks;
    const int chunk = blockIdx.x - (int)(row * r.chunks);
    const long long per = (r.elems + r.chunks - 1) / r.chunks;
    const long long lo = (long long)chunk * per;
    long long hi = lo + per;
    if (hi > r.elems) hi = r.elems;
    const C *__restrict__ a = reinterpret_cast<const C *>(r.a) + row * r.a_bstride;
    R ax = R(0), ay = R(0);
    double2 acc = make_double2(0.0, 0.0);
    int cnt = 0;
    for (long long e = lo + threadIdx.x; e < hi; e += 256) {
        const C x = __ldcs(a + e);
        if (OP == OP_NORM) {
            ax = fma(x.x, x.x, ax);
            ax = fma(x.y, x.y, ax);
        } else if (OP == OP_DIAG) {
            const R d = __ldcs(reinterpret_cast<const R *>(r.b) + row * r.b_bstride + e);
            ax = fma(d, x.x * x.x + x.y * x.y, ax);
        } else {
            const C y = __ldcs(reinterpret_cast<const C *>(r.b) + row * r.b_bstride + e);
            // conj(x) * y
            ax = fma(x.x, y.x, ax); ax = fma(x.y, y.y, ax);
            ay = fma(x.x, y.y, ay); ay = fma(-x.y, y.x, ay);
        }
        if (sizeof(R) == 4 && ++cnt == 16) {   // spill the fp32 running sum into fp64
            acc.x += (double)ax; acc.y += (double)ay;
            ax = R(0); ay = R(0); cnt = 0;
        }
    }
    acc.x += (double)ax; acc.y += (double)ay;
    acc = block_reduce_256(acc);
    if (threadIdx.x == 0) {
        if (r.chunks == 1) {
            if (OP == OP_INNER) reinterpret_cast<C *>(r.out)[row] = mk((R)acc.x, (R)acc.y);
            else reinterpret_cast<R *>(r.out)[row] = (R)acc.x;
        } else {
            r.partial[blockIdx.x] = acc;
        }
    }
}

template <typename R, int OP>
__global__ void __launch_bounds__(256) reduce_final_kernel(const ReduceArgs r) {
    using C = typename CplxOf<R>::type;
    const long long row = blockIdx.x;
    double2 acc = make_double2(0.0, 0.0);
    for (int c = threadIdx.x; c < r.chunks; c += 256) {
        const double2 p = r.partial[row * r.chunks + c];
        acc.x += p.x; acc.y += p.y;
    }
    acc = block_reduce_256(acc);
    if (threadIdx.x == 0) {
        if (OP == OP_INNER) reinterpret_cast<C *>(r.out)[row] = mk((R)acc.x, (R)acc.y);
        else reinterpret_cast<R *>(r.out)[row] = (R)acc.x;
    }
}

template <typename R, int OP>
static int run_reduce(ReduceArgs &r, long long batch, void *ws, size_t ws_bytes, cudaStream_t st,
                      const char *who) {
    r.chunks = chunks_for(batch, r.elems);
    const long long blocks = batch * r.chunks;
    if (blocks > 0x7fffffffll) { set_error("%s: grid too large", who); return UA_ERR_UNSUPPORTED; }
    if (r.chunks > 1) {
        if (!ws || ws_bytes < (size_t)blocks * sizeof(double2)) { set_error("%s: workspace too small", who); return UA_ERR_INVALID; }
        r.partial = reinterpret_cast<double2 *>(ws);
    } else {
        r.partial = nullptr;
    }
    reduce_rows_kernel<R, OP><<<(unsigned)blocks, 256, 0, st>>>(r);
    int rc = check_launch(who);
    if (rc || r.chunks == 1) return rc;
    reduce_final_kernel<R, OP><<<(unsigned)batch, 256, 0, st>>>(r);
    return check_launch(who);
}

// |psi|^2 elementwise
template <typename R>
__global__ void __launch_bounds__(256) abs2_kernel(R *__restrict__ out, const typename CplxOf<R>::type *__restrict__ in, long long count) {
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < count; i += stride) {
        const auto x = __ldcs(in + i);
        __stcs(out + i, x.x * x.x + x.y * x.y);
    }
}

__global__ void __launch_bounds__(256) abs2_vec_kernel(float2 *__restrict__ out, const float4 *__restrict__ in, long long nvec) {
    const long long i0 = (long long)blockIdx.x * 1024 + threadIdx.x;
    float4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (i0 + u * 256 < nvec) x[u] = __ldcs(in + i0 + u * 256);
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (i0 + u * 256 < nvec)
            __stcs(out + i0 + u * 256, make_float2(x[u].x * x[u].x + x[u].y * x[u].y, x[u].z * x[u].z + x[u].w * x[u].w));
}


// Backward of the real-valued reductions in ONE pass (no temporaries):
//   out[b, e] = scale * g[b*gbs + e*ges] * (diag ? diag[b*dbs + e] : 1) * in[b*ibs + e]
// abs_squared: g per amplitude; norm_squared / diag_expectation_value: g per row (ges = 0).
// (PyTorch's convention for a real loss: d|z|^2 -> 2 g z.)
template <typename R>
__global__ void __launch_bounds__(256) real_scale_kernel(typename CplxOf<R>::type *__restrict__ out,
                                                        const typename CplxOf<R>::type *__restrict__ in,
                                                        const R *__restrict__ g, const R *__restrict__ diag,
                                                        long long elems, long long batch, long long ibs,
                                                        long long gbs, long long ges, long long dbs, R scale) {
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    const long long vec_per_row = elems / APV;
    const long long total = vec_per_row * batch;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long b = i / vec_per_row, v = i - b * vec_per_row;
        const long long e = v * APV;
        const V x = ld16<true>(reinterpret_cast<const V *>(in + b * ibs + e));
        R f0 = scale * g[b * gbs + e * ges];
        if (diag) f0 *= diag[b * dbs + e];
        V y;
        if constexpr (APV == 2) {
            R f1 = scale * g[b * gbs + (e + 1) * ges];
            if (diag) f1 *= diag[b * dbs + e + 1];
            y = make_float4(f0 * x.x, f0 * x.y, f1 * x.z, f1 * x.w);
        } else {
            y = make_double2(f0 * x.x, f0 * x.y);
        }
        st16<true>(reinterpret_cast<V *>(out + b * elems + e), y);
    }
}

}  // namespace ua

using namespace ua;

extern "C" size_t ua_reduce_workspace_bytes(long long batch, long long elems) {
    if (batch < 1 || elems < 1) return 0;
    const int c = chunks_for(batch, elems);
    return c > 1 ? (size_t)batch * c * sizeof(double2) : 0;
}

extern "C" int ua_abs_squared(int dtype, void *out_real, const void *in, long long count, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!out_real || !in || count < 1) { set_error("ua_abs_squared: bad arguments"); return UA_ERR_INVALID; }
    if (dtype == UA_C64) {
        if (count % 2 == 0 && !((uintptr_t)in & 15) && !((uintptr_t)out_real & 7)) {
            const long long nvec = count / 2;
            const long long blocks = (nvec + 1023) / 1024;
            if (blocks > 0x7fffffffll) { set_error("ua_abs_squared: grid too large"); return UA_ERR_UNSUPPORTED; }
            abs2_vec_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float2 *>(out_real), reinterpret_cast<const float4 *>(in), nvec);
            return check_launch("abs2_vec_kernel");
        }
        long long blocks = (count + 255) / 256; { const long long cap = (long long)sm_count() * 64; if (blocks > cap) blocks = cap; }
        abs2_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float *>(out_real), reinterpret_cast<const float2 *>(in), count);
    } else if (dtype == UA_C128) {
        long long blocks = (count + 255) / 256; { const long long cap = (long long)sm_count() * 64; if (blocks > cap) blocks = cap; }
        abs2_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<double *>(out_real), reinterpret_cast<const double2 *>(in), count);
    } else { set_error("ua_abs_squared: bad dtype"); return UA_ERR_INVALID; }
    return check_launch("abs2_kernel");
}

extern "C" int ua_norm_squared(int dtype, void *out_real, const void *in, long long elems, long long batch,
                               void *workspace, size_t workspace_bytes, void *stream) {
    if (!out_real || !in || elems < 1 || batch < 1) { set_error("ua_norm_squared: bad arguments"); return UA_ERR_INVALID; }
    ReduceArgs r{};
    r.a = in; r.b = nullptr; r.out = out_real; r.elems = elems; r.a_bstride = elems; r.b_bstride = 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == UA_C64) return run_reduce<float, OP_NORM>(r, batch, workspace, workspace_bytes, st, "ua_norm_squared");
    if (dtype == UA_C128) return run_reduce<double, OP_NORM>(r, batch, workspace, workspace_bytes, st, "ua_norm_squared");
    set_error("ua_norm_squared: bad dtype"); return UA_ERR_INVALID;
}

extern "C" int ua_diag_expectation(int dtype, void *out_real, const void *diag_real, const void *in,
                                   long long elems, long long batch, long long diag_batch_stride,
                                   long long in_batch_stride, void *workspace, size_t workspace_bytes,
                                   void *stream) {
    if (!out_real || !in || !diag_real || elems < 1 || batch < 1) { set_error("ua_diag_expectation: bad arguments"); return UA_ERR_INVALID; }
    if ((diag_batch_stride != 0 && diag_batch_stride != elems) || (in_batch_stride != 0 && in_batch_stride != elems)) {
        set_error("ua_diag_expectation: batch strides must be 0 or elems"); return UA_ERR_INVALID;
    }
    ReduceArgs r{};
    r.a = in; r.b = diag_real; r.out = out_real; r.elems = elems; r.a_bstride = in_batch_stride; r.b_bstride = diag_batch_stride;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == UA_C64) return run_reduce<float, OP_DIAG>(r, batch, workspace, workspace_bytes, st, "ua_diag_expectation");
    if (dtype == UA_C128) return run_reduce<double, OP_DIAG>(r, batch, workspace, workspace_bytes, st, "ua_diag_expectation");
    set_error("ua_diag_expectation: bad dtype"); return UA_ERR_INVALID;
}

extern "C" int ua_inner_product(int dtype, void *out_complex, const void *a, const void *b,
                                long long elems, long long batch, long long a_batch_stride,
                                long long b_batch_stride, void *workspace, size_t workspace_bytes,
                                void *stream) {
    if (!out_complex || !a || !b || elems < 1 || batch < 1) { set_error("ua_inner_product: bad arguments"); return UA_ERR_INVALID; }
    if ((a_batch_stride != 0 && a_batch_stride != elems) || (b_batch_stride != 0 && b_batch_stride != elems)) {
        set_error("ua_inner_product: batch strides must be 0 or elems"); return UA_ERR_INVALID;
    }
    ReduceArgs r{};
    r.a = a; r.b = b; r.out = out_complex; r.elems = elems; r.a_bstride = a_batch_stride; r.b_bstride = b_batch_stride;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == UA_C64) return run_reduce<float, OP_INNER>(r, batch, workspace, workspace_bytes, st, "ua_inner_product");
    if (dtype == UA_C128) return run_reduce<double, OP_INNER>(r, batch, workspace, workspace_bytes, st, "ua_inner_product");
    set_error("ua_inner_product: bad dtype"); return UA_ERR_INVALID;
}

extern "C" int ua_real_scale(int dtype, void *out, const void *in, const void *g_real, const void *diag_real,
                             long long elems, long long batch, long long in_batch_stride,
                             long long g_batch_stride, long long g_elem_stride, long long diag_batch_stride,
                             double scale, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!out || !in || !g_real || elems < 1 || batch < 1) { set_error("ua_real_scale: bad arguments"); return UA_ERR_INVALID; }
    if ((in_batch_stride != 0 && in_batch_stride != elems) || (diag_batch_stride != 0 && diag_batch_stride != elems) ||
        (g_elem_stride != 0 && g_elem_stride != 1)) {
        set_error("ua_real_scale: unsupported strides"); return UA_ERR_INVALID;
    }
    if ((((uintptr_t)out | (uintptr_t)in) & 15) || (dtype == UA_C64 && (elems & 1))) {
        set_error("ua_real_scale: pointers must be 16-byte aligned (complex64 rows of even length)"); return UA_ERR_UNSUPPORTED;
    }
    const long long vecs = (dtype == UA_C64 ? elems / 2 : elems) * batch;
    long long blocks = (vecs + 255) / 256;
    { const long long cap = (long long)sm_count() * 32; if (blocks > cap) blocks = cap; }
    if (dtype == UA_C64)
        real_scale_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float2 *>(out), reinterpret_cast<const float2 *>(in),
            reinterpret_cast<const float *>(g_real), reinterpret_cast<const float *>(diag_real), elems, batch, in_batch_stride,
            g_batch_stride, g_elem_stride, diag_batch_stride, (float)scale);
    else if (dtype == UA_C128)
        real_scale_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<double2 *>(out), reinterpret_cast<const double2 *>(in),
            reinterpret_cast<const double *>(g_real), reinterpret_cast<const double *>(diag_real), elems, batch, in_batch_stride,
            g_batch_stride, g_elem_stride, diag_batch_stride, scale);
    else { set_error("ua_real_scale: bad dtype"); return UA_ERR_INVALID; }
    return check_launch("real_scale_kernel");
}
