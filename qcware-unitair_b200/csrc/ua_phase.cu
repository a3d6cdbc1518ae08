// Fused diagonal phase: psi_k <- exp(-i theta_k) psi_k in ONE pass.
//
// Replaces `torch.exp(-1.j * angles) * state` (src/unitair/simulation/operations.py:41-42),
// which is exp + 2 mul + a complex temporary (5 elementwise passes).  Broadcasting between
// `angles` and `state` is expressed with (batch stride, element stride) pairs; the host
// shim materialises only patterns that cannot be written that way.
//
// Algorithmic traffic per amplitude: 8+4+8 = 20 B (complex64 + f32 angles, full-size
// angles), 16+8+16 = 40 B (complex128 + f64); 16/32 B when the angle is broadcast.
#include "ua_common.cuh"

namespace ua {

__device__ __forceinline__ void sincos_acc(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_acc(double x, double *s, double *c) { sincos(x, s, c); }

struct PhaseArgs {
    void *out;          // forward: out ; backward: grad_state (may be null)
    void *out2;         // backward only: grad_angle (real, may be null)
    const void *in;     // forward: state ; backward: grad_out
    const void *psi;    // backward only: psi_in
    const void *angles;
    long long elems;    // row length
    long long total;    // batch * elems
    long long in_bstride, ang_bstride, ang_estride;
    int elems_shift;    // log2(elems) if a power of two, else -1
    int conj_phase;
};

template <typename R>
__device__ __forceinline__ void split_index(const PhaseArgs &a, long long i, long long &b, long long &e) {
    if (a.elems_shift >= 0) {
        b = i >> a.elems_shift;
        e = i & (a.elems - 1);
    } else {
        b = i / a.elems;
        e = i - b * a.elems;
    }
}

// one amplitude per thread-iteration (any shape / alignment)
template <typename R, bool BWD>
__global__ void __launch_bounds__(256) phase_scalar_kernel(const PhaseArgs a) {
    using C = typename CplxOf<R>::type;
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < a.total; i += stride) {
        long long b, e;
        split_index<R>(a, i, b, e);
        const R th = reinterpret_cast<const R *>(a.angles)[b * a.ang_bstride + e * a.ang_estride];
        R s, c;
        sincos_acc(th, &s, &c);
        if (!BWD) {
            const C x = reinterpret_cast<const C *>(a.in)[b * a.in_bstride + e];
            if (a.conj_phase) s = -s;
            // (c - i s)(x + i y)
            reinterpret_cast<C *>(a.out)[i] = mk(c * x.x + s * x.y, c * x.y - s * x.x);
        } else {
            const C g = reinterpret_cast<const C *>(a.in)[i];
            if (a.out) reinterpret_cast<C *>(a.out)[i] = mk(c * g.x - s * g.y, c * g.y + s * g.x);
            if (a.out2) {
                const C x = reinterpret_cast<const C *>(a.psi)[b * a.in_bstride + e];
                const R wx = c * x.x + s * x.y, wy = c * x.y - s * x.x;   // exp(-i th) psi
                reinterpret_cast<R *>(a.out2)[i] = g.x * wy - g.y * wx;   // Im(conj(g) w)
            }
        }
    }
}

// complex64 forward, two amplitudes (one float4) per thread-iteration, 4 in flight
template <int UNR>
__global__ void __launch_bounds__(256) phase_vec2_kernel(const PhaseArgs a) {
    const long long nvec = a.total >> 1;
    const long long i0 = ((long long)blockIdx.x * UNR) * 256 + threadIdx.x;
    float4 x[UNR];
    float th0[UNR], th1[UNR];
    bool ok[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const long long v = i0 + (long long)u * 256;
        ok[u] = v < nvec;
        if (ok[u]) {
            long long b, e;
            split_index<float>(a, v * 2, b, e);
            x[u] = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(a.in) + b * a.in_bstride + e));
            const float *ang = reinterpret_cast<const float *>(a.angles) + b * a.ang_bstride + e * a.ang_estride;
            th0[u] = ang[0];
            th1[u] = ang[a.ang_estride];
        }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        if (!ok[u]) continue;
        float s0, c0, s1, c1;
        sincosf(th0[u], &s0, &c0);
        sincosf(th1[u], &s1, &c1);
        if (a.conj_phase) { s0 = -s0; s1 = -s1; }
        float4 r;
        r.x = c0 * x[u].x + s0 * x[u].y;
        r.y = c0 * x[u].y - s0 * x[u].x;
        r.z = c1 * x[u].z + s1 * x[u].w;
        r.w = c1 * x[u].w - s1 * x[u].z;
        __stcs(reinterpret_cast<float4 *>(a.out) + i0 + (long long)u * 256, r);
    }
}

static int fill_args(PhaseArgs &a, int dtype, long long elems, long long batch,
                     long long in_bs, long long a_bs, long long a_es, const char *who) {
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("%s: bad dtype", who); return UA_ERR_INVALID; }
    if (elems < 1 || batch < 1) { set_error("%s: bad sizes", who); return UA_ERR_INVALID; }
    if (in_bs != 0 && in_bs != elems) { set_error("%s: in_batch_stride must be 0 or elems", who); return UA_ERR_INVALID; }
    if (a_es != 0 && a_es != 1) { set_error("%s: angle_elem_stride must be 0 or 1", who); return UA_ERR_INVALID; }
    if (a_bs < 0) { set_error("%s: negative angle stride", who); return UA_ERR_INVALID; }
    a.elems = elems;
    a.total = elems * batch;
    a.in_bstride = in_bs; a.ang_bstride = a_bs; a.ang_estride = a_es;
    a.elems_shift = is_pow2(elems) ? ilog2(elems) : -1;
    return UA_OK;
}

static unsigned scalar_grid(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 64;
    return (unsigned)(blocks < cap ? blocks : cap);
}

}  // namespace ua

using namespace ua;

extern "C" int ua_apply_phase(int dtype, void *out, const void *in, const void *angles,
                              long long elems, long long batch, long long in_batch_stride,
                              long long angle_batch_stride, long long angle_elem_stride,
                              int conj_phase, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!out || !in || !angles) { set_error("ua_apply_phase: null pointer"); return UA_ERR_INVALID; }
    PhaseArgs a{};
    int rc = fill_args(a, dtype, elems, batch, in_batch_stride, angle_batch_stride, angle_elem_stride, "ua_apply_phase");
    if (rc) return rc;
    a.out = out; a.out2 = nullptr; a.in = in; a.psi = nullptr; a.angles = angles; a.conj_phase = conj_phase ? 1 : 0;
    if (dtype == UA_C64) {
        const bool vec_ok = (elems % 2 == 0) && !(((uintptr_t)out | (uintptr_t)in) & 15);
        if (vec_ok) {
            constexpr int UNR = 4;
            const long long nvec = a.total >> 1;
            const long long blocks = (nvec + 256 * UNR - 1) / (256 * UNR);
            if (blocks > 0x7fffffffll) { set_error("ua_apply_phase: grid too large"); return UA_ERR_UNSUPPORTED; }
            phase_vec2_kernel<UNR><<<(unsigned)blocks, 256, 0, st>>>(a);
            return check_launch("phase_vec2_kernel");
        }
        phase_scalar_kernel<float, false><<<scalar_grid(a.total), 256, 0, st>>>(a);
    } else {
        phase_scalar_kernel<double, false><<<scalar_grid(a.total), 256, 0, st>>>(a);
    }
    return check_launch("phase_scalar_kernel");
}

extern "C" int ua_phase_backward(int dtype, void *grad_state, void *grad_angle, const void *grad_out,
                                 const void *psi_in, const void *angles, long long elems,
                                 long long batch, long long in_batch_stride,
                                 long long angle_batch_stride, long long angle_elem_stride,
                                 void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!grad_out || !angles || (grad_angle && !psi_in)) { set_error("ua_phase_backward: null pointer"); return UA_ERR_INVALID; }
    PhaseArgs a{};
    int rc = fill_args(a, dtype, elems, batch, in_batch_stride, angle_batch_stride, angle_elem_stride, "ua_phase_backward");
    if (rc) return rc;
    a.out = grad_state; a.out2 = grad_angle; a.in = grad_out; a.psi = psi_in; a.angles = angles; a.conj_phase = 0;
    if (dtype == UA_C64) phase_scalar_kernel<float, true><<<scalar_grid(a.total), 256, 0, st>>>(a);
    else phase_scalar_kernel<double, true><<<scalar_grid(a.total), 256, 0, st>>>(a);
    return check_launch("phase_scalar_kernel<bwd>");
}
