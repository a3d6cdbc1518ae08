// Register-blocked gate phase of the fused pass (see ClusterArgs in ua_tile.cu).
//
// A thread holds the 16 amplitudes of one cluster group as float2 v[16]; member m has cluster
// bit i set iff bit i of m is set.  Gates act on compile-time member indices, so v[] never
// leaves the register file; the matrix `M` points into the kernel parameters and is indexed
// with warp-uniform offsets only (-> LDCU into uniform registers, UR operands in the FMAs).
#pragma once
#include "ua_common.cuh"

namespace ua {

__device__ __forceinline__ f32x2_t fmul2(f32x2_t a, f32x2_t b) {
    f32x2_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2_t bcast2(float g) { return pack2(g, g); }

// y[r] = sum_c G[r][c] * x[c] for D complex inputs.  A matrix travels in the kernel parameters
// row by row as  [gr(r,0) .. gr(r,D-1)] [(-gi, gi)(r,0) .. (-gi, gi)(r,D-1)]  (3 D floats per row:
// 48 bytes = three 16-byte uniform loads for a 2-qubit gate, 24 bytes = three 8-byte loads for a
// 1-qubit gate).
// ARITH 0: scalar FFMA (4 per complex MAC).  ARITH 1: packed FFMA2, two per complex MAC and
// nothing else:  acc += (gr, gr) * (xr, xi);  acc += (-gi, gi) * (xi, xr).  The swapped input is
// an operand selector of the instruction (R.F32x2.LO_HI), the sign pair a packed uniform-register
// operand (UR.F32x2): no moves, no combine step (the earlier P/Q form needed two FADD per output).
template <int D>
__device__ __forceinline__ void load_row(const float *__restrict__ M, int r, float (&gr)[D], float2 (&gp)[D]) {
    if constexpr (D == 4) {
        const float4 *M4 = reinterpret_cast<const float4 *>(M);
        const float4 g = M4[3 * r], a = M4[3 * r + 1], b = M4[3 * r + 2];
        gr[0] = g.x; gr[1] = g.y; gr[2] = g.z; gr[3] = g.w;
        gp[0] = make_float2(a.x, a.y); gp[1] = make_float2(a.z, a.w);
        gp[2] = make_float2(b.x, b.y); gp[3] = make_float2(b.z, b.w);
    } else {
        static_assert(D == 2, "1- and 2-qubit gates");
        const float2 *M2 = reinterpret_cast<const float2 *>(M);
        const float2 g = M2[3 * r];
        gr[0] = g.x; gr[1] = g.y;
        gp[0] = M2[3 * r + 1]; gp[1] = M2[3 * r + 2];
    }
}

template <int D, int ARITH>
__device__ __forceinline__ void cmatvec(const float *__restrict__ M, const float2 (&x)[D], float2 (&y)[D]) {
    if constexpr (ARITH == 0) {
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float gr[D];
            float2 gp[D];
            load_row<D>(M, r, gr, gp);
            float yr = gr[0] * x[0].x, yi = gr[0] * x[0].y;
            yr = fmaf(gp[0].x, x[0].y, yr);
            yi = fmaf(gp[0].y, x[0].x, yi);
#pragma unroll
            for (int c = 1; c < D; ++c) {
                yr = fmaf(gr[c], x[c].x, yr);
                yr = fmaf(gp[c].x, x[c].y, yr);
                yi = fmaf(gr[c], x[c].y, yi);
                yi = fmaf(gp[c].y, x[c].x, yi);
            }
            y[r] = make_float2(yr, yi);
        }
    } else {
        f32x2_t X[D], Xs[D];
#pragma unroll
        for (int c = 0; c < D; ++c) { X[c] = pack2(x[c].x, x[c].y); Xs[c] = pack2(x[c].y, x[c].x); }
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float gr[D];
            float2 gp[D];
            load_row<D>(M, r, gr, gp);
            f32x2_t acc = fmul2(bcast2(gr[0]), X[0]);
            acc = ffma2(pack2(gp[0].x, gp[0].y), Xs[0], acc);
#pragma unroll
            for (int c = 1; c < D; ++c) {
                acc = ffma2(bcast2(gr[c]), X[c], acc);
                acc = ffma2(pack2(gp[c].x, gp[c].y), Xs[c], acc);
            }
            y[r] = unpack2(acc);
        }
    }
}

// 2-qubit gate on cluster bits I < J (matrix index bit 0 <-> I, bit 1 <-> J)
template <int I, int J, int ARITH>
__device__ __forceinline__ void reg_gate2(float2 (&v)[16], const float *__restrict__ M) {
    static_assert(I < J && J < 4, "cluster bits");
    constexpr int OTHERS = 0xF & ~((1 << I) | (1 << J));
    constexpr int O0 = (OTHERS & 1) ? 0 : (OTHERS & 2) ? 1 : (OTHERS & 4) ? 2 : 3;          // lowest other bit
    constexpr int O1 = (OTHERS & 8) ? 3 : (OTHERS & 4) ? 2 : (OTHERS & 2) ? 1 : 0;          // highest other bit
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
        const int base = ((gi & 1) << O0) | ((gi >> 1) << O1);
        float2 x[4], y[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = v[base | ((c & 1) << I) | ((c >> 1) << J)];
        cmatvec<4, ARITH>(M, x, y);
#pragma unroll
        for (int c = 0; c < 4; ++c) v[base | ((c & 1) << I) | ((c >> 1) << J)] = y[c];
    }
}

// 1-qubit gate on cluster bit I
template <int I, int ARITH>
__device__ __forceinline__ void reg_gate1(float2 (&v)[16], const float *__restrict__ M) {
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) {
        const int lo = gi & ((1 << I) - 1);
        const int base = ((gi >> I) << (I + 1)) | lo;
        float2 x[2], y[2];
        x[0] = v[base];
        x[1] = v[base | (1 << I)];
        cmatvec<2, ARITH>(M, x, y);
        v[base] = y[0];
        v[base | (1 << I)] = y[1];
    }
}

template <int ARITH>
__device__ __forceinline__ void reg_gate_dispatch(float2 (&v)[16], int type, const float *__restrict__ M) {
    switch (type) {
        case 0: reg_gate2<0, 1, ARITH>(v, M); break;
        case 1: reg_gate2<0, 2, ARITH>(v, M); break;
        case 2: reg_gate2<0, 3, ARITH>(v, M); break;
        case 3: reg_gate2<1, 2, ARITH>(v, M); break;
        case 4: reg_gate2<1, 3, ARITH>(v, M); break;
        case 5: reg_gate2<2, 3, ARITH>(v, M); break;
        case 6: reg_gate1<0, ARITH>(v, M); break;
        case 7: reg_gate1<1, ARITH>(v, M); break;
        case 8: reg_gate1<2, ARITH>(v, M); break;
        default: reg_gate1<3, ARITH>(v, M); break;
    }
}

}  // namespace ua
