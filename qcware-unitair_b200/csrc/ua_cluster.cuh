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

// y[r] = sum_c M[r*D + c] * x[c] for D complex inputs.
// ARITH 0: scalar FFMA (4 per complex MAC).  ARITH 1: packed FFMA2 -- with P = sum (gr,gr)*(xr,xi)
// and Q = sum (gi,gi)*(xr,xi) the product is (P.x - Q.y, P.y + Q.x): two FFMA2 per complex MAC
// plus two FADD per output; half the issue slots of the scalar form.
template <int D, int ARITH>
__device__ __forceinline__ void cmatvec(const float2 *__restrict__ M, const float2 (&x)[D], float2 (&y)[D]) {
    if constexpr (ARITH == 0) {
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float2 m = M[r * D];
            float yr = m.x * x[0].x, yi = m.x * x[0].y;
            yr = fmaf(-m.y, x[0].y, yr);
            yi = fmaf(m.y, x[0].x, yi);
#pragma unroll
            for (int c = 1; c < D; ++c) {
                m = M[r * D + c];
                yr = fmaf(m.x, x[c].x, yr);
                yr = fmaf(-m.y, x[c].y, yr);
                yi = fmaf(m.x, x[c].y, yi);
                yi = fmaf(m.y, x[c].x, yi);
            }
            y[r] = make_float2(yr, yi);
        }
    } else {
        f32x2_t X[D];
#pragma unroll
        for (int c = 0; c < D; ++c) X[c] = pack2(x[c].x, x[c].y);
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float2 m = M[r * D];
            f32x2_t P = fmul2(bcast2(m.x), X[0]);
            f32x2_t Q = fmul2(bcast2(m.y), X[0]);
#pragma unroll
            for (int c = 1; c < D; ++c) {
                m = M[r * D + c];
                P = ffma2(bcast2(m.x), X[c], P);
                Q = ffma2(bcast2(m.y), X[c], Q);
            }
            const float2 p = unpack2(P), q = unpack2(Q);
            y[r] = make_float2(p.x - q.y, p.y + q.x);
        }
    }
}

// 2-qubit gate on cluster bits I < J (matrix index bit 0 <-> I, bit 1 <-> J)
template <int I, int J, int ARITH>
__device__ __forceinline__ void reg_gate2(float2 (&v)[16], const float2 *__restrict__ M) {
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
__device__ __forceinline__ void reg_gate1(float2 (&v)[16], const float2 *__restrict__ M) {
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
__device__ __forceinline__ void reg_gate_dispatch(float2 (&v)[16], int type, const float2 *__restrict__ M) {
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
