// Gradient of a real loss with respect to a dense gate:
//     grad_U[a, b] = sum_r grad_out[a, r] * conj(psi_in[b, r])
// (a, b in the `qubits` order; r runs over the 2^(n-k) assignments of the other bits and,
// for a gate shared by the whole batch, over the batch as well).
//
// The reference has no source for this: it is what torch's tape does for the bmm at
// src/unitair/simulation/operations.py:322 (BmmBackward0 + the permute/clone backwards,
// SURVEY.md 3.4), at the price of one saved state copy and 2-3 extra passes.  Here it is one
// read of grad_out and psi_in (16 B / 32 B per amplitude), a shared-memory tile of
// 2^k x TILE_R columns of each, and a deterministic fp64 two-stage reduction.
#include "ua_common.cuh"

namespace ua {

struct GradArgs {
    const void *g;        // grad_out  [batch][2^n]
    const void *psi;      // psi_in    [batch or 1][2^n]
    void *out;            // grad_gate [nseg][4^k]
    double2 *partial;     // [nseg][bps][4^k]
    long long psi_bstride;      // amplitudes (0 = broadcast)
    long long dim;              // 2^n
    long long tiles_per_state;
    long long tiles_per_seg;    // tiles a segment has to reduce
    int bps;                    // blocks per segment
    int per_batch;              // 1: one segment per batch entry, 0: one segment for all
    int spos[UA_MAX_GATE_QUBITS];   // ascending target bit positions
    int pos[UA_MAX_GATE_QUBITS];    // bit position of qubits[j] (gate order)
    long long cols_per_state;   // 2^(n-k)
};

template <typename R> struct GradTile { static constexpr int ELEMS = sizeof(R) == 4 ? 2048 : 1024; };

template <typename R, int K>
__global__ void __launch_bounds__(256) gate_grad_kernel(const GradArgs a) {
    using C = typename CplxOf<R>::type;
    constexpr int D = 1 << K;
    constexpr int E = D * D;
    constexpr int TR = GradTile<R>::ELEMS / D;      // columns per tile
    constexpr int LD = TR + 1;                      // padded row stride
    constexpr int RL = (E >= 256) ? 1 : 256 / E;    // column lanes per entry
    constexpr int EPT = (E >= 256) ? E / 256 : 1;   // entries per thread
    __shared__ C sG[D * LD];
    __shared__ C sP[D * LD];
    __shared__ double2 sAcc[256];

    const long long seg = blockIdx.x / a.bps;
    const int blk = blockIdx.x - (int)(seg * a.bps);

    uint64_t rowoff[D];   // amplitude offset of gate-order row index a
#pragma unroll
    for (int r = 0; r < D; ++r) {
        uint64_t o = 0;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if ((r >> (K - 1 - j)) & 1) o |= 1ull << a.pos[j];
        rowoff[r] = o;
    }

    // entry / lane owned by this thread
    const int rl = threadIdx.x % RL;
    const int e0 = threadIdx.x / RL;   // EPT == 1: the entry; EPT > 1: entries e0 + j*256
    double2 acc[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) acc[j] = make_double2(0.0, 0.0);

    for (long long t = blk; t < a.tiles_per_seg; t += a.bps) {
        long long b, tt;
        if (a.per_batch) { b = seg; tt = t; }
        else { b = t / a.tiles_per_state; tt = t - b * a.tiles_per_state; }
        const C *__restrict__ g = reinterpret_cast<const C *>(a.g) + b * a.dim;
        const C *__restrict__ p = reinterpret_cast<const C *>(a.psi) + b * a.psi_bstride;
        const long long c0 = tt * TR;
        __syncthreads();   // previous tile fully consumed
#pragma unroll
        for (int r = 0; r < D; ++r) {
            for (int cc = threadIdx.x; cc < TR; cc += 256) {
                const long long col = c0 + cc;
                C gv = mk(R(0), R(0)), pv = mk(R(0), R(0));
                if (col < a.cols_per_state) {
                    uint64_t idx = (uint64_t)col;
#pragma unroll
                    for (int i = 0; i < K; ++i) idx = insert_zero(idx, a.spos[i]);
                    idx |= rowoff[r];
                    gv = __ldcs(g + idx);
                    pv = __ldcs(p + idx);
                }
                sG[r * LD + cc] = gv;
                sP[r * LD + cc] = pv;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int e = e0 + j * 256;
            const int ra = e >> K, rb = e & (D - 1);
            R sx = R(0), sy = R(0);
            for (int cc = rl; cc < TR; cc += RL) {
                const C gv = sG[ra * LD + cc];
                const C pv = sP[rb * LD + cc];
                // g * conj(p)
                sx = fma(gv.x, pv.x, sx); sx = fma(gv.y, pv.y, sx);
                sy = fma(gv.y, pv.x, sy); sy = fma(-gv.x, pv.y, sy);
            }
            acc[j].x += (double)sx;
            acc[j].y += (double)sy;
        }
    }

    // reduce the RL column lanes of every entry, then emit one partial per block
    double2 *dst = a.partial + ((long long)seg * a.bps + blk) * E;
    if (RL == 1) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) dst[e0 + j * 256] = acc[j];
    } else {
        __syncthreads();
        sAcc[threadIdx.x] = acc[0];
        __syncthreads();
        if (threadIdx.x < E) {
            double2 s = make_double2(0.0, 0.0);
            for (int l = 0; l < RL; ++l) {
                const double2 v = sAcc[threadIdx.x * RL + l];
                s.x += v.x; s.y += v.y;
            }
            dst[threadIdx.x] = s;
        }
    }
}


// Register variant for 1- and 2-qubit gates (the ones training circuits are made of).  Every
// thread walks complete amplitude groups of grad_out and psi_in with 16-byte streaming loads (a
// complex64 vector carries two independent groups unless a target is index bit 0), forms the
// D x D outer product in fp32 registers over CH groups, folds the chunk into fp64 accumulators,
// and the block reduces them with warp shuffles: no shared-memory staging, 16 B / 32 B of HBM
// traffic per amplitude at streaming rate.
template <typename R, int K, bool LOW>
__global__ void __launch_bounds__(256, 2) gate_grad_reg_kernel(const GradArgs a) {
    using C = typename CplxOf<R>::type;
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    constexpr int APVLOG = APV == 2 ? 1 : 0;
    constexpr int D = 1 << K;
    constexpr int E = D * D;
    constexpr int KH = LOW ? K - 1 : K;             // vector-level target bits
    constexpr int NV = 1 << KH;                     // vectors per group
    constexpr int NG = (APV == 2 && !LOW) ? 2 : 1;  // independent groups per vector set
    constexpr int CH = NV >= 2 ? 4 : 8;             // vector sets per fp32 chunk (>= 8 loads in flight)
    __shared__ double2 sRed[8][E];

    const long long seg = blockIdx.x / a.bps;
    const int blk = blockIdx.x - (int)(seg * a.bps);
    // vector-level positions of the targets, ascending; gate-order row of every member
    int vb[KH > 0 ? KH : 1];
#pragma unroll
    for (int i = 0; i < KH; ++i) vb[i] = a.spos[i + (LOW ? 1 : 0)] - APVLOG;
    int row_of[D];                                   // member (sorted-bit order) -> gate-order row index
#pragma unroll
    for (int m = 0; m < D; ++m) {
        int r = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            // sorted target i has bit position spos[i]; find its place j in gate order
#pragma unroll
            for (int j = 0; j < K; ++j)
                if (a.pos[j] == a.spos[i] && ((m >> i) & 1)) r |= 1 << (K - 1 - j);
        }
        row_of[m] = r;
    }
    const int nbits = 63 - __clzll((unsigned long long)a.dim);        // n
    const long long sets_per_state = a.dim >> (K + (LOW ? 0 : APVLOG));   // vector sets per state row
    const long long sets = a.per_batch ? sets_per_state : sets_per_state * a.tiles_per_seg;   // tiles_per_seg = batch here
    const V *__restrict__ gv = reinterpret_cast<const V *>(a.g);
    const V *__restrict__ pv = reinterpret_cast<const V *>(a.psi);

    double2 acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = make_double2(0.0, 0.0);

    const long long stride = (long long)a.bps * 256;
    for (long long i0 = (long long)blk * 256 + threadIdx.x; i0 < sets; i0 += stride * CH) {
        // complex128 accumulates straight into the fp64 registers (no fp32 chunk to fold)
        constexpr bool CHUNK = sizeof(R) == 4;
        R px[CHUNK ? E : 1], py[CHUNK ? E : 1];
        if constexpr (CHUNK) {
#pragma unroll
            for (int e = 0; e < E; ++e) { px[e] = R(0); py[e] = R(0); }
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const long long i = i0 + u * stride;
            if (i >= sets) break;
            long long b, s;
            if (a.per_batch) { b = seg; s = i; }
            else { b = i / sets_per_state; s = i - b * sets_per_state; }
            uint64_t base = (uint64_t)s;
#pragma unroll
            for (int j = 0; j < KH; ++j) base = insert_zero(base, vb[j]);
            const uint64_t gofs = ((uint64_t)b << (nbits - APVLOG)) + base;
            const uint64_t pofs = (uint64_t)b * (uint64_t)(a.psi_bstride >> APVLOG) + base;
            V g[NV], p[NV];
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                uint64_t o = 0;
#pragma unroll
                for (int j = 0; j < KH; ++j)
                    if ((c >> j) & 1) o |= 1ull << vb[j];
                g[c] = ld16<true>(gv + gofs + o);
                p[c] = ld16<true>(pv + pofs + o);
            }
#pragma unroll
            for (int h = 0; h < NG; ++h) {
                C gm[D], pm[D];                      // members in sorted-bit order
#pragma unroll
                for (int m = 0; m < D; ++m) {
                    if constexpr (APV == 1) { gm[m] = g[m]; pm[m] = p[m]; }
                    else if constexpr (LOW) {
                        gm[m] = (m & 1) ? mk(g[m >> 1].z, g[m >> 1].w) : mk(g[m >> 1].x, g[m >> 1].y);
                        pm[m] = (m & 1) ? mk(p[m >> 1].z, p[m >> 1].w) : mk(p[m >> 1].x, p[m >> 1].y);
                    } else {
                        gm[m] = h ? mk(g[m].z, g[m].w) : mk(g[m].x, g[m].y);
                        pm[m] = h ? mk(p[m].z, p[m].w) : mk(p[m].x, p[m].y);
                    }
                }
#pragma unroll
                for (int ma = 0; ma < D; ++ma)
#pragma unroll
                    for (int mb = 0; mb < D; ++mb) {
                        const int e = ma * D + mb;   // sorted order; mapped to gate order at the end
                        if constexpr (CHUNK) {
                            px[e] = fma(gm[ma].x, pm[mb].x, px[e]); px[e] = fma(gm[ma].y, pm[mb].y, px[e]);
                            py[e] = fma(gm[ma].y, pm[mb].x, py[e]); py[e] = fma(-gm[ma].x, pm[mb].y, py[e]);
                        } else {
                            acc[e].x = fma(gm[ma].x, pm[mb].x, acc[e].x); acc[e].x = fma(gm[ma].y, pm[mb].y, acc[e].x);
                            acc[e].y = fma(gm[ma].y, pm[mb].x, acc[e].y); acc[e].y = fma(-gm[ma].x, pm[mb].y, acc[e].y);
                        }
                    }
            }
        }
        if constexpr (CHUNK) {
#pragma unroll
            for (int e = 0; e < E; ++e) { acc[e].x += (double)px[e]; acc[e].y += (double)py[e]; }
        }
    }
    // block reduction: shuffles inside the warp, shared memory across the 8 warps
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        double x = acc[e].x, y = acc[e].y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, o);
            y += __shfl_xor_sync(0xffffffffu, y, o);
        }
        if (lane == 0) sRed[warp][e] = make_double2(x, y);
    }
    __syncthreads();
    if (threadIdx.x < E) {
        double2 t = make_double2(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < 8; ++w) { t.x += sRed[w][threadIdx.x].x; t.y += sRed[w][threadIdx.x].y; }
        const int ma = threadIdx.x / D, mb = threadIdx.x % D;
        a.partial[((long long)seg * a.bps + blk) * E + row_of[ma] * D + row_of[mb]] = t;
    }
}

template <typename R>
__global__ void __launch_bounds__(256) gate_grad_final_kernel(void *out, const double2 *partial, int bps, int E) {
    using C = typename CplxOf<R>::type;
    const long long seg = blockIdx.x;
    for (int e = threadIdx.x; e < E; e += 256) {
        double2 s = make_double2(0.0, 0.0);
        for (int b = 0; b < bps; ++b) {
            const double2 v = partial[((long long)seg * bps + b) * E + e];
            s.x += v.x; s.y += v.y;
        }
        reinterpret_cast<C *>(out)[seg * E + e] = mk((R)s.x, (R)s.y);
    }
}

struct GradPlan { long long nseg, tiles_per_state, tiles_per_seg; int bps; };

static bool grad_reg_path(int dtype, int n, int k) {
    // the register kernel needs at least one whole vector set per state row
    return k <= 2 && n >= k + (dtype == UA_C64 ? 1 : 0);
}

static GradPlan plan_grad(int dtype, int n, int k, long long batch, long long gate_bstride) {
    GradPlan p;
    if (grad_reg_path(dtype, n, k)) {
        const bool per_batch = gate_bstride != 0;
        p.nseg = per_batch ? batch : 1;
        p.tiles_per_state = 0;
        p.tiles_per_seg = batch;                       // register kernel: the batch size
        const long long sets = ((1ll << n) >> (k + (dtype == UA_C64 ? 1 : 0))) * (per_batch ? 1 : batch);
        long long want = ((long long)sm_count() * 2) / p.nseg;     // 2 CTAs per SM
        const long long most = (sets + 256 * 4 - 1) / (256 * 4);
        if (want > most) want = most;
        if (want < 1) want = 1;
        p.bps = (int)want;
        return p;
    }
    const int tile_elems = (dtype == UA_C64) ? 2048 : 1024;
    const long long tr = tile_elems >> k;
    const long long cols = 1ll << (n - k);
    p.tiles_per_state = (cols + tr - 1) / tr;
    const bool per_batch = gate_bstride != 0;
    p.nseg = per_batch ? batch : 1;
    p.tiles_per_seg = per_batch ? p.tiles_per_state : p.tiles_per_state * batch;
    long long want = ((long long)sm_count() * 4) / p.nseg;
    if (want < 1) want = 1;
    if (want > p.tiles_per_seg) want = p.tiles_per_seg;
    p.bps = (int)want;
    return p;
}

template <typename R>
static int launch_grad(int k, const GradArgs &a, unsigned grid, cudaStream_t st) {
    switch (k) {
        case 1: gate_grad_kernel<R, 1><<<grid, 256, 0, st>>>(a); break;
        case 2: gate_grad_kernel<R, 2><<<grid, 256, 0, st>>>(a); break;
        case 3: gate_grad_kernel<R, 3><<<grid, 256, 0, st>>>(a); break;
        case 4: gate_grad_kernel<R, 4><<<grid, 256, 0, st>>>(a); break;
        case 5: gate_grad_kernel<R, 5><<<grid, 256, 0, st>>>(a); break;
        default: set_error("ua_gate_grad: k=%d unsupported", k); return UA_ERR_UNSUPPORTED;
    }
    return check_launch("gate_grad_kernel");
}

}  // namespace ua

using namespace ua;

extern "C" size_t ua_gate_grad_workspace_bytes(int dtype, int num_qubits, int k, long long batch,
                                               long long gate_batch_stride) {
    if (k < 1 || k > UA_MAX_GATE_QUBITS || num_qubits < k || batch < 1) return 0;
    const GradPlan p = plan_grad(dtype, num_qubits, k, batch, gate_batch_stride);
    return (size_t)p.nseg * p.bps * (1ull << (2 * k)) * sizeof(double2);
}

extern "C" int ua_gate_grad(int dtype, void *grad_gate, const void *grad_out, const void *psi_in,
                            int num_qubits, int k, const int *host_qubits, long long batch,
                            long long psi_batch_stride, long long gate_batch_stride,
                            void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int n = num_qubits;
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("ua_gate_grad: bad dtype"); return UA_ERR_INVALID; }
    if (!grad_gate || !grad_out || !psi_in || !host_qubits || !workspace) { set_error("ua_gate_grad: null pointer"); return UA_ERR_INVALID; }
    if (n < 1 || n > 48 || k < 1 || k > n || batch < 1) { set_error("ua_gate_grad: bad sizes"); return UA_ERR_INVALID; }
    if (k > UA_MAX_GATE_QUBITS) { set_error("ua_gate_grad: k=%d > %d unsupported", k, UA_MAX_GATE_QUBITS); return UA_ERR_UNSUPPORTED; }
    const long long dim = 1ll << n;
    if (psi_batch_stride != 0 && psi_batch_stride != dim) { set_error("ua_gate_grad: psi_batch_stride must be 0 or 2^n"); return UA_ERR_INVALID; }
    if (gate_batch_stride != 0 && gate_batch_stride != (1ll << (2 * k))) { set_error("ua_gate_grad: gate_batch_stride must be 0 or 4^k"); return UA_ERR_INVALID; }
    GradArgs a{};
    uint64_t seen = 0;
    for (int j = 0; j < k; ++j) {
        const int q = host_qubits[j];
        if (q < 0 || q >= n || (seen & (1ull << q))) { set_error("ua_gate_grad: bad qubit list"); return UA_ERR_INVALID; }
        seen |= 1ull << q;
        a.pos[j] = n - 1 - q;
        a.spos[j] = a.pos[j];
    }
    for (int i = 1; i < k; ++i)
        for (int j = i; j > 0 && a.spos[j] < a.spos[j - 1]; --j) { int t = a.spos[j]; a.spos[j] = a.spos[j - 1]; a.spos[j - 1] = t; }
    const GradPlan p = plan_grad(dtype, n, k, batch, gate_batch_stride);
    const size_t need = (size_t)p.nseg * p.bps * (1ull << (2 * k)) * sizeof(double2);
    if (workspace_bytes < need) { set_error("ua_gate_grad: workspace too small (%zu < %zu)", workspace_bytes, need); return UA_ERR_INVALID; }
    a.g = grad_out; a.psi = psi_in; a.out = grad_gate; a.partial = reinterpret_cast<double2 *>(workspace);
    a.psi_bstride = psi_batch_stride; a.dim = dim;
    a.tiles_per_state = p.tiles_per_state; a.tiles_per_seg = p.tiles_per_seg; a.bps = p.bps;
    a.per_batch = gate_batch_stride != 0;
    a.cols_per_state = 1ll << (n - k);
    const long long grid = p.nseg * p.bps;
    if (grid > 0x7fffffffll) { set_error("ua_gate_grad: grid too large"); return UA_ERR_UNSUPPORTED; }
    int rc;
    if (grad_reg_path(dtype, n, k)) {
        const bool low = dtype == UA_C64 && a.spos[0] == 0;
        const bool aligned = !(((uintptr_t)grad_out | (uintptr_t)psi_in) & 15);
        if (!aligned) { set_error("ua_gate_grad: pointers must be 16-byte aligned"); return UA_ERR_INVALID; }
        const unsigned gr = (unsigned)grid;
        if (dtype == UA_C64) {
            if (k == 1 && low) gate_grad_reg_kernel<float, 1, true><<<gr, 256, 0, st>>>(a);
            else if (k == 1) gate_grad_reg_kernel<float, 1, false><<<gr, 256, 0, st>>>(a);
            else if (low) gate_grad_reg_kernel<float, 2, true><<<gr, 256, 0, st>>>(a);
            else gate_grad_reg_kernel<float, 2, false><<<gr, 256, 0, st>>>(a);
        } else {
            if (k == 1) gate_grad_reg_kernel<double, 1, false><<<gr, 256, 0, st>>>(a);
            else gate_grad_reg_kernel<double, 2, false><<<gr, 256, 0, st>>>(a);
        }
        rc = check_launch("gate_grad_reg_kernel");
    } else {
        rc = (dtype == UA_C64) ? launch_grad<float>(k, a, (unsigned)grid, st) : launch_grad<double>(k, a, (unsigned)grid, st);
    }
    if (rc) return rc;
    const int E = 1 << (2 * k);
    if (dtype == UA_C64) gate_grad_final_kernel<float><<<(unsigned)p.nseg, 256, 0, st>>>(grad_gate, a.partial, p.bps, E);
    else gate_grad_final_kernel<double><<<(unsigned)p.nseg, 256, 0, st>>>(grad_gate, a.partial, p.bps, E);
    return check_launch("gate_grad_final_kernel");
}
