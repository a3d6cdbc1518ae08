// Dense k-qubit gate application: out = U . in on arbitrary target bits, one HBM pass.
//
// Replaces the reference's permute -> contiguous -> einsum(bmm) -> permute -> contiguous
// pipeline (src/unitair/simulation/operations.py:151-186, 258-329, 626-654).
//
// gate_direct_kernel: every thread owns `U` items; an item is the 2^KH 16-byte vectors
// (2 complex64 or 1 complex128 each) that differ only in the target bits, so one item
// holds complete 2^K-amplitude groups in registers.  Loads are issued before the gate
// matrix is fetched, the (bit-order permuted, optionally adjoint) matrix sits in shared
// memory, and results are streamed out with evict-first stores.  A target on index bit
// 0 of a complex64 state lives inside the float4 (LOW = true): the pair is mixed in
// registers and no extra vector is loaded.
//
// Algorithmic traffic: 16 B (complex64) / 32 B (complex128) per amplitude (read + write).
#include "ua_common.cuh"

namespace ua {

int launch_gate_tc5(void *out, const void *in, const void *gate, int total_bits, long long batch, const int *spos,
                    const int *gbit, int adjoint, cudaStream_t st);      // ua_tc5.cu

struct GateArgs {
    const void *in;
    void *out;
    const void *gate;
    long long items_per_seg;    // items in one segment
    long long seg_in_stride;    // 16-byte vectors between segments of `in` (0 = broadcast)
    long long seg_out_stride;   // 16-byte vectors between segments of `out`
    long long gate_seg_stride;  // complex elements between per-segment gates (0 = shared)
    unsigned blocks_per_seg;
    int vpos[UA_MAX_GATE_QUBITS];  // ascending target positions in vector-index space
    int gbit[UA_MAX_GATE_QUBITS];  // gate-index bit of the i-th register-order target bit
    int adjoint;
};

// PAIR: the lowest vector-level target is vector bit 0, so members v and v|1 of an item are
// adjacent in memory and move as one 32-byte access (otherwise a warp's 16-byte accesses
// would be strided by 32 bytes and use half of every sector per instruction).
template <typename R, int K, bool LOW, int U, bool STREAM, bool PAIR>
__global__ void __launch_bounds__(256) gate_direct_kernel(const GateArgs a) {
    using C = typename CplxOf<R>::type;
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    static_assert(!LOW || APV == 2, "an in-vector target exists only for complex64");
    constexpr int KH = LOW ? K - 1 : K;   // targets that select WHICH vector
    constexpr int NV = 1 << KH;           // vectors per item
    constexpr int D = 1 << K;             // gate dimension
    __shared__ C sU[D * D];

    const unsigned bps = a.blocks_per_seg;
    const long long seg = blockIdx.x / bps;
    const unsigned chunk = blockIdx.x - (unsigned)seg * bps;

    const V *__restrict__ in = reinterpret_cast<const V *>(a.in) + seg * a.seg_in_stride;
    V *out = reinterpret_cast<V *>(a.out) + seg * a.seg_out_stride;

    uint64_t off[KH > 0 ? KH : 1];
#pragma unroll
    for (int i = 0; i < KH; ++i) off[i] = 1ull << a.vpos[i];

    // ---- 1. issue all state loads first (independent of the gate) -------------------
    V x[U][NV];
    uint64_t base[U];
    bool valid[U];
    const long long item0 = (long long)chunk * (256 * U) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long item = item0 + (long long)u * 256;
        valid[u] = item < a.items_per_seg;
        uint64_t b = (uint64_t)item;
#pragma unroll
        for (int i = 0; i < KH; ++i) b = insert_zero(b, a.vpos[i]);
        base[u] = b;
        if (valid[u]) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                uint64_t idx = b;
#pragma unroll
                for (int i = 0; i < KH; ++i)
                    if ((v >> i) & 1) idx |= off[i];
                if constexpr (PAIR) {
                    if ((v & 1) == 0) ld32<STREAM>(in + idx, x[u][v], x[u][v | 1]);
                } else {
                    x[u][v] = ld16<STREAM>(in + idx);
                }
            }
        }
    }

    // ---- 2. gate -> shared memory in register order (and adjoint if asked) -----------
    {
        const C *__restrict__ G = reinterpret_cast<const C *>(a.gate) + seg * a.gate_seg_stride;
        for (int e = threadIdx.x; e < D * D; e += 256) {
            const int s = e >> K, t = e & (D - 1);
            int gi = 0, gj = 0;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                gi |= ((s >> i) & 1) << a.gbit[i];
                gj |= ((t >> i) & 1) << a.gbit[i];
            }
            C val;
            if (a.adjoint) val = cconj(G[gj * D + gi]);
            else val = G[gi * D + gj];
            sU[e] = val;
        }
    }
    __syncthreads();

    // small gates live in registers
    constexpr bool UREG = (K <= 2);
    C ur[UREG ? D * D : 1];
    if (UREG) {
#pragma unroll
        for (int e = 0; e < D * D; ++e) ur[e] = sU[e];
    }
    auto Uat = [&](int s, int t) -> C { return UREG ? ur[s * D + t] : sU[s * D + t]; };

    // ---- 3. multiply and stream out -------------------------------------------------
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!valid[u]) continue;
        V prev;
        // K >= 4: keep the row loop rolled, the fully unrolled body (8192 FMAs for K = 5)
        // overflows the instruction cache (ncu: stall_no_instruction dominated)
#pragma unroll(K >= 4 ? ((K == 5 && sizeof(R) == 8) ? 1 : 2) : NV)
        for (int ov = 0; ov < NV; ++ov) {
            V res;
            if constexpr (APV == 2 && K >= 3) {
                // complex64, FMA-bound sizes: packed FFMA2 (P = sum gr*(xr,xi), Q = sum gi*(xr,xi),
                // result (P.x - Q.y, P.y + Q.x)); the matrix scalar is broadcast by the instruction
                f32x2_t Pa = pack2(0.f, 0.f), Qa = Pa, Pb = Pa, Qb = Pa;
                if constexpr (LOW) {
#pragma unroll
                    for (int t = 0; t < D; ++t) {
                        const V xv = x[u][t >> 1];
                        const f32x2_t X = (t & 1) ? pack2(xv.z, xv.w) : pack2(xv.x, xv.y);
                        const C ga = Uat(2 * ov, t), gb = Uat(2 * ov + 1, t);
                        Pa = ffma2(pack2(ga.x, ga.x), X, Pa);
                        Qa = ffma2(pack2(ga.y, ga.y), X, Qa);
                        Pb = ffma2(pack2(gb.x, gb.x), X, Pb);
                        Qb = ffma2(pack2(gb.y, gb.y), X, Qb);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < D; ++t) {
                        const C g = Uat(ov, t);
                        const f32x2_t gr = pack2(g.x, g.x), gi = pack2(g.y, g.y);
                        const f32x2_t X0 = pack2(x[u][t].x, x[u][t].y), X1 = pack2(x[u][t].z, x[u][t].w);
                        Pa = ffma2(gr, X0, Pa);
                        Qa = ffma2(gi, X0, Qa);
                        Pb = ffma2(gr, X1, Pb);
                        Qb = ffma2(gi, X1, Qb);
                    }
                }
                const float2 pa = unpack2(Pa), qa = unpack2(Qa), pb = unpack2(Pb), qb = unpack2(Qb);
                res = make_float4(pa.x - qa.y, pa.y + qa.x, pb.x - qb.y, pb.y + qb.x);
            } else if constexpr (APV == 1) {
                C acc = mk(R(0), R(0));
#pragma unroll
                for (int t = 0; t < D; ++t) cfma(acc, Uat(ov, t), x[u][t]);
                res = acc;
            } else if constexpr (LOW) {
                C acc0 = mk(R(0), R(0)), acc1 = mk(R(0), R(0));
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const V xv = x[u][t >> 1];
                    const C amp = (t & 1) ? mk(xv.z, xv.w) : mk(xv.x, xv.y);
                    cfma(acc0, Uat(2 * ov, t), amp);
                    cfma(acc1, Uat(2 * ov + 1, t), amp);
                }
                res = make_float4(acc0.x, acc0.y, acc1.x, acc1.y);
            } else {
                C acc0 = mk(R(0), R(0)), acc1 = mk(R(0), R(0));
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const C g = Uat(ov, t);
                    cfma(acc0, g, mk(x[u][t].x, x[u][t].y));
                    cfma(acc1, g, mk(x[u][t].z, x[u][t].w));
                }
                res = make_float4(acc0.x, acc0.y, acc1.x, acc1.y);
            }
            uint64_t idx = base[u];
#pragma unroll
            for (int i = 0; i < KH; ++i)
                if ((ov >> i) & 1) idx |= off[i];
            if constexpr (PAIR) {
                if (ov & 1) st32<STREAM>(out + (idx ^ 1ull), prev, res);
                else prev = res;
            } else {
                st16<STREAM>(out + idx, res);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// complex128, k = 4 / 5: FP64 tensor cores (DMMA, mma.sync m8n8k4).  A dense 5-qubit block is a
// genuine GEMM, Y(64 x N) = W(64 x 64) X(64 x N) in real arithmetic (W = [[Ur,-Ui],[Ui,Ur]]
// interleaved), and on B200 the DFMA version is bound by the shared-memory broadcast of the
// matrix (1 LDS.128 per 4 DFMA) at 18 TFLOP/s; DMMA needs one 8-byte fragment load per 256 MACs
// and peaks at 37 TFLOP/s (tools/micro/dfma_rate.cu).  One warp owns 8 groups ("columns") at a
// time: B fragments come straight from global memory (lane 4n+r holds real row 4j+r of group n),
// A fragments stream from shared memory in fragment order, D fragments are paired across lanes
// (shfl.xor 4) into 16-byte amplitudes and streamed out.
struct DmmaArgs {
    const void *in;
    void *out;
    const void *gate;
    long long num_batches;         // batches of 8 groups
    int spos[UA_MAX_GATE_QUBITS];  // ascending target bit positions (amplitude index)
    int gbit[UA_MAX_GATE_QUBITS];
    int adjoint;
};

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int K>
__global__ void __launch_bounds__(256) gate_dmma_kernel(const DmmaArgs a) {
    constexpr int D = 1 << K;        // complex dimension
    constexpr int MT = 2 * D / 8;    // 8-row tiles of the real matrix
    constexpr int KS = 2 * D / 4;    // k-steps of 4 real rows
    __shared__ double sW[MT * KS * 32];   // A fragments, fragment order
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // ---- matrix -> real fragments ------------------------------------------------------
    {
        const double2 *__restrict__ G = reinterpret_cast<const double2 *>(a.gate);
        for (int e = threadIdx.x; e < MT * KS * 32; e += 256) {
            const int l = e & 31, j = (e >> 5) % KS, i = (e >> 5) / KS;
            const int row = 8 * i + (l >> 2), col = 4 * j + (l & 3);
            const int s = row >> 1, cr = row & 1, t = col >> 1, cc = col & 1;
            int gi = 0, gj = 0;
#pragma unroll
            for (int b = 0; b < K; ++b) {
                gi |= ((s >> b) & 1) << a.gbit[b];
                gj |= ((t >> b) & 1) << a.gbit[b];
            }
            double2 u;
            if (a.adjoint) { u = G[gj * D + gi]; u.y = -u.y; }
            else u = G[gi * D + gj];
            sW[e] = (cr == cc) ? u.x : (cr == 0 ? -u.y : u.y);
        }
    }
    __syncthreads();

    uint64_t off[K];
#pragma unroll
    for (int b = 0; b < K; ++b) off[b] = 1ull << a.spos[b];
    const double *__restrict__ in = reinterpret_cast<const double *>(a.in);
    double2 *out = reinterpret_cast<double2 *>(a.out);
    const int n_load = lane >> 2, r = lane & 3;             // B fragment: group n_load, real row 4j + r
    const int comp = (lane >> 2) & 1;                        // D fragment: real row parity
    const int n_store = 2 * (lane & 3) + comp;               // group this lane stores after pairing
    const long long warps_total = (long long)gridDim.x * 8;
    for (long long bt = (long long)blockIdx.x * 8 + warp; bt < a.num_batches; bt += warps_total) {
        uint64_t base_l = (uint64_t)(bt * 8 + n_load), base_s = (uint64_t)(bt * 8 + n_store);
#pragma unroll
        for (int b = 0; b < K; ++b) { base_l = insert_zero(base_l, a.spos[b]); base_s = insert_zero(base_s, a.spos[b]); }
        // ---- B fragments: this lane's component (r & 1) of amplitudes 2j + (r >> 1) ------------
        double xb[KS];
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            const int t = 2 * j + (r >> 1);
            uint64_t idx = base_l;
#pragma unroll
            for (int b = 0; b < K; ++b)
                if ((t >> b) & 1) idx |= off[b];
            xb[j] = __ldcs(in + 2 * idx + (r & 1));
        }
        // ---- 8 row tiles x KS k-steps ------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int j = 0; j < KS; ++j) dmma_m8n8k4(d0, d1, sW[(i * KS + j) * 32 + lane], xb[j]);
            // lane holds real row 8i + (lane>>2) of groups 2(lane&3), 2(lane&3)+1; the lane 4 above
            // or below holds the other component of the same amplitude
            const double send = comp ? d0 : d1;
            const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
            const double2 amp = comp ? make_double2(recv, d1) : make_double2(d0, recv);
            const int s = 4 * i + (lane >> 3);
            uint64_t idx = base_s;
#pragma unroll
            for (int b = 0; b < K; ++b)
                if ((s >> b) & 1) idx |= off[b];
            __stcs(out + idx, amp);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Generic out-of-place kernel for 5 < k <= 10: one thread per output amplitude.  Slow
// (re-reads through L1/L2) but complete; the reference accepts any k.
struct GenericArgs {
    const void *in;
    void *out;
    const void *gate;
    long long total;           // batch * 2^n output amplitudes
    long long in_batch_stride; // amplitudes
    long long gate_batch_stride;
    int n, k;
    int pos[UA_MAX_GENERIC_GATE_QUBITS];  // bit position of qubits[j]
    int adjoint;
};

template <typename R>
__global__ void __launch_bounds__(256) gate_generic_kernel(const GenericArgs a) {
    using C = typename CplxOf<R>::type;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.total) return;
    const long long b = i >> a.n;
    const uint64_t e = (uint64_t)i & ((1ull << a.n) - 1ull);
    const int D = 1 << a.k;
    uint64_t mask = 0;
    int row = 0;
    for (int j = 0; j < a.k; ++j) {
        mask |= 1ull << a.pos[j];
        row |= (int)((e >> a.pos[j]) & 1ull) << (a.k - 1 - j);
    }
    const uint64_t ebase = e & ~mask;
    const C *__restrict__ in = reinterpret_cast<const C *>(a.in) + b * a.in_batch_stride;
    const C *__restrict__ G = reinterpret_cast<const C *>(a.gate) + b * a.gate_batch_stride;
    C acc = mk(R(0), R(0));
    for (int t = 0; t < D; ++t) {
        uint64_t idx = ebase;
        for (int j = 0; j < a.k; ++j)
            idx |= (uint64_t)((t >> (a.k - 1 - j)) & 1) << a.pos[j];
        const C g = a.adjoint ? cconj(G[t * D + row]) : G[row * D + t];
        cfma(acc, g, in[idx]);
    }
    reinterpret_cast<C *>(a.out)[i] = acc;
}

// ---------------------------------------------------------------------------------------
static int cache_policy() {
    static int policy = -1;
    if (policy < 0) {
        const char *e = getenv("UA_CACHE_POLICY");  // 0 auto, 1 always stream, 2 never
        policy = e ? atoi(e) : 0;
    }
    return policy;
}

struct LaunchOpts { bool low, stream_hint, pair; };

template <typename R, int K, bool LOW, int U>
static int launch_direct(const GateArgs &a, unsigned grid, const LaunchOpts &o, cudaStream_t st) {
    constexpr int KH = LOW ? K - 1 : K;
    if constexpr (KH >= 1) {
        if (o.pair) {
            if (o.stream_hint) gate_direct_kernel<R, K, LOW, U, true, true><<<grid, 256, 0, st>>>(a);
            else gate_direct_kernel<R, K, LOW, U, false, true><<<grid, 256, 0, st>>>(a);
            return check_launch("gate_direct_kernel");
        }
    }
    if (o.stream_hint) gate_direct_kernel<R, K, LOW, U, true, false><<<grid, 256, 0, st>>>(a);
    else gate_direct_kernel<R, K, LOW, U, false, false><<<grid, 256, 0, st>>>(a);
    return check_launch("gate_direct_kernel");
}

template <typename R, int K, int U>
static int launch_direct_low(const GateArgs &a, unsigned grid, const LaunchOpts &o, cudaStream_t st) {
    if constexpr (VecOf<R>::APV == 2) {
        if (o.low) return launch_direct<R, K, true, U>(a, grid, o, st);
    }
    return launch_direct<R, K, false, U>(a, grid, o, st);
}

template <int K> struct UnrollFor { static constexpr int value = K == 1 ? 4 : (K <= 3 ? 2 : 1); };

template <typename R>
static int dispatch_direct(int k, const GateArgs &a, unsigned grid, const LaunchOpts &o, cudaStream_t st) {
    switch (k) {
        case 1: return launch_direct_low<R, 1, UnrollFor<1>::value>(a, grid, o, st);
        case 2: return launch_direct_low<R, 2, UnrollFor<2>::value>(a, grid, o, st);
        case 3: return launch_direct_low<R, 3, UnrollFor<3>::value>(a, grid, o, st);
        case 4: return launch_direct_low<R, 4, UnrollFor<4>::value>(a, grid, o, st);
        case 5: return launch_direct_low<R, 5, UnrollFor<5>::value>(a, grid, o, st);
    }
    set_error("ua_apply_gate: k=%d out of range", k);
    return UA_ERR_INVALID;
}

static int unroll_for(int k) { return k == 1 ? 4 : (k <= 3 ? 2 : 1); }

}  // namespace ua

using namespace ua;

extern "C" int ua_apply_gate(int dtype, void *out, const void *in, const void *gate,
                             int num_qubits, int k, const int *host_qubits, long long batch,
                             long long in_batch_stride, long long gate_batch_stride,
                             int adjoint, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int n = num_qubits;
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("ua_apply_gate: bad dtype %d", dtype); return UA_ERR_INVALID; }
    if (!out || !in || !gate || !host_qubits) { set_error("ua_apply_gate: null pointer"); return UA_ERR_INVALID; }
    if (n < 1 || n > 48 || k < 1 || k > n) { set_error("ua_apply_gate: bad n=%d k=%d", n, k); return UA_ERR_INVALID; }
    if (k > UA_MAX_GENERIC_GATE_QUBITS) { set_error("ua_apply_gate: k=%d > %d unsupported", k, UA_MAX_GENERIC_GATE_QUBITS); return UA_ERR_UNSUPPORTED; }
    if (batch < 1) { set_error("ua_apply_gate: batch=%lld", batch); return UA_ERR_INVALID; }
    const long long dim = 1ll << n;
    const long long gdim2 = 1ll << (2 * k);
    if (in_batch_stride != 0 && in_batch_stride != dim) { set_error("ua_apply_gate: in_batch_stride must be 0 or 2^n"); return UA_ERR_INVALID; }
    if (gate_batch_stride != 0 && gate_batch_stride != gdim2) { set_error("ua_apply_gate: gate_batch_stride must be 0 or 4^k"); return UA_ERR_INVALID; }
    if (((uintptr_t)out | (uintptr_t)in) & 15) { set_error("ua_apply_gate: state pointers must be 16-byte aligned"); return UA_ERR_INVALID; }
    if (in_batch_stride == 0 && batch > 1 && out == in) { set_error("ua_apply_gate: broadcast input cannot alias output"); return UA_ERR_INVALID; }
    uint64_t seen = 0;
    int pos[UA_MAX_GENERIC_GATE_QUBITS];
    for (int j = 0; j < k; ++j) {
        const int q = host_qubits[j];
        if (q < 0 || q >= n) { set_error("ua_apply_gate: qubit %d out of range [0,%d)", q, n); return UA_ERR_INVALID; }
        if (seen & (1ull << q)) { set_error("ua_apply_gate: duplicate qubit %d", q); return UA_ERR_INVALID; }
        seen |= 1ull << q;
        pos[j] = n - 1 - q;
    }

    if (k > UA_MAX_GATE_QUBITS) {
        if (out == in) { set_error("ua_apply_gate: k>%d is out-of-place only", UA_MAX_GATE_QUBITS); return UA_ERR_UNSUPPORTED; }
        GenericArgs g;
        g.in = in; g.out = out; g.gate = gate;
        g.total = batch * dim;
        g.in_batch_stride = in_batch_stride; g.gate_batch_stride = gate_batch_stride;
        g.n = n; g.k = k; g.adjoint = adjoint;
        for (int j = 0; j < k; ++j) g.pos[j] = pos[j];
        const long long blocks = (g.total + 255) / 256;
        if (blocks > 0x7fffffffll) { set_error("ua_apply_gate: grid too large"); return UA_ERR_UNSUPPORTED; }
        if (dtype == UA_C64) gate_generic_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(g);
        else gate_generic_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(g);
        return check_launch("gate_generic_kernel");
    }

    // (the DMMA path below needs the sorted targets; it is dispatched after the sort)
    // sort targets by bit position (ascending), remember their gate-index bit
    int order[UA_MAX_GATE_QUBITS];
    for (int j = 0; j < k; ++j) order[j] = j;
    for (int i = 1; i < k; ++i)
        for (int j = i; j > 0 && pos[order[j]] < pos[order[j - 1]]; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }

    // complex128 dense 4-/5-qubit blocks with a shared gate: FP64 tensor cores
    {
        static int use_dmma = -1;
        if (use_dmma < 0) { const char *e = getenv("UA_DMMA"); use_dmma = e ? atoi(e) : 1; }
        const bool flat_c128 = dtype == UA_C128 && gate_batch_stride == 0 && (in_batch_stride == dim || batch == 1);
        if (use_dmma && flat_c128 && (k == 4 || k == 5) && n - k >= 3) {
            DmmaArgs d{};
            d.in = in; d.out = out; d.gate = gate; d.adjoint = adjoint ? 1 : 0;
            for (int i = 0; i < k; ++i) { d.spos[i] = pos[order[i]]; d.gbit[i] = k - 1 - order[i]; }
            d.num_batches = (batch * (dim >> k)) / 8;
            long long blocks = (d.num_batches + 7) / 8;
            const long long cap = (long long)sm_count() * 8;
            if (blocks > cap) blocks = cap;
            if (k == 4) gate_dmma_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(d);
            else gate_dmma_kernel<5><<<(unsigned)blocks, 256, 0, st>>>(d);
            return check_launch("gate_dmma_kernel");
        }
    }

    // complex64 dense 5-qubit blocks with a shared gate: tcgen05 tensor cores (3xTF32), ua_tc5.cu
    {
        static int use_tc = -1;
        if (use_tc < 0) { const char *e = getenv("UA_TC5"); use_tc = e ? atoi(e) : 1; }
        const bool flat_c64 = dtype == UA_C64 && gate_batch_stride == 0 && (in_batch_stride == dim || batch == 1);
        if (use_tc && flat_c64 && k == 5) {
            int sp[5], gb[5];
            for (int i = 0; i < 5; ++i) { sp[i] = pos[order[i]]; gb[i] = k - 1 - order[i]; }
            const int rc = launch_gate_tc5(out, in, gate, n, batch, sp, gb, adjoint ? 1 : 0, st);
            if (rc != UA_ERR_UNSUPPORTED) return rc;      // shapes it does not cover fall through to the CUDA cores
        }
    }

    GateArgs a;
    a.in = in; a.out = out; a.gate = gate; a.adjoint = adjoint ? 1 : 0;
    const int apv_log = (dtype == UA_C64) ? 1 : 0;       // log2(amplitudes per vector)
    const bool low = (dtype == UA_C64) && pos[order[0]] == 0;
    const int kh = low ? k - 1 : k;
    for (int i = 0; i < UA_MAX_GATE_QUBITS; ++i) { a.vpos[i] = 0; a.gbit[i] = 0; }
    for (int i = 0; i < k; ++i) a.gbit[i] = k - 1 - order[i];
    for (int i = 0; i < kh; ++i) a.vpos[i] = pos[order[i + (low ? 1 : 0)]] - apv_log;

    const long long vec_per_state = dim >> apv_log;
    const long long items_per_state = vec_per_state >> kh;
    const bool flat = (gate_batch_stride == 0) && (in_batch_stride == dim || batch == 1);
    long long segs;
    if (flat) {
        segs = 1;
        a.items_per_seg = items_per_state * batch;
        a.seg_in_stride = 0; a.seg_out_stride = 0; a.gate_seg_stride = 0;
    } else {
        segs = batch;
        a.items_per_seg = items_per_state;
        a.seg_in_stride = in_batch_stride >> apv_log;
        a.seg_out_stride = vec_per_state;
        a.gate_seg_stride = gate_batch_stride;
    }
    const int U = unroll_for(k);
    const long long bps = (a.items_per_seg + 256ll * U - 1) / (256ll * U);
    const long long grid = bps * segs;
    if (grid > 0x7fffffffll || bps > 0x7fffffffll) { set_error("ua_apply_gate: grid too large"); return UA_ERR_UNSUPPORTED; }
    a.blocks_per_seg = (unsigned)bps;

    const long long bytes = batch * dim * ((dtype == UA_C64) ? 8ll : 16ll);
    const int pol = cache_policy();
    LaunchOpts o;
    o.low = low;
    o.stream_hint = pol == 1 || (pol == 0 && bytes >= (96ll << 20));
    // 32-byte accesses when members v, v|1 are adjacent: needs 32-byte aligned segments
    const bool aligned32 = !(((uintptr_t)out | (uintptr_t)in) & 31) &&
                           (flat || ((a.seg_in_stride % 2 == 0) && (a.seg_out_stride % 2 == 0)));
    o.pair = kh >= 1 && a.vpos[0] == 0 && aligned32;
    if (dtype == UA_C64) return dispatch_direct<float>(k, a, (unsigned)grid, o, st);
    return dispatch_direct<double>(k, a, (unsigned)grid, o, st);
}
