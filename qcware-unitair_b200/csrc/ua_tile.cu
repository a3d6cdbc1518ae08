// Fused shared-memory pass: stage a tile of 2^T amplitudes once, apply a whole list of
// dense 1-3 qubit gates to it in shared memory, write it back.  One HBM read + write
// (16 B / 32 B per amplitude) for the entire list instead of one per gate.
//
// A tile is the set of amplitudes that agree on every index bit outside the tile's T bit
// positions: the low L bits (contiguous in memory -> 2^L * 8/16 B runs) plus H chosen
// higher positions.  Gates whose target bits all lie in that set never need another pass.
// This is the engine behind apply_all_qubits (src/unitair/simulation/operations.py:332-413,
// n strided einsum passes in the reference) and behind the circuit API (the reference's own
// fusion idea, apply_to_qubits, operations.py:416-503, generalised from same-qubit 2x2
// products to whole gate lists).
//
// Kernel structure: persistent CTAs (one per SM), NSTAGE tile buffers in shared memory.
//   * Tiles move with the bulk asynchronous copy engine (TMA, cp.async.bulk): warp 0 issues
//     one bulk copy per contiguous run (2^H runs of 2^L amplitudes), completion is tracked
//     by an mbarrier (global -> shared) or a bulk group (shared -> global).  No registers
//     are used for staging, and the load of tile i+1 and the store of tile i-1 overlap the
//     gate phase of tile i.
//   * Gate phase: each thread takes groups of 2^KH 16-byte vectors (a complex64 target on
//     local bit 0 lives inside the float4), multiplies by the gate (matrix in shared memory
//     in register order) and writes the group back, LDS.128/STS.128 throughout.  When a
//     target sits on one of the three lowest vector bits the 8 lanes of a shared-memory
//     phase would hit only half/quarter of the banks; those gates take the SWZ path where
//     lanes read the group members in a lane-dependent order (XOR on the member index) and
//     undo it with register selects: conflict-free for every target position.
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "ua_common.cuh"
#include "ua_cluster.cuh"

namespace ua {

struct FusedGate {
    long long goff;          // offset of the matrix in `mats` (complex elements)
    unsigned short smoff;    // offset of the register-order copy in shared memory
    unsigned char k;
    unsigned char sb[3];     // ascending tile-local target bits
    unsigned char gb[3];     // gate-index bit of sb[i]
    unsigned char pad[5];
};

struct FusedArgs {
    const void *in;
    void *out;
    const void *mats;
    long long mats_row_stride;   // complex elements between rows' matrix sets (0 = shared)
    long long num_tiles;
    long long tiles_per_row;
    int total_bits, T, L, H;
    int high[UA_MAX_TILE_BITS];  // ascending global positions of tile-local bits L..T-1
    int num_gates;
    int adjoint;
    int nstage;
    // TMA tensor path: the state seen as a rank-`trank` tensor of 8-byte elements whose
    // dimension j spans element-index bits [tstart[j], tstart[j+1]); a tile is the box made
    // of the low bits of every dimension, moved by ONE cp.async.bulk.tensor instruction.
    int swizzle;                 // 1: bank-conflict-free member swizzle for low targets
    int use_f2;                  // 1: complex64 gate phase with packed FFMA2
    int l2_prefetch;             // 1: L2-prefetch the tile this CTA will load next
    unsigned long long *trace;   // debug: per-CTA phase timestamps (globaltimer ns), or null
    int stagger_ns;              // start delay per co-resident CTA index (breaks lockstep)
    int num_sms;
    int trank;                   // 0 = tensor path off (per-run bulk copies instead)
    int tstart[6];
    // Scatter store (global-qubit exchange folded into the pass, ua_apply_fused_pass_scatter):
    // scatter_m index bits vpos[] (ascending, none of them a tile bit) are removed from the
    // output index; their values select one of 2^m destination buffers (peer GPUs' memory
    // mapped into this process).  tstart_out = tstart in the compressed index.
    int scatter_m;
    unsigned long long tile_xor; // flips scatter bits of every tile's base: rank-dependent visiting order
    int nins;                    // scatter pass: tile counter -> base inserts zeros at ins[] (tile high bits and
    int ins[UA_MAX_TILE_BITS + UA_MAX_SCATTER_BITS];   // scatter bits, ascending); its low m bits are the scatter bits
    int vpos[UA_MAX_SCATTER_BITS];
    int tstart_out[6];
    void *dst[1 << UA_MAX_SCATTER_BITS];
    alignas(64) CUtensorMap tmap_in;
    alignas(64) CUtensorMap tmap_out;
    alignas(64) CUtensorMap tmap_dst[1 << UA_MAX_SCATTER_BITS];
    FusedGate gates[UA_MAX_FUSED_GATES];
};

constexpr int FUSED_MAX_MAT_ELEMS = 2048;   // complex elements of gate matrices per pass

// ---------------------------------------------------------------------------------------
// Register-blocked gate phase ("cluster" path, complex64, shared 1-/2-qubit gates whose matrix
// VALUES are known on the host).  The pass's gate list is cut into clusters: runs of gates whose
// target bits all lie inside a set of FOUR tile bits.  A thread loads the 16 amplitudes of one
// group (all values of the 4 cluster bits) from the tile once, applies every gate of the cluster
// to them in registers and stores them back: one shared-memory round trip and one barrier per
// CLUSTER instead of per gate.  The matrices travel in the kernel parameters (constant bank):
// the gate index is warp-uniform, so ptxas keeps them in UNIFORM registers (LDCU) and the FMAs
// take them as UR operands -- no vector registers and no shared-memory traffic for the matrix,
// and FFMA with a UR operand issues at full rate (3-vector-register FFMA does not:
// tools/micro/gate_reg_rate.cu, 64 vs 50 TFLOP/s).
constexpr int CL_BITS = 4;
constexpr int CL_TAB = 8;            // (cluster, sweep) slots of the per-thread group-offset table
constexpr int CL_MAX_GATES = UA_MAX_FUSED_GATES;
constexpr int CL_MAX_MAT_ELEMS = 16 * CL_MAX_GATES;

struct ClusterDesc {
    unsigned char cb[CL_BITS];   // ascending tile-local bit positions
    unsigned char gbeg, gend;    // gates [gbeg, gend) of ClusterArgs::g
    unsigned char vec16;         // cb[0] == 0: members 2m, 2m+1 are one 16-byte vector
    unsigned char pad;
    // bit k of a thread's group number lands on tile bit fb[k] (the non-cluster bits, ordered so
    // that the lanes of one shared-memory wavefront fall into distinct banks)
    unsigned char fb[UA_MAX_TILE_BITS - CL_BITS + 2];
    // member m sits at byte (group base ^ po8[m]) of the tile buffer (swizzle applied)
    unsigned short po8[1 << CL_BITS];
};
struct ClusterGate {
    unsigned short moff;         // offset of the matrix in ClusterArgs::mats (target-bit order, adjoint applied)
    unsigned short type;         // 0..5: 2-qubit gate on cluster bits (0,1) (0,2) (0,3) (1,2) (1,3) (2,3); 6..9: 1-qubit on bit type-6
};
struct ClusterArgs {
    int ncl;
    ClusterDesc cl[CL_MAX_GATES];
    ClusterGate g[CL_MAX_GATES];
    float2 mats[CL_MAX_MAT_ELEMS];
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void *src_gmem, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, unsigned src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load(int rank, unsigned dst, const CUtensorMap *tm, const int *c, unsigned bar) {
    const unsigned long long t = reinterpret_cast<unsigned long long>(tm);
    switch (rank) {
        case 1: asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];"
                             ::"r"(dst), "l"(t), "r"(c[0]), "r"(bar) : "memory"); break;
        case 2: asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(dst), "l"(t), "r"(c[0]), "r"(c[1]), "r"(bar) : "memory"); break;
        case 3: asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(dst), "l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(bar) : "memory"); break;
        case 4: asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                             ::"r"(dst), "l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(bar) : "memory"); break;
        default: asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                             ::"r"(dst), "l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(bar) : "memory"); break;
    }
}
__device__ __forceinline__ void tma_prefetch_l2(int rank, const CUtensorMap *tm, const int *c) {
    const unsigned long long t = reinterpret_cast<unsigned long long>(tm);
    switch (rank) {
        case 1: asm volatile("cp.async.bulk.prefetch.tensor.1d.L2.global.tile [%0, {%1}];" ::"l"(t), "r"(c[0]) : "memory"); break;
        case 2: asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(t), "r"(c[0]), "r"(c[1]) : "memory"); break;
        case 3: asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory"); break;
        case 4: asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory"); break;
        default: asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory"); break;
    }
}
__device__ __forceinline__ void tma_store(int rank, const CUtensorMap *tm, const int *c, unsigned src) {
    const unsigned long long t = reinterpret_cast<unsigned long long>(tm);
    switch (rank) {
        case 1: asm volatile("cp.async.bulk.tensor.1d.global.shared::cta.tile.bulk_group [%0, {%1}], [%2];"
                             ::"l"(t), "r"(c[0]), "r"(src) : "memory"); break;
        case 2: asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(src) : "memory"); break;
        case 3: asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                             ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(src) : "memory"); break;
        case 4: asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                             ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(src) : "memory"); break;
        default: asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                             ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(src) : "memory"); break;
    }
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_but_one() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr int TRACE_TILES = 16;     // tiles traced per CTA, 4 timestamps each

__device__ __forceinline__ unsigned insert_zero32(unsigned x, int p) {
    const unsigned lo = x & ((1u << p) - 1u);
    return ((x >> p) << (p + 1)) | lo;
}

template <typename V> __device__ __forceinline__ void cond_swap(V &a, V &b, bool doit) {
    const V ta = a, tb = b;
    a = doit ? tb : ta;
    b = doit ? ta : tb;
}

// One gate on the whole tile.  K qubits; LOW: (complex64 only) the lowest target is local
// bit 0, i.e. inside the float4; SWZ: some vector-level target is among the 3 lowest vector
// bits (bank-conflict avoiding member swizzle on).  `tv` is the tile as 16-byte vectors,
// TV = log2 of their count.  GU groups are processed together for memory-level parallelism.
template <typename R, int K, bool LOW, bool SWZ, bool LEAN = false>
__device__ __forceinline__ void apply_gate_smem(typename VecOf<R>::type *tv,
                                                const typename CplxOf<R>::type *M,
                                                const FusedGate &gd, int TV, int nthreads,
                                                unsigned tid = threadIdx.x) {
    using C = typename CplxOf<R>::type;
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    constexpr int APVLOG = APV == 2 ? 1 : 0;
    constexpr int KH = LOW ? K - 1 : K;
    constexpr int NV = 1 << KH;
    constexpr int D = 1 << K;
    constexpr int INFLIGHT = sizeof(R) == 4 ? 4 : 2;  // 16-byte vectors in flight per thread
    constexpr int GU = (LEAN || K >= 2 || NV >= INFLIGHT) ? 1 : INFLIGHT / NV;   // K >= 2: the matrix fills the registers

    int vb[KH > 0 ? KH : 1];          // ascending vector-bit positions of the vector-level targets
    unsigned off[KH > 0 ? KH : 1];
    int m = 0;                        // how many of them are among the 3 lowest vector bits
#pragma unroll
    for (int i = 0; i < KH; ++i) {
        vb[i] = (int)gd.sb[i + (LOW ? 1 : 0)] - APVLOG;
        off[i] = 1u << vb[i];
        m += (vb[i] < 3) ? 1 : 0;
    }
    unsigned msk = 0;
    if (SWZ) {
        // lane-dependent member swizzle: low target i is selected by lane bit (3 - m + i)
        const unsigned lane = threadIdx.x & 31u;
#pragma unroll
        for (int i = 0; i < KH; ++i)
            if (i < m) msk |= ((lane >> (3 - m + i)) & 1u) << i;
    }

    // LEAN (3 CTAs x 256 threads per SM, <= 80 registers): complex128 2-qubit matrices stay in
    // shared memory
    constexpr bool MREG = LEAN ? (K <= (sizeof(R) == 4 ? 2 : 1)) : (K <= 2);
    C mr[MREG ? D * D : 1];
    if (MREG) {
#pragma unroll
        for (int e = 0; e < D * D; ++e) mr[e] = M[e];
    }
    auto Mat = [&](int s, int t) -> C { return MREG ? mr[s * D + t] : M[s * D + t]; };

    const unsigned groups = 1u << (TV - KH);
    for (unsigned g0 = tid; g0 < groups; g0 += nthreads * GU) {
        unsigned gbase[GU];
        V x[GU][NV];
        bool ok[GU];
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            const unsigned g = g0 + u * nthreads;
            ok[u] = g < groups;
            unsigned b = g;
#pragma unroll
            for (int i = 0; i < KH; ++i) b = insert_zero32(b, vb[i]);
            gbase[u] = b;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                const unsigned cc = SWZ ? ((unsigned)c ^ msk) : (unsigned)c;
                unsigned idx = b;
#pragma unroll
                for (int i = 0; i < KH; ++i)
                    if ((cc >> i) & 1u) idx |= off[i];
                if (ok[u]) x[u][c] = tv[idx];
            }
        }
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            if (!ok[u]) continue;
            if (SWZ) {   // undo the swizzle: x[c] currently holds member (c ^ msk)
#pragma unroll
                for (int i = 0; i < KH; ++i) {
                    const bool sw = (msk >> i) & 1u;
#pragma unroll
                    for (int c = 0; c < NV; ++c)
                        if (!((c >> i) & 1)) cond_swap(x[u][c], x[u][c | (1 << i)], sw);
                }
            }
            V y[SWZ ? NV : 1];
#pragma unroll
            for (int ov = 0; ov < NV; ++ov) {
                V res;
                if constexpr (APV == 1) {
                    C acc = mk(R(0), R(0));
#pragma unroll
                    for (int t = 0; t < D; ++t) cfma(acc, Mat(ov, t), x[u][t]);
                    res = acc;
                } else if constexpr (LOW) {
                    C acc0 = mk(R(0), R(0)), acc1 = mk(R(0), R(0));
#pragma unroll
                    for (int t = 0; t < D; ++t) {
                        const V xv = x[u][t >> 1];
                        const C amp = (t & 1) ? mk(xv.z, xv.w) : mk(xv.x, xv.y);
                        cfma(acc0, Mat(2 * ov, t), amp);
                        cfma(acc1, Mat(2 * ov + 1, t), amp);
                    }
                    res = make_float4(acc0.x, acc0.y, acc1.x, acc1.y);
                } else {
                    C acc0 = mk(R(0), R(0)), acc1 = mk(R(0), R(0));
#pragma unroll
                    for (int t = 0; t < D; ++t) {
                        const C gm = Mat(ov, t);
                        cfma(acc0, gm, mk(x[u][t].x, x[u][t].y));
                        cfma(acc1, gm, mk(x[u][t].z, x[u][t].w));
                    }
                    res = make_float4(acc0.x, acc0.y, acc1.x, acc1.y);
                }
                if constexpr (SWZ) {
                    y[ov] = res;
                } else {       // every member is already in registers: store right away
                    unsigned idx = gbase[u];
#pragma unroll
                    for (int i = 0; i < KH; ++i)
                        if ((ov >> i) & 1) idx |= off[i];
                    tv[idx] = res;
                }
            }
            if constexpr (SWZ) {   // y[c] must become member (c ^ msk) again for the store slots
#pragma unroll
                for (int i = 0; i < KH; ++i) {
                    const bool sw = (msk >> i) & 1u;
#pragma unroll
                    for (int c = 0; c < NV; ++c)
                        if (!((c >> i) & 1)) cond_swap(y[c], y[c | (1 << i)], sw);
                }
#pragma unroll
                for (int c = 0; c < NV; ++c) {
                    const unsigned cc = (unsigned)c ^ msk;
                    unsigned idx = gbase[u];
#pragma unroll
                    for (int i = 0; i < KH; ++i)
                        if ((cc >> i) & 1u) idx |= off[i];
                    tv[idx] = y[c];
                }
            }
        }
    }
}

// complex64 gate on the tile with packed FFMA2.  With P = sum (gr,gr)*(xr,xi) and
// Q = sum (gi,gi)*(xr,xi) the product is (P.x - Q.y, P.y + Q.x): two FFMA2 per complex MAC, the
// amplitude pair comes straight from the 16-byte load and the matrix scalar is broadcast by the
// instruction itself, so there are no operand shuffles.
__device__ __forceinline__ float2 combine_pq(f32x2_t P, f32x2_t Q) {
    const float2 p = unpack2(P), q = unpack2(Q);
    return make_float2(p.x - q.y, p.y + q.x);
}

// scatter the bits of x around the (ascending) zero-insertion points vb[]: a bit permutation,
// so scatter(x | y) = scatter(x) | scatter(y) for disjoint x, y
template <int KH>
__device__ __forceinline__ unsigned scatter_bits(unsigned x, const int (&vb)[KH > 0 ? KH : 1]) {
#pragma unroll
    for (int i = 0; i < KH; ++i) {
        const unsigned himask = ~0u << vb[i];
        x += x & himask;                 // bits >= vb[i] move up by one
    }
    return x;
}

// Fast path: the tile has at least NT*GU groups (true for the production tile sizes), so there is
// no tail predicate, the thread part of every address is computed once per gate and the group
// part is warp-uniform.
template <int K, bool LOW, int NT, int VIF = 8>
__device__ __forceinline__ void apply_gate_smem_f2(ulonglong2 *tv, const float2 *M,
                                                   const FusedGate &gd, int TV) {
    constexpr int KH = LOW ? K - 1 : K;
    constexpr int NV = 1 << KH;
    constexpr int D = 1 << K;
    constexpr int GU = (NV >= VIF) ? 1 : VIF / NV;   // VIF vectors (2 VIF amplitudes) per thread in flight
    int vb[KH > 0 ? KH : 1];
    unsigned off[KH > 0 ? KH : 1];
#pragma unroll
    for (int i = 0; i < KH; ++i) {
        vb[i] = (int)gd.sb[i + (LOW ? 1 : 0)] - 1;
        off[i] = 1u << vb[i];
    }
    // the matrix stays unexpanded (re, im): FFMA2 takes a scalar register broadcast to both
    // lanes (SASS "R.F32" operand), so pack2(g, g) costs nothing
    constexpr bool MREG = (K <= 2);
    float2 mr[MREG ? D * D : 1];
    if (MREG) {
#pragma unroll
        for (int e = 0; e < D * D; ++e) mr[e] = M[e];
    }
    // !MREG (3-qubit gates): the rows needed for one output vector are fetched with 16-byte
    // broadcast loads right before they are used (D/2 LDS.128 per row)
    float2 rowa[MREG ? 1 : D], rowb[(MREG || !LOW) ? 1 : D];
    auto load_row = [&](float2 (&dst)[MREG ? 1 : D], int r) {
        if constexpr (!MREG) {
            const float4 *src = reinterpret_cast<const float4 *>(M + r * D);
#pragma unroll
            for (int q = 0; q < D / 2; ++q) {
                const float4 v = src[q];
                dst[2 * q] = make_float2(v.x, v.y);
                dst[2 * q + 1] = make_float2(v.z, v.w);
            }
        }
    };
    const f32x2_t zero = pack2(0.f, 0.f);

    const unsigned groups = 1u << (TV - KH);
    const unsigned tbase = scatter_bits<KH>(threadIdx.x, vb);          // per thread, once per gate
    for (unsigned g0 = 0; g0 < groups; g0 += NT * GU) {                 // warp-uniform
        ulonglong2 *p[GU];
        ulonglong2 x[GU][NV];
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            p[u] = tv + (tbase | scatter_bits<KH>(g0 + u * NT, vb));
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                unsigned o = 0;
#pragma unroll
                for (int i = 0; i < KH; ++i)
                    if ((c >> i) & 1) o |= off[i];
                x[u][c] = p[u][o];
            }
        }
#pragma unroll
        for (int ov = 0; ov < NV; ++ov) {
            f32x2_t Pa[GU], Qa[GU], Pb[GU], Qb[GU];
#pragma unroll
            for (int u = 0; u < GU; ++u) { Pa[u] = zero; Qa[u] = zero; Pb[u] = zero; Qb[u] = zero; }
            if constexpr (LOW) {
                // rows 2ov, 2ov+1; member t is half (t & 1) of vector t >> 1
                if constexpr (!MREG) { load_row(rowa, 2 * ov); load_row(rowb, 2 * ov + 1); }
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    float2 ga, gb;
                    if constexpr (MREG) { ga = mr[(2 * ov) * D + t]; gb = mr[(2 * ov + 1) * D + t]; }
                    else { ga = rowa[t]; gb = rowb[t]; }
                    const f32x2_t gar = pack2(ga.x, ga.x), gai = pack2(ga.y, ga.y);
                    const f32x2_t gbr = pack2(gb.x, gb.x), gbi = pack2(gb.y, gb.y);
#pragma unroll
                    for (int u = 0; u < GU; ++u) {
                        const f32x2_t X = (t & 1) ? x[u][t >> 1].y : x[u][t >> 1].x;
                        Pa[u] = ffma2(gar, X, Pa[u]);
                        Qa[u] = ffma2(gai, X, Qa[u]);
                        Pb[u] = ffma2(gbr, X, Pb[u]);
                        Qb[u] = ffma2(gbi, X, Qb[u]);
                    }
                }
            } else {
                // row ov for both amplitudes of every vector
                if constexpr (!MREG) load_row(rowa, ov);
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    float2 gm;
                    if constexpr (MREG) gm = mr[ov * D + t];
                    else gm = rowa[t];
                    const f32x2_t gr = pack2(gm.x, gm.x), gi = pack2(gm.y, gm.y);
#pragma unroll
                    for (int u = 0; u < GU; ++u) {
                        Pa[u] = ffma2(gr, x[u][t].x, Pa[u]);
                        Qa[u] = ffma2(gi, x[u][t].x, Qa[u]);
                        Pb[u] = ffma2(gr, x[u][t].y, Pb[u]);
                        Qb[u] = ffma2(gi, x[u][t].y, Qb[u]);
                    }
                }
            }
            unsigned o = 0;
#pragma unroll
            for (int i = 0; i < KH; ++i)
                if ((ov >> i) & 1) o |= off[i];
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                const float2 r0 = combine_pq(Pa[u], Qa[u]), r1 = combine_pq(Pb[u], Qb[u]);
                reinterpret_cast<float4 *>(p[u])[o] = make_float4(r0.x, r0.y, r1.x, r1.y);
            }
        }
    }
}

template <typename R, int K, bool LOW, int NT, bool LEAN>
__device__ __forceinline__ void apply_gate_swz(typename VecOf<R>::type *tv,
                                               const typename CplxOf<R>::type *M,
                                               const FusedGate &gd, int TV, bool allow_swz, bool use_f2) {
    constexpr int nthreads = NT;
    constexpr int APVLOG = VecOf<R>::APV == 2 ? 1 : 0;
    constexpr int KH = LOW ? K - 1 : K;
    bool swz = false;
    if constexpr (KH > 0) swz = allow_swz && ((int)gd.sb[LOW ? 1 : 0] - APVLOG) < 3;   // lowest vector-level target
    if constexpr (LEAN) {
        // register-lean build: plain scalar path only.  (The packed-FFMA2 path was measured in
        // this build too: same pass time -- FFMA2 halves the issue slots but occupies the FMA
        // pipe for two cycles, and the gate phase is latency- not issue-bound; it costs spills
        // at 80 registers, so it is not compiled in.)
        apply_gate_smem<R, K, LOW, false, true>(tv, M, gd, TV, nthreads);
        return;
    }
    if constexpr (sizeof(R) == 4) {      // complex64: packed FFMA2 fast path on full-size tiles
        constexpr int NVf = 1 << KH;
        constexpr int GUf = (NVf >= 8) ? 1 : 8 / NVf;
        if (use_f2 && !swz && (1u << (TV - KH)) >= (unsigned)(NT * GUf)) {
            apply_gate_smem_f2<K, LOW, NT>(reinterpret_cast<ulonglong2 *>(tv),
                                           reinterpret_cast<const float2 *>(M), gd, TV);
            return;
        }
    }
    if constexpr (KH > 0 && K <= 2) {   // 3-qubit gates (rare after merging) keep the plain path
        if (swz) { apply_gate_smem<R, K, LOW, true>(tv, M, gd, TV, nthreads); return; }
    }
    apply_gate_smem<R, K, LOW, false>(tv, M, gd, TV, nthreads);
}

template <typename R, int NT, bool LEAN>
__device__ __forceinline__ void apply_any_gate(typename VecOf<R>::type *tv,
                                               const typename CplxOf<R>::type *M,
                                               const FusedGate &gd, int TV, bool swz, bool f2) {
    constexpr int APV = VecOf<R>::APV;
    if constexpr (APV == 2) {
        if (gd.sb[0] == 0) {
            if (gd.k == 1) apply_gate_swz<R, 1, true, NT, LEAN>(tv, M, gd, TV, swz, f2);
            else if (gd.k == 2) apply_gate_swz<R, 2, true, NT, LEAN>(tv, M, gd, TV, swz, f2);
            else apply_gate_swz<R, 3, true, NT, LEAN>(tv, M, gd, TV, swz, f2);
            return;
        }
    }
    if (gd.k == 1) apply_gate_swz<R, 1, false, NT, LEAN>(tv, M, gd, TV, swz, f2);
    else if (gd.k == 2) apply_gate_swz<R, 2, false, NT, LEAN>(tv, M, gd, TV, swz, f2);
    else apply_gate_swz<R, 3, false, NT, LEAN>(tv, M, gd, TV, swz, f2);
}

// Persistent kernel: CTA b processes tiles b, b + gridDim.x, ...  Shared memory layout:
// [nstage tile buffers][gate matrices]; mbarriers are static.
template <typename R, int FUSED_THREADS, int MINB, bool SCATTER = false>
__global__ void __launch_bounds__(FUSED_THREADS, MINB) fused_pass_kernel(const __grid_constant__ FusedArgs a) {
    constexpr bool LEAN = (FUSED_THREADS * MINB > 512);
    using C = typename CplxOf<R>::type;
    using V = typename VecOf<R>::type;
    constexpr int APVLOG = VecOf<R>::APV == 2 ? 1 : 0;
    constexpr int MAXSTAGE = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[MAXSTAGE];

    const unsigned tile_bytes = (1u << a.T) * (unsigned)sizeof(C);
    const int nstage = a.nstage;
    C *sM = reinterpret_cast<C *>(smem_raw + (size_t)nstage * tile_bytes);
    const int TV = a.T - APVLOG;
    const unsigned run_bytes = (1u << a.L) * (unsigned)sizeof(C);
    const unsigned runs = 1u << a.H;
    const unsigned lane = threadIdx.x & 31u;
    const bool mover = threadIdx.x < 32;      // warp 0 drives the copy engine
    const char *in = reinterpret_cast<const char *>(a.in);
    char *out = reinterpret_cast<char *>(a.out);

    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    auto tile_base = [&](long long tile_id, long long &row) -> uint64_t {
        row = tile_id / a.tiles_per_row;
        const long long j = tile_id - row * a.tiles_per_row;
        uint64_t base;
        if constexpr (SCATTER) {
            // consecutive tiles go to different destinations (the low m bits of the counter are the
            // scatter bits): every GPU writes to all its peers all the time, like a ring-less
            // all-to-all, instead of one peer after the other (incast when ranks drift apart)
            base = (uint64_t)(j >> a.scatter_m) << a.L;
            for (int i = 0; i < a.nins; ++i) base = insert_zero(base, a.ins[i]);
            for (int i = 0; i < a.scatter_m; ++i) base |= (uint64_t)((j >> i) & 1) << a.vpos[i];
            base ^= a.tile_xor;      // a bijection on tiles (only scatter bits flip)
        } else {
            base = (uint64_t)j << a.L;
            for (int i = 0; i < a.H; ++i) base = insert_zero(base, a.high[i]);
        }
        return base + ((uint64_t)row << a.total_bits);
    };
    auto run_offset = [&](unsigned r) -> uint64_t {   // amplitude offset of run r inside a tile
        uint64_t o = 0;
        for (int i = 0; i < a.H; ++i)
            if ((r >> i) & 1u) o |= 1ull << a.high[i];
        return o;
    };
    constexpr int EBITS = sizeof(C) == 8 ? 0 : 1;    // 8-byte TMA elements per amplitude (log2)
    auto tensor_coords = [&](uint64_t base, int *c) {
        const uint64_t e = base << EBITS;
        for (int j = 0; j < a.trank; ++j) {
            uint64_t v = e >> a.tstart[j];
            if (j + 1 < a.trank) v &= (1ull << (a.tstart[j + 1] - a.tstart[j])) - 1ull;
            c[j] = (int)v;
        }
    };
    auto issue_load = [&](long long tile_id, int s) {   // warp 0 only
        long long row;
        const uint64_t base = tile_base(tile_id, row);
        const unsigned bar = smem_u32(&bars[s]);
        if (lane == 0) mbar_arrive_expect_tx(bar, tile_bytes);
        __syncwarp();
        const unsigned dst = smem_u32(smem_raw + (size_t)s * tile_bytes);
        if (a.trank > 0) {
            if (lane == 0) {
                int c[5];
                tensor_coords(base, c);
                tma_load(a.trank, dst, &a.tmap_in, c, bar);
            }
        } else {
            for (unsigned r = lane; r < runs; r += 32)
                bulk_g2s(dst + r * run_bytes, in + (base + run_offset(r)) * sizeof(C), run_bytes, bar);
        }
    };
    auto compress = [&](uint64_t x) -> uint64_t {      // drop the scatter bits from an index
        for (int j = a.scatter_m - 1; j >= 0; --j) {
            const int v = a.vpos[j];
            x = ((x >> (v + 1)) << v) | (x & ((1ull << v) - 1ull));
        }
        return x;
    };
    auto issue_store = [&](long long tile_id, int s) {  // warp 0 only
        long long row;
        const uint64_t base = tile_base(tile_id, row);
        const unsigned src = smem_u32(smem_raw + (size_t)s * tile_bytes);
        if constexpr (SCATTER) {
            // the tile goes to ONE destination buffer (no scatter bit is a tile bit), at the
            // position its index has once the scatter bits are squeezed out
            unsigned b = 0;
            for (int j = 0; j < a.scatter_m; ++j) b |= (unsigned)((base >> a.vpos[j]) & 1ull) << j;
            if (a.trank > 0) {
                if (lane == 0) {
                    const uint64_t e = compress(base) << EBITS;
                    int c[5];
                    for (int j = 0; j < a.trank; ++j) {
                        uint64_t v = e >> a.tstart_out[j];
                        if (j + 1 < a.trank) v &= (1ull << (a.tstart_out[j + 1] - a.tstart_out[j])) - 1ull;
                        c[j] = (int)v;
                    }
                    tma_store(a.trank, &a.tmap_dst[b], c, src);
                }
            } else {
                char *dstp = reinterpret_cast<char *>(a.dst[b]);
                for (unsigned r = lane; r < runs; r += 32)
                    bulk_s2g(dstp + compress(base + run_offset(r)) * sizeof(C), src + r * run_bytes, run_bytes);
            }
            bulk_commit();
            return;
        }
        if (a.trank > 0) {
            if (lane == 0) {
                int c[5];
                tensor_coords(base, c);
                tma_store(a.trank, &a.tmap_out, c, src);
            }
        } else {
            for (unsigned r = lane; r < runs; r += 32)
                bulk_s2g(out + (base + run_offset(r)) * sizeof(C), src + r * run_bytes, run_bytes);
        }
        bulk_commit();
    };

    // co-resident CTAs start in lockstep (all load, then all compute) unless they are offset
    if (a.stagger_ns > 0) {
        const unsigned k = blockIdx.x / (unsigned)a.num_sms;
        for (unsigned i = 0; i < k; ++i) __nanosleep((unsigned)a.stagger_ns);
    }
    const long long first = blockIdx.x;
    const long long step = gridDim.x;
    // prologue: fill nstage-1 stages
    if (mover) {
        for (int p = 0; p < nstage - 1; ++p) {
            const long long t = first + p * step;
            if (t < a.num_tiles) issue_load(t, p);
        }
    }

    bool mats_loaded = false;
    long long it = 0;
    for (long long tile_id = first; tile_id < a.num_tiles; tile_id += step, ++it) {
        const int s = (int)(it % nstage);
        const unsigned parity = (unsigned)((it / nstage) & 1);
        // nstage <= 2: prefetch tile it+nstage-1 into the stage that tile it-1 used (its store
        // was issued at the end of the previous iteration: wait until the engine has read it).
        // nstage >= 3 refills at the END of the iteration instead (below), when that store has
        // had a whole gate phase to drain, so warp 0 never stalls on it.
        if (mover && nstage <= 2) {
            const long long tn = tile_id + (long long)(nstage - 1) * step;
            if (tn < a.num_tiles) {
                bulk_wait_read_all();
                if (a.trace && threadIdx.x == 0 && it < TRACE_TILES) a.trace[((size_t)blockIdx.x * TRACE_TILES + it) * 4 + 3] = global_ns();
                __syncwarp();
                issue_load(tn, (int)((it + nstage - 1) % nstage));
            }
        }
        // gate matrices -> shared memory, register order (once, or per row for batched gates)
        if (!mats_loaded || a.mats_row_stride != 0) {
            const long long row = tile_id / a.tiles_per_row;
            const C *__restrict__ mats = reinterpret_cast<const C *>(a.mats) + row * a.mats_row_stride;
            if (mats_loaded) __syncthreads();          // previous tile's gates are done with sM
            for (int g = 0; g < a.num_gates; ++g) {
                const FusedGate &gd = a.gates[g];
                const int K = gd.k, D = 1 << K;
                for (int e = threadIdx.x; e < D * D; e += FUSED_THREADS) {
                    const int sr = e >> K, t = e & (D - 1);
                    int gi = 0, gj = 0;
                    for (int i = 0; i < K; ++i) {
                        gi |= ((sr >> i) & 1) << gd.gb[i];
                        gj |= ((t >> i) & 1) << gd.gb[i];
                    }
                    C val;
                    if (a.adjoint) val = cconj(mats[gd.goff + gj * D + gi]);
                    else val = mats[gd.goff + gi * D + gj];
                    sM[gd.smoff + e] = val;
                }
            }
            mats_loaded = true;
            __syncthreads();
        }
        if (a.trace && threadIdx.x == 0 && it < TRACE_TILES) a.trace[((size_t)blockIdx.x * TRACE_TILES + it) * 4 + 0] = global_ns();
        mbar_wait(smem_u32(&bars[s]), parity);
        if (a.trace && threadIdx.x == 0 && it < TRACE_TILES) a.trace[((size_t)blockIdx.x * TRACE_TILES + it) * 4 + 1] = global_ns();
        // single-buffered CTAs cannot load ahead: at least pull the next tile into L2 now so the
        // real load after this tile's store is an L2 hit
        if (a.l2_prefetch && a.trank > 0 && threadIdx.x == 0) {
            const long long tn = tile_id + (long long)nstage * step;
            if (tn < a.num_tiles) {
                long long row;
                int c[5];
                tensor_coords(tile_base(tn, row), c);
                tma_prefetch_l2(a.trank, &a.tmap_in, c);
            }
        }

        V *tv = reinterpret_cast<V *>(smem_raw + (size_t)s * tile_bytes);
        for (int g = 0; g < a.num_gates; ++g) {
            const FusedGate &gd = a.gates[g];
            apply_any_gate<R, FUSED_THREADS, LEAN>(tv, sM + gd.smoff, gd, TV, a.swizzle != 0, a.use_f2 != 0);
            if (g + 1 < a.num_gates) __syncthreads();
        }
        fence_proxy_async();        // make the generic-proxy writes visible to the copy engine
        __syncthreads();
        if (a.trace && threadIdx.x == 0 && it < TRACE_TILES) a.trace[((size_t)blockIdx.x * TRACE_TILES + it) * 4 + 2] = global_ns();
        if (mover) {
            issue_store(tile_id, s);
            if (nstage >= 3) {
                const long long tn = tile_id + (long long)(nstage - 1) * step;
                if (tn < a.num_tiles) {
                    bulk_wait_read_but_one();      // store of tile it-1 has left shared memory
                    __syncwarp();
                    issue_load(tn, (int)((it + nstage - 1) % nstage));
                }
            }
        }
    }
    if (mover) bulk_wait_all();
}

// ---------------------------------------------------------------------------------------
// Register-blocked pass (ClusterArgs above): persistent CTAs, 3 x 256 threads per SM, one 64 KiB
// tile buffer per CTA moved by ONE TMA tensor copy each way; the co-resident CTAs overlap each
// other's load / gate / store phases.  Complex64, shared 1-/2-qubit gates, TMA tensor path only.
// ARITH 0: scalar FFMA with UR operands, 1: packed FFMA2 with UR-broadcast operands.
struct ClusterGeom {
    const void *in;
    long long num_tiles, tiles_per_row;
    int total_bits, T, L, H;
    int high[UA_MAX_TILE_BITS];
    int trank;
    int nbuf;                    // tile buffers in the ring
    int tab_front, tab_bytes;    // offset table in front of / behind the ring, its size
    int tstart[6];
    int scatter_m;
    unsigned long long tile_xor;
    int nins;
    int ins[UA_MAX_TILE_BITS + UA_MAX_SCATTER_BITS];
    int vpos[UA_MAX_SCATTER_BITS];
    int tstart_out[6];
    alignas(64) CUtensorMap tmap_in;
    alignas(64) CUtensorMap tmap_out;
    alignas(64) CUtensorMap tmap_dst[1 << UA_MAX_SCATTER_BITS];
};

__device__ __forceinline__ float2 lds64(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(unsigned addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(unsigned addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void team_barrier(int team) {
    asm volatile("bar.sync %0, 256;" ::"r"(1 + team) : "memory");
}

#ifdef UA_RING_DEBUG
__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_dbg(unsigned bar, unsigned parity, int role, int j, int b) {
    unsigned iters = 0;
    while (!mbar_try(bar, parity)) {
        if (++iters > (1u << 20)) {
            if ((threadIdx.x & 31) == 0) {
                unsigned long long st;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(st) : "r"(bar));
                printf("STUCK block %d warp %d role %d tile %d buf %d parity %u state %llx\n", blockIdx.x, threadIdx.x / 32, role, j, b, parity, st);
            }
            __nanosleep(1000000);
            __trap();
        }
    }
}
#define RING_WAIT(bar, par, role, j, b) mbar_wait_dbg(bar, par, role, j, b)
#else
#define RING_WAIT(bar, par, role, j, b) mbar_wait(bar, par)
#endif

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ONE CTA per SM: TEAMS independent 256-thread teams (named barriers), one producer warp and a
// ring of a.nbuf tile buffers.  Tile j of the CTA's sequence is processed by team j % TEAMS in
// buffer j % nbuf.  The producer warp owns the copy engine: when a team reports a tile done
// (mbarrier) it stores the tile with one TMA copy and, as soon as the store has left shared memory,
// refills the buffer with tile j + nbuf.  The buffers no team computes on are therefore always
// in flight to or from HBM, teams never wait for each other, and one team's shared-memory phases
// (LDS / STS / barrier) overlap another's FMA phase.  nbuf is a multiple of TEAMS (launcher), so a
// buffer and its two mbarriers are only ever used by one team and the producer.
template <int TEAMS, bool SCATTER, int ARITH, bool SWZ>
__global__ void __launch_bounds__(TEAMS * 256 + 32, 1) cluster_ring_kernel(const __grid_constant__ ClusterGeom a,
                                                                          const __grid_constant__ ClusterArgs ca) {
    constexpr int TT = 256, MAXB = 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_full[MAXB], bar_done[MAXB];
    const unsigned tile_bytes = (1u << a.T) * 8u;
    // buffers are aligned to their own size (member addresses are formed with XOR) and to 1 KiB
    // (the 128-byte swizzle pattern); the offset table sits in the alignment slack in front of the
    // ring when it fits there, behind the ring otherwise (a.tab_front, decided by the launcher)
    const unsigned buf_bytes = tile_bytes < 1024u ? 1024u : tile_bytes;
    const unsigned dyn_s = smem_u32(smem_raw);
    const unsigned tile_s = (dyn_s + (a.tab_front ? (unsigned)a.tab_bytes : 0u) + buf_bytes - 1u) & ~(buf_bytes - 1u);
    const int NB = a.nbuf;
    unsigned char *ring = smem_raw + (tile_s - dyn_s);
    unsigned short *tab = reinterpret_cast<unsigned short *>(a.tab_front ? smem_raw : ring + (size_t)NB * buf_bytes);
    // the shuffle tells the compiler that the warp number is warp-uniform: everything derived from
    // it (team, tile counter, buffer, the gate loop) stays on the uniform datapath
    const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x / 32), 0);
    const int team = wid / (TT / 32);
    const unsigned ttid = threadIdx.x % TT;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NB; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_done[s]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }

    // tile counter -> (tensor coordinates of the source box, of the destination box, destination)
    const int tpr_bits = a.total_bits - a.T;       // log2(tiles per row)
    auto coords = [&](long long tile_id, int *cin, int *cout, int &dst) {
        const long long row = tile_id >> tpr_bits;
        const long long j = tile_id - (row << tpr_bits);
        uint64_t base;
        if constexpr (SCATTER) {
            base = (uint64_t)(j >> a.scatter_m) << a.L;
            for (int i = 0; i < a.nins; ++i) base = insert_zero(base, a.ins[i]);
            for (int i = 0; i < a.scatter_m; ++i) base |= (uint64_t)((j >> i) & 1) << a.vpos[i];
            base ^= a.tile_xor;
        } else {
            base = (uint64_t)j << a.L;
            for (int i = 0; i < a.H; ++i) base = insert_zero(base, a.high[i]);
        }
        base += (uint64_t)row << a.total_bits;
        for (int d = 0; d < a.trank; ++d) {
            uint64_t v = base >> a.tstart[d];
            if (d + 1 < a.trank) v &= (1ull << (a.tstart[d + 1] - a.tstart[d])) - 1ull;
            cin[d] = (int)v;
        }
        dst = 0;
        if constexpr (SCATTER) {
            uint64_t x = base;
            for (int i = a.scatter_m - 1; i >= 0; --i) {
                const int v = a.vpos[i];
                dst |= (int)((base >> v) & 1ull) << i;
                x = ((x >> (v + 1)) << v) | (x & ((1ull << v) - 1ull));
            }
            for (int d = 0; d < a.trank; ++d) {
                uint64_t v = x >> a.tstart_out[d];
                if (d + 1 < a.trank) v &= (1ull << (a.tstart_out[d + 1] - a.tstart_out[d])) - 1ull;
                cout[d] = (int)v;
            }
        }
    };
    const int count = (int)((a.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);    // tiles of this CTA
    const int gbits = a.T - CL_BITS;
    const unsigned groups = 1u << gbits;
    const int ncl = ca.ncl;
    // byte offset of a thread's group inside a tile buffer, per cluster and per sweep of the team
    // over the groups: computed once (the bit deposit is ~40 instructions) and kept in shared
    // memory for the first clusters of the pass
    const int sweeps = (int)((groups + TT - 1) / TT);
    auto group_base8 = [&](int c, unsigned g) -> unsigned {
        unsigned base = 0;
        for (int k = 0; k < gbits; ++k) base |= ((g >> k) & 1u) << ca.cl[c].fb[k];
        if constexpr (SWZ) base ^= ((base >> 4) & 7u) << 1;
        return base * 8u;
    };
    const int ntab = ncl * sweeps <= CL_TAB ? ncl : CL_TAB / sweeps;
    if (team == 0) {
        for (int c = 0; c < ntab; ++c)
            for (int w = 0; w < sweeps; ++w) {
                const unsigned g = (unsigned)w * TT + ttid;
                tab[(c * sweeps + w) * TT + ttid] = (unsigned short)(g < groups ? group_base8(c, g) : 0u);
            }
    }
    __syncthreads();

    if (team >= TEAMS) {
        // ---------------------------------------------------------------- producer warp
        if ((threadIdx.x & 31u) == 0) {
            auto issue_load = [&](int j) {
                int cin[5], cout[5], dst;
                coords(blockIdx.x + j * (long long)gridDim.x, cin, cout, dst);
                const int b = j % NB;
                const unsigned bar = smem_u32(&bar_full[b]);
                mbar_arrive_expect_tx(bar, tile_bytes);
                tma_load(a.trank, tile_s + (unsigned)b * buf_bytes, &a.tmap_in, cin, bar);
            };
            for (int j = 0; j < NB && j < count; ++j) issue_load(j);
            for (int j = 0; j < count; ++j) {
                const int b = j % NB;
                // coordinates first: the arithmetic overlaps the wait for the team
                int cin[5], cout[5], dst;
                coords(blockIdx.x + j * (long long)gridDim.x, cin, cout, dst);
                RING_WAIT(smem_u32(&bar_done[b]), (unsigned)((j / NB) & 1), 9, j, b);
                const unsigned src = tile_s + (unsigned)b * buf_bytes;
                if constexpr (SCATTER) tma_store(a.trank, &a.tmap_dst[dst], cout, src);
                else tma_store(a.trank, &a.tmap_out, cin, src);
                bulk_commit();
                // refill the buffer of the PREVIOUS store (it has had a whole tile time to leave
                // shared memory, so this wait does not stall the next store)
                if (j >= 1 && j - 1 + NB < count) {
                    bulk_wait_read_but_one();
                    issue_load(j - 1 + NB);
                }
            }
            bulk_wait_all();
        }
        return;
    }

    // -------------------------------------------------------------------- compute teams
    for (int j = team; j < count; j += TEAMS) {
        const int b = j % NB;
        RING_WAIT(smem_u32(&bar_full[b]), (unsigned)((j / NB) & 1), team, j, b);
        const unsigned tile_sb = tile_s + (unsigned)b * buf_bytes;
        for (int c = 0; c < ncl; ++c) {
            const ClusterDesc &cd = ca.cl[c];
            const int gbeg = cd.gbeg, gend = cd.gend;
            const bool vec16 = cd.vec16 != 0;
            // warp-uniform trip count: the gate loop must stay convergent or ptxas moves the
            // matrices from uniform to vector registers; tiles with fewer groups than threads
            // predicate the loads and stores instead
            for (int w = 0; w < sweeps; ++w) {
                const unsigned g = (unsigned)w * TT + ttid;
                const bool act = g < groups;
                const unsigned base8 = tile_sb + (c < ntab ? (unsigned)tab[(c * sweeps + w) * TT + ttid]
                                                           : (act ? group_base8(c, g) : 0u));
                float2 v[16];
                if (vec16) {
#pragma unroll
                    for (int m = 0; m < 16; m += 2) {
                        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (act) t = lds128(base8 ^ cd.po8[m]);
                        v[m] = make_float2(t.x, t.y);
                        v[m + 1] = make_float2(t.z, t.w);
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        v[m] = make_float2(0.f, 0.f);
                        if (act) v[m] = lds64(base8 ^ cd.po8[m]);
                    }
                }
                for (int q = gbeg; q < gend; ++q) {
                    const ClusterGate cg = ca.g[q];
                    reg_gate_dispatch<ARITH>(v, cg.type, ca.mats + cg.moff);
                }
                if (vec16) {
#pragma unroll
                    for (int m = 0; m < 16; m += 2)
                        if (act) sts128(base8 ^ cd.po8[m], make_float4(v[m].x, v[m].y, v[m + 1].x, v[m + 1].y));
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m)
                        if (act) sts64(base8 ^ cd.po8[m], v[m]);
                }
            }
            if (c + 1 < ncl) team_barrier(team);
        }
        fence_proxy_async();        // make the generic-proxy writes visible to the copy engine
        team_barrier(team);
        if (ttid == 0) mbar_arrive(smem_u32(&bar_done[b]));
    }
}

// ---------------------------------------------------------------------------------------
// Team variant of the forward pass (shared gates, TMA tensor path): ONE 512-thread CTA per SM
// split into two independent 256-thread teams that share THREE tile buffers.  Tile j of the
// CTA's sequence is processed by team j & 1 in buffer j % 3; after a team has stored tile j
// it refills that buffer with tile j + 3 a couple of gates into its next tile (by then the
// store has drained), so every team always finds its next tile already in shared memory:
// the ~3 us store drain + ~1-2 us load wait per tile (measured with ua_debug_set_fused_trace)
// no longer idle a CTA, while the two teams still interleave their LDS and FMA phases.

template <typename R>
__global__ void __launch_bounds__(512, 1) fused_pass_team_kernel(const __grid_constant__ FusedArgs a) {
    using C = typename CplxOf<R>::type;
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    constexpr int APVLOG = APV == 2 ? 1 : 0;
    constexpr int NB = 3;
    constexpr int EBITS = sizeof(C) == 8 ? 0 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[NB];
    const unsigned tile_bytes = (1u << a.T) * (unsigned)sizeof(C);
    C *sM = reinterpret_cast<C *>(smem_raw + (size_t)NB * tile_bytes);
    const int TV = a.T - APVLOG;
    const int team = threadIdx.x >> 8;
    const unsigned ttid = threadIdx.x & 255u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NB; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    // gate matrices -> shared memory, register order (shared by both teams)
    {
        const C *__restrict__ mats = reinterpret_cast<const C *>(a.mats);
        for (int g = 0; g < a.num_gates; ++g) {
            const FusedGate &gd = a.gates[g];
            const int K = gd.k, D = 1 << K;
            for (int e = threadIdx.x; e < D * D; e += 512) {
                const int sr = e >> K, t = e & (D - 1);
                int gi = 0, gj = 0;
                for (int i = 0; i < K; ++i) {
                    gi |= ((sr >> i) & 1) << gd.gb[i];
                    gj |= ((t >> i) & 1) << gd.gb[i];
                }
                C val;
                if (a.adjoint) val = cconj(mats[gd.goff + gj * D + gi]);
                else val = mats[gd.goff + gi * D + gj];
                sM[gd.smoff + e] = val;
            }
        }
    }
    __syncthreads();

    const long long count = (a.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // tiles of this CTA
    auto coords_of = [&](long long j, int *c) {
        const long long tile_id = blockIdx.x + j * (long long)gridDim.x;
        const long long row = tile_id / a.tiles_per_row;
        const long long jj = tile_id - row * a.tiles_per_row;
        uint64_t base = (uint64_t)jj << a.L;
        for (int i = 0; i < a.H; ++i) base = insert_zero(base, a.high[i]);
        base += (uint64_t)row << a.total_bits;
        const uint64_t e = base << EBITS;
        for (int d = 0; d < a.trank; ++d) {
            uint64_t v = e >> a.tstart[d];
            if (d + 1 < a.trank) v &= (1ull << (a.tstart[d + 1] - a.tstart[d])) - 1ull;
            c[d] = (int)v;
        }
    };
    auto issue_load = [&](long long j) {          // one thread
        const int b = (int)(j % NB);
        const unsigned bar = smem_u32(&bars[b]);
        int c[5];
        coords_of(j, c);
        mbar_arrive_expect_tx(bar, tile_bytes);
        tma_load(a.trank, smem_u32(smem_raw + (size_t)b * tile_bytes), &a.tmap_in, c, bar);
    };
    if (threadIdx.x == 0)
        for (long long j = 0; j < NB && j < count; ++j) issue_load(j);
    __syncthreads();

    long long pending = -1;                         // tile whose load waits for this team's last store
    const int refill_after = a.num_gates >= 2 ? 1 : 0;
    for (long long j = team; j < count; j += 2) {
        const int b = (int)(j % NB);
        mbar_wait(smem_u32(&bars[b]), (unsigned)((j / NB) & 1));
        V *tv = reinterpret_cast<V *>(smem_raw + (size_t)b * tile_bytes);
        for (int g = 0; g < a.num_gates; ++g) {
            const FusedGate &gd = a.gates[g];
            const C *M = sM + gd.smoff;
            const bool low = (APV == 2) && gd.sb[0] == 0;
            if constexpr (APV == 2) {
                if (low) {
                    if (gd.k == 1) apply_gate_smem<R, 1, true, false>(tv, M, gd, TV, 256, ttid);
                    else if (gd.k == 2) apply_gate_smem<R, 2, true, false>(tv, M, gd, TV, 256, ttid);
                    else apply_gate_smem<R, 3, true, false>(tv, M, gd, TV, 256, ttid);
                }
            }
            if (!low) {
                if (gd.k == 1) apply_gate_smem<R, 1, false, false>(tv, M, gd, TV, 256, ttid);
                else if (gd.k == 2) apply_gate_smem<R, 2, false, false>(tv, M, gd, TV, 256, ttid);
                else apply_gate_smem<R, 3, false, false>(tv, M, gd, TV, 256, ttid);
            }
            if (g == refill_after && ttid == 0 && pending >= 0) {
                bulk_wait_read_all();               // this team's previous store has left its buffer
                issue_load(pending);
                pending = -1;
            }
            if (g + 1 < a.num_gates) team_barrier(team);
        }
        fence_proxy_async();
        team_barrier(team);
        if (ttid == 0) {
            int c[5];
            coords_of(j, c);
            tma_store(a.trank, &a.tmap_out, c, smem_u32(smem_raw + (size_t)b * tile_bytes));
            bulk_commit();
            if (j + NB < count) pending = j + NB;
        }
    }
    if (ttid == 0) {
        if (pending >= 0) { bulk_wait_read_all(); issue_load(pending); }
        bulk_wait_all();
    }
}

// =======================================================================================
// Fused BACKWARD pass (adjoint method): psi and grad tiles are staged together; walking the
// pass's gates in reverse, every gate does
//     psi <- U^H psi            (recomputes the gate's input; gates are unitary)
//     grad_U += g psi_in^H      (if the gate needs a gradient; summed in registers over the
//                                thread's groups, over the warp by shuffles, over the CTA in
//                                fp64 shared-memory accumulators)
//     g   <- U^H g
// and both tiles go back.  One read + one write of psi and of g (32 B / 64 B per amplitude)
// for the whole pass instead of three passes per gate.  Gates are 1- or 2-qubit.
struct BwdArgs {
    FusedArgs f;                               // f.in = psi (in-out), f.out = grad (in-out)
    double2 *acc;                              // [rows or 1][mat_elems] fp64 gradient accumulators
    int mat_elems;
    unsigned long long needs_grad;             // bit g: gate g needs a gradient
};

template <typename R, int K, bool LOW>
__device__ __forceinline__ void bwd_gate_smem(typename VecOf<R>::type *tpsi, typename VecOf<R>::type *tg,
                                              const typename CplxOf<R>::type *M /* U^H, register order */,
                                              const FusedGate &gd, int TV, int nthreads, bool needs_grad,
                                              double2 *sAcc /* this gate's D*D accumulators, register order */) {
    using C = typename CplxOf<R>::type;
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    constexpr int APVLOG = APV == 2 ? 1 : 0;
    constexpr int KH = LOW ? K - 1 : K;
    constexpr int NV = 1 << KH;
    constexpr int D = 1 << K;
    constexpr int NG = (APV == 2 && !LOW) ? 2 : 1;     // independent groups per item
    int vb[KH > 0 ? KH : 1];
    unsigned off[KH > 0 ? KH : 1];
#pragma unroll
    for (int i = 0; i < KH; ++i) {
        vb[i] = (int)gd.sb[i + (LOW ? 1 : 0)] - APVLOG;
        off[i] = 1u << vb[i];
    }
    C mr[D * D];
#pragma unroll
    for (int e = 0; e < D * D; ++e) mr[e] = M[e];
    R accx[D * D], accy[D * D];                    // grad[a][b] partial sums (register order)
#pragma unroll
    for (int e = 0; e < D * D; ++e) { accx[e] = R(0); accy[e] = R(0); }

    const unsigned groups = 1u << (TV - KH);
    for (unsigned g = threadIdx.x; g < groups; g += nthreads) {
        unsigned b = g;
#pragma unroll
        for (int i = 0; i < KH; ++i) b = insert_zero32(b, vb[i]);
        V xv[NV], yv[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            unsigned idx = b;
#pragma unroll
            for (int i = 0; i < KH; ++i)
                if ((c >> i) & 1) idx |= off[i];
            xv[c] = tpsi[idx];
            yv[c] = tg[idx];
        }
        V xo[NV], yo[NV];
#pragma unroll
        for (int h = 0; h < NG; ++h) {
            // gather the 2^K amplitudes of this group
            C x[D], y[D];
#pragma unroll
            for (int t = 0; t < D; ++t) {
                if constexpr (APV == 1) { x[t] = xv[t]; y[t] = yv[t]; }
                else if constexpr (LOW) {
                    const V a = xv[t >> 1], bb = yv[t >> 1];
                    x[t] = (t & 1) ? mk(a.z, a.w) : mk(a.x, a.y);
                    y[t] = (t & 1) ? mk(bb.z, bb.w) : mk(bb.x, bb.y);
                } else {
                    x[t] = h ? mk(xv[t].z, xv[t].w) : mk(xv[t].x, xv[t].y);
                    y[t] = h ? mk(yv[t].z, yv[t].w) : mk(yv[t].x, yv[t].y);
                }
            }
            C xin[D], yin[D];
#pragma unroll
            for (int s = 0; s < D; ++s) {
                C ax = mk(R(0), R(0)), ay = mk(R(0), R(0));
#pragma unroll
                for (int t = 0; t < D; ++t) { cfma(ax, mr[s * D + t], x[t]); cfma(ay, mr[s * D + t], y[t]); }
                xin[s] = ax; yin[s] = ay;
            }
            if (needs_grad) {
#pragma unroll
                for (int aa = 0; aa < D; ++aa)
#pragma unroll
                    for (int bb = 0; bb < D; ++bb) {
                        // y[aa] * conj(xin[bb])
                        accx[aa * D + bb] = fma(y[aa].x, xin[bb].x, accx[aa * D + bb]);
                        accx[aa * D + bb] = fma(y[aa].y, xin[bb].y, accx[aa * D + bb]);
                        accy[aa * D + bb] = fma(y[aa].y, xin[bb].x, accy[aa * D + bb]);
                        accy[aa * D + bb] = fma(-y[aa].x, xin[bb].y, accy[aa * D + bb]);
                    }
            }
            // scatter back into vectors
#pragma unroll
            for (int t = 0; t < D; ++t) {
                if constexpr (APV == 1) { xo[t] = xin[t]; yo[t] = yin[t]; }
                else if constexpr (LOW) {
                    if (t & 1) { xo[t >> 1].z = xin[t].x; xo[t >> 1].w = xin[t].y; yo[t >> 1].z = yin[t].x; yo[t >> 1].w = yin[t].y; }
                    else { xo[t >> 1].x = xin[t].x; xo[t >> 1].y = xin[t].y; yo[t >> 1].x = yin[t].x; yo[t >> 1].y = yin[t].y; }
                } else {
                    if (h) { xo[t].z = xin[t].x; xo[t].w = xin[t].y; yo[t].z = yin[t].x; yo[t].w = yin[t].y; }
                    else { xo[t].x = xin[t].x; xo[t].y = xin[t].y; yo[t].x = yin[t].x; yo[t].y = yin[t].y; }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            unsigned idx = b;
#pragma unroll
            for (int i = 0; i < KH; ++i)
                if ((c >> i) & 1) idx |= off[i];
            tpsi[idx] = xo[c];
            tg[idx] = yo[c];
        }
    }
    if (needs_grad) {
        // warp reduce, then one fp64 shared-memory atomic per entry per warp
#pragma unroll
        for (int e = 0; e < D * D; ++e) {
            R vx = accx[e], vy = accy[e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vx += __shfl_xor_sync(0xffffffffu, vx, o);
                vy += __shfl_xor_sync(0xffffffffu, vy, o);
            }
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&sAcc[e].x, (double)vx);
                atomicAdd(&sAcc[e].y, (double)vy);
            }
        }
    }
}

template <typename R>
__global__ void __launch_bounds__(256, sizeof(R) == 4 ? 2 : 1) fused_bwd_kernel(const __grid_constant__ BwdArgs ba) {
    using C = typename CplxOf<R>::type;
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    constexpr int APVLOG = APV == 2 ? 1 : 0;
    constexpr int NT = 256;
    const FusedArgs &a = ba.f;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned tile_bytes = (1u << a.T) * (unsigned)sizeof(C);
    unsigned char *buf_psi = smem_raw, *buf_g = smem_raw + tile_bytes;
    C *sM = reinterpret_cast<C *>(smem_raw + 2 * (size_t)tile_bytes);
    double2 *sAcc = reinterpret_cast<double2 *>(smem_raw + 2 * (size_t)tile_bytes +
                                                (((size_t)ba.mat_elems * sizeof(C) + 127) & ~(size_t)127));
    const int TV = a.T - APVLOG;
    const unsigned run_bytes = (1u << a.L) * (unsigned)sizeof(C);
    const unsigned runs = 1u << a.H;
    const unsigned lane = threadIdx.x & 31u;
    const bool mover = threadIdx.x < 32;
    constexpr int EBITS = sizeof(C) == 8 ? 0 : 1;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    for (int e = threadIdx.x; e < ba.mat_elems; e += NT) sAcc[e] = make_double2(0.0, 0.0);
    __syncthreads();

    auto tile_base = [&](long long tile_id, long long &row) -> uint64_t {
        row = tile_id / a.tiles_per_row;
        const long long j = tile_id - row * a.tiles_per_row;
        uint64_t base = (uint64_t)j << a.L;
        for (int i = 0; i < a.H; ++i) base = insert_zero(base, a.high[i]);
        return base + ((uint64_t)row << a.total_bits);
    };
    auto run_offset = [&](unsigned r) -> uint64_t {
        uint64_t o = 0;
        for (int i = 0; i < a.H; ++i)
            if ((r >> i) & 1u) o |= 1ull << a.high[i];
        return o;
    };
    auto tensor_coords = [&](uint64_t base, int *c) {
        const uint64_t e = base << EBITS;
        for (int j = 0; j < a.trank; ++j) {
            uint64_t v = e >> a.tstart[j];
            if (j + 1 < a.trank) v &= (1ull << (a.tstart[j + 1] - a.tstart[j])) - 1ull;
            c[j] = (int)v;
        }
    };
    char *gp_psi = reinterpret_cast<char *>(const_cast<void *>(a.in));
    char *gp_g = reinterpret_cast<char *>(a.out);

    unsigned phase = 0;
    long long last_row = -1;
    for (long long tile_id = blockIdx.x; tile_id < a.num_tiles; tile_id += gridDim.x) {
        long long row;
        const uint64_t base = tile_base(tile_id, row);
        // ---- load both tiles (previous stores must have left shared memory) ---------------
        if (mover) {
            bulk_wait_read_all();
            __syncwarp();
            const unsigned b32 = smem_u32(&bar);
            if (lane == 0) mbar_arrive_expect_tx(b32, 2 * tile_bytes);
            __syncwarp();
            if (a.trank > 0) {
                if (lane == 0) {
                    int c[5];
                    tensor_coords(base, c);
                    tma_load(a.trank, smem_u32(buf_psi), &a.tmap_in, c, b32);
                    tma_load(a.trank, smem_u32(buf_g), &a.tmap_out, c, b32);
                }
            } else {
                for (unsigned r = lane; r < runs; r += 32) {
                    const uint64_t o = (base + run_offset(r)) * sizeof(C);
                    bulk_g2s(smem_u32(buf_psi) + r * run_bytes, gp_psi + o, run_bytes, b32);
                    bulk_g2s(smem_u32(buf_g) + r * run_bytes, gp_g + o, run_bytes, b32);
                }
            }
        }
        // ---- adjoint matrices -> shared memory (register order); per-row gates reload per row
        if (last_row < 0 || (a.mats_row_stride != 0 && row != last_row)) {
            if (last_row >= 0 && a.mats_row_stride != 0) {
                // flush the finished row's gradient accumulators
                __syncthreads();
                for (int e = threadIdx.x; e < ba.mat_elems; e += NT) {
                    const double2 v = sAcc[e];
                    atomicAdd(&ba.acc[last_row * ba.mat_elems + e].x, v.x);
                    atomicAdd(&ba.acc[last_row * ba.mat_elems + e].y, v.y);
                    sAcc[e] = make_double2(0.0, 0.0);
                }
            }
            const C *__restrict__ mats = reinterpret_cast<const C *>(a.mats) + row * a.mats_row_stride;
            for (int g = 0; g < a.num_gates; ++g) {
                const FusedGate &gd = a.gates[g];
                const int K = gd.k, D = 1 << K;
                for (int e = threadIdx.x; e < D * D; e += NT) {
                    const int sr = e >> K, t = e & (D - 1);
                    int gi = 0, gj = 0;
                    for (int i = 0; i < K; ++i) {
                        gi |= ((sr >> i) & 1) << gd.gb[i];
                        gj |= ((t >> i) & 1) << gd.gb[i];
                    }
                    sM[gd.smoff + e] = cconj(mats[gd.goff + gj * D + gi]);     // U^H
                }
            }
            last_row = row;
            __syncthreads();
        }
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1u;

        V *tpsi = reinterpret_cast<V *>(buf_psi);
        V *tg = reinterpret_cast<V *>(buf_g);
        for (int g = a.num_gates - 1; g >= 0; --g) {
            const FusedGate &gd = a.gates[g];
            const bool ng = (ba.needs_grad >> g) & 1ull;
            const C *M = sM + gd.smoff;
            double2 *acc = sAcc + gd.smoff;
            const bool low = (APV == 2) && gd.sb[0] == 0;
            if constexpr (APV == 2) {
                if (low) {
                    if (gd.k == 1) bwd_gate_smem<R, 1, true>(tpsi, tg, M, gd, TV, NT, ng, acc);
                    else bwd_gate_smem<R, 2, true>(tpsi, tg, M, gd, TV, NT, ng, acc);
                }
            }
            if (!low) {
                if (gd.k == 1) bwd_gate_smem<R, 1, false>(tpsi, tg, M, gd, TV, NT, ng, acc);
                else bwd_gate_smem<R, 2, false>(tpsi, tg, M, gd, TV, NT, ng, acc);
            }
            __syncthreads();
        }
        fence_proxy_async();
        __syncthreads();
        if (mover) {
            if (a.trank > 0) {
                if (lane == 0) {
                    int c[5];
                    tensor_coords(base, c);
                    tma_store(a.trank, &a.tmap_in, c, smem_u32(buf_psi));
                    tma_store(a.trank, &a.tmap_out, c, smem_u32(buf_g));
                }
            } else {
                for (unsigned r = lane; r < runs; r += 32) {
                    const uint64_t o = (base + run_offset(r)) * sizeof(C);
                    bulk_s2g(gp_psi + o, smem_u32(buf_psi) + r * run_bytes, run_bytes);
                    bulk_s2g(gp_g + o, smem_u32(buf_g) + r * run_bytes, run_bytes);
                }
            }
            bulk_commit();
        }
    }
    if (mover) bulk_wait_all();
    __syncthreads();
    // flush the gradient accumulators (entries are in register order per gate; the host shim
    // un-permutes them)
    if (last_row >= 0) {
        const long long dst_row = a.mats_row_stride != 0 ? last_row : 0;
        for (int e = threadIdx.x; e < ba.mat_elems; e += NT) {
            const double2 v = sAcc[e];
            atomicAdd(&ba.acc[dst_row * ba.mat_elems + e].x, v.x);
            atomicAdd(&ba.acc[dst_row * ba.mat_elems + e].y, v.y);
        }
    }
}

static int max_tile_bits(int dtype) { return dtype == UA_C64 ? 14 : 13; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// Describe the tile set {0..L-1} U high[] (amplitude bits) as TMA boxes.  Works in 8-byte
// element bits (complex128 = 2 elements).  Returns false when more than 5 dimensions would be
// needed or the encoder is unavailable; the kernel then moves tiles run by run.
// swizzle128: the first dimension is exactly the 4 lowest element bits (a 128-byte row) and the
// box lands in shared memory with the 128-byte swizzle (16-byte chunk index ^= row index mod 8):
// the register-blocked gate phase then reads any 4-bit cluster without bank conflicts.
static bool setup_tensor_maps(FusedArgs &a, int ebits, long long total_amps, bool swizzle128 = false) {
    if (getenv("UA_FUSED_NO_TENSOR")) return false;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    // windows of consecutive element bits inside the tile, each at most 8 bits (box <= 256)
    int wstart[16], wlen[16], nw = 0;
    int pos[UA_MAX_TILE_BITS + 2], np = 0;
    for (int b = 0; b < a.L + ebits; ++b) pos[np++] = b;
    for (int i = 0; i < a.H; ++i) pos[np++] = a.high[i] + ebits;
    if (swizzle128 && a.L + ebits < 4) return false;
    for (int i = 0; i < np; ++i) {
        const bool split = swizzle128 && pos[i] == 4;
        if (nw > 0 && !split && pos[i] == wstart[nw - 1] + wlen[nw - 1] && wlen[nw - 1] < 8) wlen[nw - 1]++;
        else { if (nw == 16) return false; wstart[nw] = pos[i]; wlen[nw] = 1; nw++; }
    }
    if (nw > 5) return false;
    const unsigned long long total_elems = (unsigned long long)total_amps << ebits;
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t box[5], estr[5];
    for (int j = 0; j < nw; ++j) {
        a.tstart[j] = wstart[j];
        box[j] = 1u << wlen[j];
        estr[j] = 1;
        if (j + 1 < nw) gdim[j] = 1ull << (wstart[j + 1] - wstart[j]);
        else gdim[j] = total_elems >> wstart[j];
        if (gdim[j] > 0xffffffffull || gdim[j] < box[j]) return false;
        if (j > 0) gstride[j - 1] = (8ull << wstart[j]);
    }
    a.tstart[nw] = 0;
    for (int which = 0; which < 2; ++which) {
        void *addr = const_cast<void *>(which == 0 ? a.in : (const void *)a.out);
        CUtensorMap *tm = which == 0 ? &a.tmap_in : &a.tmap_out;
        const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)nw, addr, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    if (a.scatter_m > 0) {
        // destination geometry: the same windows in the index with the scatter bits removed
        // (no scatter bit lies inside a window, so every window stays contiguous)
        const unsigned long long dst_elems = total_elems >> a.scatter_m;
        int dstart[6];
        for (int j = 0; j < nw; ++j) {
            int below = 0;
            for (int i = 0; i < a.scatter_m; ++i) below += (a.vpos[i] + ebits < wstart[j]) ? 1 : 0;
            dstart[j] = wstart[j] - below;
        }
        for (int j = 0; j < nw; ++j) {
            a.tstart_out[j] = dstart[j];
            if (j + 1 < nw) gdim[j] = 1ull << (dstart[j + 1] - dstart[j]);
            else gdim[j] = dst_elems >> dstart[j];
            if (gdim[j] > 0xffffffffull || gdim[j] < box[j]) return false;
            if (j > 0) gstride[j - 1] = (8ull << dstart[j]);
        }
        a.tstart_out[nw] = 0;
        for (int b = 0; b < (1 << a.scatter_m); ++b) {
            const CUresult r = enc(&a.tmap_dst[b], CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)nw, a.dst[b], gdim,
                                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return false;
        }
    }
    a.trank = nw;
    return true;
}

static unsigned long long *g_trace_ptr = nullptr;

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <typename R, int FUSED_THREADS, int MINB = 512 / FUSED_THREADS, bool SCATTER = false>
static int launch_fused(FusedArgs &a, size_t tile_bytes, size_t mat_bytes, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = fused_pass_kernel<R, FUSED_THREADS, MINB, SCATTER>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { set_error("ua_apply_fused_pass: cannot raise shared memory limit: %s", cudaGetErrorString(e)); return UA_ERR_CUDA; }
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // shared memory budget per CTA: up to 512/FUSED_THREADS CTAs share an SM (fewer when
    // UA_FUSED_CTAS says so or the tiles are too large)
    int ctas = MINB;
    const int want_ctas = env_int("UA_FUSED_CTAS", 0);
    if (want_ctas >= 1 && want_ctas < ctas) ctas = want_ctas;
    while (ctas > 1 && (size_t)(224 * 1024) / ctas - 1024 < tile_bytes + mat_bytes) --ctas;
    const size_t budget = (size_t)(224 * 1024) / ctas - 1024;
    int nstage = (budget > mat_bytes) ? (int)((budget - mat_bytes) / tile_bytes) : 0;
    if (nstage > 3) nstage = 3;
    if (nstage < 1) { set_error("ua_apply_fused_pass: tile does not fit in shared memory"); return UA_ERR_UNSUPPORTED; }
    const int want = env_int("UA_FUSED_STAGES", 0);
    if (want >= 1 && want <= nstage) nstage = want;
    a.nstage = nstage;
    a.swizzle = env_int("UA_FUSED_SWZ", 0);
    a.use_f2 = env_int("UA_FUSED_F2", 0);
    // L2 prefetch of the tile after next: measured neutral on time (147.9 vs 148.3 ms per bench step)
    // but 15-20 % of the prefetched lines are evicted before use and read from HBM twice
    // (ncu: dram read 9.7-10.3 GB per pass instead of 8.59 GB), so it is off by default
    a.l2_prefetch = env_int("UA_FUSED_L2PF", 0);
    a.stagger_ns = env_int("UA_FUSED_STAGGER_NS", 0);
    a.trace = g_trace_ptr;
    a.num_sms = sms;
    const size_t smem = (size_t)nstage * tile_bytes + mat_bytes;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FUSED_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) { set_error("ua_apply_fused_pass: occupancy query failed (%s), smem=%zu", cudaGetErrorString(e), smem); cudaGetLastError(); return UA_ERR_CUDA; }
    long long grid = (long long)per_sm * sms;
    if (grid > a.num_tiles) grid = a.num_tiles;
    kern<<<(unsigned)grid, FUSED_THREADS, smem, st>>>(a);
    return check_launch("fused_pass_kernel");
}


// ------------------------------------------------------------------ cluster path (host side)
// Where a thread's group lives and how its 16 members are addressed.  Tile element i sits at
// shared-memory element phys(i) = i ^ (((i >> 4) & 7) << 1) when the tile was written with the
// 128-byte TMA swizzle (phys(i) = i otherwise); phys is linear over GF(2), so the address of a
// member is phys(group base) ^ phys(member offset) = base' ^ po[m].
// One shared-memory wavefront serves 128 bytes: 16 lanes of an 8-byte access (8 lanes of a 16-byte
// one), and is conflict-free when those lanes cover all 16 (8) slots of a 128-byte row, i.e. when
// the low group-number bits reach every slot bit.  Slot bit 0 is element bit 0; slot bit s = 1..3
// is element bit s XOR element bit s + 3 under the swizzle, so it can be driven from whichever of
// the two is not a cluster bit.
static void fill_cluster_layout(ClusterDesc &cd, unsigned mask, int T, bool swz) {
    auto phys = [&](unsigned i) -> unsigned { return swz ? (i ^ (((i >> 4) & 7u) << 1)) : i; };
    for (int m = 0; m < (1 << CL_BITS); ++m) {
        unsigned off = 0;
        for (int i = 0; i < CL_BITS; ++i)
            if ((m >> i) & 1) off |= 1u << cd.cb[i];
        cd.po8[m] = (unsigned short)(phys(off) * 8u);
    }
    cd.vec16 = (mask & 1u) ? 1 : 0;
    bool used[32] = {};
    int nf = 0;
    auto take = [&](int b) { cd.fb[nf++] = (unsigned char)b; used[b] = true; };
    auto is_free = [&](int b) { return b >= 0 && b < T && !((mask >> b) & 1u) && !used[b]; };
    if (!cd.vec16 && is_free(0)) take(0);
    for (int sbit = 1; sbit <= 3; ++sbit) {
        if (is_free(sbit)) take(sbit);
        else if (swz && is_free(sbit + 3)) take(sbit + 3);
    }
    for (int b = 0; b < T; ++b)
        if (is_free(b)) take(b);
}

// Cut the pass's gates (a.gates[], tile-local target bits ascending in sb[]) into clusters of at
// most CL_BITS tile bits and copy the HOST matrices into the kernel parameters in target-bit
// order (adjoint applied).  A gate joins an earlier cluster only across clusters it shares no
// bit with (gates on disjoint bits commute), so the product is unchanged.  Returns false when
// the pass cannot use the cluster path (a gate with k > 2, tile smaller than a cluster).
static bool build_clusters(const FusedArgs &a, const float2 *host_mats, ClusterArgs &ca, bool swz) {
    if (a.T < CL_BITS || a.num_gates < 1 || a.num_gates > CL_MAX_GATES) return false;
    struct Cl { unsigned mask; int gates[CL_MAX_GATES]; int ng; };
    static thread_local Cl cls[CL_MAX_GATES];
    int ncl = 0;
    for (int g = 0; g < a.num_gates; ++g) {
        const FusedGate &gd = a.gates[g];
        if (gd.k > 2) return false;
        unsigned gm = 0;
        for (int i = 0; i < gd.k; ++i) gm |= 1u << gd.sb[i];
        int best = -1, best_growth = 99;
        for (int j = ncl - 1; j >= 0; --j) {
            const unsigned u = cls[j].mask | gm;
            const int growth = __builtin_popcount(u) - __builtin_popcount(cls[j].mask);
            if (__builtin_popcount(u) <= CL_BITS && growth < best_growth) { best = j; best_growth = growth; }
            if (cls[j].mask & gm) break;          // cannot move in front of a gate sharing a bit
        }
        if (best < 0) {
            best = ncl++;
            cls[best].mask = 0;
            cls[best].ng = 0;
        }
        cls[best].mask |= gm;
        cls[best].gates[cls[best].ng++] = g;
    }
    ca.ncl = ncl;
    int q = 0, moff = 0;
    for (int c = 0; c < ncl; ++c) {
        // pad to CL_BITS bits from the top of the tile: the thread index then lands on the lowest
        // free bits, i.e. on consecutive shared-memory addresses
        unsigned mask = cls[c].mask;
        for (int b = a.T - 1; b >= 0 && __builtin_popcount(mask) < CL_BITS; --b)
            if (!((mask >> b) & 1u)) mask |= 1u << b;
        int pos_of[32];
        int nb = 0;
        for (int b = 0; b < a.T; ++b)
            if ((mask >> b) & 1u) { ca.cl[c].cb[nb] = (unsigned char)b; pos_of[b] = nb; ++nb; }
        fill_cluster_layout(ca.cl[c], mask, a.T, swz);
        ca.cl[c].gbeg = (unsigned char)q;
        for (int i = 0; i < cls[c].ng; ++i) {
            const FusedGate &gd = a.gates[cls[c].gates[i]];
            const int K = gd.k, D = 1 << K;
            ClusterGate &cg = ca.g[q++];
            cg.moff = (unsigned short)moff;
            if (K == 1) {
                cg.type = (unsigned char)(6 + pos_of[gd.sb[0]]);
            } else {
                static const int pair_type[4][4] = {{-1, 0, 1, 2}, {-1, -1, 3, 4}, {-1, -1, -1, 5}, {-1, -1, -1, -1}};
                cg.type = (unsigned char)pair_type[pos_of[gd.sb[0]]][pos_of[gd.sb[1]]];
            }
            const float2 *src = host_mats + gd.goff;
            for (int e = 0; e < D * D; ++e) {
                const int sr = e >> K, t = e & (D - 1);
                int gi = 0, gj = 0;
                for (int i2 = 0; i2 < K; ++i2) {
                    gi |= ((sr >> i2) & 1) << gd.gb[i2];
                    gj |= ((t >> i2) & 1) << gd.gb[i2];
                }
                float2 val;
                if (a.adjoint) { val = src[gj * D + gi]; val.y = -val.y; }
                else val = src[gi * D + gj];
                ca.mats[moff + e] = val;
            }
            moff += D * D;
        }
        ca.cl[c].gend = (unsigned char)q;
    }
    return true;
}

static void fill_cluster_geom(ClusterGeom &g, const FusedArgs &a) {
    g.in = a.in; g.num_tiles = a.num_tiles; g.tiles_per_row = a.tiles_per_row;
    g.total_bits = a.total_bits; g.T = a.T; g.L = a.L; g.H = a.H;
    for (int i = 0; i < UA_MAX_TILE_BITS; ++i) g.high[i] = a.high[i];
    g.trank = a.trank;
    g.nbuf = 1;
    for (int i = 0; i < 6; ++i) { g.tstart[i] = a.tstart[i]; g.tstart_out[i] = a.tstart_out[i]; }
    g.scatter_m = a.scatter_m; g.tile_xor = a.tile_xor; g.nins = a.nins;
    for (int i = 0; i < UA_MAX_TILE_BITS + UA_MAX_SCATTER_BITS; ++i) g.ins[i] = a.ins[i];
    for (int i = 0; i < UA_MAX_SCATTER_BITS; ++i) g.vpos[i] = a.vpos[i];
    g.tmap_in = a.tmap_in; g.tmap_out = a.tmap_out;
    for (int i = 0; i < (1 << UA_MAX_SCATTER_BITS); ++i) g.tmap_dst[i] = a.tmap_dst[i];
}

static int cluster_knob(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
static bool cluster_swizzle() {
    static int v = -1;
    if (v < 0) v = cluster_knob("UA_CLUSTER_SWZ", 1);
    return v != 0;
}

template <int TEAMS, bool SCATTER, int ARITH, bool SWZ>
static int launch_ring(ClusterGeom &g, const ClusterArgs &ca, int want_buf, cudaStream_t st) {
    auto kern = cluster_ring_kernel<TEAMS, SCATTER, ARITH, SWZ>;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static bool attr_set[64] = {};
    static size_t dyn_limit[64] = {};
    const int di = (dev >= 0 && dev < 64) ? dev : 0;
    if (dev != di || !attr_set[di]) {
        int optin = 0;
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kern);
        if (e == cudaSuccess) {
            dyn_limit[di] = (size_t)optin - fa.sharedSizeBytes;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_limit[di]);
        }
        if (e != cudaSuccess) { set_error("cluster pass: cannot raise shared memory limit: %s", cudaGetErrorString(e)); cudaGetLastError(); return UA_ERR_CUDA; }
        attr_set[di] = true;
    }
    // Dynamic shared memory starts <= 2 KiB into the SM's window (1 KiB reserved + the static
    // mbarriers).  The ring starts at the next multiple of the buffer size; a large buffer leaves
    // room for the offset table in front of it, otherwise the table follows the ring.
    size_t buf_bytes = ((size_t)1 << g.T) * 8;
    if (buf_bytes < 1024) buf_bytes = 1024;            // swizzle atoms are 1 KiB
    const size_t tab_bytes = (size_t)CL_TAB * 256 * sizeof(unsigned short);
    const size_t limit = dyn_limit[di];                 // opt-in maximum minus the static part
    const bool front = buf_bytes >= tab_bytes + 2048;
    // front: ring occupies window [buf_bytes, (nbuf + 1) * buf_bytes); dynamic size counted from
    // an (unknown, >= 1 KiB) start: ask for (nbuf + 1) * buf_bytes - 1024
    int nbuf = front ? (int)((limit + 1024) / buf_bytes) - 1 : (int)((limit - tab_bytes) / buf_bytes) - 1;
    if (nbuf > 8) nbuf = 8;
    if (want_buf >= 1 && want_buf < nbuf) nbuf = want_buf;
    // The ring length must be a multiple of the team count: a buffer then belongs to ONE team for
    // the whole pass.  The mbarrier waits are parity-based; if a buffer were handed from team to
    // team, a fast team could ask for phase k+1 of a barrier whose phase k load has not landed yet
    // (TMA loads complete out of order) and the parity test would pass one phase early.
    nbuf -= nbuf % TEAMS;
    if (nbuf < 2 * TEAMS) { set_error("cluster pass: %d tile buffers of %zu bytes do not fit", 2 * TEAMS, buf_bytes); return UA_ERR_UNSUPPORTED; }
    g.nbuf = nbuf;
    g.tab_front = front ? 1 : 0;
    g.tab_bytes = (int)tab_bytes;
    const size_t smem = front ? (size_t)(nbuf + 1) * buf_bytes - 1024 : (size_t)(nbuf + 1) * buf_bytes + tab_bytes;
    long long grid = sms;
    if (grid > g.num_tiles) grid = g.num_tiles;
    kern<<<(unsigned)grid, TEAMS * 256 + 32, smem, st>>>(g, ca);
    return check_launch("cluster_ring_kernel");
}

// knobs (read once): UA_CLUSTER_TEAMS 1..3 (default 3), UA_CLUSTER_NBUF (default: all that fit),
// UA_CLUSTER_ARITH 0 = scalar FFMA / 1 = packed FFMA2 (default), UA_CLUSTER_SWZ (default 1)
template <bool SCATTER>
static int launch_cluster(const FusedArgs &a, const ClusterArgs &ca, cudaStream_t st) {
    static thread_local ClusterGeom g;
    fill_cluster_geom(g, a);
    static int teams = -1, arith = -1, nbuf = -1;
    if (teams < 0) {
        teams = cluster_knob("UA_CLUSTER_TEAMS", 3);
        arith = cluster_knob("UA_CLUSTER_ARITH", 1);
        nbuf = cluster_knob("UA_CLUSTER_NBUF", 0);
    }
    const size_t tile_bytes = ((size_t)1 << g.T) * 8;
    int t = teams;
    while (t > 1 && (size_t)(2 * t + 1) * tile_bytes > (size_t)227 * 1024 + 768) --t;    // two buffers per team
    const bool swz = cluster_swizzle();
#define UA_RING(T_, A_, S_) return launch_ring<T_, SCATTER, A_, S_>(g, ca, nbuf, st)
    if (arith == 0) {
        if (swz) { if (t >= 3) UA_RING(3, 0, true); if (t == 2) UA_RING(2, 0, true); UA_RING(1, 0, true); }
        if (t >= 3) UA_RING(3, 0, false); if (t == 2) UA_RING(2, 0, false); UA_RING(1, 0, false);
    }
    if (swz) { if (t >= 3) UA_RING(3, 1, true); if (t == 2) UA_RING(2, 1, true); UA_RING(1, 1, true); }
    if (t >= 3) UA_RING(3, 1, false); if (t == 2) UA_RING(2, 1, false); UA_RING(1, 1, false);
#undef UA_RING
}

}  // namespace ua

using namespace ua;

/* debug hook (not part of the public header): record per-CTA phase timestamps of the next fused
 * passes into a device buffer of gridDim * 16 * 4 u64 (load issued / tile landed / gates done /
 * previous store drained); null switches it off */
extern "C" int ua_debug_set_fused_trace(void *device_buffer) {
    g_trace_ptr = reinterpret_cast<unsigned long long *>(device_buffer);
    return UA_OK;
}

extern "C" int ua_fused_limits(int dtype, int *max_tile_bits_out, int *max_matrix_elems_out) {
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("ua_fused_limits: bad dtype"); return UA_ERR_INVALID; }
    if (max_tile_bits_out) *max_tile_bits_out = max_tile_bits(dtype);
    if (max_matrix_elems_out) *max_matrix_elems_out = FUSED_MAX_MAT_ELEMS;
    return UA_OK;
}

static int fill_scatter_args(FusedArgs &a, const char *who, int total_bits, int num_scatter_bits,
                             const int *host_scatter_pos, void *const *host_dst_ptrs, int visit_xor);

// Validate a pass description and fill the kernel arguments shared by the forward and the
// backward pass (geometry, gate descriptors).  max_k: largest gate the caller's kernel handles.
static int fill_fused_args(FusedArgs &a, const char *who, int dtype, void *out, const void *in,
                           long long total_amps, int total_bits, int tile_low_bits, int num_high,
                           const int *host_high_pos, int num_gates, const int *host_gate_k,
                           const int *host_gate_bits, const long long *host_gate_offset,
                           const void *gate_mats, long long gate_row_stride, int adjoint, int max_k,
                           int *mat_elems_out) {
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("%s: bad dtype", who); return UA_ERR_INVALID; }
    if (!out || !in || (num_high > 0 && !host_high_pos) ||
        (num_gates > 0 && (!gate_mats || !host_gate_k || !host_gate_bits || !host_gate_offset))) {
        set_error("%s: null pointer", who); return UA_ERR_INVALID;
    }
    if (((uintptr_t)out | (uintptr_t)in) & 15) { set_error("%s: state pointers must be 16-byte aligned", who); return UA_ERR_INVALID; }
    const int L = tile_low_bits, H = num_high, T = L + H;
    const int min_low = (dtype == UA_C64) ? 1 : 0;
    if (total_bits < 1 || total_bits > 48 || L < min_low || H < 0 || T > total_bits || T > max_tile_bits(dtype)) {
        set_error("%s: bad tile geometry total_bits=%d L=%d H=%d", who, total_bits, L, H); return UA_ERR_INVALID;
    }
    if (num_gates < (max_k < 0 ? 0 : 1) || num_gates > UA_MAX_FUSED_GATES) { set_error("%s: num_gates=%d out of range", who, num_gates); return UA_ERR_INVALID; }
    if (max_k < 0) max_k = -max_k;          // negative max_k: a pass without gates (pure copy) is allowed
    const long long space = 1ll << total_bits;
    if (total_amps < space || total_amps % space != 0) { set_error("%s: total_amps must be a multiple of 2^total_bits", who); return UA_ERR_INVALID; }
    const long long rows = total_amps >> total_bits;

    a.in = in; a.out = out; a.mats = gate_mats; a.mats_row_stride = gate_row_stride;
    a.total_bits = total_bits; a.T = T; a.L = L; a.H = H;
    a.num_gates = num_gates; a.adjoint = adjoint ? 1 : 0;
    int local_of[64];
    for (int p = 0; p < 64; ++p) local_of[p] = (p < L) ? p : -1;
    int prev = L - 1;
    for (int i = 0; i < H; ++i) {
        const int p = host_high_pos[i];
        if (p <= prev || p >= total_bits) { set_error("%s: high positions must be ascending in [L, total_bits)", who); return UA_ERR_INVALID; }
        a.high[i] = p; local_of[p] = L + i; prev = p;
    }
    a.tiles_per_row = 1ll << (total_bits - T);
    a.num_tiles = a.tiles_per_row * rows;

    int mat_elems = 0;
    for (int g = 0; g < num_gates; ++g) {
        const int k = host_gate_k[g];
        if (k < 1 || k > max_k) { set_error("%s: gate %d has k=%d (1..%d supported)", who, g, k, max_k); return UA_ERR_UNSUPPORTED; }
        if (k > T) { set_error("%s: gate %d has more qubits than the tile", who, g); return UA_ERR_INVALID; }
        FusedGate &gd = a.gates[g];
        gd.k = (unsigned char)k;
        gd.goff = host_gate_offset[g];
        gd.smoff = (unsigned short)mat_elems;
        mat_elems += 1 << (2 * k);
        int lb[3], order[3];
        for (int j = 0; j < k; ++j) {
            const int p = host_gate_bits[g * 3 + j];
            if (p < 0 || p >= total_bits || local_of[p] < 0) { set_error("%s: gate %d bit %d is outside the tile", who, g, p); return UA_ERR_INVALID; }
            lb[j] = local_of[p]; order[j] = j;
            for (int jj = 0; jj < j; ++jj) if (lb[jj] == lb[j]) { set_error("%s: gate %d repeats a bit", who, g); return UA_ERR_INVALID; }
        }
        for (int i = 1; i < k; ++i)
            for (int j = i; j > 0 && lb[order[j]] < lb[order[j - 1]]; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
        for (int i = 0; i < k; ++i) { gd.sb[i] = (unsigned char)lb[order[i]]; gd.gb[i] = (unsigned char)(k - 1 - order[i]); }
    }
    if (mat_elems > FUSED_MAX_MAT_ELEMS) { set_error("%s: %d matrix elements exceed the %d limit", who, mat_elems, FUSED_MAX_MAT_ELEMS); return UA_ERR_INVALID; }
    *mat_elems_out = mat_elems;
    return UA_OK;
}

extern "C" int ua_apply_fused_pass(int dtype, void *out, const void *in, long long total_amps,
                                   int total_bits, int tile_low_bits, int num_high,
                                   const int *host_high_pos, int num_gates, const int *host_gate_k,
                                   const int *host_gate_bits, const long long *host_gate_offset,
                                   const void *gate_mats, long long gate_row_stride, int adjoint,
                                   void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FusedArgs a{};
    int mat_elems = 0;
    const int rc = fill_fused_args(a, "ua_apply_fused_pass", dtype, out, in, total_amps, total_bits, tile_low_bits,
                                   num_high, host_high_pos, num_gates, host_gate_k, host_gate_bits,
                                   host_gate_offset, gate_mats, gate_row_stride, adjoint, 3, &mat_elems);
    if (rc) return rc;
    const int T = a.T;

    a.trank = 0;
    setup_tensor_maps(a, dtype == UA_C64 ? 0 : 1, total_amps);
    const size_t csize = (dtype == UA_C64) ? 8 : 16;
    const size_t tile_bytes = ((size_t)1 << T) * csize;
    const size_t mat_bytes = (((size_t)mat_elems * csize) + 127) & ~(size_t)127;
    // 256-thread CTAs run two per SM (phases of the two interleave); fall back to one
    // 512-thread CTA when two tiles do not fit
    // measured best (profiles/): complex64 3 x 256-thread register-lean CTAs per SM ("768"),
    // complex128 3 x 128-thread CTAs
    // team kernel: shared gates, TMA tensor path, three tiles + matrices fit
    // (measured equal to the 3-CTA variant, profiles/: both sit at ~78 % of the shared-memory +
    // FMA bound of the gate phase, so it stays opt-in)
    if (env_int("UA_FUSED_TEAM", 0) && gate_row_stride == 0 && a.trank > 0 &&
        3 * tile_bytes + mat_bytes + 1024 <= (size_t)226 * 1024 && a.num_tiles >= 1) {
        const size_t smem = 3 * tile_bytes + mat_bytes;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        long long grid = sms;
        if (grid > a.num_tiles) grid = a.num_tiles;
        if (dtype == UA_C64) {
            static bool s64 = false;
            if (!s64) { cudaFuncSetAttribute(fused_pass_team_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); s64 = true; }
            fused_pass_team_kernel<float><<<(unsigned)grid, 512, smem, st>>>(a);
        } else {
            static bool s128 = false;
            if (!s128) { cudaFuncSetAttribute(fused_pass_team_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); s128 = true; }
            fused_pass_team_kernel<double><<<(unsigned)grid, 512, smem, st>>>(a);
        }
        return check_launch("fused_pass_team_kernel");
    }
    int threads = env_int("UA_FUSED_THREADS", dtype == UA_C64 ? 768 : 128);
    if (threads != 512 && 2 * (tile_bytes + mat_bytes) > (size_t)222 * 1024) threads = 512;
    if (threads == 768) {     // 3 CTAs x 256 threads per SM, register-lean build
        if (dtype == UA_C64) return launch_fused<float, 256, 3>(a, tile_bytes, mat_bytes, st);
        return launch_fused<double, 256, 3>(a, tile_bytes, mat_bytes, st);
    }
    if (dtype == UA_C64) {
        if (threads == 128) return launch_fused<float, 128>(a, tile_bytes, mat_bytes, st);
        if (threads == 256) return launch_fused<float, 256>(a, tile_bytes, mat_bytes, st);
        return launch_fused<float, 512>(a, tile_bytes, mat_bytes, st);
    }
    if (threads == 128) return launch_fused<double, 128>(a, tile_bytes, mat_bytes, st);
    if (threads == 256) return launch_fused<double, 256>(a, tile_bytes, mat_bytes, st);
    return launch_fused<double, 512>(a, tile_bytes, mat_bytes, st);
}

extern "C" int ua_fused_backward_pass(int dtype, void *psi, void *grad, long long total_amps,
                                      int total_bits, int tile_low_bits, int num_high,
                                      const int *host_high_pos, int num_gates, const int *host_gate_k,
                                      const int *host_gate_bits, const long long *host_gate_offset,
                                      const void *gate_mats, long long gate_row_stride,
                                      const int *host_gate_needs_grad, void *grad_acc, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    BwdArgs ba{};
    int mat_elems = 0;
    if (!host_gate_needs_grad || !grad_acc) { set_error("ua_fused_backward_pass: null pointer"); return UA_ERR_INVALID; }
    if (psi == grad) { set_error("ua_fused_backward_pass: psi and grad must be different buffers"); return UA_ERR_INVALID; }
    const int rc = fill_fused_args(ba.f, "ua_fused_backward_pass", dtype, grad, psi, total_amps, total_bits,
                                   tile_low_bits, num_high, host_high_pos, num_gates, host_gate_k,
                                   host_gate_bits, host_gate_offset, gate_mats, gate_row_stride, 1, 2, &mat_elems);
    if (rc) return rc;
    ba.f.trank = 0;
    setup_tensor_maps(ba.f, dtype == UA_C64 ? 0 : 1, total_amps);
    ba.acc = reinterpret_cast<double2 *>(grad_acc);
    ba.mat_elems = mat_elems;
    ba.needs_grad = 0;
    for (int g = 0; g < num_gates; ++g)
        if (host_gate_needs_grad[g]) ba.needs_grad |= 1ull << g;
    const size_t csize = (dtype == UA_C64) ? 8 : 16;
    const size_t tile_bytes = ((size_t)1 << ba.f.T) * csize;
    const size_t mat_bytes = (((size_t)mat_elems * csize) + 127) & ~(size_t)127;
    const size_t smem = 2 * tile_bytes + mat_bytes + (size_t)mat_elems * sizeof(double2);
    if (smem > 200 * 1024) { set_error("ua_fused_backward_pass: tile too large (%zu bytes of shared memory)", smem); return UA_ERR_UNSUPPORTED; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 0;
    cudaError_t e;
    if (dtype == UA_C64) {
        static bool set64 = false;
        if (!set64) { cudaFuncSetAttribute(fused_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); set64 = true; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_bwd_kernel<float>, 256, smem);
    } else {
        static bool set128 = false;
        if (!set128) { cudaFuncSetAttribute(fused_bwd_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); set128 = true; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_bwd_kernel<double>, 256, smem);
    }
    if (e != cudaSuccess || per_sm < 1) { set_error("ua_fused_backward_pass: occupancy query failed (%s), smem=%zu", cudaGetErrorString(e), smem); cudaGetLastError(); return UA_ERR_CUDA; }
    long long grid = (long long)per_sm * sms;
    if (grid > ba.f.num_tiles) grid = ba.f.num_tiles;
    if (dtype == UA_C64) fused_bwd_kernel<float><<<(unsigned)grid, 256, smem, st>>>(ba);
    else fused_bwd_kernel<double><<<(unsigned)grid, 256, smem, st>>>(ba);
    return check_launch("fused_bwd_kernel");
}

extern "C" int ua_apply_fused_pass_scatter(int dtype, const void *in, long long total_amps, int total_bits,
                                           int tile_low_bits, int num_high, const int *host_high_pos,
                                           int num_gates, const int *host_gate_k, const int *host_gate_bits,
                                           const long long *host_gate_offset, const void *gate_mats,
                                           int num_scatter_bits, const int *host_scatter_pos,
                                           void *const *host_dst_ptrs, int visit_xor, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const char *who = "ua_apply_fused_pass_scatter";
    if (num_scatter_bits < 1 || num_scatter_bits > UA_MAX_SCATTER_BITS || !host_scatter_pos || !host_dst_ptrs) {
        set_error("%s: num_scatter_bits=%d out of range (1..%d) or null pointer", who, num_scatter_bits, UA_MAX_SCATTER_BITS);
        return UA_ERR_INVALID;
    }
    if (total_amps != (1ll << total_bits)) { set_error("%s: one state only (total_amps must be 2^total_bits)", who); return UA_ERR_INVALID; }
    FusedArgs a{};
    int mat_elems = 0;
    // `out` is only validated for alignment: pass the first destination
    const int rc = fill_fused_args(a, who, dtype, host_dst_ptrs[0], in, total_amps, total_bits, tile_low_bits,
                                   num_high, host_high_pos, num_gates, host_gate_k, host_gate_bits,
                                   host_gate_offset, gate_mats, 0, 0, -3, &mat_elems);
    if (rc) return rc;
    {
        const int rcs = fill_scatter_args(a, who, total_bits, num_scatter_bits, host_scatter_pos, host_dst_ptrs, visit_xor);
        if (rcs) return rcs;
    }
    a.trank = 0;
    setup_tensor_maps(a, dtype == UA_C64 ? 0 : 1, total_amps);
    const size_t csize = (dtype == UA_C64) ? 8 : 16;
    const size_t tile_bytes = ((size_t)1 << a.T) * csize;
    const size_t mat_bytes = (((size_t)mat_elems * csize) + 127) & ~(size_t)127;
    if (dtype == UA_C64) return launch_fused<float, 256, 3, true>(a, tile_bytes, mat_bytes, st);
    return launch_fused<double, 128, 4, true>(a, tile_bytes, mat_bytes, st);
}

// fill the scatter fields of `a` (after fill_fused_args); shared by both scatter entry points
static int fill_scatter_args(FusedArgs &a, const char *who, int total_bits, int num_scatter_bits,
                             const int *host_scatter_pos, void *const *host_dst_ptrs, int visit_xor) {
    a.scatter_m = num_scatter_bits;
    for (int j = 0; j < num_scatter_bits; ++j) {
        const int v = host_scatter_pos[j];
        if (v < a.L || v >= total_bits || (j > 0 && v <= host_scatter_pos[j - 1])) {
            set_error("%s: scatter positions must be ascending in [tile_low_bits, total_bits)", who); return UA_ERR_INVALID;
        }
        for (int i = 0; i < a.H; ++i)
            if (a.high[i] == v) { set_error("%s: scatter bit %d is a tile bit", who, v); return UA_ERR_INVALID; }
        a.vpos[j] = v;
        if ((visit_xor >> j) & 1) a.tile_xor |= 1ull << v;
    }
    {   // merged ascending list of the tile's high bits and the scatter bits
        int i = 0, j = 0;
        a.nins = 0;
        while (i < a.H || j < num_scatter_bits) {
            if (j >= num_scatter_bits || (i < a.H && a.high[i] < a.vpos[j])) a.ins[a.nins++] = a.high[i++];
            else a.ins[a.nins++] = a.vpos[j++];
        }
    }
    if (a.T > total_bits - num_scatter_bits) { set_error("%s: tile larger than the destination blocks", who); return UA_ERR_INVALID; }
    for (int b = 0; b < (1 << num_scatter_bits); ++b) {
        if (!host_dst_ptrs[b] || ((uintptr_t)host_dst_ptrs[b] & 15)) { set_error("%s: destination %d is null or misaligned", who, b); return UA_ERR_INVALID; }
        a.dst[b] = host_dst_ptrs[b];
    }
    return UA_OK;
}

extern "C" int ua_apply_fused_pass_hostmats(int dtype, void *out, const void *in, long long total_amps,
                                            int total_bits, int tile_low_bits, int num_high,
                                            const int *host_high_pos, int num_gates, const int *host_gate_k,
                                            const int *host_gate_bits, const long long *host_gate_offset,
                                            const void *host_gate_mats, int adjoint, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const char *who = "ua_apply_fused_pass_hostmats";
    if (dtype != UA_C64) { set_error("%s: complex64 only", who); return UA_ERR_UNSUPPORTED; }
    FusedArgs a{};
    int mat_elems = 0;
    const int rc = fill_fused_args(a, who, dtype, out, in, total_amps, total_bits, tile_low_bits, num_high,
                                   host_high_pos, num_gates, host_gate_k, host_gate_bits, host_gate_offset,
                                   host_gate_mats, 0, adjoint, 2, &mat_elems);
    if (rc) return rc;
    static thread_local ClusterArgs ca;
    if (!build_clusters(a, reinterpret_cast<const float2 *>(host_gate_mats), ca, cluster_swizzle())) {
        set_error("%s: the pass does not fit the register-blocked path (tile of %d bits)", who, a.T);
        return UA_ERR_UNSUPPORTED;
    }
    if (getenv("UA_CLUSTER_DEBUG")) {
        fprintf(stderr, "cluster pass: T=%d gates=%d clusters=%d:", a.T, a.num_gates, ca.ncl);
        for (int c = 0; c < ca.ncl; ++c)
            fprintf(stderr, " [%d %d %d %d | %d gates]", ca.cl[c].cb[0], ca.cl[c].cb[1], ca.cl[c].cb[2], ca.cl[c].cb[3],
                    ca.cl[c].gend - ca.cl[c].gbeg);
        fprintf(stderr, "\n");
    }
    a.mats = nullptr;
    a.trank = 0;
    if (!setup_tensor_maps(a, 0, total_amps, cluster_swizzle())) { set_error("%s: the tile needs more than 5 TMA dimensions", who); return UA_ERR_UNSUPPORTED; }
    return launch_cluster<false>(a, ca, st);
}

extern "C" int ua_apply_fused_pass_scatter_hostmats(int dtype, const void *in, long long total_amps, int total_bits,
                                                    int tile_low_bits, int num_high, const int *host_high_pos,
                                                    int num_gates, const int *host_gate_k, const int *host_gate_bits,
                                                    const long long *host_gate_offset, const void *host_gate_mats,
                                                    int num_scatter_bits, const int *host_scatter_pos,
                                                    void *const *host_dst_ptrs, int visit_xor, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const char *who = "ua_apply_fused_pass_scatter_hostmats";
    if (dtype != UA_C64) { set_error("%s: complex64 only", who); return UA_ERR_UNSUPPORTED; }
    if (num_scatter_bits < 1 || num_scatter_bits > UA_MAX_SCATTER_BITS || !host_scatter_pos || !host_dst_ptrs) {
        set_error("%s: num_scatter_bits=%d out of range (1..%d) or null pointer", who, num_scatter_bits, UA_MAX_SCATTER_BITS);
        return UA_ERR_INVALID;
    }
    if (total_amps != (1ll << total_bits)) { set_error("%s: one state only (total_amps must be 2^total_bits)", who); return UA_ERR_INVALID; }
    if (num_gates < 1) { set_error("%s: needs at least one gate (use ua_apply_fused_pass_scatter for a pure copy)", who); return UA_ERR_INVALID; }
    FusedArgs a{};
    int mat_elems = 0;
    int rc = fill_fused_args(a, who, dtype, host_dst_ptrs[0], in, total_amps, total_bits, tile_low_bits,
                             num_high, host_high_pos, num_gates, host_gate_k, host_gate_bits,
                             host_gate_offset, host_gate_mats, 0, 0, 2, &mat_elems);
    if (rc) return rc;
    rc = fill_scatter_args(a, who, total_bits, num_scatter_bits, host_scatter_pos, host_dst_ptrs, visit_xor);
    if (rc) return rc;
    static thread_local ClusterArgs ca;
    if (!build_clusters(a, reinterpret_cast<const float2 *>(host_gate_mats), ca, cluster_swizzle())) {
        set_error("%s: the pass does not fit the register-blocked path (tile of %d bits)", who, a.T);
        return UA_ERR_UNSUPPORTED;
    }
    a.mats = nullptr;
    a.trank = 0;
    if (!setup_tensor_maps(a, 0, total_amps, cluster_swizzle())) { set_error("%s: the tile needs more than 5 TMA dimensions", who); return UA_ERR_UNSUPPORTED; }
    return launch_cluster<true>(a, ca, st);
}
