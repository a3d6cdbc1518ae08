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
#include "ua_tile.cuh"

namespace ua {

// One gate on the whole tile.  K qubits; LOW: (complex64 only) the lowest target is local
// bit 0, i.e. inside the float4.  `tv` is the tile as 16-byte vectors, TV = log2 of their count.
// GU groups are processed together for memory-level parallelism.  LEAN (3 CTAs x 256 threads per
// SM, <= 80 registers): complex128 2-qubit matrices stay in shared memory.
template <typename R, int K, bool LOW, bool LEAN>
__device__ __forceinline__ void apply_gate_smem(typename VecOf<R>::type *tv,
                                                const typename CplxOf<R>::type *M,
                                                const FusedGate &gd, int TV, int nthreads) {
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
#pragma unroll
    for (int i = 0; i < KH; ++i) {
        vb[i] = (int)gd.sb[i + (LOW ? 1 : 0)] - APVLOG;
        off[i] = 1u << vb[i];
    }
    constexpr bool MREG = LEAN ? (K <= (sizeof(R) == 4 ? 2 : 1)) : (K <= 2);
    C mr[MREG ? D * D : 1];
    if (MREG) {
#pragma unroll
        for (int e = 0; e < D * D; ++e) mr[e] = M[e];
    }
    auto Mat = [&](int s, int t) -> C { return MREG ? mr[s * D + t] : M[s * D + t]; };

    const unsigned groups = 1u << (TV - KH);
    for (unsigned g0 = threadIdx.x; g0 < groups; g0 += nthreads * GU) {
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
                unsigned idx = b;
#pragma unroll
                for (int i = 0; i < KH; ++i)
                    if ((c >> i) & 1) idx |= off[i];
                if (ok[u]) x[u][c] = tv[idx];
            }
        }
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            if (!ok[u]) continue;
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
                // every member is already in registers: store right away
                unsigned idx = gbase[u];
#pragma unroll
                for (int i = 0; i < KH; ++i)
                    if ((ov >> i) & 1) idx |= off[i];
                tv[idx] = res;
            }
        }
    }
}

template <typename R, int NT, bool LEAN>
__device__ __forceinline__ void apply_any_gate(typename VecOf<R>::type *tv,
                                               const typename CplxOf<R>::type *M,
                                               const FusedGate &gd, int TV) {
    constexpr int APV = VecOf<R>::APV;
    if constexpr (APV == 2) {
        if (gd.sb[0] == 0) {
            if (gd.k == 1) apply_gate_smem<R, 1, true, LEAN>(tv, M, gd, TV, NT);
            else if (gd.k == 2) apply_gate_smem<R, 2, true, LEAN>(tv, M, gd, TV, NT);
            else apply_gate_smem<R, 3, true, LEAN>(tv, M, gd, TV, NT);
            return;
        }
    }
    if (gd.k == 1) apply_gate_smem<R, 1, false, LEAN>(tv, M, gd, TV, NT);
    else if (gd.k == 2) apply_gate_smem<R, 2, false, LEAN>(tv, M, gd, TV, NT);
    else apply_gate_smem<R, 3, false, LEAN>(tv, M, gd, TV, NT);
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
        mbar_wait(smem_u32(&bars[s]), parity);
        V *tv = reinterpret_cast<V *>(smem_raw + (size_t)s * tile_bytes);
        for (int g = 0; g < a.num_gates; ++g) {
            const FusedGate &gd = a.gates[g];
            apply_any_gate<R, FUSED_THREADS, LEAN>(tv, sM + gd.smoff, gd, TV);
            if (g + 1 < a.num_gates) __syncthreads();
        }
        fence_proxy_async();        // make the generic-proxy writes visible to the copy engine
        __syncthreads();
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
    // register-blocked form (complex64): the gates, walked in reverse, are cut into runs whose
    // targets fit four tile bits; a thread keeps the 16 psi and 16 grad amplitudes of one group
    // in registers for the whole run (one shared-memory round trip and one barrier per run
    // instead of per gate).  ncl == 0: per-gate form.
    int ncl;
    unsigned char cl_first[UA_MAX_FUSED_GATES];    // highest gate index of the run
    unsigned char cl_count[UA_MAX_FUSED_GATES];    // gates cl_first, cl_first - 1, ...
    unsigned char cl_bits[UA_MAX_FUSED_GATES][4];  // ascending tile-local bits of the run
    unsigned char gtype[UA_MAX_FUSED_GATES];       // 0..5: pair (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) of the run's bits, 6..9: single bit
};

// ---- register-blocked backward gates (complex64).  M = U^H in target-bit order (shared memory);
//      ax/ay collect Re / Im of grad[a][b] += g[a] * conj(psi_in[b]) over the thread's group.
template <int I, bool GRAD>
__device__ __forceinline__ void bwd_reg_gate1(float2 (&vp)[16], float2 (&vg)[16], const float2 *__restrict__ M,
                                              float (&ax)[16], float (&ay)[16]) {
    const float2 m00 = M[0], m01 = M[1], m10 = M[2], m11 = M[3];
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) {
        const int lo = gi & ((1 << I) - 1);
        const int b0 = ((gi >> I) << (I + 1)) | lo, b1 = b0 | (1 << I);
        const float2 x0 = vp[b0], x1 = vp[b1], y0 = vg[b0], y1 = vg[b1];
        float2 p0 = mk(0.f, 0.f), p1 = p0, q0 = p0, q1 = p0;
        cfma(p0, m00, x0); cfma(p0, m01, x1);
        cfma(p1, m10, x0); cfma(p1, m11, x1);
        cfma(q0, m00, y0); cfma(q0, m01, y1);
        cfma(q1, m10, y0); cfma(q1, m11, y1);
        if constexpr (GRAD) {
            const float2 y[2] = {y0, y1}, xin[2] = {p0, p1};
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    ax[a * 2 + b] = fmaf(y[a].x, xin[b].x, ax[a * 2 + b]);
                    ax[a * 2 + b] = fmaf(y[a].y, xin[b].y, ax[a * 2 + b]);
                    ay[a * 2 + b] = fmaf(y[a].y, xin[b].x, ay[a * 2 + b]);
                    ay[a * 2 + b] = fmaf(-y[a].x, xin[b].y, ay[a * 2 + b]);
                }
        }
        vp[b0] = p0; vp[b1] = p1; vg[b0] = q0; vg[b1] = q1;
    }
}

template <int I, int J, bool GRAD>
__device__ __forceinline__ void bwd_reg_gate2(float2 (&vp)[16], float2 (&vg)[16], const float2 *__restrict__ M,
                                              float (&ax)[16], float (&ay)[16]) {
    constexpr int OTHERS = 0xF & ~((1 << I) | (1 << J));
    constexpr int O0 = (OTHERS & 1) ? 0 : (OTHERS & 2) ? 1 : (OTHERS & 4) ? 2 : 3;
    constexpr int O1 = (OTHERS & 8) ? 3 : (OTHERS & 4) ? 2 : (OTHERS & 2) ? 1 : 0;
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
        const int base = ((gi & 1) << O0) | ((gi >> 1) << O1);
        float2 x[4], y[4], xin[4], yin[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            x[c] = vp[base | ((c & 1) << I) | ((c >> 1) << J)];
            y[c] = vg[base | ((c & 1) << I) | ((c >> 1) << J)];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float2 a = mk(0.f, 0.f), b = a;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float2 m = M[r * 4 + c];          // shared-memory broadcast, not held in registers
                cfma(a, m, x[c]);
                cfma(b, m, y[c]);
            }
            xin[r] = a; yin[r] = b;
        }
        if constexpr (GRAD) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    ax[a * 4 + b] = fmaf(y[a].x, xin[b].x, ax[a * 4 + b]);
                    ax[a * 4 + b] = fmaf(y[a].y, xin[b].y, ax[a * 4 + b]);
                    ay[a * 4 + b] = fmaf(y[a].y, xin[b].x, ay[a * 4 + b]);
                    ay[a * 4 + b] = fmaf(-y[a].x, xin[b].y, ay[a * 4 + b]);
                }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            vp[base | ((c & 1) << I) | ((c >> 1) << J)] = xin[c];
            vg[base | ((c & 1) << I) | ((c >> 1) << J)] = yin[c];
        }
    }
}

template <bool GRAD>
__device__ __forceinline__ void bwd_reg_dispatch(float2 (&vp)[16], float2 (&vg)[16], int type, const float2 *__restrict__ M,
                                                 float (&ax)[16], float (&ay)[16]) {
    switch (type) {
        case 0: bwd_reg_gate2<0, 1, GRAD>(vp, vg, M, ax, ay); break;
        case 1: bwd_reg_gate2<0, 2, GRAD>(vp, vg, M, ax, ay); break;
        case 2: bwd_reg_gate2<0, 3, GRAD>(vp, vg, M, ax, ay); break;
        case 3: bwd_reg_gate2<1, 2, GRAD>(vp, vg, M, ax, ay); break;
        case 4: bwd_reg_gate2<1, 3, GRAD>(vp, vg, M, ax, ay); break;
        case 5: bwd_reg_gate2<2, 3, GRAD>(vp, vg, M, ax, ay); break;
        case 6: bwd_reg_gate1<0, GRAD>(vp, vg, M, ax, ay); break;
        case 7: bwd_reg_gate1<1, GRAD>(vp, vg, M, ax, ay); break;
        case 8: bwd_reg_gate1<2, GRAD>(vp, vg, M, ax, ay); break;
        default: bwd_reg_gate1<3, GRAD>(vp, vg, M, ax, ay); break;
    }
}

// Sum N per-lane values over the warp; lane l ends up with the total of value number
// (l >> (5 - log2 N)) (the lanes sharing those top bits all hold it).  N + log2(32 / N) - 1 shuffles.
template <int N>
__device__ __forceinline__ float warp_reduce_scatter(float (&v)[N], unsigned lane) {
    int off = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, off >>= 1) {
        const bool up = (lane & (unsigned)off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
#pragma unroll
    for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    return v[0];
}

// One run of gates on the two tiles, all threads of the CTA.  The tiles were written by TMA with
// the 128-byte swizzle (element i sits at i ^ (((i >> 4) & 7) << 1)): consecutive lanes walk the
// lowest non-run bits, and the swizzle spreads them over the banks whatever the run's bits are.
// When bit 0 belongs to the run, members 2m and 2m+1 are one 16-byte access.
__device__ __forceinline__ unsigned bwd_phys(unsigned i) { return i ^ (((i >> 4) & 7u) << 1); }

__device__ __forceinline__ void bwd_cluster_run(float2 *tpsi, float2 *tg, const BwdArgs &ba, int c, const float2 *sM,
                                                double2 *sAcc, int T, int nthreads) {
    const int b0 = ba.cl_bits[c][0], b1 = ba.cl_bits[c][1], b2 = ba.cl_bits[c][2], b3 = ba.cl_bits[c][3];
    const bool vec16 = b0 == 0;
    const unsigned groups = 1u << (T - 4);
    const int first = ba.cl_first[c], count = ba.cl_count[c];
    for (unsigned g0 = 0; g0 < groups; g0 += nthreads) {
        const unsigned grp = g0 + threadIdx.x;
        const bool act = grp < groups;
        unsigned base = act ? grp : 0u;
        base = insert_zero32(base, b0);
        base = insert_zero32(base, b1);
        base = insert_zero32(base, b2);
        base = insert_zero32(base, b3);
        auto member = [&](int m) -> unsigned {
            return bwd_phys(base | ((m & 1) << b0) | (((m >> 1) & 1) << b1) | (((m >> 2) & 1) << b2) | (((m >> 3) & 1) << b3));
        };
        float2 vp[16], vg[16];
        if (vec16) {
#pragma unroll
            for (int m = 0; m < 16; m += 2) {
                const unsigned idx = member(m);
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (act) {
                    a = *reinterpret_cast<const float4 *>(tpsi + idx);
                    b = *reinterpret_cast<const float4 *>(tg + idx);
                }
                vp[m] = mk(a.x, a.y); vp[m + 1] = mk(a.z, a.w);
                vg[m] = mk(b.x, b.y); vg[m + 1] = mk(b.z, b.w);
            }
        } else {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const unsigned idx = member(m);
                vp[m] = act ? tpsi[idx] : mk(0.f, 0.f);
                vg[m] = act ? tg[idx] : mk(0.f, 0.f);
            }
        }
        for (int q = first; q > first - count; --q) {
            const FusedGate &gd = ba.f.gates[q];
            const float2 *M = sM + gd.smoff;
            const bool grad = (ba.needs_grad >> q) & 1ull;
            float ax[16], ay[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) { ax[e] = 0.f; ay[e] = 0.f; }
            if (grad) {
                bwd_reg_dispatch<true>(vp, vg, ba.gtype[q], M, ax, ay);
                // warp reduce-scatter (9 shuffles for the 8 sums of a 1-qubit gate instead of 40:
                // every step halves the number of values a lane carries), then the lanes that hold
                // the totals add them to the fp64 shared-memory accumulators in one go
                const unsigned lane = threadIdx.x & 31u;
                if (gd.k == 1) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { v[2 * e] = ax[e]; v[2 * e + 1] = ay[e]; }
                    const float tot = warp_reduce_scatter<8>(v, lane);
                    if ((lane & 3u) == 0) {
                        const unsigned i = lane >> 2;
                        double *dst = reinterpret_cast<double *>(&sAcc[gd.smoff + (i >> 1)]) + (i & 1u);
                        atomicAdd(dst, (double)tot);
                    }
                } else {
                    float v[32];
#pragma unroll
                    for (int e = 0; e < 16; ++e) { v[2 * e] = ax[e]; v[2 * e + 1] = ay[e]; }
                    const float tot = warp_reduce_scatter<32>(v, lane);
                    double *dst = reinterpret_cast<double *>(&sAcc[gd.smoff + (lane >> 1)]) + (lane & 1u);
                    atomicAdd(dst, (double)tot);
                }
            } else {
                bwd_reg_dispatch<false>(vp, vg, ba.gtype[q], M, ax, ay);
            }
        }
        if (act) {
            if (vec16) {
#pragma unroll
                for (int m = 0; m < 16; m += 2) {
                    const unsigned idx = member(m);
                    *reinterpret_cast<float4 *>(tpsi + idx) = make_float4(vp[m].x, vp[m].y, vp[m + 1].x, vp[m + 1].y);
                    *reinterpret_cast<float4 *>(tg + idx) = make_float4(vg[m].x, vg[m].y, vg[m + 1].x, vg[m + 1].y);
                }
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const unsigned idx = member(m);
                    tpsi[idx] = vp[m];
                    tg[idx] = vg[m];
                }
            }
        }
    }
}

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
    // the swizzled tiles of the register-blocked form need 1 KiB alignment (the launcher adds the slack)
    unsigned char *smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *buf_psi = smem_al, *buf_g = smem_al + (tile_bytes < 1024u ? 1024u : tile_bytes);
    unsigned char *after = buf_g + (tile_bytes < 1024u ? 1024u : tile_bytes);
    C *sM = reinterpret_cast<C *>(after);
    double2 *sAcc = reinterpret_cast<double2 *>(after + (((size_t)ba.mat_elems * sizeof(C) + 127) & ~(size_t)127));
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
        if constexpr (sizeof(R) == 4) {
            if (ba.ncl > 0) {
                for (int c = 0; c < ba.ncl; ++c) {
                    bwd_cluster_run(reinterpret_cast<float2 *>(buf_psi), reinterpret_cast<float2 *>(buf_g), ba, c,
                                    reinterpret_cast<const float2 *>(sM), sAcc, a.T, NT);
                    __syncthreads();
                }
            }
        }
        for (int g = (sizeof(R) == 4 && ba.ncl > 0) ? -1 : a.num_gates - 1; g >= 0; --g) {
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
// swizzle128 (ua_cluster.cu): the first dimension is exactly the 4 lowest element bits (a 128-byte
// row) and the box lands in shared memory with the 128-byte swizzle (16-byte chunk index ^= row
// index mod 8): the register-blocked gate phase then reads any 4-bit cluster without bank conflicts.
bool setup_tensor_maps(FusedArgs &a, int ebits, long long total_amps, bool swizzle128) {
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

template <typename R, int FUSED_THREADS, int MINB, bool SCATTER = false>
static int launch_fused(FusedArgs &a, size_t tile_bytes, size_t mat_bytes, cudaStream_t st) {
    auto kern = fused_pass_kernel<R, FUSED_THREADS, MINB, SCATTER>;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static bool attr_set[64] = {};            // the attribute is per device
    const int di = (dev >= 0 && dev < 64) ? dev : 0;
    if (dev != di || !attr_set[di]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { set_error("ua_apply_fused_pass: cannot raise shared memory limit: %s", cudaGetErrorString(e)); cudaGetLastError(); return UA_ERR_CUDA; }
        attr_set[di] = true;
    }
    // shared memory budget per CTA: up to MINB CTAs share an SM (fewer when the tiles are too large)
    int ctas = MINB;
    while (ctas > 1 && (size_t)(224 * 1024) / ctas - 1024 < tile_bytes + mat_bytes) --ctas;
    const size_t budget = (size_t)(224 * 1024) / ctas - 1024;
    int nstage = (budget > mat_bytes) ? (int)((budget - mat_bytes) / tile_bytes) : 0;
    if (nstage > 3) nstage = 3;
    if (nstage < 1) { set_error("ua_apply_fused_pass: tile does not fit in shared memory"); return UA_ERR_UNSUPPORTED; }
    a.nstage = nstage;
    const size_t smem = (size_t)nstage * tile_bytes + mat_bytes;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FUSED_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) { set_error("ua_apply_fused_pass: occupancy query failed (%s), smem=%zu", cudaGetErrorString(e), smem); cudaGetLastError(); return UA_ERR_CUDA; }
    long long grid = (long long)per_sm * sms;
    if (grid > a.num_tiles) grid = a.num_tiles;
    kern<<<(unsigned)grid, FUSED_THREADS, smem, st>>>(a);
    return check_launch("fused_pass_kernel");
}

}  // namespace ua

using namespace ua;

extern "C" int ua_fused_limits(int dtype, int *max_tile_bits_out, int *max_matrix_elems_out) {
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("ua_fused_limits: bad dtype"); return UA_ERR_INVALID; }
    if (max_tile_bits_out) *max_tile_bits_out = max_tile_bits(dtype);
    if (max_matrix_elems_out) *max_matrix_elems_out = FUSED_MAX_MAT_ELEMS;
    return UA_OK;
}

namespace ua {

// Validate a pass description and fill the kernel arguments shared by the forward and the
// backward pass (geometry, gate descriptors).  max_k: largest gate the caller's kernel handles.
int fill_fused_args(FusedArgs &a, const char *who, int dtype, void *out, const void *in,
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

int fill_scatter_args(FusedArgs &a, const char *who, int total_bits, int num_scatter_bits,
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

}  // namespace ua

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
    a.trank = 0;
    setup_tensor_maps(a, dtype == UA_C64 ? 0 : 1, total_amps);
    const size_t csize = (dtype == UA_C64) ? 8 : 16;
    const size_t tile_bytes = ((size_t)1 << a.T) * csize;
    const size_t mat_bytes = (((size_t)mat_elems * csize) + 127) & ~(size_t)127;
    // measured best (profiles/r01_tune_fused_c64_n30.json): complex64 3 x 256-thread register-lean
    // CTAs per SM, complex128 4 x 128-thread CTAs; one 512-thread CTA when two tiles do not fit
    const bool big = 2 * (tile_bytes + mat_bytes) > (size_t)222 * 1024;
    if (dtype == UA_C64) {
        if (big) return launch_fused<float, 512, 1>(a, tile_bytes, mat_bytes, st);
        return launch_fused<float, 256, 3>(a, tile_bytes, mat_bytes, st);
    }
    if (big) return launch_fused<double, 512, 1>(a, tile_bytes, mat_bytes, st);
    return launch_fused<double, 128, 4>(a, tile_bytes, mat_bytes, st);
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
    ba.acc = reinterpret_cast<double2 *>(grad_acc);
    ba.mat_elems = mat_elems;
    ba.needs_grad = 0;
    for (int g = 0; g < num_gates; ++g)
        if (host_gate_needs_grad[g]) ba.needs_grad |= 1ull << g;
    // complex64: cut the reversed gate list into runs of consecutive gates on <= 4 tile bits
    ba.ncl = 0;
    { const char *e = getenv("UA_BWD_CLUSTER");
      if (dtype == UA_C64 && ba.f.T >= 4 && !(e && atoi(e) == 0)) {
        int g = num_gates - 1;
        while (g >= 0) {
            unsigned mask = 0;
            const int first = g;
            while (g >= 0) {
                const FusedGate &gd = ba.f.gates[g];
                unsigned gm = 0;
                for (int i = 0; i < gd.k; ++i) gm |= 1u << gd.sb[i];
                if (__builtin_popcount(mask | gm) > 4) break;
                mask |= gm;
                --g;
            }
            // pad to four bits from the top of the tile: the thread index then runs over the low bits
            for (int b = ba.f.T - 1; b >= 0 && __builtin_popcount(mask) < 4; --b)
                if (!((mask >> b) & 1u)) mask |= 1u << b;
            const int c = ba.ncl++;
            ba.cl_first[c] = (unsigned char)first;
            ba.cl_count[c] = (unsigned char)(first - g);
            int pos_of[32], nb = 0;
            for (int b = 0; b < ba.f.T; ++b)
                if ((mask >> b) & 1u) { ba.cl_bits[c][nb] = (unsigned char)b; pos_of[b] = nb; ++nb; }
            for (int q = first; q > g; --q) {
                const FusedGate &gd = ba.f.gates[q];
                static const int pair_type[4][4] = {{-1, 0, 1, 2}, {-1, -1, 3, 4}, {-1, -1, -1, 5}, {-1, -1, -1, -1}};
                ba.gtype[q] = (unsigned char)(gd.k == 1 ? 6 + pos_of[gd.sb[0]] : pair_type[pos_of[gd.sb[0]]][pos_of[gd.sb[1]]]);
            }
        }
      } }
    // the register-blocked form works on 128-byte-swizzled tiles; when the tile does not fit a
    // swizzled tensor map (more than 5 dimensions with the first one fixed to bits 0..3) the pass
    // runs in the per-gate form on plain tiles
    ba.f.trank = 0;
    if (ba.ncl > 0 && !setup_tensor_maps(ba.f, 0, total_amps, true)) { ba.ncl = 0; ba.f.trank = 0; }
    if (ba.ncl == 0) setup_tensor_maps(ba.f, dtype == UA_C64 ? 0 : 1, total_amps);
    const size_t csize = (dtype == UA_C64) ? 8 : 16;
    size_t tile_bytes = ((size_t)1 << ba.f.T) * csize;
    if (tile_bytes < 1024) tile_bytes = 1024;
    const size_t mat_bytes = (((size_t)mat_elems * csize) + 127) & ~(size_t)127;
    const size_t smem = 2 * tile_bytes + mat_bytes + (size_t)mat_elems * sizeof(double2) + 1024;
    if (smem > 200 * 1024) { set_error("ua_fused_backward_pass: tile too large (%zu bytes of shared memory)", smem); return UA_ERR_UNSUPPORTED; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 0;
    cudaError_t e;
    if (dtype == UA_C64) {
        static bool set64[64] = {};
        if (!set64[dev & 63]) { cudaFuncSetAttribute(fused_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); set64[dev & 63] = true; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_bwd_kernel<float>, 256, smem);
    } else {
        static bool set128[64] = {};
        if (!set128[dev & 63]) { cudaFuncSetAttribute(fused_bwd_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); set128[dev & 63] = true; }
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

