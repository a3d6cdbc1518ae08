// Register-blocked fused pass (complex64, shared 1-/2-qubit gates whose matrix VALUES are known
// on the host): ua_apply_fused_pass_hostmats / ua_apply_fused_pass_scatter_hostmats.
//
// The pass's gate list is cut into clusters: runs of gates whose target bits all lie inside a
// set of FOUR tile bits.  A thread loads the 16 amplitudes of one group (all values of the 4
// cluster bits) from the tile once, applies every gate of the cluster to them in registers and
// stores them back: one shared-memory round trip and one barrier per CLUSTER instead of per
// gate.  The matrices travel in the kernel parameters (constant bank): the gate index is
// warp-uniform, so ptxas keeps them in UNIFORM registers (LDCU) and the packed FMAs (FFMA2) take
// them as UR-broadcast operands -- no vector registers and no shared-memory traffic for the
// matrix, and an FMA with a UR operand issues at full rate where the 3-vector-register form does
// not (tools/micro/gate_reg_rate.cu: 64 vs 50 TFLOP/s).
#include "ua_tile.cuh"
#include "ua_cluster.cuh"

namespace ua {

#ifndef UA_CL_TEAMS
#define UA_CL_TEAMS 3
#endif
#ifndef UA_CL_TT
#define UA_CL_TT 256
#endif
constexpr int CL_TEAMS = UA_CL_TEAMS;   // teams per CTA (+ one producer warp)
constexpr int CL_TT = UA_CL_TT;         // threads per team
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
    unsigned short moff;         // offset of the matrix in ClusterArgs::mats in 16-byte units (target-bit order, adjoint applied)
    unsigned short type;         // 0..5: 2-qubit gate on cluster bits (0,1) (0,2) (0,3) (1,2) (1,3) (2,3); 6..9: 1-qubit on bit type-6
};
struct ClusterArgs {
    int ncl;
    ClusterDesc cl[CL_MAX_GATES];
    ClusterGate g[CL_MAX_GATES];
    alignas(16) float mats[3 * CL_MAX_MAT_ELEMS];   // per gate, row by row: D x gr, D x (-gi, gi)  (ua_cluster.cuh)
};


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
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(CL_TT) : "memory");
}

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
template <bool SCATTER>
__global__ void __launch_bounds__(CL_TEAMS * CL_TT + 32, 1) cluster_ring_kernel(const __grid_constant__ ClusterGeom a,
                                                                          const __grid_constant__ ClusterArgs ca) {
    constexpr int TEAMS = CL_TEAMS, TT = CL_TT, MAXB = 8;
    constexpr int ARITH = 1;         // packed FFMA2 (the scalar FFMA form measured 13-35 % slower, profiles/)
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
    const bool full_tile = (groups % TT) == 0;      // no idle threads in any sweep
    auto group_base8 = [&](int c, unsigned g) -> unsigned {
        unsigned base = 0;
        for (int k = 0; k < gbits; ++k) base |= ((g >> k) & 1u) << ca.cl[c].fb[k];
        base ^= ((base >> 4) & 7u) << 1;          // 128-byte TMA swizzle
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
                mbar_wait(smem_u32(&bar_done[b]), (unsigned)((j / NB) & 1));
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
        mbar_wait(smem_u32(&bar_full[b]), (unsigned)((j / NB) & 1));
        const unsigned tile_sb = tile_s + (unsigned)b * buf_bytes;
        for (int c = 0; c < ncl; ++c) {
            const ClusterDesc &cd = ca.cl[c];
            const int gbeg = cd.gbeg, gend = cd.gend;
            const bool vec16 = cd.vec16 != 0;
            // warp-uniform trip count: the gate loop must stay convergent or ptxas moves the
            // matrices from uniform to vector registers; tiles with fewer groups than threads
            // predicate the loads and stores instead
            // member addresses: base ^ (XOR of the byte offsets of the member's set cluster bits)
            const unsigned p0 = cd.po8[1], p1 = cd.po8[2], p2 = cd.po8[4], p3 = cd.po8[8];
            auto member_addr = [&](unsigned base8, int m) -> unsigned {
                unsigned x = base8;
                if (m & 1) x ^= p0;
                if (m & 2) x ^= p1;
                if (m & 4) x ^= p2;
                if (m & 8) x ^= p3;
                return x;
            };
            for (int w = 0; w < sweeps; ++w) {
                const unsigned g = (unsigned)w * TT + ttid;
                const bool act = full_tile || g < groups;
                const unsigned base8 = tile_sb + (c < ntab ? (unsigned)tab[(c * sweeps + w) * TT + ttid]
                                                           : (act ? group_base8(c, g) : 0u));
                float2 v[16];
                // four straight-line variants of the load phase (16- / 8-byte members, full tile or
                // predicated tail); the branch is warp-uniform
                if (vec16) {
                    if (full_tile) {
#pragma unroll
                        for (int m = 0; m < 16; m += 2) {
                            const float4 t = lds128(member_addr(base8, m));
                            v[m] = make_float2(t.x, t.y);
                            v[m + 1] = make_float2(t.z, t.w);
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < 16; m += 2) {
                            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (act) t = lds128(member_addr(base8, m));
                            v[m] = make_float2(t.x, t.y);
                            v[m + 1] = make_float2(t.z, t.w);
                        }
                    }
                } else {
                    if (full_tile) {
#pragma unroll
                        for (int m = 0; m < 16; ++m) v[m] = lds64(member_addr(base8, m));
                    } else {
#pragma unroll
                        for (int m = 0; m < 16; ++m) {
                            v[m] = make_float2(0.f, 0.f);
                            if (act) v[m] = lds64(member_addr(base8, m));
                        }
                    }
                }
                for (int q = gbeg; q < gend; ++q) {
                    const ClusterGate cg = ca.g[q];
                    // moff counts 16-byte units: the compiler can prove the alignment of the row loads
                    reg_gate_dispatch<ARITH>(v, cg.type, reinterpret_cast<const float *>(
                        reinterpret_cast<const float4 *>(ca.mats) + cg.moff));
                }
                if (vec16) {
                    if (full_tile) {
#pragma unroll
                        for (int m = 0; m < 16; m += 2)
                            sts128(member_addr(base8, m), make_float4(v[m].x, v[m].y, v[m + 1].x, v[m + 1].y));
                    } else {
#pragma unroll
                        for (int m = 0; m < 16; m += 2)
                            if (act) sts128(member_addr(base8, m), make_float4(v[m].x, v[m].y, v[m + 1].x, v[m + 1].y));
                    }
                } else {
                    if (full_tile) {
#pragma unroll
                        for (int m = 0; m < 16; ++m) sts64(member_addr(base8, m), v[m]);
                    } else {
#pragma unroll
                        for (int m = 0; m < 16; ++m)
                            if (act) sts64(member_addr(base8, m), v[m]);
                    }
                }
            }
            if (c + 1 < ncl) team_barrier(team);
        }
        fence_proxy_async();        // make the generic-proxy writes visible to the copy engine
        team_barrier(team);
        if (ttid == 0) mbar_arrive(smem_u32(&bar_done[b]));
    }
}


// ------------------------------------------------------------------ cluster path (host side)
// Where a thread's group lives and how its 16 members are addressed.  Tile element i sits at
// shared-memory element phys(i) = i ^ (((i >> 4) & 7) << 1): the tile is written with the
// 128-byte TMA swizzle; phys is linear over GF(2), so the address of a
// member is phys(group base) ^ phys(member offset) = base' ^ po8[m] (in bytes).
// One shared-memory wavefront serves 128 bytes: 16 lanes of an 8-byte access (8 lanes of a 16-byte
// one), and is conflict-free when those lanes cover all 16 (8) slots of a 128-byte row, i.e. when
// the low group-number bits reach every slot bit.  Slot bit 0 is element bit 0; slot bit s = 1..3
// is element bit s XOR element bit s + 3 under the swizzle, so it can be driven from whichever of
// the two is not a cluster bit.
static void fill_cluster_layout(ClusterDesc &cd, unsigned mask, int T) {
    auto phys = [&](unsigned i) -> unsigned { return i ^ (((i >> 4) & 7u) << 1); };
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
        else if (is_free(sbit + 3)) take(sbit + 3);
    }
    for (int b = 0; b < T; ++b)
        if (is_free(b)) take(b);
}

// Cut the pass's gates (a.gates[], tile-local target bits ascending in sb[]) into clusters of at
// most CL_BITS tile bits and copy the HOST matrices into the kernel parameters in target-bit
// order (adjoint applied).  A gate joins an earlier cluster only across clusters it shares no
// bit with (gates on disjoint bits commute), so the product is unchanged.  Returns false when
// the pass cannot use the cluster path (a gate with k > 2, tile smaller than a cluster).
static bool build_clusters(const FusedArgs &a, const float2 *host_mats, ClusterArgs &ca) {
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
        fill_cluster_layout(ca.cl[c], mask, a.T);
        ca.cl[c].gbeg = (unsigned char)q;
        for (int i = 0; i < cls[c].ng; ++i) {
            const FusedGate &gd = a.gates[cls[c].gates[i]];
            const int K = gd.k, D = 1 << K;
            ClusterGate &cg = ca.g[q++];
            cg.moff = (unsigned short)(moff / 4);          // moff counts floats; 3 D^2 per gate = 48 or 12
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
                float *row = ca.mats + moff + 3 * D * sr;
                row[t] = val.x;
                row[D + 2 * t] = -val.y;
                row[D + 2 * t + 1] = val.y;
            }
            moff += 3 * D * D;
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

template <bool SCATTER>
static int launch_cluster(const FusedArgs &a, const ClusterArgs &ca, cudaStream_t st) {
    static thread_local ClusterGeom g;
    fill_cluster_geom(g, a);
    auto kern = cluster_ring_kernel<SCATTER>;
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
        if (e != cudaSuccess) { set_error("register-blocked pass: cannot raise shared memory limit: %s", cudaGetErrorString(e)); cudaGetLastError(); return UA_ERR_CUDA; }
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
    // front: the ring occupies window [buf_bytes, (nbuf + 1) * buf_bytes); the dynamic region starts
    // at an unknown offset >= 1 KiB, so (nbuf + 1) * buf_bytes - 1024 bytes always cover it
    int nbuf = front ? (int)((limit + 1024) / buf_bytes) - 1 : (int)((limit - tab_bytes) / buf_bytes) - 1;
    if (nbuf > 8) nbuf = 8;
    // The ring length must be a multiple of the team count: a buffer then belongs to ONE team for
    // the whole pass.  The mbarrier waits are parity-based; if a buffer were handed from team to
    // team, a fast team could ask for phase k+1 of a barrier whose phase k load has not landed yet
    // (TMA loads complete out of order) and the parity test would pass one phase early (seen as a
    // hang with 5 buffers, profiles/r02_ring_hang_nbuf5.txt).
    nbuf -= nbuf % CL_TEAMS;
    if (nbuf < (CL_TEAMS > 3 ? CL_TEAMS : 2 * CL_TEAMS)) { set_error("register-blocked pass: %d tile buffers of %zu bytes do not fit", 2 * CL_TEAMS, buf_bytes); return UA_ERR_UNSUPPORTED; }
    g.nbuf = nbuf;
    g.tab_front = front ? 1 : 0;
    g.tab_bytes = (int)tab_bytes;
    const size_t smem = front ? (size_t)(nbuf + 1) * buf_bytes - 1024 : (size_t)(nbuf + 1) * buf_bytes + tab_bytes;
    long long grid = sms;
    if (grid > g.num_tiles) grid = g.num_tiles;
    kern<<<(unsigned)grid, CL_TEAMS * CL_TT + 32, smem, st>>>(g, ca);
    return check_launch("cluster_ring_kernel");
}

}  // namespace ua

using namespace ua;

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
    if (!build_clusters(a, reinterpret_cast<const float2 *>(host_gate_mats), ca)) {
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
    if (!setup_tensor_maps(a, 0, total_amps, true)) { set_error("%s: the tile needs more than 5 TMA dimensions", who); return UA_ERR_UNSUPPORTED; }
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
    if (!build_clusters(a, reinterpret_cast<const float2 *>(host_gate_mats), ca)) {
        set_error("%s: the pass does not fit the register-blocked path (tile of %d bits)", who, a.T);
        return UA_ERR_UNSUPPORTED;
    }
    a.mats = nullptr;
    a.trank = 0;
    if (!setup_tensor_maps(a, 0, total_amps, true)) { set_error("%s: the tile needs more than 5 TMA dimensions", who); return UA_ERR_UNSUPPORTED; }
    return launch_cluster<true>(a, ca, st);
}
