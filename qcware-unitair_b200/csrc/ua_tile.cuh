// Shared declarations of the fused tile passes (ua_tile.cu: shared-memory-matrix gate phase and
// the adjoint backward pass; ua_cluster.cu: register-blocked gate phase): kernel argument
// structs, the PTX wrappers of the bulk-copy engine (TMA) and the host helpers that validate a
// pass description and encode its tensor maps.
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "ua_common.cuh"

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

__device__ __forceinline__ unsigned insert_zero32(unsigned x, int p) {
    const unsigned lo = x & ((1u << p) - 1u);
    return ((x >> p) << (p + 1)) | lo;
}


// ------------------------------------------------------------------ host helpers (ua_tile.cu)
// Validate a pass description and fill the kernel arguments shared by all tile passes
// (geometry, gate descriptors with tile-local target bits).  max_k: largest gate the caller's
// kernel handles; a negative max_k also allows a pass without gates.
int fill_fused_args(FusedArgs &a, const char *who, int dtype, void *out, const void *in,
                    long long total_amps, int total_bits, int tile_low_bits, int num_high,
                    const int *host_high_pos, int num_gates, const int *host_gate_k,
                    const int *host_gate_bits, const long long *host_gate_offset,
                    const void *gate_mats, long long gate_row_stride, int adjoint, int max_k,
                    int *mat_elems_out);
// fill the scatter fields of `a` (after fill_fused_args)
int fill_scatter_args(FusedArgs &a, const char *who, int total_bits, int num_scatter_bits,
                      const int *host_scatter_pos, void *const *host_dst_ptrs, int visit_xor);
// Describe the tile set {0..L-1} U high[] (amplitude bits) as TMA boxes; false when more than 5
// dimensions would be needed or the encoder is unavailable.  swizzle128: see ua_cluster.cu.
bool setup_tensor_maps(FusedArgs &a, int ebits, long long total_amps, bool swizzle128 = false);

}  // namespace ua
