// Dense 5-qubit complex64 gate on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// fp32 accumulator in tensor memory), with the 3xTF32 split that keeps the 1e-5 tolerance.
//
// The reference contracts a 32 x 32 complex matrix with the state through einsum -> cuBLAS
// Cgemm after two permute copies (src/unitair/simulation/operations.py:165-186, :322).  Here
// the block is a real GEMM per tile and runs in ONE pass over the state:
//
//   tile   = {index bits 0..3} U {the 5 target bits} U {2 passenger bits}  = 2^11 amplitudes,
//            brought to shared memory by TMA with the 128-byte / 32-byte-atom swizzle in the order
//            [passenger p (2 bits)] [target index g (5 bits)] [low (4 bits)] [re/im]
//            (passengers = the lowest free bits: the tile is made of 512-byte runs of the state)
//   A      = the tile itself, read MN-major: row m' = (p, low, re/im) (128 rows), column k = g (32)
//            -- exactly the canonical SWIZZLE_128B_BASE32B MN-major operand layout (the only one
//            the tensor core takes for transposed 32-bit operands), no repacking
//   B      = [Re U^T | Im U^T] (32 x 64), K-major, built once per CTA from the gate in global memory
//   D      = A . B  (128 x 64 fp32 in TMEM):  D[(p,low,re), n] = sum_g x_re(g) Ur(n,g), ...
//   y_re(n) = D[(.,re), n] - D[(.,im), 32+n],  y_im(n) = D[(.,re), 32+n] + D[(.,im), n]
//            (epilogue: tcgen05.ld, one shuffle with the neighbouring lane, store to the tile)
//
// Precision: a TF32 operand keeps 10 mantissa bits.  Both operands are split x = hi + lo with
// hi = x truncated to TF32 (exactly representable, so the tensor core's own conversion cannot
// change it) and D = A_hi B_hi + A_hi B_lo + A_lo B_hi: 12 MMAs of 128 x 64 x 8 per tile, error
// ~2^-21 per product, like fp32.  tools/micro/tcgen05_tf32_rate.cu: 479 TFLOP/s dense at this
// shape = 160 TFLOP/s after the split, against 64 TFLOP/s for FFMA2 -- the pass becomes HBM-bound.
#include "ua_tile.cuh"

namespace ua {

// Passenger bits: with 2 the tile is 2^11 amplitudes = ONE M = 128 block (16 KiB, 3 CTAs per SM);
// with 3 it is two M = 128 halves (32 KiB, 1 KiB runs, 2 CTAs per SM) -- measured slower (5.6-7.8 ms
// against 3.7-4.4 ms per 30-qubit pass): the pass is bound by the serial chain of a CTA, not by
// the length of the DRAM runs, so more resident CTAs win.
// L2 prefetch of tiles further ahead (cp.async.bulk.prefetch.tensor) was measured and lost:
// 4.7-5.6 ms per 30-qubit pass against 3.6-4.2 ms without.  So did a ring version (one CTA per SM,
// producer warp + 3-4 teams of 128 threads on a ring of 8-9 tile buffers, the structure of
// cluster_ring_kernel): 5.3-8.2 ms, 4.8-7.5 ms with the tensor-core work switched off -- with
// 16 KiB tiles made of 512-byte runs one producer warp per SM cannot issue the copies fast enough,
// three CTAs with their own issuing warps can (profiles/r02_tc5_variants.txt).
constexpr int TC_AHEAD = 0;
constexpr int TC_PASS = 2;
constexpr int TC_HALVES = 1 << (TC_PASS - 2);
constexpr int TC_TILE_BITS = 4 + 5 + TC_PASS;
constexpr int TC_THREADS = 128 * TC_HALVES;     // warps 4h..4h+3 read the accumulator of half h

struct Tc5Args {
    const float2 *gate;          // 32 x 32 complex, gate order (row-major)
    int gbit[5];                 // gate-index bit of the i-th lowest target bit
    int adjoint;
    long long num_tiles;
    int total_bits;              // index bits of one state (n); rows of a batch follow each other
    int nfree;                   // non-tile bits
    int free_pos[48];            // their positions, ascending
    int nd;                      // TMA dimensions (<= 5)
    int dstart[5], dlen[5];      // coordinate d = (index >> dstart[d]) & ((1 << dlen[d]) - 1)
    int nloop;                   // tile bits iterated by separate TMA copies (highest shared-memory order)
    int loop_pos[8];             // their index-bit positions
    int box_bytes;               // bytes per TMA copy = 16 KiB >> nloop
    int debug;                   // UA_TC5_DEBUG: 1 = copy the (truncated) tile through, no MMA
    alignas(64) CUtensorMap tmap_in;
    alignas(64) CUtensorMap tmap_out;
};

// layout: 2 = SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks: the only
// shared-memory layout the tensor core accepts for MN-major 32-bit operands)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address
    d |= (uint64_t)(lbo_bytes >> 4) << 16;              // leading-dimension byte offset
    d |= (uint64_t)(sbo_bytes >> 4) << 32;              // stride-dimension byte offset
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_prefetch_l2(int rank, const CUtensorMap *tm, const int *c) {
    const unsigned long long t = reinterpret_cast<unsigned long long>(tm);
    switch (rank) {
        case 1: asm volatile("cp.async.bulk.prefetch.tensor.1d.L2.global.tile [%0, {%1}];" ::"l"(t), "r"(c[0]) : "memory"); break;
        case 2: asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(t), "r"(c[0]), "r"(c[1]) : "memory"); break;
        case 3: asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory"); break;
        case 4: asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory"); break;
        default: asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory"); break;
    }
}
// byte offset inside a 128B-swizzled buffer (1 KiB aligned): 16-byte chunk index ^= row index mod 8
__device__ __forceinline__ uint32_t swz128(uint32_t b) { return b ^ (((b >> 7) & 7u) << 4); }
// ... and inside a 128B / 32-byte-atom swizzled buffer: 32-byte chunk index ^= row index mod 4
__device__ __forceinline__ uint32_t swz128_32(uint32_t b) { return b ^ (((b >> 7) & 3u) << 5); }

__global__ void __launch_bounds__(TC_THREADS, TC_PASS == 2 ? 3 : 2) gate_tc5_kernel(const __grid_constant__ Tc5Args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_load[2], bar_mma;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base_s = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *base_p = smem_raw + (base_s - smem_u32(smem_raw));
    constexpr uint32_t TILE = 8u << TC_TILE_BITS, BT = 8192;
    // [tile buffer 0][tile buffer 1][lo part of the current tile][B hi][B lo]
    const uint32_t sAlo = base_s + 2 * TILE, sBhi = base_s + 3 * TILE, sBlo = base_s + 3 * TILE + BT;
    float *pBhi = reinterpret_cast<float *>(base_p + 3 * TILE), *pBlo = reinterpret_cast<float *>(base_p + 3 * TILE + BT);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar_load[0]), 1);
        mbar_init(smem_u32(&bar_load[1]), 1);
        mbar_init(smem_u32(&bar_mma), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(64 * TC_HALVES));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // ---- B = [Re U^T | Im U^T], hi and lo parts, K-major rows of 128 bytes with the 128B swizzle:
    //      row nn (0..63) holds k = 0..31; (nn < 32: Re U[nn][k], else Im U[nn-32][k]), U in
    //      target-bit order with the adjoint applied
    for (int e = threadIdx.x; e < 64 * 32; e += TC_THREADS) {
        const int nn = e >> 5, k = e & 31;
        const int s = nn & 31;
        int gi = 0, gj = 0;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            gi |= ((s >> b) & 1) << a.gbit[b];
            gj |= ((k >> b) & 1) << a.gbit[b];
        }
        float2 u;
        if (a.adjoint) { u = a.gate[gj * 32 + gi]; u.y = -u.y; }
        else u = a.gate[gi * 32 + gj];
        const float x = nn < 32 ? u.x : u.y;
        const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const uint32_t off = swz128((uint32_t)(nn * 128 + k * 4));
        pBhi[off >> 2] = hi;
        pBlo[off >> 2] = x - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_d = tmem_base_s;
    // D = f32, A = B = tf32, A MN-major (bit 15), B K-major, N = 64, M = 128
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    if (a.debug & 2) idesc &= ~(1u << 15);           // debug: K-major A (wrong layout, tests the plumbing)

    // tile counter -> index of the tile's first amplitude: lane i deposits counter bit i (and
    // i + 32), one warp OR-reduction instead of a 20-40 step serial loop (warp-collective)
    auto tile_index = [&](long long tile_id) -> uint64_t {
        uint64_t part = 0;
        for (int i = lane; i < a.nfree; i += 32) part |= (uint64_t)((tile_id >> i) & 1) << a.free_pos[i];
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)part);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(part >> 32));
        return (((uint64_t)hi << 32) | lo) | ((uint64_t)(tile_id >> a.nfree) << a.total_bits);      // + batch row
    };
    // the TMA copies of a tile: copy j covers the values of the looped bits, lanes issue them in parallel
    auto issue_copies = [&](uint64_t base_idx, int mode, int buf) {      // mode 0: load, 1: store, 2: L2 prefetch
        const uint32_t sA = base_s + (uint32_t)buf * TILE;
        const int ncopies = 1 << a.nloop;
        for (int j = lane; j < ncopies; j += 32) {
            uint64_t idx = base_idx;
            for (int i = 0; i < a.nloop; ++i) idx |= (uint64_t)((j >> i) & 1) << a.loop_pos[i];
            int c[5];
            for (int d = 0; d < a.nd; ++d) c[d] = (int)((idx >> a.dstart[d]) & ((a.dlen[d] >= 63 ? 0ull : (1ull << a.dlen[d])) - 1ull));
            if (mode == 0) tma_load(a.nd, sA + (uint32_t)j * a.box_bytes, &a.tmap_in, c, smem_u32(&bar_load[buf]));
            else if (mode == 1) tma_store(a.nd, &a.tmap_out, c, sA + (uint32_t)j * a.box_bytes);
            else tc_prefetch_l2(a.nd, &a.tmap_in, c);
        }
    };

    // two tile buffers: the load of the next tile is issued while the tensor core works on this one
    if (warp == 0 && (long long)blockIdx.x < a.num_tiles) {
        if (lane == 0) mbar_arrive_expect_tx(smem_u32(&bar_load[0]), TILE);
        __syncwarp();
        issue_copies(tile_index(blockIdx.x), 0, 0);
        for (int ahead = 1; ahead <= TC_AHEAD; ++ahead)
            if (blockIdx.x + (long long)ahead * gridDim.x < a.num_tiles) issue_copies(tile_index(blockIdx.x + (long long)ahead * gridDim.x), 2, 0);
    }
    uint32_t it = 0;
    for (long long tile_id = blockIdx.x; tile_id < a.num_tiles; tile_id += gridDim.x, ++it) {
        const int buf = (int)(it & 1u);
        const uint32_t sA = base_s + (uint32_t)buf * TILE;
        unsigned char *pA = base_p + (size_t)buf * TILE;
        const uint64_t base_idx = tile_index(tile_id);
        mbar_wait(smem_u32(&bar_load[buf]), (it >> 1) & 1u);
        // ---- split the tile in place: A <- hi (TF32-exact), Alo <- x - hi ------------------------
        {
            float4 *A4 = reinterpret_cast<float4 *>(pA);
            float4 *L4 = reinterpret_cast<float4 *>(base_p + 2 * TILE);
#pragma unroll 4
            for (int v = threadIdx.x; v < (int)(TILE / 16); v += TC_THREADS) {
                const float4 x = A4[v];
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
                h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
                h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
                h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
                A4[v] = h;
                L4[v] = l;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (a.debug == 1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (warp == 0) { issue_copies(base_idx, 1, buf); bulk_commit(); }
            if (warp == 0 && tile_id + gridDim.x < a.num_tiles) {
                bulk_wait_read_all();
                if (lane == 0) mbar_arrive_expect_tx(smem_u32(&bar_load[buf ^ 1]), TILE);
                __syncwarp();
                issue_copies(tile_index(tile_id + gridDim.x), 0, buf ^ 1);
            }
            continue;
        }
        // ---- 12 MMAs: K = 32 in four steps of 8 (one 1 KiB swizzle atom of A, 32 bytes of a B row) --
        if (threadIdx.x == 0) {
#pragma unroll
            for (int h = 0; h < ((a.debug & 16) ? 0 : TC_HALVES); ++h) {      // M = 128 blocks (top passenger bit when there are two)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    // A (MN-major): 32 rows m' per 128-byte line, next 32 rows (next passenger value)
                    // 4 KiB further (LBO), four-column swizzle atoms of 512 bytes (SBO)
                    const uint64_t ahi = tc_desc(sA + h * (TILE / 2) + ks * 1024, 4096, 512, 1);
                    const uint64_t alo = tc_desc(sAlo + h * (TILE / 2) + ks * 1024, 4096, 512, 1);
                    // B (K-major): 8-row groups 1 KiB apart (SBO)
                    const uint64_t bhi = tc_desc(sBhi + ks * 32, 16, 1024), blo = tc_desc(sBlo + ks * 32, 16, 1024);
                    tc_mma(tmem_d + 64 * h, ahi, bhi, idesc, ks ? 1u : 0u);
                    tc_mma(tmem_d + 64 * h, ahi, blo, idesc, 1u);
                    tc_mma(tmem_d + 64 * h, alo, bhi, idesc, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
        }
        // while the MMAs run: the other buffer's store (previous tile) has long left shared memory,
        // refill it with the next tile
        if (warp == 0 && tile_id + gridDim.x < a.num_tiles) {
            bulk_wait_read_all();
            if (lane == 0) mbar_arrive_expect_tx(smem_u32(&bar_load[buf ^ 1]), TILE);
            __syncwarp();
            issue_copies(tile_index(tile_id + gridDim.x), 0, buf ^ 1);
            // ... and pull a tile further ahead into L2: the two 16 KiB buffers alone keep too few
            // bytes in flight per SM to cover the HBM latency
            if (TC_AHEAD > 0) {
                const long long tp = tile_id + (long long)(1 + TC_AHEAD) * gridDim.x;
                if (tp < a.num_tiles) issue_copies(tile_index(tp), 2, 0);
            }
        }
        mbar_wait(smem_u32(&bar_mma), it & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (!(a.debug & 8))
        // ---- epilogue: lane = row m' = (p = warp & 3, low = lane >> 1, re/im = lane & 1) of half warp >> 2 ----
        {
            uint32_t r[64];
            const int half = warp >> 2, pq = warp & 3;
            const uint32_t taddr = tmem_d + ((uint32_t)(pq * 32) << 16) + 64u * half;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[16 * q + 0]), "=r"(r[16 * q + 1]), "=r"(r[16 * q + 2]), "=r"(r[16 * q + 3]),
                      "=r"(r[16 * q + 4]), "=r"(r[16 * q + 5]), "=r"(r[16 * q + 6]), "=r"(r[16 * q + 7]),
                      "=r"(r[16 * q + 8]), "=r"(r[16 * q + 9]), "=r"(r[16 * q + 10]), "=r"(r[16 * q + 11]),
                      "=r"(r[16 * q + 12]), "=r"(r[16 * q + 13]), "=r"(r[16 * q + 14]), "=r"(r[16 * q + 15])
                    : "r"(taddr + 16 * q));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            if ((a.debug & 4) && blockIdx.x == 0 && tile_id == blockIdx.x && lane < 2)
                printf("warp %d lane %d idesc %08x tmem %08x D[0..3] %g %g %g %g D[32..33] %g %g\n", warp, lane, idesc, tmem_d,
                       __uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]),
                       __uint_as_float(r[32]), __uint_as_float(r[33]));
            const int c = lane & 1, low = lane >> 1;
            float *outp = reinterpret_cast<float *>(pA);
#pragma unroll
            for (int nn = 0; nn < 32; ++nn) {
                const float mine = __uint_as_float(r[nn]);
                const float other = __shfl_xor_sync(0xffffffffu, __uint_as_float(r[32 + nn]), 1);
                const float y = c ? mine + other : mine - other;
                // element (p = 4 half + pq, g = nn, low), component c
                const uint32_t b = (uint32_t)((((half * 4 + pq) * 512 + nn * 16 + low) * 2 + c) * 4);
                outp[swz128_32(b) >> 2] = y;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            issue_copies(base_idx, 1, buf);
            bulk_commit();
        }
    }
    if (warp == 0) bulk_wait_all();
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(64 * TC_HALVES));
}

typedef CUresult (*EncodeTiledFnTc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

// Dense complex64 5-qubit gate through the tensor cores.  spos: ascending target bit positions;
// gbit: gate-index bit of each.  Returns UA_ERR_UNSUPPORTED when the shape does not fit this
// path (a target among the 4 lowest bits, fewer than 11 index bits): the caller then uses the
// CUDA-core kernel.
int launch_gate_tc5(void *out, const void *in, const void *gate, int total_bits, long long batch, const int *spos,
                    const int *gbit, int adjoint, cudaStream_t st) {
    if (total_bits < TC_TILE_BITS || spos[0] < 4) return UA_ERR_UNSUPPORTED;
    static EncodeTiledFnTc enc = nullptr;
    if (!enc) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return UA_ERR_UNSUPPORTED; }
        enc = reinterpret_cast<EncodeTiledFnTc>(p);
    }
    static thread_local Tc5Args a;
    a.gate = reinterpret_cast<const float2 *>(gate);
    a.adjoint = adjoint ? 1 : 0;
    a.total_bits = total_bits;
    { const char *e = getenv("UA_TC5_DEBUG"); a.debug = e ? atoi(e) : 0; }
    unsigned long long tmask = 0xFull;
    for (int i = 0; i < 5; ++i) { a.gbit[i] = gbit[i]; tmask |= 1ull << spos[i]; }
    int pas[TC_PASS], np = 0;
    for (int b = 4; b < total_bits && np < TC_PASS; ++b)
        if (!((tmask >> b) & 1ull)) { pas[np++] = b; tmask |= 1ull << b; }
    if (np < TC_PASS) return UA_ERR_UNSUPPORTED;
    a.nfree = 0;
    for (int b = 0; b < total_bits; ++b)
        if (!((tmask >> b) & 1ull)) a.free_pos[a.nfree++] = b;
    a.num_tiles = batch << a.nfree;
    // tile bits above the low four, in shared-memory order: targets ascending, then passengers
    int order[5 + TC_PASS];
    for (int i = 0; i < 5; ++i) order[i] = spos[i];
    for (int i = 0; i < TC_PASS; ++i) order[5 + i] = pas[i];
    // windows: runs that are consecutive in this order AND in the index; at most 4 become TMA
    // dimensions (after the 128-byte row), the rest are iterated bit by bit
    int wstart[5 + TC_PASS], wlen[5 + TC_PASS], nw = 0;
    for (int i = 0; i < 5 + TC_PASS; ++i) {
        if (nw > 0 && order[i] == wstart[nw - 1] + wlen[nw - 1]) wlen[nw - 1]++;
        else { wstart[nw] = order[i]; wlen[nw] = 1; nw++; }
    }
    const int mapped = nw < 4 ? nw : 4;
    a.nloop = 0;
    for (int w = mapped; w < nw; ++w)
        for (int i = 0; i < wlen[w]; ++i) a.loop_pos[a.nloop++] = wstart[w] + i;
    if (a.nloop > 5) return UA_ERR_UNSUPPORTED;          // copies of at least one 1 KiB swizzle atom
    a.box_bytes = (8 << TC_TILE_BITS) >> a.nloop;
    // every index bit belongs to the field of the mapped dimension that starts at or below it
    int starts[5], nd = 1 + mapped;
    starts[0] = 0;
    for (int w = 0; w < mapped; ++w) starts[1 + w] = wstart[w];
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t box[5], estr[5];
    for (int d = 0; d < nd; ++d) {
        int end = total_bits;                      // field end: the next higher mapped start
        for (int e = 0; e < nd; ++e)
            if (starts[e] > starts[d] && starts[e] < end) end = starts[e];
        a.dstart[d] = starts[d];
        a.dlen[d] = end - starts[d];
        gdim[d] = 1ull << a.dlen[d];
        if (end == total_bits) {                   // the highest field also spans the batch rows
            a.dlen[d] = 63;
            gdim[d] = (cuuint64_t)batch << (total_bits - starts[d]);
        }
        box[d] = d == 0 ? 16u : (1u << wlen[d - 1]);
        estr[d] = 1;
        if (d > 0) gstride[d - 1] = 8ull << starts[d];
        if (gdim[d] > 0xffffffffull || gdim[d] < box[d]) return UA_ERR_UNSUPPORTED;
    }
    a.nd = nd;
    for (int which = 0; which < 2; ++which) {
        const CUresult r = enc(which ? &a.tmap_out : &a.tmap_in, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)nd,
                               which ? out : const_cast<void *>(in), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return UA_ERR_UNSUPPORTED;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = 3 * (8 << TC_TILE_BITS) + 2 * 8192 + 1024;
    static bool attr_set[64] = {};
    if (!attr_set[dev & 63]) {
        if (cudaFuncSetAttribute(gate_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return UA_ERR_UNSUPPORTED;
        }
        attr_set[dev & 63] = true;
    }
    long long grid = (long long)sm_count() * (TC_PASS == 2 ? 3 : 2);
    if (grid > a.num_tiles) grid = a.num_tiles;
    gate_tc5_kernel<<<(unsigned)grid, TC_THREADS, smem, st>>>(a);
    return check_launch("gate_tc5_kernel");
}

}  // namespace ua
