// Index bit-permutation (qubit swap / roll / permute) in one pass.
//
// Replaces torch.permute(...).contiguous() on the tensor layout
// (src/unitair/simulation/operations.py:506-654): out[b, i] = in[b, j] where bit
// src_bit[p] of j equals bit p of i.  Pure data movement (bit-exact).
#include "ua_common.cuh"

namespace ua {

// Both sides of the copy see long contiguous runs.  A tile is the set of index
// bits {low A input bits} U {sources of the low B output bits} (at most 12 bits = 2^12
// amplitudes in shared memory).  The block reads it with runs of 2^A contiguous amplitudes,
// and writes it with runs of 2^B contiguous amplitudes of the output; the permutation happens
// in shared memory.  (Round 1's one-amplitude-per-thread kernel gathered 8/16-byte pieces at
// random addresses -- every read touched a whole 32-byte sector; numbers in profiles/.)
struct PermTileArgs {
    const void *in;
    void *out;
    long long num_tiles;         // batch * 2^(n - t)
    int n, t;
    int tin[12];                 // ascending INPUT positions of the tile bits
    int tout_sorted[12];         // ascending OUTPUT positions of the tile bits
    int lsrc[12];                // tile-local: output-order bit j comes from input-order bit lsrc[j]
    int outer_in[48];            // ascending input positions of the non-tile bits
    int outer_out[48];           // output position of outer bit i (dst of outer_in[i])
};

template <typename T>
__global__ void __launch_bounds__(256) permute_tile_kernel(const PermTileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *tile = reinterpret_cast<T *>(smem_raw);
    // byte-wise tables of the three bit deposits (t <= 12: two tables of 64 entries each)
    __shared__ unsigned long long in_lo[64], in_hi[64], out_lo[64], out_hi[64];
    __shared__ unsigned short perm_lo[64], perm_hi[64];
    const int t = a.t;
    if (threadIdx.x < 128) {
        const int half = threadIdx.x >> 6, v = threadIdx.x & 63;
        unsigned long long di = 0, dq = 0;
        unsigned pl = 0;
        for (int bit = 0; bit < 6; ++bit) {
            const int j = half * 6 + bit;
            if (j < t && ((v >> bit) & 1)) {
                di |= 1ull << a.tin[j];              // input-order local bit j -> input position
                dq |= 1ull << a.tout_sorted[j];      // output-order local bit j -> output position
                pl |= 1u << a.lsrc[j];               // output-order local bit j -> input-order local bit
            }
        }
        if (half == 0) { in_lo[v] = di; out_lo[v] = dq; perm_lo[v] = (unsigned short)pl; }
        else { in_hi[v] = di; out_hi[v] = dq; perm_hi[v] = (unsigned short)pl; }
    }
    __syncthreads();
    const T *__restrict__ in = reinterpret_cast<const T *>(a.in);
    T *out = reinterpret_cast<T *>(a.out);
    const int nouter = a.n - t;
    const unsigned tile_elems = 1u << t;
    const unsigned lo = threadIdx.x & 63u, hi0 = threadIdx.x >> 6;
    const unsigned long long ilo = in_lo[lo], olo = out_lo[lo];
    const unsigned plo = perm_lo[lo];
    for (long long tile_id = blockIdx.x; tile_id < a.num_tiles; tile_id += gridDim.x) {
        const unsigned long long row = (unsigned long long)tile_id >> nouter;
        const unsigned long long o = (unsigned long long)tile_id - (row << nouter);
        unsigned long long in_base = row << a.n, out_base = row << a.n;
        for (int i = 0; i < nouter; ++i) {
            const unsigned long long bit = (o >> i) & 1ull;
            in_base |= bit << a.outer_in[i];
            out_base |= bit << a.outer_out[i];
        }
        // l = tid + 256*it: the low-table part of every address is loop-invariant; four
        // independent loads are in flight per thread
        if (tile_elems >= 1024) {
            const unsigned iters = tile_elems >> 8;
            for (unsigned it = 0; it < iters; it += 4) {
                T v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldcs(in + (in_base | ilo | in_hi[hi0 + 4 * (it + u)]));
#pragma unroll
                for (int u = 0; u < 4; ++u) tile[threadIdx.x + 256 * (it + u)] = v[u];
            }
            __syncthreads();
            for (unsigned it = 0; it < iters; it += 4) {
                T v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = tile[plo | perm_hi[hi0 + 4 * (it + u)]];
#pragma unroll
                for (int u = 0; u < 4; ++u) __stcs(out + (out_base | olo | out_hi[hi0 + 4 * (it + u)]), v[u]);
            }
        } else {
            for (unsigned l = threadIdx.x; l < tile_elems; l += 256)
                tile[l] = __ldcs(in + (in_base | in_lo[l & 63] | in_hi[l >> 6]));
            __syncthreads();
            for (unsigned m = threadIdx.x; m < tile_elems; m += 256)
                __stcs(out + (out_base | out_lo[m & 63] | out_hi[m >> 6]), tile[perm_lo[m & 63] | perm_hi[m >> 6]]);
        }
        __syncthreads();
    }
}

}  // namespace ua

using namespace ua;

extern "C" int ua_permute_bits(int dtype, void *out, const void *in, int num_bits, long long batch,
                               const int *host_src_bit, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("ua_permute_bits: bad dtype"); return UA_ERR_INVALID; }
    if (!out || !in || !host_src_bit || out == in) { set_error("ua_permute_bits: bad pointers (must be out-of-place)"); return UA_ERR_INVALID; }
    if (num_bits < 1 || num_bits > 48 || batch < 1) { set_error("ua_permute_bits: bad sizes"); return UA_ERR_INVALID; }
    int src[48];
    unsigned long long seen = 0;
    for (int p = 0; p < num_bits; ++p) {
        const int sb = host_src_bit[p];
        if (sb < 0 || sb >= num_bits || (seen & (1ull << sb))) { set_error("ua_permute_bits: not a permutation"); return UA_ERR_INVALID; }
        seen |= 1ull << sb;
        src[p] = sb;
    }
    {
        // tiled path: low 6 input bits + sources of the low 6 output bits
        const int n = num_bits;
        const int A = n < 6 ? n : 6, B = n < 6 ? n : 6;
        unsigned long long tmask = (1ull << A) - 1ull;
        for (int p = 0; p < B; ++p) tmask |= 1ull << src[p];
        // permutations that leave the low bits alone give a small tile: grow it with the next
        // input bits (longer runs on both sides) up to 2^12 amplitudes
        for (int q = A; q < n && __builtin_popcountll(tmask) < 12; ++q) tmask |= 1ull << q;
        PermTileArgs ta{};
        int dst[48];
        for (int p = 0; p < n; ++p) dst[src[p]] = p;
        int t = 0, no = 0;
        unsigned long long omask = 0;
        for (int q = 0; q < n; ++q) {
            if ((tmask >> q) & 1ull) { ta.tin[t++] = q; omask |= 1ull << dst[q]; }
            else { ta.outer_in[no] = q; ta.outer_out[no] = dst[q]; ++no; }
        }
        int j = 0;
        for (int p = 0; p < n; ++p)
            if ((omask >> p) & 1ull) {
                ta.tout_sorted[j] = p;
                int li = 0;
                while (ta.tin[li] != src[p]) ++li;
                ta.lsrc[j] = li;
                ++j;
            }
        ta.in = in; ta.out = out; ta.n = n; ta.t = t;
        ta.num_tiles = batch << (n - t);
        const size_t esize = dtype == UA_C64 ? 8 : 16;
        const size_t smem = esize << t;
        long long blocks = ta.num_tiles;
        { const long long cap = (long long)sm_count() * 8; if (blocks > cap) blocks = cap; }
        if (dtype == UA_C64) {
            static bool set64[64] = {};
            int dev = 0; cudaGetDevice(&dev);
            if (smem > 48 * 1024 && !set64[dev & 63]) { cudaFuncSetAttribute(permute_tile_kernel<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); set64[dev & 63] = true; }
            permute_tile_kernel<float2><<<(unsigned)blocks, 256, smem, st>>>(ta);
        } else {
            static bool set128[64] = {};
            int dev = 0; cudaGetDevice(&dev);
            if (smem > 48 * 1024 && !set128[dev & 63]) { cudaFuncSetAttribute(permute_tile_kernel<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); set128[dev & 63] = true; }
            permute_tile_kernel<double2><<<(unsigned)blocks, 256, smem, st>>>(ta);
        }
        return check_launch("permute_tile_kernel");
    }
}
