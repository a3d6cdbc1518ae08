// Fused shared-memory pass: stage a tile of 2^T amplitudes once, apply a whole list of
// dense 1-3 qubit gates to it in shared memory, write it back.  One HBM read + write
// (16 B / 32 B per amplitude) for the entire list instead of one per gate.
//
// A tile is the set of amplitudes that agree on every index bit outside the tile's T bit
// positions: the low L bits (contiguous in memory -> 2^L * 8/16 B coalesced runs) plus H
// chosen higher positions.  Gates whose target bits all lie in that set never need
// another pass.  This is the engine behind apply_all_qubits
// (src/unitair/simulation/operations.py:332-413, n strided einsum passes in the reference)
// and behind the circuit API (the reference's own fusion idea, apply_to_qubits,
// operations.py:416-503, generalised from same-qubit 2x2 products to whole gate lists).
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
    FusedGate gates[UA_MAX_FUSED_GATES];
};

constexpr int FUSED_MAX_MAT_ELEMS = 2048;   // complex elements of gate matrices per pass

template <typename R, int K>
__device__ __forceinline__ void apply_gate_smem(typename CplxOf<R>::type *tile,
                                                const typename CplxOf<R>::type *M,
                                                const FusedGate &gd, int T, int nthreads) {
    using C = typename CplxOf<R>::type;
    constexpr int D = 1 << K;
    unsigned off[K];
#pragma unroll
    for (int i = 0; i < K; ++i) off[i] = 1u << gd.sb[i];
    constexpr bool MREG = (K <= 2);
    C mr[MREG ? D * D : 1];
    if (MREG) {
#pragma unroll
        for (int e = 0; e < D * D; ++e) mr[e] = M[e];
    }
    const unsigned groups = 1u << (T - K);
    for (unsigned g = threadIdx.x; g < groups; g += nthreads) {
        unsigned b = g;
#pragma unroll
        for (int i = 0; i < K; ++i) b = (unsigned)insert_zero(b, gd.sb[i]);
        C x[D];
#pragma unroll
        for (int s = 0; s < D; ++s) {
            unsigned idx = b;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if ((s >> i) & 1) idx |= off[i];
            x[s] = tile[idx];
        }
#pragma unroll
        for (int s = 0; s < D; ++s) {
            C acc = mk(R(0), R(0));
#pragma unroll
            for (int t = 0; t < D; ++t) cfma(acc, MREG ? mr[s * D + t] : M[s * D + t], x[t]);
            unsigned idx = b;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if ((s >> i) & 1) idx |= off[i];
            tile[idx] = acc;
        }
    }
}

template <typename R, int THREADS>
__global__ void __launch_bounds__(THREADS) fused_pass_kernel(const __grid_constant__ FusedArgs a) {
    using C = typename CplxOf<R>::type;
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *tile = reinterpret_cast<C *>(smem_raw);
    C *sM = tile + (1u << a.T);
    const unsigned nvec = (1u << a.T) / APV;
    const unsigned lowmask = (1u << a.L) - 1u;

    bool mats_loaded = false;
    for (long long tile_id = blockIdx.x; tile_id < a.num_tiles; tile_id += gridDim.x) {
        const long long row = tile_id / a.tiles_per_row;
        const long long j = tile_id - row * a.tiles_per_row;
        uint64_t base = (uint64_t)j << a.L;
        for (int i = 0; i < a.H; ++i) base = insert_zero(base, a.high[i]);
        base += (uint64_t)row << a.total_bits;

        // ---- tile loads first (coalesced 2^L-amplitude runs) -------------------------
        const V *__restrict__ in = reinterpret_cast<const V *>(a.in);
        for (unsigned v = threadIdx.x; v < nvec; v += THREADS) {
            const unsigned la = v * APV;
            uint64_t idx = base + (la & lowmask);
            const unsigned hi = la >> a.L;
            for (int i = 0; i < a.H; ++i)
                if ((hi >> i) & 1) idx |= 1ull << a.high[i];
            reinterpret_cast<V *>(tile)[v] = __ldcs(in + idx / APV);
        }
        // ---- gate matrices -> shared memory, register order ---------------------------
        if (!mats_loaded || a.mats_row_stride != 0) {
            const C *__restrict__ mats = reinterpret_cast<const C *>(a.mats) + row * a.mats_row_stride;
            for (int g = 0; g < a.num_gates; ++g) {
                const FusedGate &gd = a.gates[g];
                const int K = gd.k, D = 1 << K;
                for (int e = threadIdx.x; e < D * D; e += THREADS) {
                    const int s = e >> K, t = e & (D - 1);
                    int gi = 0, gj = 0;
                    for (int i = 0; i < K; ++i) {
                        gi |= ((s >> i) & 1) << gd.gb[i];
                        gj |= ((t >> i) & 1) << gd.gb[i];
                    }
                    C val;
                    if (a.adjoint) val = cconj(mats[gd.goff + gj * D + gi]);
                    else val = mats[gd.goff + gi * D + gj];
                    sM[gd.smoff + e] = val;
                }
            }
            mats_loaded = true;
        }
        __syncthreads();

        // ---- all gates in shared memory ------------------------------------------------
        for (int g = 0; g < a.num_gates; ++g) {
            const FusedGate &gd = a.gates[g];
            const C *M = sM + gd.smoff;
            if (gd.k == 1) apply_gate_smem<R, 1>(tile, M, gd, a.T, THREADS);
            else if (gd.k == 2) apply_gate_smem<R, 2>(tile, M, gd, a.T, THREADS);
            else apply_gate_smem<R, 3>(tile, M, gd, a.T, THREADS);
            __syncthreads();
        }

        // ---- write back ------------------------------------------------------------------
        V *out = reinterpret_cast<V *>(a.out);
        for (unsigned v = threadIdx.x; v < nvec; v += THREADS) {
            const unsigned la = v * APV;
            uint64_t idx = base + (la & lowmask);
            const unsigned hi = la >> a.L;
            for (int i = 0; i < a.H; ++i)
                if ((hi >> i) & 1) idx |= 1ull << a.high[i];
            __stcs(out + idx / APV, reinterpret_cast<V *>(tile)[v]);
        }
        __syncthreads();   // tile buffer is reused by the next iteration
    }
}

static int max_tile_bits(int dtype) { return dtype == UA_C64 ? 14 : 13; }

template <typename R, int THREADS>
static int launch_fused(const FusedArgs &a, size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = fused_pass_kernel<R, THREADS>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("ua_apply_fused_pass: cannot raise shared memory limit: %s", cudaGetErrorString(e)); return UA_ERR_CUDA; }
        attr_set = true;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) { set_error("ua_apply_fused_pass: occupancy query failed (%s), smem=%zu", cudaGetErrorString(e), smem); cudaGetLastError(); return UA_ERR_CUDA; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (long long)per_sm * sms;
    if (grid > a.num_tiles) grid = a.num_tiles;
    kern<<<(unsigned)grid, THREADS, smem, st>>>(a);
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

extern "C" int ua_apply_fused_pass(int dtype, void *out, const void *in, long long total_amps,
                                   int total_bits, int tile_low_bits, int num_high,
                                   const int *host_high_pos, int num_gates, const int *host_gate_k,
                                   const int *host_gate_bits, const long long *host_gate_offset,
                                   const void *gate_mats, long long gate_row_stride, int adjoint,
                                   void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("ua_apply_fused_pass: bad dtype"); return UA_ERR_INVALID; }
    if (!out || !in || !gate_mats || !host_gate_k || !host_gate_bits || !host_gate_offset || (num_high > 0 && !host_high_pos)) {
        set_error("ua_apply_fused_pass: null pointer"); return UA_ERR_INVALID;
    }
    if (((uintptr_t)out | (uintptr_t)in) & 15) { set_error("ua_apply_fused_pass: state pointers must be 16-byte aligned"); return UA_ERR_INVALID; }
    const int L = tile_low_bits, H = num_high, T = L + H;
    const int min_low = (dtype == UA_C64) ? 1 : 0;
    if (total_bits < 1 || total_bits > 48 || L < min_low || H < 0 || T > total_bits || T > max_tile_bits(dtype)) {
        set_error("ua_apply_fused_pass: bad tile geometry total_bits=%d L=%d H=%d", total_bits, L, H); return UA_ERR_INVALID;
    }
    if (num_gates < 1 || num_gates > UA_MAX_FUSED_GATES) { set_error("ua_apply_fused_pass: num_gates=%d out of range", num_gates); return UA_ERR_INVALID; }
    const long long space = 1ll << total_bits;
    if (total_amps < space || total_amps % space != 0) { set_error("ua_apply_fused_pass: total_amps must be a multiple of 2^total_bits"); return UA_ERR_INVALID; }
    const long long rows = total_amps >> total_bits;

    FusedArgs a{};
    a.in = in; a.out = out; a.mats = gate_mats; a.mats_row_stride = gate_row_stride;
    a.total_bits = total_bits; a.T = T; a.L = L; a.H = H;
    a.num_gates = num_gates; a.adjoint = adjoint ? 1 : 0;
    int local_of[64];
    for (int p = 0; p < 64; ++p) local_of[p] = (p < L) ? p : -1;
    int prev = L - 1;
    for (int i = 0; i < H; ++i) {
        const int p = host_high_pos[i];
        if (p <= prev || p >= total_bits) { set_error("ua_apply_fused_pass: high positions must be ascending in [L, total_bits)"); return UA_ERR_INVALID; }
        a.high[i] = p; local_of[p] = L + i; prev = p;
    }
    a.tiles_per_row = 1ll << (total_bits - T);
    a.num_tiles = a.tiles_per_row * rows;

    int mat_elems = 0;
    for (int g = 0; g < num_gates; ++g) {
        const int k = host_gate_k[g];
        if (k < 1 || k > 3) { set_error("ua_apply_fused_pass: gate %d has k=%d (1..3 supported)", g, k); return UA_ERR_UNSUPPORTED; }
        FusedGate &gd = a.gates[g];
        gd.k = (unsigned char)k;
        gd.goff = host_gate_offset[g];
        gd.smoff = (unsigned short)mat_elems;
        mat_elems += 1 << (2 * k);
        int lb[3], order[3];
        for (int j = 0; j < k; ++j) {
            const int p = host_gate_bits[g * 3 + j];
            if (p < 0 || p >= total_bits || local_of[p] < 0) { set_error("ua_apply_fused_pass: gate %d bit %d is outside the tile", g, p); return UA_ERR_INVALID; }
            lb[j] = local_of[p]; order[j] = j;
            for (int jj = 0; jj < j; ++jj) if (lb[jj] == lb[j]) { set_error("ua_apply_fused_pass: gate %d repeats a bit", g); return UA_ERR_INVALID; }
        }
        for (int i = 1; i < k; ++i)
            for (int j = i; j > 0 && lb[order[j]] < lb[order[j - 1]]; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
        for (int i = 0; i < k; ++i) { gd.sb[i] = (unsigned char)lb[order[i]]; gd.gb[i] = (unsigned char)(k - 1 - order[i]); }
    }
    if (mat_elems > FUSED_MAX_MAT_ELEMS) { set_error("ua_apply_fused_pass: %d matrix elements exceed the %d limit", mat_elems, FUSED_MAX_MAT_ELEMS); return UA_ERR_INVALID; }

    const size_t csize = (dtype == UA_C64) ? 8 : 16;
    const size_t smem = ((size_t)(1u << T) + (size_t)mat_elems) * csize;
    if (dtype == UA_C64) {
        if (T <= 12) return launch_fused<float, 256>(a, smem, st);
        if (T == 13) return launch_fused<float, 512>(a, smem, st);
        return launch_fused<float, 1024>(a, smem, st);
    }
    if (T <= 11) return launch_fused<double, 256>(a, smem, st);
    if (T == 12) return launch_fused<double, 512>(a, smem, st);
    return launch_fused<double, 1024>(a, smem, st);
}
