// Index bit-permutation (qubit swap / roll / permute) in one pass.
//
// Replaces torch.permute(...).contiguous() on the tensor layout
// (src/unitair/simulation/operations.py:506-654): out[b, i] = in[b, j] where bit
// src_bit[p] of j equals bit p of i.  Pure data movement (bit-exact).  Every thread writes
// one 16-byte vector when bit 0 (complex64) stays in place, else one amplitude; writes are
// coalesced, reads gather through byte-wise lookup tables built once per block.
#include "ua_common.cuh"

namespace ua {

struct PermArgs {
    const void *in;
    void *out;
    long long total;       // batch * 2^n
    int n;
    int src[48];           // bit p of the OUTPUT index comes from bit src[p] of the INPUT index
};

template <typename T>   // T = element moved per thread (float2, double2/float4)
__global__ void __launch_bounds__(256) permute_bits_kernel(const PermArgs a, int elem_shift) {
    // lut[c][v]: contribution of output-index byte c (value v) to the input index
    __shared__ unsigned long long lut[6][256];
    const int nb = a.n - elem_shift;          // permuted bits above the in-element bits
    const int nbytes = (nb + 7) / 8;
    for (int e = threadIdx.x; e < nbytes * 256; e += 256) {
        const int c = e >> 8, v = e & 255;
        unsigned long long j = 0;
        for (int bit = 0; bit < 8; ++bit) {
            const int p = c * 8 + bit;      // output bit (in element units)
            if (p < nb && ((v >> bit) & 1)) j |= 1ull << (a.src[p + elem_shift] - elem_shift);
        }
        lut[c][v] = j;
    }
    __syncthreads();
    const long long count = a.total >> elem_shift;
    const unsigned long long inner = (1ull << nb) - 1ull;
    const long long stride = (long long)gridDim.x * 256;
    const T *__restrict__ in = reinterpret_cast<const T *>(a.in);
    T *out = reinterpret_cast<T *>(a.out);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < count; i += stride) {
        const unsigned long long e = (unsigned long long)i & inner;
        unsigned long long j = 0;
        for (int c = 0; c < nbytes; ++c) j |= lut[c][(e >> (8 * c)) & 255];
        out[i] = in[((unsigned long long)i & ~inner) | j];
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
    PermArgs a{};
    unsigned long long seen = 0;
    for (int p = 0; p < num_bits; ++p) {
        const int s = host_src_bit[p];
        if (s < 0 || s >= num_bits || (seen & (1ull << s))) { set_error("ua_permute_bits: not a permutation"); return UA_ERR_INVALID; }
        seen |= 1ull << s;
        a.src[p] = s;
    }
    a.in = in; a.out = out; a.n = num_bits; a.total = batch << num_bits;
    const bool vec = dtype == UA_C64 && a.src[0] == 0 && !(((uintptr_t)out | (uintptr_t)in) & 15);
    const int shift = vec ? 1 : 0;
    const long long count = a.total >> shift;
    long long blocks = (count + 255) / 256;
    { const long long cap = (long long)sm_count() * 16; if (blocks > cap) blocks = cap; }
    if (dtype == UA_C64) {
        if (vec) permute_bits_kernel<float4><<<(unsigned)blocks, 256, 0, st>>>(a, 1);
        else permute_bits_kernel<float2><<<(unsigned)blocks, 256, 0, st>>>(a, 0);
    } else {
        permute_bits_kernel<double2><<<(unsigned)blocks, 256, 0, st>>>(a, 0);
    }
    return check_launch("permute_bits_kernel");
}
