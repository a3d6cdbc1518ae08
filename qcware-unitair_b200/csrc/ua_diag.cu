// Sign-mask diagonal gates (CZ, C...CZ and products of them) in one streaming pass:
//   out[b, i] = in[b, i] * prod_m (-1)^[ (i & mask_m) == mask_m ]
// Replaces multi_cz / multi_controlled_z (src/unitair/simulation/operations.py:657-783), which
// materialise a 2^n phase vector (arange + bitwise_and + prod, :733-748) or permute the whole
// state twice (:781-783).  The masks are tested in-kernel; traffic is the 16 B / 32 B per
// amplitude of one read + one write.
#include "ua_common.cuh"

namespace ua {

constexpr int MAX_MASKS = 64;
struct DiagArgs {
    const void *in;
    void *out;
    long long nvec;       // 16-byte vectors in total
    int n;                // qubits (masks apply to the low n bits of the flat index)
    int num_masks;
    unsigned long long masks[MAX_MASKS];
};

template <typename R>
__global__ void __launch_bounds__(256) diag_masks_kernel(const DiagArgs a) {
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    const V *__restrict__ in = reinterpret_cast<const V *>(a.in);
    V *out = reinterpret_cast<V *>(a.out);
    const unsigned long long inner = (1ull << a.n) - 1ull;
    const long long i0 = (long long)blockIdx.x * 1024 + threadIdx.x;
    V x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (i0 + u * 256 < a.nvec) x[u] = __ldcs(in + i0 + u * 256);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const long long v = i0 + u * 256;
        if (v >= a.nvec) continue;
        const unsigned long long e0 = ((unsigned long long)v * APV) & inner;
        bool neg0 = false, neg1 = false;
        for (int m = 0; m < a.num_masks; ++m) {
            const unsigned long long mk = a.masks[m];
            neg0 ^= ((e0 & mk) == mk);
            if (APV == 2) neg1 ^= (((e0 | 1ull) & mk) == mk);
        }
        V r = x[u];
        if constexpr (APV == 2) {
            if (neg0) { r.x = -r.x; r.y = -r.y; }
            if (neg1) { r.z = -r.z; r.w = -r.w; }
        } else {
            if (neg0) { r.x = -r.x; r.y = -r.y; }
        }
        __stcs(out + v, r);
    }
}

}  // namespace ua

using namespace ua;

extern "C" int ua_apply_sign_masks(int dtype, void *out, const void *in, int num_qubits,
                                   long long batch, int num_masks,
                                   const unsigned long long *host_masks, void *stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("ua_apply_sign_masks: bad dtype"); return UA_ERR_INVALID; }
    if (!out || !in || (num_masks > 0 && !host_masks)) { set_error("ua_apply_sign_masks: null pointer"); return UA_ERR_INVALID; }
    if (num_qubits < 1 || num_qubits > 48 || batch < 1) { set_error("ua_apply_sign_masks: bad sizes"); return UA_ERR_INVALID; }
    if (num_masks < 0 || num_masks > MAX_MASKS) { set_error("ua_apply_sign_masks: at most %d masks per call", MAX_MASKS); return UA_ERR_UNSUPPORTED; }
    if (((uintptr_t)out | (uintptr_t)in) & 15) { set_error("ua_apply_sign_masks: state pointers must be 16-byte aligned"); return UA_ERR_INVALID; }
    DiagArgs a{};
    a.in = in; a.out = out; a.n = num_qubits; a.num_masks = num_masks;
    for (int m = 0; m < num_masks; ++m) {
        if (host_masks[m] == 0 || (host_masks[m] >> num_qubits)) { set_error("ua_apply_sign_masks: mask %d out of range", m); return UA_ERR_INVALID; }
        a.masks[m] = host_masks[m];
    }
    const long long amps = batch << num_qubits;
    a.nvec = (dtype == UA_C64) ? amps / 2 : amps;
    const long long blocks = (a.nvec + 1023) / 1024;
    if (blocks > 0x7fffffffll) { set_error("ua_apply_sign_masks: grid too large"); return UA_ERR_UNSUPPORTED; }
    if (dtype == UA_C64) diag_masks_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(a);
    else diag_masks_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(a);
    return check_launch("diag_masks_kernel");
}
