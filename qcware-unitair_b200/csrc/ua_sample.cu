// Sampling in the computational basis without materialising the probability vector.
//
// The reference's measure() (src/unitair/simulation/measurement.py:40-43) builds
// probs = abs_squared(state) (a full-size real tensor), hands it to
// torch.distributions.Categorical (normalisation + cumulative sums, more full-size passes and
// temporaries) and samples.  Here sampling is inverse-CDF in two small steps:
//   1. sample_block_sums_kernel: ONE read of the state, fp64 sum of |psi|^2 per block of
//      2^block_log2 amplitudes (2^18 numbers for a 30-qubit state);
//   2. (host shim: cumulative sum of the block sums, uniform draws, searchsorted -- tiny tensors)
//   3. sample_locate_kernel: one warp per sample walks the one block its draw fell into
//      (32 KiB) and returns the amplitude index whose cumulative probability crosses the draw.
// Algorithmic traffic: 8 B / 16 B per amplitude once, plus one block per sample.
#include "ua_common.cuh"

namespace ua {

template <typename R>
__global__ void __launch_bounds__(256) sample_block_sums_kernel(const typename CplxOf<R>::type *__restrict__ in,
                                                                double *__restrict__ out, long long elems,
                                                                int block_log2, long long num_blocks) {
    using V = typename VecOf<R>::type;
    constexpr int APV = VecOf<R>::APV;
    const int warps_per_cta = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const long long stride = (long long)gridDim.x * warps_per_cta;
    const long long bsize = 1ll << block_log2;
    for (long long b = warp0; b < num_blocks; b += stride) {      // one warp per block
        const long long lo = b * bsize;
        const long long hi = (lo + bsize < elems) ? lo + bsize : elems;
        double acc = 0.0;
        if (((lo | bsize) & (APV - 1)) == 0 && hi == lo + bsize) {
            const V *v = reinterpret_cast<const V *>(in + lo);
            const long long nv = bsize / APV;
            for (long long i = lane; i < nv; i += 32) {
                const V x = ld16<true>(v + i);
                if constexpr (APV == 2) {
                    // same fp64 arithmetic as sample_locate_kernel: the block sums and the walk
                    // inside a block must agree
                    acc += (double)x.x * (double)x.x + (double)x.y * (double)x.y;
                    acc += (double)x.z * (double)x.z + (double)x.w * (double)x.w;
                } else {
                    acc += x.x * x.x + x.y * x.y;
                }
            }
        } else {
            for (long long i = lo + lane; i < hi; i += 32) {
                const auto x = in[i];
                acc += (double)x.x * (double)x.x + (double)x.y * (double)x.y;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[b] = acc;
    }
}

// One warp per sample.  targets[s] is the draw scaled to [0, total probability); block_cdf is the
// INCLUSIVE cumulative sum of the block sums.  The block is found by binary search, then every
// lane sums a contiguous slice of the block, a warp scan finds the slice, and that lane walks it.
template <typename R>
__global__ void __launch_bounds__(256) sample_locate_kernel(const typename CplxOf<R>::type *__restrict__ in,
                                                            long long *__restrict__ out, long long elems,
                                                            int block_log2, long long num_blocks,
                                                            const double *__restrict__ block_cdf,
                                                            const double *__restrict__ targets,
                                                            long long num_samples) {
    const int lane = threadIdx.x & 31;
    const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= num_samples) return;
    const double t = targets[s];
    // first block whose inclusive cdf exceeds t
    long long lo = 0, hi = num_blocks - 1;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (block_cdf[mid] > t) hi = mid; else lo = mid + 1;
    }
    const long long b = lo;
    const double before = (b > 0) ? block_cdf[b - 1] : 0.0;
    double r = t - before;                      // residual inside the block
    const long long bsize = 1ll << block_log2;
    const long long start = b * bsize;
    const long long end = (start + bsize < elems) ? start + bsize : elems;
    const long long len = end - start;
    const long long per = (len + 31) / 32;
    const long long mylo = start + (long long)lane * per;
    const long long myhi = (mylo + per < end) ? mylo + per : end;
    double mine = 0.0;
    for (long long i = mylo; i < myhi; ++i) {
        const auto x = in[i];
        mine += (double)x.x * (double)x.x + (double)x.y * (double)x.y;
    }
    // inclusive scan over lanes
    double incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, incl > r && myhi > mylo);
    // rounding can leave r marginally above the block's own sum: fall back to the last
    // non-empty lane
    int owner;
    if (ballot) owner = __ffs(ballot) - 1;
    else {
        const unsigned nonempty = __ballot_sync(0xffffffffu, myhi > mylo);
        owner = 31 - __clz(nonempty);
    }
    const double excl = incl - mine;
    const double r_owner = r - __shfl_sync(0xffffffffu, excl, owner);
    if (lane == owner) {
        // if rounding leaves r_owner at or above the lane's own sum, fall back to the LAST amplitude
        // with non-zero probability (never an amplitude that cannot be measured)
        double acc = 0.0;
        long long idx = -1, last_nonzero = myhi - 1;
        for (long long i = mylo; i < myhi; ++i) {
            const auto x = in[i];
            const double pr = (double)x.x * (double)x.x + (double)x.y * (double)x.y;
            acc += pr;
            if (pr > 0.0) last_nonzero = i;
            if (acc > r_owner) { idx = i; break; }
        }
        out[s] = idx >= 0 ? idx : last_nonzero;
    }
}

}  // namespace ua

using namespace ua;

extern "C" int ua_sample_block_sums(int dtype, void *out_f64, const void *in, long long elems,
                                    int block_log2, void *stream) {
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("ua_sample_block_sums: bad dtype"); return UA_ERR_INVALID; }
    if (!out_f64 || !in || elems < 1 || block_log2 < 5 || block_log2 > 20) {
        set_error("ua_sample_block_sums: bad argument (elems=%lld block_log2=%d)", elems, block_log2);
        return UA_ERR_INVALID;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long num_blocks = (elems + (1ll << block_log2) - 1) >> block_log2;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (num_blocks + 7) / 8;
    if (grid > (long long)sms * 8) grid = (long long)sms * 8;
    if (dtype == UA_C64)
        sample_block_sums_kernel<float><<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const float2 *>(in),
                                                                      reinterpret_cast<double *>(out_f64), elems,
                                                                      block_log2, num_blocks);
    else
        sample_block_sums_kernel<double><<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const double2 *>(in),
                                                                       reinterpret_cast<double *>(out_f64), elems,
                                                                       block_log2, num_blocks);
    return check_launch("sample_block_sums_kernel");
}

extern "C" int ua_sample_locate(int dtype, void *out_index_i64, const void *in, long long elems,
                                int block_log2, const void *block_cdf_f64, const void *targets_f64,
                                long long num_samples, void *stream) {
    if (dtype != UA_C64 && dtype != UA_C128) { set_error("ua_sample_locate: bad dtype"); return UA_ERR_INVALID; }
    if (!out_index_i64 || !in || !block_cdf_f64 || !targets_f64 || elems < 1 || num_samples < 0 ||
        block_log2 < 5 || block_log2 > 20) {
        set_error("ua_sample_locate: bad argument"); return UA_ERR_INVALID;
    }
    if (num_samples == 0) return UA_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long num_blocks = (elems + (1ll << block_log2) - 1) >> block_log2;
    const long long grid = (num_samples + 7) / 8;
    if (grid > 0x7fffffffll) { set_error("ua_sample_locate: too many samples"); return UA_ERR_INVALID; }
    if (dtype == UA_C64)
        sample_locate_kernel<float><<<(unsigned)grid, 256, 0, st>>>(
            reinterpret_cast<const float2 *>(in), reinterpret_cast<long long *>(out_index_i64), elems, block_log2,
            num_blocks, reinterpret_cast<const double *>(block_cdf_f64), reinterpret_cast<const double *>(targets_f64),
            num_samples);
    else
        sample_locate_kernel<double><<<(unsigned)grid, 256, 0, st>>>(
            reinterpret_cast<const double2 *>(in), reinterpret_cast<long long *>(out_index_i64), elems, block_log2,
            num_blocks, reinterpret_cast<const double *>(block_cdf_f64), reinterpret_cast<const double *>(targets_f64),
            num_samples);
    return check_launch("sample_locate_kernel");
}
