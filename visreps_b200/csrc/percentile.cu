// Percentile confidence interval of the bootstrap replicates, on the device.
//
// Replaces `np.percentile(scores, [2.5, 97.5])` at visreps/evals.py:371-372, :520-521 and visreps/analysis/rsa.py:262-263
// (NumPy's default 'linear' method), bit for bit: virtual index q / 100 * (B - 1), its floor k and fraction t, and NumPy's
// _lerp -- lower + diff * t below t = 0.5, upper - diff * (1 - t) from 0.5 on -- with every product and sum rounded on its own
// (no FMA contraction).  A row that holds a NaN gives NaN for both bounds, as NumPy does.
// One CTA per row of replicates: the row is sorted in shared memory (bitonic, order-preserving 64-bit keys, NaN last).
#include "common.cuh"

namespace vr {

constexpr int kPercentileMaxB = 16384;       // 128 KB of keys in shared memory

__device__ __forceinline__ unsigned long long double_key(double v) {
  if (v != v) return ~0ull;
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_double(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256)
percentile_ci_kernel(const double* __restrict__ scores, int64_t row_stride, int B, int P, double q_lo, double q_hi,
                     double* __restrict__ ci_lo, double* __restrict__ ci_hi) {
  extern __shared__ __align__(16) unsigned long long keys[];
  __shared__ int s_nan;
  const double* row = scores + (int64_t)blockIdx.x * row_stride;
  if (threadIdx.x == 0) s_nan = 0;
  __syncthreads();
  for (int k = threadIdx.x; k < P; k += 256) {
    unsigned long long key = ~0ull;                       // padding sorts behind everything (with the NaNs)
    if (k < B) {
      const double v = row[k];
      if (v != v) s_nan = 1;
      key = double_key(v);
    }
    keys[k] = key;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += 256) {
        const int lo = (t / stride) * (stride << 1) + (t % stride), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  if (threadIdx.x < 2) {
    const double q = threadIdx.x == 0 ? q_lo : q_hi;
    double* out = threadIdx.x == 0 ? ci_lo : ci_hi;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (s_nan) { out[blockIdx.x] = nan; return; }
    const double virt = __dmul_rn(__ddiv_rn(q, 100.0), (double)(B - 1));
    const double kf = floor(virt);
    const int k = (int)kf;
    const double t = __dsub_rn(virt, kf);
    const double lower = key_double(keys[k]), upper = key_double(keys[k + 1 < B ? k + 1 : B - 1]);
    const double diff = __dsub_rn(upper, lower);
    out[blockIdx.x] = t >= 0.5 ? __dsub_rn(upper, __dmul_rn(diff, __dsub_rn(1.0, t))) : __dadd_rn(lower, __dmul_rn(diff, t));
  }
}

}  // namespace vr

using namespace vr;

extern "C" int vr_percentile_ci_max_replicates(void) { return kPercentileMaxB; }

extern "C" int vr_percentile_ci_f64(const double* scores, int rows, int B, int64_t row_stride, double q_lo, double q_hi,
                                    double* ci_lo, double* ci_hi, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(scores && ci_lo && ci_hi, "vr_percentile_ci_f64: null pointer");
  VR_REQUIRE(rows >= 0 && B >= 1 && B <= kPercentileMaxB && row_stride >= B, "vr_percentile_ci_f64: bad sizes rows=%d B=%d stride=%lld (B <= %d)",
             rows, B, (long long)row_stride, kPercentileMaxB);
  VR_REQUIRE(q_lo >= 0.0 && q_lo <= 100.0 && q_hi >= 0.0 && q_hi <= 100.0, "vr_percentile_ci_f64: percentiles must lie in [0, 100]");
  if (rows == 0) return VR_OK;
  int P = 2;
  while (P < B) P <<= 1;
  const size_t smem = (size_t)P * 8;
  static unsigned long long attr = 0;
  if (smem > 40 * 1024) VR_CHECK_CUDA(ensure_dynamic_smem(percentile_ci_kernel, kPercentileMaxB * 8, &attr));
  percentile_ci_kernel<<<rows, 256, smem, stream>>>(scores, row_stride, B, P, q_lo, q_hi, ci_lo, ci_hi);
  VR_CHECK_LAUNCH();
  return VR_OK;
}
