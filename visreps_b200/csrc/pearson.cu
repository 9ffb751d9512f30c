// Grouped Pearson comparison of RDM upper triangles over subsample masks.
//
// Replaces `rdm[idx][:, idx]` + triu gather + scipy.stats.pearsonr
// (visreps/evals.py:365-368 with correlation="Pearson"; visreps/analysis/rsa.py:43-47,111-122).
// The kept pairs of replicate r are exactly {(i, j): i < j, bit i and bit j set in mask r}, so the
// kernel walks the strict upper triangle of the FULL matrices in place (coalesced along j) and
// accumulates the five moments in fp64, shifted by one matrix entry to avoid cancellation.
// (SciPy evaluates pearsonr in float32 for float32 input; fp64 here is well inside the 1e-5 bar.)
#include "common.cuh"

namespace vr {

constexpr int kPearThreads = 256;
constexpr int kPearTile = 4;
constexpr int kPearMaxMats = 64;
struct MatList { const float* p[kPearMaxMats]; };

__global__ void __launch_bounds__(kPearThreads)
pearson_kernel(const __grid_constant__ MatList mats, int nA, int nB, int n, int64_t ld,
               const uint32_t* __restrict__ masks, int W, int R, double* __restrict__ out) {
  extern __shared__ uint32_t s_mask[];
  constexpr int T = kPearTile;
  __shared__ double s_red[kPearThreads / 32][4 * T + T * T];
  const int tilesB = (nB + T - 1) / T;
  const int ta = blockIdx.x / tilesB, tb = blockIdx.x % tilesB;
  const int r = blockIdx.y;
  for (int w = threadIdx.x; w < W; w += kPearThreads) s_mask[w] = masks[(int64_t)r * W + w];
  __syncthreads();

  const float* A[T];
  const float* B[T];
  float ca[T], cb[T];
#pragma unroll
  for (int k = 0; k < T; ++k) {
    A[k] = mats.p[min(ta * T + k, nA - 1)];
    B[k] = mats.p[nA + min(tb * T + k, nB - 1)];
    ca[k] = __ldg(A[k] + 1);  // entry (0,1): a typical value, used as the shift
    cb[k] = __ldg(B[k] + 1);
  }
  double s1a[T], s2a[T], s1b[T], s2b[T], sab[T][T];
#pragma unroll
  for (int k = 0; k < T; ++k) {
    s1a[k] = s2a[k] = s1b[k] = s2b[k] = 0.0;
#pragma unroll
    for (int l = 0; l < T; ++l) sab[k][l] = 0.0;
  }
  long long cnt = 0;
  for (int i = 0; i < n - 1; ++i) {
    if (!((s_mask[i >> 5] >> (i & 31)) & 1u)) continue;
    for (int j = i + 1 + threadIdx.x; j < n; j += kPearThreads) {
      if (!((s_mask[j >> 5] >> (j & 31)) & 1u)) continue;
      const int64_t off = (int64_t)i * ld + j;
      double xa[T], xb[T];
#pragma unroll
      for (int k = 0; k < T; ++k) {
        xa[k] = (double)(__ldg(A[k] + off) - ca[k]);
        xb[k] = (double)(__ldg(B[k] + off) - cb[k]);
      }
      ++cnt;
#pragma unroll
      for (int k = 0; k < T; ++k) {
        s1a[k] += xa[k]; s2a[k] += xa[k] * xa[k];
        s1b[k] += xb[k]; s2b[k] += xb[k] * xb[k];
#pragma unroll
        for (int l = 0; l < T; ++l) sab[k][l] += xa[k] * xb[l];
      }
    }
  }
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  cnt = warp_sum_ll(cnt);
#pragma unroll
  for (int k = 0; k < T; ++k) {
    double v0 = warp_sum(s1a[k]), v1 = warp_sum(s2a[k]), v2 = warp_sum(s1b[k]), v3 = warp_sum(s2b[k]);
    if (lane == 0) {
      s_red[warp][k] = v0; s_red[warp][T + k] = v1; s_red[warp][2 * T + k] = v2; s_red[warp][3 * T + k] = v3;
    }
#pragma unroll
    for (int l = 0; l < T; ++l) {
      double v = warp_sum(sab[k][l]);
      if (lane == 0) s_red[warp][4 * T + k * T + l] = v;
    }
  }
  __shared__ long long s_cnt[kPearThreads / 32];
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x < T * T) {
    const int k = threadIdx.x / T, l = threadIdx.x % T;
    double S1a = 0, S2a = 0, S1b = 0, S2b = 0, Sab = 0;
    long long c = 0;
    for (int w = 0; w < kPearThreads / 32; ++w) {
      S1a += s_red[w][k]; S2a += s_red[w][T + k]; S1b += s_red[w][2 * T + l]; S2b += s_red[w][3 * T + l];
      Sab += s_red[w][4 * T + k * T + l];
      c += s_cnt[w];
    }
    const int a = ta * T + k, b = tb * T + l;
    if (a < nA && b < nB) {
      const double Nn = (double)c;
      const double va = S2a - S1a * S1a / Nn, vb = S2b - S1b * S1b / Nn;
      double rho = (Sab - S1a * S1b / Nn) / (sqrt(va) * sqrt(vb));
      if (c < 2 || !(va > 0.0) || !(vb > 0.0)) rho = __longlong_as_double(0x7ff8000000000000LL);
      if (rho > 1.0) rho = 1.0;
      if (rho < -1.0) rho = -1.0;
      out[((int64_t)a * nB + b) * R + r] = rho;
    }
  }
}

}  // namespace vr

using namespace vr;

extern "C" size_t vr_bootstrap_pearson_workspace_bytes(int, int, int, int) { return 256; }

extern "C" int vr_bootstrap_pearson(const float* const* A, int nA, const float* const* B, int nB, int n, int64_t ld,
                                    const uint32_t* masks, int R, double* out, void*, size_t, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(A && B && masks && out, "vr_bootstrap_pearson: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && ld >= n, "vr_bootstrap_pearson: bad n=%d ld=%lld", n, (long long)ld);
  VR_REQUIRE(nA >= 1 && nB >= 1 && nA + nB <= kPearMaxMats, "vr_bootstrap_pearson: nA=%d nB=%d (max %d in total)", nA, nB, kPearMaxMats);
  VR_REQUIRE(R >= 1 && R <= 65535, "vr_bootstrap_pearson: R=%d outside [1, 65535]", R);
  MatList ml{};
  for (int i = 0; i < nA; ++i) { VR_REQUIRE(A[i], "null matrix"); ml.p[i] = A[i]; }
  for (int i = 0; i < nB; ++i) { VR_REQUIRE(B[i], "null matrix"); ml.p[nA + i] = B[i]; }
  const int W = vr_mask_words(n);
  const int tiles = ((nA + kPearTile - 1) / kPearTile) * ((nB + kPearTile - 1) / kPearTile);
  dim3 grid(tiles, R);
  ProfScope prof(F_PEARSON, stream);
  pearson_kernel<<<grid, kPearThreads, (size_t)W * 4, stream>>>(ml, nA, nB, n, ld, masks, W, R, out);
  VR_CHECK_LAUNCH();
  return VR_OK;
}
