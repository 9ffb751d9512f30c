// Column-wise Pearson correlation of two (n x V) matrices over subsample masks, averaged over the columns:
// the bootstrap of the encoding score (visreps/analysis/encoding_score.py:228-240),
//
//     for i in range(n_bootstrap):
//         boot_idx = rng.choice(n_test, size=int(0.9 n_test), replace=False)
//         scores[i] = correlation_score(Y_test[boot_idx], pred_test[boot_idx]).mean()
//
// where himalaya.scoring.correlation_score (himalaya 0.4.9, absent from this image) is the per-column Pearson r:
// mean over rows of zscore(y_true) * zscore(y_pred) with population standard deviations.  SURVEY.md 8f-4: the
// same 90 % draw pattern as the RSA bootstrap, over cached predictions.
//
// A replicate's column moments are sums over its kept ROWS, taken by exclusion of its dropped rows:
//     sum_kept = sum_all - sum_{i in D},   D = n - m rows (100 of 1000),
// on columns centred by their full-sample mean (fp64 throughout, so the subtraction costs nothing in accuracy).
// A column without variance in a replicate makes that replicate NaN, as the division by a zero std does upstream.
#include "common.cuh"

namespace vr {

constexpr int kColThreads = 256;
constexpr int kColMaxDropped = 4096;

// One thread per column: column means, then the five centred totals.  totals: [5][V] = Sy, Sp, Syy, Spp, Syp.
__global__ void __launch_bounds__(kColThreads)
colcorr_totals_kernel(const float* __restrict__ Y, const float* __restrict__ P, int n, int V, int64_t ldy, int64_t ldp,
                      double* __restrict__ means, double* __restrict__ totals) {
  const int v = blockIdx.x * kColThreads + threadIdx.x;
  if (v >= V) return;
  double sy = 0.0, sp = 0.0;
  for (int i = 0; i < n; ++i) { sy += (double)Y[(int64_t)i * ldy + v]; sp += (double)P[(int64_t)i * ldp + v]; }
  const double my = sy / n, mp = sp / n;
  double a = 0.0, b = 0.0, aa = 0.0, bb = 0.0, ab = 0.0;
  for (int i = 0; i < n; ++i) {
    const double y = (double)Y[(int64_t)i * ldy + v] - my, p = (double)P[(int64_t)i * ldp + v] - mp;
    a += y; b += p; aa += y * y; bb += p * p; ab += y * p;
  }
  means[v] = my; means[V + v] = mp;
  totals[v] = a; totals[V + v] = b; totals[2 * (int64_t)V + v] = aa; totals[3 * (int64_t)V + v] = bb; totals[4 * (int64_t)V + v] = ab;
}

// grid (column blocks, R).  Writes r per (replicate, column) when asked, and one partial sum per (replicate, block).
__global__ void __launch_bounds__(kColThreads)
colcorr_subsample_kernel(const float* __restrict__ Y, const float* __restrict__ P, int n, int V, int64_t ldy, int64_t ldp,
                         const uint32_t* __restrict__ masks, int W, const double* __restrict__ means,
                         const double* __restrict__ totals, float* __restrict__ per_column, double* __restrict__ partial) {
  extern __shared__ uint32_t s_mask[];
  __shared__ uint16_t s_rows[kColMaxDropped];
  __shared__ int s_wcnt[kColThreads / 32];
  __shared__ double s_red[kColThreads / 32];
  const int r = blockIdx.y;
  for (int k = threadIdx.x; k < W; k += kColThreads) s_mask[k] = masks[(int64_t)r * W + k];
  __syncthreads();
  // rows to visit, ascending: the dropped rows if they fit the list (exclusion), else the kept rows (direct sum)
  int ndrop = 0;
  for (int i = threadIdx.x; i < n; i += kColThreads) ndrop += !((s_mask[i >> 5] >> (i & 31)) & 1u);
  ndrop = (int)warp_sum_ll((long long)ndrop);
  __syncthreads();
  if (lane_id() == 0) s_wcnt[threadIdx.x >> 5] = ndrop;
  __syncthreads();
  ndrop = 0;
  for (int w = 0; w < kColThreads / 32; ++w) ndrop += s_wcnt[w];
  __syncthreads();
  const bool by_dropped = ndrop <= kColMaxDropped;
  const int m = n - ndrop;
  int listed = 0;
  if (by_dropped) {
    for (int i0 = 0; i0 < n; i0 += kColThreads) {
      const int i = i0 + (int)threadIdx.x;
      const bool drop = i < n && !((s_mask[i >> 5] >> (i & 31)) & 1u);
      const unsigned bal = __ballot_sync(0xffffffffu, drop);
      if (lane_id() == 0) s_wcnt[threadIdx.x >> 5] = __popc(bal);
      __syncthreads();
      int off = listed, tot = 0;
      for (int k = 0; k < kColThreads / 32; ++k) { const int c = s_wcnt[k]; if (k < (int)(threadIdx.x >> 5)) off += c; tot += c; }
      if (drop) s_rows[off + __popc(bal & lanemask_lt())] = (uint16_t)i;
      listed += tot;
      __syncthreads();
    }
  }
  const int v = blockIdx.x * kColThreads + threadIdx.x;
  double rv = 0.0;
  if (v < V) {
    const double my = means[v], mp = means[V + v];
    double a = 0.0, b = 0.0, aa = 0.0, bb = 0.0, ab = 0.0;
    auto add_row = [&](int i) {
      const double y = (double)__ldg(Y + (int64_t)i * ldy + v) - my, p = (double)__ldg(P + (int64_t)i * ldp + v) - mp;
      a += y; b += p; aa += y * y; bb += p * p; ab += y * p;
    };
    if (by_dropped) {
      for (int t = 0; t < listed; ++t) add_row(s_rows[t]);
      a = totals[v] - a; b = totals[V + v] - b; aa = totals[2 * (int64_t)V + v] - aa; bb = totals[3 * (int64_t)V + v] - bb;
      ab = totals[4 * (int64_t)V + v] - ab;
    } else {
      for (int i = 0; i < n; ++i) if ((s_mask[i >> 5] >> (i & 31)) & 1u) add_row(i);
    }
    const double mm = (double)m;
    const double va = aa - a * a / mm, vb = bb - b * b / mm;
    rv = (ab - a * b / mm) / (sqrt(va) * sqrt(vb));
    if (m < 1 || !(va > 0.0) || !(vb > 0.0)) rv = __longlong_as_double(0x7ff8000000000000LL);
    if (per_column) per_column[(int64_t)r * V + v] = (float)rv;
  }
  double s = warp_sum(rv);
  if (lane_id() == 0) s_red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kColThreads / 32; ++w) t += s_red[w];
    partial[(int64_t)r * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void colcorr_mean_kernel(const double* __restrict__ partial, int blocks, int V, int R, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double t = 0.0;
  for (int b = 0; b < blocks; ++b) t += partial[(int64_t)r * blocks + b];     // fixed order: reproducible
  out[r] = t / (double)V;
}

}  // namespace vr

using namespace vr;

extern "C" size_t vr_bootstrap_column_corr_workspace_bytes(int n, int V, int R) {
  if (n < 1 || V < 1 || R < 1) return 256;
  const size_t blocks = ((size_t)V + kColThreads - 1) / kColThreads;
  return arena_bytes((size_t)2 * V, 8) + arena_bytes((size_t)5 * V, 8) + arena_bytes((size_t)R * blocks, 8) + 256;
}

extern "C" int vr_bootstrap_column_corr_f32(const float* y_true, const float* y_pred, int n, int V, int64_t ld_true, int64_t ld_pred,
                                            const uint32_t* masks, int R, double* mean_out, float* per_column,
                                            void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(y_true && y_pred && masks && mean_out, "vr_bootstrap_column_corr_f32: null pointer");
  VR_REQUIRE(n >= 1 && n <= VR_MAX_STIMULI && V >= 1 && ld_true >= V && ld_pred >= V && R >= 1 && R <= 65535,
             "vr_bootstrap_column_corr_f32: bad sizes n=%d V=%d R=%d", n, V, R);
  if (!ws || ws_bytes < vr_bootstrap_column_corr_workspace_bytes(n, V, R)) {
    set_error("vr_bootstrap_column_corr_f32: workspace %zu < %zu", ws_bytes, vr_bootstrap_column_corr_workspace_bytes(n, V, R));
    return VR_ERR_WORKSPACE;
  }
  const int blocks = (V + kColThreads - 1) / kColThreads;
  Arena a(ws, ws_bytes);
  double* means = a.take<double>((size_t)2 * V);
  double* totals = a.take<double>((size_t)5 * V);
  double* partial = a.take<double>((size_t)R * blocks);
  if (!a.ok) { set_error("vr_bootstrap_column_corr_f32: internal workspace plan overflow"); return VR_ERR_WORKSPACE; }
  const int W = vr_mask_words(n);
  ProfScope prof(F_COLCORR, stream);
  colcorr_totals_kernel<<<blocks, kColThreads, 0, stream>>>(y_true, y_pred, n, V, ld_true, ld_pred, means, totals);
  VR_CHECK_LAUNCH();
  colcorr_subsample_kernel<<<dim3(blocks, R), kColThreads, (size_t)W * 4, stream>>>(y_true, y_pred, n, V, ld_true, ld_pred, masks, W,
                                                                                    means, totals, per_column, partial);
  VR_CHECK_LAUNCH();
  colcorr_mean_kernel<<<(R + 127) / 128, 128, 0, stream>>>(partial, blocks, V, R, mean_out);
  VR_CHECK_LAUNCH();
  return VR_OK;
}
