// Grouped Pearson comparison of RDM upper triangles over subsample masks.
//
// Replaces `rdm[idx][:, idx]` + triu gather + scipy.stats.pearsonr
// (visreps/evals.py:365-368 with correlation="Pearson"; visreps/analysis/rsa.py:43-47,111-122).
//
// The kept pairs of replicate r are {(i, j): i < j, both kept}.  Every moment the coefficient needs is a sum
// of a fixed matrix M (entries x, x^2, x*y ...) over those pairs, and the 90 % subsamples of the reference drop
// only D = n - m stimuli, so the sum is taken by inclusion-exclusion over the DROPPED stimuli:
//
//     sum_{kept pairs} M = sum_{all pairs} M  -  sum_{i in D} rowsum_i(M)  +  sum_{i<j in D} M_ij
//
// (a pair with exactly one dropped end is removed once, a pair with two dropped ends is removed twice and
// added back once).  Totals and row sums are built once per call; a replicate then touches D(D-1)/2 entries
// per matrix instead of m(m-1)/2: 4,950 instead of 404,550 at n = 1000.  NaN entries are summed as zero and
// counted with the same identity; a replicate whose kept pairs hold a NaN yields NaN (pearsonr would).
// A mask that drops more than kMaxDropped stimuli falls back to the direct sum over its kept pairs.
// All moments are fp64, shifted by one matrix entry to avoid cancellation.
// (SciPy evaluates pearsonr in float32 for float32 input; fp64 here is well inside the 1e-5 bar.)
#include "common.cuh"

namespace vr {

constexpr int kPearThreads = 256;
constexpr int kPearWarps = kPearThreads / 32;
constexpr int kPearB = 8;            // neural matrices handled together by one CTA
constexpr int kPearMaxMats = 64;
constexpr int kMaxDropped = 2048;    // dropped-stimulus list kept in shared memory
struct MatList { const float* p[kPearMaxMats]; };

// per-vector arrays: [vector][3][n] = row sums of x, x^2 and NaN counts; cross: [a][b][n] row sums of x*y;
// totals: [vector][3] and [a][b] (sums over i < j).
struct PearsonWork {
  double* rows;     // (nA + nB) * 3 * n
  double* cross;    // nA * nB * n
  double* tot;      // (nA + nB) * 3
  double* totx;     // nA * nB
};

// A typical value of the matrix, used as the shift: entry (0,1), or 0 if that entry is NaN.
__device__ __forceinline__ float shift_of(const float* m) { const float c = __ldg(m + 1); return c == c ? c : 0.f; }

__device__ __forceinline__ double block_sum(double v, double* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if (lane_id() == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kPearWarps; ++w) t += s_red[w];
  return t;
}

// Only the strict upper triangle of a matrix is ever read (rsa.py:111-113): entry {i, j} lives at (min, max).
__device__ __forceinline__ int64_t triu_off(int i, int j, int64_t ld) { return i < j ? (int64_t)i * ld + j : (int64_t)j * ld + i; }

// Row sums.  grid (n, nA + nB): vector v = blockIdx.y; model vectors (v < nA) also produce the cross row sums
// against every neural vector.
__global__ void __launch_bounds__(kPearThreads)
pearson_rows_kernel(const __grid_constant__ MatList mats, int nA, int nB, int n, int64_t ld, PearsonWork w) {
  __shared__ double s_red[kPearWarps];
  const int i = blockIdx.x, v = blockIdx.y;
  const float* __restrict__ X = mats.p[v];
  const float cx = shift_of(X);
  double s1 = 0.0, s2 = 0.0, sn = 0.0;
  for (int j = threadIdx.x; j < n; j += kPearThreads) {
    if (j == i) continue;
    const float x = __ldg(X + triu_off(i, j, ld));
    if (x != x) { sn += 1.0; continue; }
    const double d = (double)(x - cx);
    s1 += d; s2 += d * d;
  }
  s1 = block_sum(s1, s_red); s2 = block_sum(s2, s_red); sn = block_sum(sn, s_red);
  if (threadIdx.x == 0) {
    w.rows[((int64_t)v * 3 + 0) * n + i] = s1;
    w.rows[((int64_t)v * 3 + 1) * n + i] = s2;
    w.rows[((int64_t)v * 3 + 2) * n + i] = sn;
  }
  if (v >= nA) return;
  for (int b0 = 0; b0 < nB; b0 += kPearB) {
    double sx[kPearB];
    const float* Y[kPearB];
    float cy[kPearB];
#pragma unroll
    for (int k = 0; k < kPearB; ++k) { sx[k] = 0.0; Y[k] = mats.p[nA + min(b0 + k, nB - 1)]; cy[k] = shift_of(Y[k]); }
    for (int j = threadIdx.x; j < n; j += kPearThreads) {
      if (j == i) continue;
      const int64_t off = triu_off(i, j, ld);
      const float x = __ldg(X + off);
      if (x != x) continue;
      const double d = (double)(x - cx);
#pragma unroll
      for (int k = 0; k < kPearB; ++k) {
        const float y = __ldg(Y[k] + off);
        if (y == y) sx[k] += d * (double)(y - cy[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < kPearB; ++k) {
      const double t = block_sum(sx[k], s_red);
      if (threadIdx.x == 0 && b0 + k < nB) w.cross[((int64_t)v * nB + b0 + k) * n + i] = t;
    }
  }
}

// Totals over i < j = half the sum of the row sums.  One CTA per array.
__global__ void __launch_bounds__(kPearThreads)
pearson_totals_kernel(int nA, int nB, int n, PearsonWork w) {
  __shared__ double s_red[kPearWarps];
  const int nv3 = (nA + nB) * 3;
  const double* src = (int)blockIdx.x < nv3 ? w.rows + (int64_t)blockIdx.x * n : w.cross + (int64_t)(blockIdx.x - nv3) * n;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += kPearThreads) s += src[i];
  s = block_sum(s, s_red);
  if (threadIdx.x == 0) {
    if ((int)blockIdx.x < nv3) w.tot[blockIdx.x] = 0.5 * s;
    else w.totx[blockIdx.x - nv3] = 0.5 * s;
  }
}

// grid (R, nA): one model vector against all neural vectors for one mask.
__global__ void __launch_bounds__(kPearThreads)
pearson_subsample_kernel(const __grid_constant__ MatList mats, int nA, int nB, int n, int64_t ld,
                         const uint32_t* __restrict__ masks, int W, int R, PearsonWork w, double* __restrict__ out) {
  extern __shared__ uint32_t s_mask[];                 // W mask words
  __shared__ uint16_t s_drop[kMaxDropped];
  __shared__ double s_red[kPearWarps];
  __shared__ int s_wcnt[kPearWarps];
  const int r = blockIdx.x, a = blockIdx.y;
  for (int k = threadIdx.x; k < W; k += kPearThreads) s_mask[k] = masks[(int64_t)r * W + k];
  __syncthreads();
  // dropped stimuli in ascending order (a fixed order keeps the fp64 sums reproducible)
  int ndrop = 0;
  for (int i0 = 0; i0 < n; i0 += kPearThreads) {
    const int i = i0 + (int)threadIdx.x;
    const bool drop = i < n && !((s_mask[i >> 5] >> (i & 31)) & 1u);
    const unsigned bal = __ballot_sync(0xffffffffu, drop);
    if (lane_id() == 0) s_wcnt[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    int off = ndrop, tot = 0;
#pragma unroll
    for (int k = 0; k < kPearWarps; ++k) { const int c = s_wcnt[k]; if (k < (int)(threadIdx.x >> 5)) off += c; tot += c; }
    off += __popc(bal & lanemask_lt());
    if (drop && off < kMaxDropped) s_drop[off] = (uint16_t)i;
    ndrop += tot;
    __syncthreads();
  }
  const bool by_dropped = ndrop <= kMaxDropped;
  const long long m = n - ndrop;
  const double Nk = (double)(m * (m - 1) / 2);
  const double nan = __longlong_as_double(0x7ff8000000000000LL);

  const float* __restrict__ X = mats.p[a];
  const float cx = shift_of(X);
  for (int b0 = 0; b0 < nB; b0 += kPearB) {
    const float* Y[kPearB];
    float cy[kPearB];
#pragma unroll
    for (int k = 0; k < kPearB; ++k) { Y[k] = mats.p[nA + min(b0 + k, nB - 1)]; cy[k] = shift_of(Y[k]); }
    // signed corrections to the totals (by_dropped) or the kept sums themselves (direct)
    double ax1 = 0.0, ax2 = 0.0, axn = 0.0, ay1[kPearB], ay2[kPearB], ayn[kPearB], axy[kPearB];
#pragma unroll
    for (int k = 0; k < kPearB; ++k) ay1[k] = ay2[k] = ayn[k] = axy[k] = 0.0;
    auto add_pair = [&](int i, int j) {
      const int64_t off = (int64_t)i * ld + j;
      const float x = __ldg(X + off);
      const bool xn = x != x;
      const double dx = xn ? 0.0 : (double)(x - cx);
      ax1 += dx; ax2 += dx * dx; axn += xn ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < kPearB; ++k) {
        const float y = __ldg(Y[k] + off);
        const bool yn = y != y;
        const double dy = yn ? 0.0 : (double)(y - cy[k]);
        ay1[k] += dy; ay2[k] += dy * dy; ayn[k] += yn ? 1.0 : 0.0;
        axy[k] += dx * dy;
      }
    };
    if (by_dropped) {
      for (int t = threadIdx.x; t < ndrop; t += kPearThreads) {       // - row sums of the dropped stimuli
        const int i = s_drop[t];
        ax1 -= w.rows[((int64_t)a * 3 + 0) * n + i];
        ax2 -= w.rows[((int64_t)a * 3 + 1) * n + i];
        axn -= w.rows[((int64_t)a * 3 + 2) * n + i];
#pragma unroll
        for (int k = 0; k < kPearB; ++k) {
          const int b = min(b0 + k, nB - 1);
          ay1[k] -= w.rows[((int64_t)(nA + b) * 3 + 0) * n + i];
          ay2[k] -= w.rows[((int64_t)(nA + b) * 3 + 1) * n + i];
          ayn[k] -= w.rows[((int64_t)(nA + b) * 3 + 2) * n + i];
          axy[k] -= w.cross[((int64_t)a * nB + b) * n + i];
        }
      }
      const int npair = ndrop * (ndrop - 1) / 2;                      // + the pairs dropped at both ends
      for (int t = threadIdx.x; t < npair; t += kPearThreads) {
        int q = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
        while (q * (q - 1) / 2 > t) --q;
        while ((q + 1) * q / 2 <= t) ++q;
        const int p = t - q * (q - 1) / 2;                            // 0 <= p < q < ndrop
        const int i = s_drop[p], j = s_drop[q];
        add_pair(min(i, j), max(i, j));
      }
    } else {
      for (int i = 0; i < n - 1; ++i) {
        if (!((s_mask[i >> 5] >> (i & 31)) & 1u)) continue;
        for (int j = i + 1 + threadIdx.x; j < n; j += kPearThreads)
          if ((s_mask[j >> 5] >> (j & 31)) & 1u) add_pair(i, j);
      }
    }
    const double base = by_dropped ? 1.0 : 0.0;
    const double Sx1 = base * w.tot[a * 3 + 0] + block_sum(ax1, s_red);
    const double Sx2 = base * w.tot[a * 3 + 1] + block_sum(ax2, s_red);
    const double Sxn = base * w.tot[a * 3 + 2] + block_sum(axn, s_red);
#pragma unroll
    for (int k = 0; k < kPearB; ++k) {
      const int b = min(b0 + k, nB - 1);
      const double Sy1 = base * w.tot[(nA + b) * 3 + 0] + block_sum(ay1[k], s_red);
      const double Sy2 = base * w.tot[(nA + b) * 3 + 1] + block_sum(ay2[k], s_red);
      const double Syn = base * w.tot[(nA + b) * 3 + 2] + block_sum(ayn[k], s_red);
      const double Sxy = base * w.totx[a * nB + b] + block_sum(axy[k], s_red);
      if (threadIdx.x == 0 && b0 + k < nB) {
        const double va = Sx2 - Sx1 * Sx1 / Nk, vb = Sy2 - Sy1 * Sy1 / Nk;
        double rho = (Sxy - Sx1 * Sy1 / Nk) / (sqrt(va) * sqrt(vb));
        if (m < 2 || Nk < 2.0 || !(va > 0.0) || !(vb > 0.0) || Sxn > 0.5 || Syn > 0.5) rho = nan;
        if (rho > 1.0) rho = 1.0;
        if (rho < -1.0) rho = -1.0;
        out[((int64_t)a * nB + b) * R + r] = rho;
      }
    }
  }
}

static size_t pearson_ws_bytes(int n, int nA, int nB) {
  return arena_bytes((size_t)(nA + nB) * 3 * n, 8) + arena_bytes((size_t)nA * nB * n, 8) + arena_bytes((size_t)(nA + nB) * 3, 8) +
         arena_bytes((size_t)nA * nB, 8) + 256;
}

}  // namespace vr

using namespace vr;

extern "C" size_t vr_bootstrap_pearson_workspace_bytes(int n, int nA, int nB, int) {
  if (n < 2 || nA < 1 || nB < 1) return 256;
  return pearson_ws_bytes(n, nA, nB);
}

extern "C" int vr_bootstrap_pearson(const float* const* A, int nA, const float* const* B, int nB, int n, int64_t ld,
                                    const uint32_t* masks, int R, double* out, void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(A && B && masks && out, "vr_bootstrap_pearson: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && ld >= n, "vr_bootstrap_pearson: bad n=%d ld=%lld", n, (long long)ld);
  VR_REQUIRE(nA >= 1 && nB >= 1 && nA + nB <= kPearMaxMats, "vr_bootstrap_pearson: nA=%d nB=%d (max %d in total)", nA, nB, kPearMaxMats);
  VR_REQUIRE(R >= 1 && R <= 65535, "vr_bootstrap_pearson: R=%d outside [1, 65535]", R);
  if (!ws || ws_bytes < pearson_ws_bytes(n, nA, nB)) {
    set_error("vr_bootstrap_pearson: workspace %zu < %zu", ws_bytes, pearson_ws_bytes(n, nA, nB));
    return VR_ERR_WORKSPACE;
  }
  MatList ml{};
  for (int i = 0; i < nA; ++i) { VR_REQUIRE(A[i], "null matrix"); ml.p[i] = A[i]; }
  for (int i = 0; i < nB; ++i) { VR_REQUIRE(B[i], "null matrix"); ml.p[nA + i] = B[i]; }
  Arena a(ws, ws_bytes);
  PearsonWork w;
  w.rows = a.take<double>((size_t)(nA + nB) * 3 * n);
  w.cross = a.take<double>((size_t)nA * nB * n);
  w.tot = a.take<double>((size_t)(nA + nB) * 3);
  w.totx = a.take<double>((size_t)nA * nB);
  if (!a.ok) { set_error("vr_bootstrap_pearson: internal workspace plan overflow"); return VR_ERR_WORKSPACE; }
  const int W = vr_mask_words(n);
  ProfScope prof(F_PEARSON, stream);
  pearson_rows_kernel<<<dim3(n, nA + nB), kPearThreads, 0, stream>>>(ml, nA, nB, n, ld, w);
  VR_CHECK_LAUNCH();
  pearson_totals_kernel<<<(nA + nB) * 3 + nA * nB, kPearThreads, 0, stream>>>(nA, nB, n, w);
  VR_CHECK_LAUNCH();
  pearson_subsample_kernel<<<dim3(R, nA), kPearThreads, (size_t)W * 4, stream>>>(ml, nA, nB, n, ld, masks, W, R, w, out);
  VR_CHECK_LAUNCH();
  return VR_OK;
}
