// Pearson correlation-distance RDM: replaces compute_rdm(), visreps/analysis/rsa.py:59-93.
//
//   row_center_kernel   per-stimulus mean / std (fp64 accumulation) and the centred copy Z
//                       (zero-padded to the SYRK tile grid)                  rsa.py:80-87
//   syrk kernel         upper-triangular tiles of Z Z^T, split over K        rsa.py:89
//   rdm_epilogue_kernel sum of K-splits, /d, /(std_i std_j + c), clamp, diag,
//                       1 - corr, mirrored to the lower triangle              rsa.py:89-92
//
// The SYRK is the tcgen05 / TMEM / TMA kernel of syrk_tc.cuh (3xTF32, fp32-accurate).
#include "common.cuh"
#include "syrk_tc.cuh"
#include "syrk_tc2.cuh"

namespace vr {

struct RdmPlan {
  bool pair;        // 256 x 256 pair tiles on CTA pairs (syrk_tc2.cuh) instead of 128 x 128 tiles on single CTAs
  int n_pad;        // multiple of 128 (tensor tile rows), of 256 with pair tiles
  int64_t d_pad;    // multiple of 32 (one TMA box row)
  int nb128;        // 128-row blocks
  int kb_total;     // 32-column K blocks
  int kb_per_split;
  int tc_splits;
};

inline RdmPlan rdm_plan(int n, int64_t d) {
  RdmPlan p;
  const int pair_mode = option_syrk_pair();
  // pair tiles from four 128-row blocks on, and for K ranges long enough to fill whole waves of CTA pairs with
  // at least ~100 K-blocks each (short rows are launch- and wave-bound: the single-CTA kernel has more, smaller units)
  p.pair = pair_mode != 0 && n > 384 && d >= 16384;
  p.n_pad = (int)round_up(n, p.pair ? 256 : 128);
  p.d_pad = round_up(d, 32);
  // K splits only to fill and balance the machine; each CTA chains its whole K range in-kernel
  p.nb128 = p.n_pad / tc::kBM;
  p.kb_total = (int)(p.d_pad / tc::kBK);
  const int nb2 = p.n_pad / tc2::kPM;
  const int64_t tiles128 = p.pair ? (int64_t)nb2 * (nb2 + 1) : (int64_t)p.nb128 * (p.nb128 + 1) / 2;   // CTAs per K split
  // One wave: tiles x splits <= number of SMs (one CTA is resident per SM).  Longer K chains per CTA mean fewer
  // pipeline fills, fewer partial planes for the epilogue to add, and no partially filled last wave; measured at the
  // NSD shape with the pair-tile kernel: 1 wave 2.40 + 0.19 ms (SYRK + epilogue), 3 waves 2.46 + 0.24, 6 waves 2.67 + 0.29.
  const int waves = option_syrk_waves();
  int64_t want_splits = ((int64_t)waves * kNumSMs) / tiles128;
  int64_t max_splits = (p.kb_total + 7) / 8;               // at least 8 K-blocks (256 columns) per CTA
  if (want_splits > max_splits) want_splits = max_splits;
  if (want_splits < 1) want_splits = 1;
  p.kb_per_split = (int)((p.kb_total + want_splits - 1) / want_splits);
  p.tc_splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  return p;
}

// One CTA per (padded) row.  Rows >= n and columns >= d are written as zeros.
//
// The row is shifted by a value c close to its mean, split into tf32 hi/lo planes, and the statistics of the
// shifted values v = x - c are kept: delta = mean(v) (the part of the mean c missed) and the variance
// mean(v^2) - delta^2.  The epilogue removes d * delta_i * delta_j from the Gram entries, so the result is the
// covariance of exactly centred rows whatever c was; c only has to be close enough that v carries no large
// common offset (which would cost tf32 mantissa bits).
//   kSampledShift = false: c = the row mean (fp64 sum, rounded to fp32): two passes over the row.
//   kSampledShift = true : c = the mean of 2048 samples strided across the row: ONE pass over the row.  Used for
//                          wide rows, where the second pass no longer hits L2 and costs a full HBM read.
constexpr int kShiftSamples = 2048;
constexpr int64_t kSinglePassMinCols = 16384;

//   given_shift != nullptr: c = given_shift[row] (K-sharded rows: the shift comes from the moments of the WHOLE row).
template <bool kSampledShift>
__global__ void __launch_bounds__(256)
row_center_kernel(const float* __restrict__ X, int n, int64_t d, int64_t ldx, float correction,
                  float* __restrict__ Z, float* __restrict__ Zlo, int64_t d_pad, float* __restrict__ stdv,
                  double* __restrict__ delta, int* __restrict__ zero_rows, const float* __restrict__ given_shift = nullptr) {
  __shared__ double s_red[8];
  __shared__ double s_red2[8];
  __shared__ float s_shift;
  const int row = blockIdx.x;
  // Z / Zlo receive the tf32 split of the shifted row: hi = rn_tf32(v), lo = rn_tf32(v - hi)
  // (v - hi is exact in fp32).
  float* z = Z + (int64_t)row * d_pad;
  float* zl = Zlo + (int64_t)row * d_pad;
  if (row >= n) {
    for (int64_t c = threadIdx.x; c < d_pad; c += 256) { z[c] = 0.f; zl[c] = 0.f; }
    if (threadIdx.x == 0) { stdv[row] = 1.f; delta[row] = 0.0; }
    return;
  }
  const float* x = X + (int64_t)row * ldx;
  // 16-byte path when the row start is aligned: float4 loads / stores, 4 values folded in fp32 before each
  // fp64 accumulation.
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  const int64_t d4 = vec ? (d & ~(int64_t)3) : 0;           // elements covered by the vector path
  double acc = 0.0;
  double count = (double)d;
  if (given_shift) {
    // nothing to estimate
  } else if (kSampledShift) {
    const int64_t stride = d / kShiftSamples;                 // >= 8 (d >= kSinglePassMinCols)
    for (int k = threadIdx.x; k < kShiftSamples; k += 256) acc += (double)__ldg(x + (int64_t)k * stride);
    count = (double)kShiftSamples;
  } else {
    for (int64_t c = (int64_t)threadIdx.x * 4; c < d4; c += 1024) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + c));
      acc += (double)((v.x + v.y) + (v.z + v.w));
    }
    for (int64_t c = d4 + threadIdx.x; c < d; c += 256) acc += (double)x[c];
  }
  acc = warp_sum(acc);
  if (lane_id() == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    const float c = given_shift ? given_shift[row] : (float)(t / count);
    s_shift = (c == c && fabsf(c) <= 3.0e38f) ? c : 0.f;
  }
  __syncthreads();
  const float shift = s_shift;
  auto split = [](float v, float& hi, float& lo) {
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    hi = __uint_as_float(hb);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(v - hi));
    lo = __uint_as_float(lb);
  };
  acc = 0.0;
  double acc1 = 0.0;
  auto emit4 = [&](const float4 xv, int64_t c) {
    const float v0 = xv.x - shift, v1 = xv.y - shift, v2 = xv.z - shift, v3 = xv.w - shift;
    acc += (double)((v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3));
    acc1 += (double)((v0 + v1) + (v2 + v3));
    float4 h, l;
    split(v0, h.x, l.x); split(v1, h.y, l.y); split(v2, h.z, l.z); split(v3, h.w, l.w);
    *reinterpret_cast<float4*>(z + c) = h;
    *reinterpret_cast<float4*>(zl + c) = l;
  };
  int64_t c = (int64_t)threadIdx.x * 4;
  for (; c + 3072 < d4; c += 4096) {                     // four independent 16-byte loads in flight per thread
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(x + c));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(x + c + 1024));
    const float4 a2 = __ldg(reinterpret_cast<const float4*>(x + c + 2048));
    const float4 a3 = __ldg(reinterpret_cast<const float4*>(x + c + 3072));
    emit4(a0, c); emit4(a1, c + 1024); emit4(a2, c + 2048); emit4(a3, c + 3072);
  }
  for (; c < d4; c += 1024) emit4(__ldg(reinterpret_cast<const float4*>(x + c)), c);
  for (int64_t c = d4 + threadIdx.x; c < d_pad; c += 256) {
    float v = 0.f;
    if (c < d) { v = x[c] - shift; acc += (double)v * (double)v; acc1 += (double)v; }
    float hi, lo;
    split(v, hi, lo);
    z[c] = hi;
    zl[c] = lo;
  }
  acc = warp_sum(acc);
  acc1 = warp_sum(acc1);
  __syncthreads();
  if (lane_id() == 0) { s_red[threadIdx.x >> 5] = acc; s_red2[threadIdx.x >> 5] = acc1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) { t += s_red[w]; t1 += s_red2[w]; }
    const double dl = t1 / (double)d;
    double var = t / (double)d - dl * dl;
    if (var < 0.0) var = 0.0;
    float sd = (float)sqrt(var + (double)correction);
    if (sd < correction * 10.f) { sd = 1.f; if (zero_rows) atomicAdd(zero_rows, 1); }
    stdv[row] = sd;
    delta[row] = dl;
  }
}

// One thread per (i <= j).  Mirrors to (j, i).
__global__ void __launch_bounds__(256)
rdm_epilogue_kernel(const float* __restrict__ partial, int splits, int n_pad, int n, int64_t d,
                    const float* __restrict__ stdv, const double* __restrict__ delta, float correction,
                    float* __restrict__ rdm, float* __restrict__ triu) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n || j < i) return;
  if (i == j) { if (rdm) rdm[(int64_t)i * n + i] = 0.f; return; }
  double g = 0.0;
  const int64_t plane = (int64_t)n_pad * n_pad;
  for (int s = 0; s < splits; ++s) g += (double)partial[s * plane + (int64_t)i * n_pad + j];
  const float cov = (float)(g / (double)d - delta[i] * delta[j]);     // rows were shifted, not exactly centred
  float corr = cov / (stdv[i] * stdv[j] + correction);
  corr = fminf(fmaxf(corr, -1.f), 1.f);
  const float v = 1.0f - corr;
  if (rdm) {
    rdm[(int64_t)i * n + j] = v;
    rdm[(int64_t)j * n + i] = v;
  }
  if (triu) triu[triu_pos((uint32_t)i, (uint32_t)j, (uint32_t)n)] = v;      // row-major strict upper triangle (rsa.py:111-113)
}

// Strict upper triangle of an (n, n) matrix in the row-major order of torch.triu_indices(n, n, 1).
__global__ void __launch_bounds__(256)
triu_pack_kernel(const float* __restrict__ m, int n, int64_t ld, float* __restrict__ out) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int i = blockIdx.y;
  if (j < n && j > i) out[triu_pos((uint32_t)i, (uint32_t)j, (uint32_t)n)] = m[(int64_t)i * ld + j];
}

// ---- staged form for rows split over feature shards (multi-GPU K split) -----------------------------
// Per row: sum and sum of squares of these d columns, fp64.  One CTA per row.
__global__ void __launch_bounds__(256)
row_moments_kernel(const float* __restrict__ X, int n, int64_t d, int64_t ldx, double* __restrict__ sum, double* __restrict__ sumsq) {
  __shared__ double s_a[8], s_b[8];
  const int row = blockIdx.x;
  const float* x = X + (int64_t)row * ldx;
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  const int64_t d4 = vec ? (d & ~(int64_t)3) : 0;
  double a = 0.0, b = 0.0;
  for (int64_t c = (int64_t)threadIdx.x * 4; c < d4; c += 1024) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + c));
    a += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    b += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
  }
  for (int64_t c = d4 + threadIdx.x; c < d; c += 256) { const double v = x[c]; a += v; b += v * v; }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane_id() == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < 8; ++w) { ta += s_a[w]; tb += s_b[w]; }
    sum[row] = ta;
    sumsq[row] = tb;
  }
}
// fp32 shift per row from the moments of the whole row (what row_center_kernel would have used: the rounded mean).
__global__ void shift_from_moments_kernel(const double* __restrict__ sum, int n, double d_total, float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c = (float)(sum[i] / d_total);
  shift[i] = (c == c && fabsf(c) <= 3.0e38f) ? c : 0.f;
}
// Gram of the shifted shard: fp64 sum of the K-split partial planes, upper triangle (j >= i), zero below it.
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const float* __restrict__ partial, int splits, int n_pad, int n, float* __restrict__ gram, int64_t ldg) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n) return;
  double g = 0.0;
  if (j >= i) {
    const int64_t plane = (int64_t)n_pad * n_pad;
    for (int s = 0; s < splits; ++s) g += (double)partial[s * plane + (int64_t)i * n_pad + j];
  }
  gram[(int64_t)i * ldg + j] = (float)g;
}
// The epilogue of compute_rdm from the reduced Gram and the whole-row moments.
__global__ void __launch_bounds__(256)
rdm_finish_kernel(const float* __restrict__ gram, int64_t ldg, int n, double d_total, const double* __restrict__ sum,
                  const double* __restrict__ sumsq, const float* __restrict__ shift, float correction, float* __restrict__ rdm,
                  int* __restrict__ zero_rows) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n || j < i) return;
  auto stats = [&](int r, double& delta, float& sd, bool& flat) {
    const double mean = sum[r] / d_total;
    delta = mean - (double)shift[r];
    double var = sumsq[r] / d_total - mean * mean;
    if (var < 0.0) var = 0.0;
    sd = (float)sqrt(var + (double)correction);
    flat = sd < correction * 10.f;
    if (flat) sd = 1.f;
  };
  double di, dj;
  float si, sj;
  bool fi, fj;
  stats(i, di, si, fi);
  if (i == j) {
    rdm[(int64_t)i * n + i] = 0.f;
    if (fi && zero_rows) atomicAdd(zero_rows, 1);
    return;
  }
  stats(j, dj, sj, fj);
  const float cov = (float)((double)gram[(int64_t)i * ldg + j] / d_total - di * dj);
  float corr = cov / (si * sj + correction);
  corr = fminf(fmaxf(corr, -1.f), 1.f);
  const float v = 1.0f - corr;
  rdm[(int64_t)i * n + j] = v;
  rdm[(int64_t)j * n + i] = v;
}

// centre (own statistics, or a given shift) + tensor-core SYRK of one (n, d) block -> K-split partial planes
static int center_and_syrk(const float* X, int n, int64_t d, int64_t ldx, float correction, const float* given_shift, const RdmPlan& p,
                           float* Zhi, float* Zlo, float* stdv, double* delta, float* partial, int* zero_var_rows, cudaStream_t stream) {
  {
    ProfScope prof(F_RDM_CENTER, stream);
    if (d >= kSinglePassMinCols || given_shift)
      row_center_kernel<true><<<p.n_pad, 256, 0, stream>>>(X, n, d, ldx, correction, Zhi, Zlo, p.d_pad, stdv, delta, zero_var_rows, given_shift);
    else
      row_center_kernel<false><<<p.n_pad, 256, 0, stream>>>(X, n, d, ldx, correction, Zhi, Zlo, p.d_pad, stdv, delta, zero_var_rows, given_shift);
    VR_CHECK_LAUNCH();
  }
  CUtensorMap map_hi, map_lo;
  int rc = tc::make_plane_map(&map_hi, Zhi, p.n_pad, p.d_pad);
  if (rc != VR_OK) return rc;
  rc = tc::make_plane_map(&map_lo, Zlo, p.n_pad, p.d_pad);
  if (rc != VR_OK) return rc;
  ProfScope prof(p.pair ? F_RDM_SYRK_PAIR : F_RDM_SYRK, stream);
  if (p.pair) {
    static unsigned long long attr_done2 = 0;
    VR_CHECK_CUDA(ensure_dynamic_smem(tc2::syrk_tc2_kernel, (int)tc2::kSmemBytes2, &attr_done2));
    const int nb2 = p.n_pad / tc2::kPM;
    dim3 grid(nb2 * (nb2 + 1), p.tc_splits);             // 2 CTAs per upper-triangular pair tile
    tc2::syrk_tc2_kernel<<<grid, tc2::kThreads2, tc2::kSmemBytes2, stream>>>(map_hi, map_lo, nb2, p.kb_total, p.kb_per_split,
                                                                           partial, p.n_pad);
  } else {
    static unsigned long long attr_done = 0;
    VR_CHECK_CUDA(ensure_dynamic_smem(tc::syrk_tc_kernel, (int)tc::kSmemBytes, &attr_done));
    const int super = tc::kSuper;
    dim3 grid(tc::grid_slots(p.nb128, super), p.tc_splits);
    tc::syrk_tc_kernel<<<grid, tc::kThreads, tc::kSmemBytes, stream>>>(map_hi, map_lo, p.nb128, p.kb_total, p.kb_per_split,
                                                                     partial, p.n_pad, super);
  }
  VR_CHECK_LAUNCH();
  return VR_OK;
}

}  // namespace vr

using namespace vr;

extern "C" int vr_rdm_moments_f32(const float* X, int n, int64_t d, int64_t ldx, double* sum, double* sumsq, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(X && sum && sumsq, "vr_rdm_moments_f32: null pointer");
  VR_REQUIRE(n >= 1 && n <= VR_MAX_STIMULI && d >= 1 && ldx >= d, "vr_rdm_moments_f32: bad shape n=%d d=%lld ldx=%lld", n, (long long)d, (long long)ldx);
  ProfScope prof(F_RDM_CENTER, stream);
  row_moments_kernel<<<n, 256, 0, stream>>>(X, n, d, ldx, sum, sumsq);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

extern "C" int vr_rdm_shift_from_moments(const double* sum, int n, int64_t d_total, float* shift, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(sum && shift && n >= 1 && d_total >= 1, "vr_rdm_shift_from_moments: bad arguments");
  shift_from_moments_kernel<<<(n + 255) / 256, 256, 0, stream>>>(sum, n, (double)d_total, shift);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

extern "C" int vr_rdm_gram_f32(const float* X, int n, int64_t d, int64_t ldx, const float* shift, float* gram, int64_t ldg,
                               void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(X && shift && gram, "vr_rdm_gram_f32: null pointer");
  VR_REQUIRE(n >= 1 && n <= VR_MAX_STIMULI && d >= 1 && ldx >= d && ldg >= n, "vr_rdm_gram_f32: bad shape n=%d d=%lld ldx=%lld ldg=%lld",
             n, (long long)d, (long long)ldx, (long long)ldg);
  if (!ws || ws_bytes < vr_rdm_workspace_bytes(n, d)) {
    set_error("vr_rdm_gram_f32: workspace %zu < %zu", ws_bytes, vr_rdm_workspace_bytes(n, d));
    return VR_ERR_WORKSPACE;
  }
  RdmPlan p = rdm_plan(n, d);
  Arena a(ws, ws_bytes);
  float* Zhi = a.take<float>((size_t)p.n_pad * p.d_pad);
  float* Zlo = a.take<float>((size_t)p.n_pad * p.d_pad);
  float* stdv = a.take<float>(p.n_pad);
  double* delta = a.take<double>(p.n_pad);
  float* partial = a.take<float>((size_t)p.tc_splits * p.n_pad * p.n_pad);
  int rc = center_and_syrk(X, n, d, ldx, 1e-12f, shift, p, Zhi, Zlo, stdv, delta, partial, nullptr, stream);
  if (rc != VR_OK) return rc;
  dim3 eg((n + 255) / 256, n);
  ProfScope prof(F_RDM_EPILOGUE, stream);
  gram_reduce_kernel<<<eg, 256, 0, stream>>>(partial, p.tc_splits, p.n_pad, n, gram, ldg);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

extern "C" int vr_rdm_finish_f32(const float* gram, int64_t ldg, int n, int64_t d_total, const double* sum, const double* sumsq,
                                 const float* shift, float correction, float* rdm, int* zero_var_rows, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(gram && sum && sumsq && shift && rdm, "vr_rdm_finish_f32: null pointer");
  VR_REQUIRE(n >= 1 && n <= VR_MAX_STIMULI && d_total >= 1 && ldg >= n, "vr_rdm_finish_f32: bad shape n=%d d=%lld ldg=%lld", n, (long long)d_total, (long long)ldg);
  if (zero_var_rows) VR_CHECK_CUDA(cudaMemsetAsync(zero_var_rows, 0, sizeof(int), stream));
  dim3 eg((n + 255) / 256, n);
  ProfScope prof(F_RDM_EPILOGUE, stream);
  rdm_finish_kernel<<<eg, 256, 0, stream>>>(gram, ldg, n, (double)d_total, sum, sumsq, shift, correction, rdm, zero_var_rows);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

extern "C" size_t vr_rdm_workspace_bytes(int n, int64_t d) {
  if (n < 1 || d < 1) return 256;
  RdmPlan p = rdm_plan(n, d);
  return 2 * arena_bytes((size_t)p.n_pad * p.d_pad, 4) + arena_bytes(p.n_pad, 4) + arena_bytes(p.n_pad, 8) +
         arena_bytes((size_t)p.tc_splits * p.n_pad * p.n_pad, 4) + 256;
}

extern "C" int vr_triu_pack_f32(const float* m, int n, int64_t ld, float* out, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(m && out, "vr_triu_pack_f32: null pointer");
  VR_REQUIRE(n >= 1 && n <= VR_MAX_STIMULI && ld >= n, "vr_triu_pack_f32: bad shape n=%d ld=%lld", n, (long long)ld);
  if (n < 2) return VR_OK;
  triu_pack_kernel<<<dim3((n + 255) / 256, n), 256, 0, stream>>>(m, n, ld, out);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

extern "C" int vr_rdm_pearson_f32(const float* X, int n, int64_t d, int64_t ldx, float correction, float* rdm,
                                  int* zero_var_rows, void* ws, size_t ws_bytes, vr_stream_t stream_) {
  return vr_rdm_pearson_triu_f32(X, n, d, ldx, correction, rdm, nullptr, zero_var_rows, ws, ws_bytes, stream_);
}

extern "C" int vr_rdm_pearson_triu_f32(const float* X, int n, int64_t d, int64_t ldx, float correction, float* rdm, float* triu,
                                       int* zero_var_rows, void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(X && (rdm || triu), "vr_rdm_pearson_f32: null pointer");
  VR_REQUIRE(n >= 1 && n <= VR_MAX_STIMULI && d >= 1 && ldx >= d, "vr_rdm_pearson_f32: bad shape n=%d d=%lld ldx=%lld",
             n, (long long)d, (long long)ldx);
  if (!ws || ws_bytes < vr_rdm_workspace_bytes(n, d)) {
    set_error("vr_rdm_pearson_f32: workspace %zu < %zu", ws_bytes, vr_rdm_workspace_bytes(n, d));
    return VR_ERR_WORKSPACE;
  }
  RdmPlan p = rdm_plan(n, d);
  Arena a(ws, ws_bytes);
  float* Zhi = a.take<float>((size_t)p.n_pad * p.d_pad);
  float* Zlo = a.take<float>((size_t)p.n_pad * p.d_pad);
  float* stdv = a.take<float>(p.n_pad);
  double* delta = a.take<double>(p.n_pad);
  const int splits = p.tc_splits;
  float* partial = a.take<float>((size_t)p.tc_splits * p.n_pad * p.n_pad);
  if (zero_var_rows) VR_CHECK_CUDA(cudaMemsetAsync(zero_var_rows, 0, sizeof(int), stream));
  {
    int rc = center_and_syrk(X, n, d, ldx, correction, nullptr, p, Zhi, Zlo, stdv, delta, partial, zero_var_rows, stream);
    if (rc != VR_OK) return rc;
  }
  dim3 eg((n + 255) / 256, n);
  ProfScope prof(F_RDM_EPILOGUE, stream);
  rdm_epilogue_kernel<<<eg, 256, 0, stream>>>(partial, splits, p.n_pad, n, d, stdv, delta, correction, rdm, triu);
  VR_CHECK_LAUNCH();
  return VR_OK;
}
