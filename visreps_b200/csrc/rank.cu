// Subsample ranking and the grouped Spearman bootstrap.
//
// Replaces, per bootstrap replicate, `rdm[idx][:, idx]` (visreps/evals.py:365-366),
// the `triu_indices` gather (visreps/analysis/rsa.py:111-113) and
// scipy.stats.spearmanr == rankdata(average) x2 + np.corrcoef (rsa.py:122).
//
// Key fact: the sorted order of a subsample's pairs is the stable compaction of the sorted
// order of ALL pairs.  So a replicate never sorts.  With C(x) = number of kept pairs at
// sorted positions < x, and [s, e) the run of equal values containing a kept pair,
//     2*rank = C(s) + C(e) + 1            (tie-averaged, 1-based, SciPy 'average')
//     dev    = 2*rank - (N_r + 1) = C(s) + C(e) - N_r      (centred, integer)
// and Spearman rho = sum(devA*devB) / sqrt(sum(devA^2) * sum(devB^2)).
//
// rank_kernel: one CTA per (vector, replicate) work item, persistent over items.
//   pass 1  (sorted order) stream the packed (i,j) words, test membership against the
//           replicate's n-bit mask in shared memory, ballot -> one 32-bit kept word per
//           32-group plus the running kept count: the table {kept, count} (8 B per group).
//           The table lives in shared memory when it fits (N <= ~880k pairs), else in an
//           L2-resident per-CTA scratch.  C(x) is then one table read + popc for ANY x.
//   pass 2  (canonical order) for each kept pair read its sorted position, find the run
//           bounds from the head-bit word of its 32-group (bit tricks; prev/next-head tables
//           only when the run leaves the group), two table reads -> dev.  Reads and the dev
//           write are fully coalesced; the only random accesses hit the small tables.
// dot_kernel: per replicate, tiles of 4x4 (model, neural) vectors, exact int64 products.
#include "common.cuh"

namespace vr {

constexpr int kRankThreads = 1024;
constexpr int kRankWarps = kRankThreads / 32;
constexpr int kMaxVectors = 64;
constexpr size_t kRankSmemTableMax = 200 * 1024;   // bytes of shared memory the table may take

struct BlobList {
  const void* p[kMaxVectors];
};

__global__ void build_masks_kernel(const int32_t* __restrict__ idx, int R, int m, int n, int W,
                                   uint32_t* __restrict__ masks) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)R * m) return;
  const int r = (int)(t / m);
  const int i = idx[t];
  if (i >= 0 && i < n) atomicOr(&masks[(int64_t)r * W + (i >> 5)], 1u << (i & 31));
}

__device__ __forceinline__ uint32_t low_mask(uint32_t bit) { return (1u << bit) - 1u; }  // bit in [0,31]

// Row of canonical pair p: largest i with i(2n-i-1)/2 <= p.
__device__ __forceinline__ uint32_t triu_row(uint32_t p, uint32_t n) {
  const float b = (float)(2u * n - 1u);
  float disc = b * b - 8.0f * (float)p;
  int i = (int)((b - sqrtf(fmaxf(disc, 0.f))) * 0.5f);
  i = max(0, min(i, (int)n - 2));
  while (i > 0 && (uint32_t)i * (2u * n - (uint32_t)i - 1u) / 2u > p) --i;
  while ((uint32_t)(i + 1) * (2u * n - (uint32_t)(i + 1) - 1u) / 2u <= p) ++i;
  return (uint32_t)i;
}

// Dynamic shared memory: mask words [Wpad] | table uint2[Gs] (only when kSmemTable).
template <bool kSmemTable>
__global__ void __launch_bounds__(kRankThreads, 1)
rank_kernel(const __grid_constant__ BlobList blobs, int n_vectors, int n, const uint32_t* __restrict__ masks,
            int W, int R, int32_t* __restrict__ dev, int64_t dev_stride, double* __restrict__ ss_out,
            int64_t* __restrict__ nkept_out, uint2* __restrict__ scratch, uint32_t Gs) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* s_mask = smem;                                   // [round_up(W, 4)]
  __shared__ uint32_t s_segbase[kRankWarps + 1];
  __shared__ double s_red[kRankWarps];

  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t N = (uint32_t)triu_len(n);
  const uint32_t G = (N + 31u) / 32u;
  const uint32_t SG = (G + 31u) / 32u;                          // super-groups of 32 groups
  const uint32_t sg_per_warp = (SG + kRankWarps - 1) / kRankWarps;
  const uint32_t sg_lo = warp * sg_per_warp;
  const uint32_t sg_hi = min(SG, sg_lo + sg_per_warp);
  uint2* table = kSmemTable ? reinterpret_cast<uint2*>(smem + ((W + 3) & ~3))
                            : scratch + (size_t)blockIdx.x * Gs;

  const int64_t items = (int64_t)n_vectors * R;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int v = (int)(item / R), r = (int)(item % R);
    PresortView blob = presort_view(const_cast<void*>(blobs.p[v]), n);
    const uint32_t* __restrict__ stream = blob.stream;
    for (int w = threadIdx.x; w < W; w += kRankThreads) s_mask[w] = masks[(int64_t)r * W + w];
    __syncthreads();

    // ---------------- pass 1: kept words + running counts (sorted order) ----------------
    uint32_t running = 0;
    for (uint32_t sg = sg_lo; sg < sg_hi; ++sg) {
      uint32_t myK = 0;
      const uint32_t gbase = sg * 32u;
#pragma unroll 8
      for (uint32_t t = 0; t < 32u; ++t) {
        const uint32_t g = gbase + t;
        bool kept = false;
        if (g < G) {
          const uint32_t wd = __ldg(stream + (size_t)g * 32u + lane);
          const uint32_t i = pair_i(wd), j = pair_j(wd);
          kept = !(wd & kPadBit) && ((s_mask[i >> 5] >> (i & 31u)) & (s_mask[j >> 5] >> (j & 31u)) & 1u);
        }
        const uint32_t K = __ballot_sync(0xffffffffu, kept);
        if (lane == t) myK = K;
      }
      const uint32_t c = __popc(myK);
      uint32_t incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += x;
      }
      table[gbase + lane] = make_uint2(myK, running + incl - c);
      running += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_segbase[warp] = running;
    __syncthreads();
    if (warp == 0) {
      const uint32_t mine = s_segbase[lane];
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += x;
      }
      s_segbase[lane] = incl - mine;
      if (lane == 31) s_segbase[kRankWarps] = incl;
    }
    __syncthreads();
    const uint32_t Ns = s_segbase[kRankWarps];
    {  // make the counts global (add each warp's segment base)
      const uint32_t mybase = s_segbase[warp];
      if (mybase)
        for (uint32_t g = sg_lo * 32u + lane; g < sg_hi * 32u; g += 32u) table[g].y += mybase;
    }
    __syncthreads();

    // kept pairs at sorted positions < x, x in [0, N]
    auto count_below = [&](uint32_t x) -> uint32_t {
      if (x >= N) return Ns;
      const uint2 e = table[x >> 5];
      return e.y + __popc(e.x & low_mask(x & 31u));
    };

    // ---------------- pass 2: canonical order, 4 pairs per thread per step ----------------
    int32_t* __restrict__ out = dev + ((int64_t)v * R + r) * dev_stride;
    const uint32_t* __restrict__ posv = blob.pos;
    const uint32_t* __restrict__ headw = blob.head_bits;
    double acc = 0.0;
    for (uint32_t p0 = threadIdx.x * 4u; p0 < (uint32_t)dev_stride; p0 += kRankThreads * 4u) {
      int32_t d4[4] = {0, 0, 0, 0};
      if (p0 < N) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(posv + p0));
        const uint32_t xs[4] = {q.x, q.y, q.z, q.w};
        uint32_t i = triu_row(p0, (uint32_t)n);
        uint32_t j = p0 - i * (2u * n - i - 1u) / 2u + i + 1u;
        bool ki = (s_mask[i >> 5] >> (i & 31u)) & 1u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (p0 + e < N) {
            if (ki && ((s_mask[j >> 5] >> (j & 31u)) & 1u)) {
              const uint32_t x = xs[e];
              const uint32_t g = x >> 5, l = x & 31u;
              const uint2 te = table[g];
              const uint32_t H = __ldg(headw + g);
              uint32_t lo, hi;
              const uint32_t below = H & (low_mask(l) | (1u << l));
              if (below) lo = te.y + __popc(te.x & low_mask(31u - __clz(below)));
              else lo = count_below(__ldg(blob.prev_head + g - 1));
              const uint32_t above = l == 31u ? 0u : (H >> (l + 1u));
              if (above) hi = te.y + __popc(te.x & low_mask(l + (uint32_t)__ffs(above)));
              else hi = count_below(__ldg(blob.next_head + g + 1));
              const int32_t d = (int32_t)(lo + hi) - (int32_t)Ns;
              d4[e] = d;
              acc += (double)d * (double)d;
            }
            if (++j == (uint32_t)n) { ++i; j = i + 1u; ki = (s_mask[i >> 5] >> (i & 31u)) & 1u; }
          }
        }
      }
      *reinterpret_cast<int4*>(out + p0) = make_int4(d4[0], d4[1], d4[2], d4[3]);
    }

    acc = warp_sum(acc);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < kRankWarps; ++w) tot += s_red[w];
      if (blob.hdr->has_nan) tot = __longlong_as_double(0x7ff8000000000000LL);
      ss_out[(int64_t)v * R + r] = tot;
      nkept_out[(int64_t)v * R + r] = (int64_t)Ns;
    }
    __syncthreads();  // table + s_mask are reused by the next item
  }
}

// out[a][b][r] = sum_p devA[a][r][p] * devB[b][r][p] / sqrt(ssA[a][r] * ssB[b][r])
constexpr int kDotThreads = 256;
constexpr int kDotTile = 4;
__global__ void __launch_bounds__(kDotThreads)
dot_kernel(const int32_t* __restrict__ dev, int64_t dev_stride, const double* __restrict__ ss, int nA, int nB,
           int R, int r0, int Rtot, double* __restrict__ out) {
  __shared__ long long s_part[kDotThreads / 32][kDotTile * kDotTile];
  __shared__ double s_partd[kDotThreads / 32][kDotTile * kDotTile];
  const int tilesB = (nB + kDotTile - 1) / kDotTile;
  const int ta = blockIdx.x / tilesB, tb = blockIdx.x % tilesB;
  const int r = blockIdx.y;
  const int a0 = ta * kDotTile, b0 = tb * kDotTile;
  const int4* pa[kDotTile];
  const int4* pb[kDotTile];
#pragma unroll
  for (int k = 0; k < kDotTile; ++k) {
    const int a = min(a0 + k, nA - 1), b = min(b0 + k, nB - 1);
    pa[k] = reinterpret_cast<const int4*>(dev + ((int64_t)a * R + r) * dev_stride);
    pb[k] = reinterpret_cast<const int4*>(dev + ((int64_t)(nA + b) * R + r) * dev_stride);
  }
  const int64_t n4 = dev_stride / 4;
  long long acc[kDotTile][kDotTile];
  double accd[kDotTile][kDotTile];
#pragma unroll
  for (int x = 0; x < kDotTile; ++x)
#pragma unroll
    for (int y = 0; y < kDotTile; ++y) { acc[x][y] = 0; accd[x][y] = 0.0; }

  int since_flush = 0;
  for (int64_t q = threadIdx.x; q < n4; q += kDotThreads) {
    int4 va[kDotTile], vb[kDotTile];
#pragma unroll
    for (int k = 0; k < kDotTile; ++k) { va[k] = __ldg(pa[k] + q); vb[k] = __ldg(pb[k] + q); }
#pragma unroll
    for (int x = 0; x < kDotTile; ++x)
#pragma unroll
      for (int y = 0; y < kDotTile; ++y) {
        acc[x][y] += (long long)va[x].x * vb[y].x + (long long)va[x].y * vb[y].y +
                     (long long)va[x].z * vb[y].z + (long long)va[x].w * vb[y].w;
      }
    // |dev| < 2^26 -> each step adds < 2^54; flush to fp64 before 2^63 can be reached
    if (++since_flush == 256) {
      since_flush = 0;
#pragma unroll
      for (int x = 0; x < kDotTile; ++x)
#pragma unroll
        for (int y = 0; y < kDotTile; ++y) { accd[x][y] += (double)acc[x][y]; acc[x][y] = 0; }
    }
  }
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  // Up to 2^20 pairs the whole sum stays below 2^60 and is carried exactly in int64; beyond
  // that the per-thread integers are folded into fp64 before any cross-thread addition.
  const bool exact = dev_stride <= (1ll << 20);
#pragma unroll
  for (int x = 0; x < kDotTile; ++x)
#pragma unroll
    for (int y = 0; y < kDotTile; ++y) {
      if (!exact) { accd[x][y] += (double)acc[x][y]; acc[x][y] = 0; }
      long long s = warp_sum_ll(acc[x][y]);
      double sd = warp_sum(accd[x][y]);
      if (lane == 0) { s_part[warp][x * kDotTile + y] = s; s_partd[warp][x * kDotTile + y] = sd; }
    }
  __syncthreads();
  if (threadIdx.x < kDotTile * kDotTile) {
    const int x = threadIdx.x / kDotTile, y = threadIdx.x % kDotTile;
    double tot = 0.0;
    long long ti = 0;
    for (int w = 0; w < kDotThreads / 32; ++w) { ti += s_part[w][threadIdx.x]; tot += s_partd[w][threadIdx.x]; }
    tot += (double)ti;
    const int a = a0 + x, b = b0 + y;
    if (a < nA && b < nB) {
      const double sa = ss[(int64_t)a * R + r], sb = ss[(int64_t)(nA + b) * R + r];
      double rho = tot / (sqrt(sa) * sqrt(sb));
      if (!(sa > 0.0) || !(sb > 0.0)) rho = __longlong_as_double(0x7ff8000000000000LL);
      out[((int64_t)a * nB + b) * Rtot + r0 + r] = rho;
    }
  }
}

static inline uint32_t table_groups(int n) { return (uint32_t)round_up(presort_groups(n) + 1, 32) + 32u; }
static inline size_t rank_smem_bytes(int n, bool smem_table) {
  const size_t mask_bytes = (size_t)round_up(vr_mask_words(n), 4) * 4;
  return mask_bytes + (smem_table ? (size_t)table_groups(n) * 8 : 0);
}
static inline bool rank_table_in_smem(int n) { return (size_t)table_groups(n) * 8 <= kRankSmemTableMax; }
static inline size_t rank_scratch_bytes(int n) {
  return rank_table_in_smem(n) ? 256 : arena_bytes((size_t)kNumSMs * table_groups(n), 8);
}

static int launch_rank(const BlobList& blobs, int nV, int n, const uint32_t* masks, int R, int32_t* dev,
                       double* ss, int64_t* nkept, void* scratch, cudaStream_t stream) {
  const int W = vr_mask_words(n);
  const bool in_smem = rank_table_in_smem(n);
  const size_t smem = rank_smem_bytes(n, in_smem);
  const int64_t items = (int64_t)nV * R;
  const int grid = (int)(items < kNumSMs ? items : kNumSMs);
  ProfScope prof(F_RANK, stream);
  if (in_smem) {
    static bool attr_set = false;
    if (!attr_set) {
      VR_CHECK_CUDA(cudaFuncSetAttribute(rank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kRankSmemTableMax + 8192)));
      attr_set = true;
    }
    rank_kernel<true><<<grid, kRankThreads, smem, stream>>>(blobs, nV, n, masks, W, R, dev, vr_dev_stride(n), ss,
                                                           nkept, nullptr, table_groups(n));
  } else {
    rank_kernel<false><<<grid, kRankThreads, smem, stream>>>(blobs, nV, n, masks, W, R, dev, vr_dev_stride(n), ss,
                                                            nkept, static_cast<uint2*>(scratch), table_groups(n));
  }
  VR_CHECK_LAUNCH();
  return VR_OK;
}

}  // namespace vr

using namespace vr;

extern "C" int vr_mask_words(int n) { return (n + 31) / 32; }

extern "C" int vr_build_masks(const int32_t* idx, int R, int m, int n, uint32_t* masks, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(masks && (idx || (int64_t)R * m == 0), "vr_build_masks: null pointer");
  VR_REQUIRE(R >= 0 && m >= 0 && n >= 1 && n <= VR_MAX_STIMULI && m <= n, "vr_build_masks: bad sizes R=%d m=%d n=%d", R, m, n);
  if (R == 0) return VR_OK;
  const int W = vr_mask_words(n);
  VR_CHECK_CUDA(cudaMemsetAsync(masks, 0, (size_t)R * W * 4, stream));
  const int64_t tot = (int64_t)R * m;
  if (tot > 0) {
    ProfScope prof(F_MASKS, stream);
    build_masks_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(idx, R, m, n, W, masks);
    VR_CHECK_LAUNCH();
  }
  return VR_OK;
}

extern "C" int64_t vr_dev_stride(int n) { return n >= 2 ? round_up(triu_len(n), 32) : 32; }

extern "C" size_t vr_subsample_rank_workspace_bytes(int n) { return n >= 2 ? rank_scratch_bytes(n) + 256 : 256; }

extern "C" int vr_subsample_rank(const void* presorted, int n, const uint32_t* masks, int R, int32_t* dev,
                                 double* ss, int64_t* nkept, void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(presorted && masks && dev && ss && nkept, "vr_subsample_rank: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && R >= 1, "vr_subsample_rank: bad sizes n=%d R=%d", n, R);
  if (!ws || ws_bytes < vr_subsample_rank_workspace_bytes(n)) {
    set_error("vr_subsample_rank: workspace %zu < %zu", ws_bytes, vr_subsample_rank_workspace_bytes(n));
    return VR_ERR_WORKSPACE;
  }
  BlobList bl{};
  bl.p[0] = presorted;
  return launch_rank(bl, 1, n, masks, R, dev, ss, nkept, ws, stream);
}

namespace {
struct SpearmanPlan {
  size_t fixed;     // scratch
  size_t per_rep;   // dev + ss + nkept for one replicate over all vectors
};
SpearmanPlan spearman_plan(int n, int nV) {
  SpearmanPlan p;
  p.fixed = rank_scratch_bytes(n) + 4096;
  p.per_rep = (size_t)nV * ((size_t)vr_dev_stride(n) * 4 + 8 + 8);
  return p;
}
constexpr size_t kSoftCap = (size_t)40 << 30;  // default upper bound on one call's workspace
}  // namespace

extern "C" size_t vr_bootstrap_spearman_workspace_bytes(int n, int nA, int nB, int R) {
  if (n < 2 || nA < 1 || nB < 1 || R < 1) return 4096;
  SpearmanPlan p = spearman_plan(n, nA + nB);
  size_t reps = (size_t)R;
  size_t cap = kSoftCap / p.per_rep;
  if (cap < 1) cap = 1;
  if (reps > cap) reps = cap;
  return p.fixed + reps * p.per_rep + 3 * 256;
}

extern "C" int vr_bootstrap_spearman(const void* const* A, int nA, const void* const* B, int nB, int n,
                                     const uint32_t* masks, int R, double* out, void* ws, size_t ws_bytes,
                                     vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(A && B && masks && out, "vr_bootstrap_spearman: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI, "vr_bootstrap_spearman: n=%d outside [2, %d]", n, VR_MAX_STIMULI);
  VR_REQUIRE(nA >= 1 && nB >= 1 && nA + nB <= kMaxVectors, "vr_bootstrap_spearman: nA=%d nB=%d (max %d in total)", nA, nB, kMaxVectors);
  VR_REQUIRE(R >= 1, "vr_bootstrap_spearman: R=%d", R);
  const int nV = nA + nB;
  SpearmanPlan plan = spearman_plan(n, nV);
  if (!ws || ws_bytes < plan.fixed + plan.per_rep + 3 * 256) {
    set_error("vr_bootstrap_spearman: workspace %zu < minimum %zu", ws_bytes, plan.fixed + plan.per_rep + 3 * 256);
    return VR_ERR_WORKSPACE;
  }
  BlobList bl{};
  for (int i = 0; i < nA; ++i) { VR_REQUIRE(A[i], "null blob"); bl.p[i] = A[i]; }
  for (int i = 0; i < nB; ++i) { VR_REQUIRE(B[i], "null blob"); bl.p[nA + i] = B[i]; }
  int64_t chunk = (int64_t)((ws_bytes - plan.fixed - 3 * 256) / plan.per_rep);
  if (chunk > R) chunk = R;
  if (chunk > 65535) chunk = 65535;  // grid.y limit of the dot kernel
  Arena a(ws, ws_bytes);
  void* scratch = a.take<char>(rank_scratch_bytes(n));
  double* ss = a.take<double>((size_t)nV * chunk);
  int64_t* nkept = a.take<int64_t>((size_t)nV * chunk);
  int32_t* dev = a.take<int32_t>((size_t)nV * chunk * vr_dev_stride(n));
  if (!a.ok) { set_error("vr_bootstrap_spearman: internal workspace plan overflow"); return VR_ERR_WORKSPACE; }
  const int W = vr_mask_words(n);
  const int tiles = ((nA + kDotTile - 1) / kDotTile) * ((nB + kDotTile - 1) / kDotTile);
  for (int64_t r0 = 0; r0 < R; r0 += chunk) {
    const int rc = (int)(R - r0 < chunk ? R - r0 : chunk);
    int st = launch_rank(bl, nV, n, masks + (size_t)r0 * W, rc, dev, ss, nkept, scratch, stream);
    if (st != VR_OK) return st;
    dim3 grid(tiles, rc);
    ProfScope prof(F_DOT, stream);
    dot_kernel<<<grid, kDotThreads, 0, stream>>>(dev, vr_dev_stride(n), ss, nA, nB, rc, (int)r0, R, out);
    VR_CHECK_LAUNCH();
  }
  return VR_OK;
}
