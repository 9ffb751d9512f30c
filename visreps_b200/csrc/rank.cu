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
// count_kernel (sorted order, bit-sliced over 32 replicates): for one vector and a batch of 32
//   replicates, lane l streams the packed (i,j) words; membership of the pair in ALL 32 replicates
//   is one AND of two words of the transposed mask (bit r of maskT[i] = stimulus i kept in
//   replicate r).  A 5-stage butterfly transposes the warp's 32x32 bit matrix so lane r holds the
//   kept word of replicate r for the 32-group, keeps that replicate's running count in a register,
//   and appends {kept word, count} to the replicate's table (8 B per 32 pairs).
// rank_kernel (canonical order): one CTA per (vector, replicate).  The replicate's table is
//   copied to shared memory (global when it does not fit); for every pair the precomputed run
//   bounds (one packed word, coalesced) give C(s) and C(e) with two table reads + popc.
//   Reads and the dev write are coalesced; the only random accesses hit the table.
// dot_kernel: per replicate, tiles of 4x4 (model, neural) vectors, exact int64 products.
#include "common.cuh"

namespace vr {

constexpr int kRankThreads = 1024;
constexpr int kRankWarps = kRankThreads / 32;
constexpr int kCountThreads = 512;
constexpr int kCountWarps = kCountThreads / 32;
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

// maskT[batch][i]: bit r = stimulus i is kept in replicate batch*32 + r (0 beyond R).
__global__ void transpose_masks_kernel(const uint32_t* __restrict__ masks, int R, int n, int W,
                                       uint32_t* __restrict__ maskT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int batch = blockIdx.y;
  if (i >= n) return;
  uint32_t w = 0;
  for (int r = 0; r < 32; ++r) {
    const int rep = batch * 32 + r;
    if (rep < R) w |= ((masks[(int64_t)rep * W + (i >> 5)] >> (i & 31)) & 1u) << r;
  }
  maskT[(int64_t)batch * n + i] = w;
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

// Transpose a 32x32 bit matrix held one row per lane (row = lane, column = bit).
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t w, unsigned lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu
                     : s == 2 ? 0x33333333u : 0x55555555u;   // columns c with (c & s) == 0
    const uint32_t x = __shfl_xor_sync(0xffffffffu, w, s);
    w = (lane & s) ? ((w & ~m) | ((x >> s) & m)) : ((w & m) | ((x << s) & ~m));
  }
  return w;
}

// tables: [vector][replicate][Gs] of {kept word, kept count before the group}.
// Dynamic shared memory: maskT of the batch, n words.
__global__ void __launch_bounds__(kCountThreads)
count_kernel(const __grid_constant__ BlobList blobs, int n, const uint32_t* __restrict__ maskT, int R,
             uint2* __restrict__ tables, uint32_t Gs, uint32_t* __restrict__ seg_tot) {
  extern __shared__ __align__(16) uint32_t s_maskT[];
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int v = blockIdx.y, batch = blockIdx.x;
  const uint32_t N = (uint32_t)triu_len(n);
  const uint32_t G = (N + 31u) / 32u;
  const uint32_t SG = (G + 31u) / 32u;
  const uint32_t sg_per_warp = (SG + kCountWarps - 1) / kCountWarps;
  const uint32_t g_lo = min(G, warp * sg_per_warp * 32u);
  const uint32_t g_hi = min(G, g_lo + sg_per_warp * 32u);
  PresortView blob = presort_view(const_cast<void*>(blobs.p[v]), n);
  const uint32_t* __restrict__ stream = blob.stream;
  for (int i = threadIdx.x; i < n; i += kCountThreads) s_maskT[i] = maskT[(int64_t)batch * n + i];
  __syncthreads();

  const int rep = batch * 32 + (int)lane;                 // the replicate this lane accumulates
  uint2* __restrict__ my_table = tables + ((int64_t)v * R + min(rep, R - 1)) * Gs;
  uint32_t running = 0;
  for (uint32_t g0 = g_lo; g0 < g_hi; g0 += 4u) {          // g_lo is a multiple of 32: 4-group steps stay aligned
    uint32_t wd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) wd[q] = (g0 + q < G) ? __ldg(stream + (size_t)(g0 + q) * 32u + lane) : kPadBit;
    uint32_t K[4], C[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t bits = (wd[q] & kPadBit) ? 0u : (s_maskT[pair_i(wd[q])] & s_maskT[pair_j(wd[q])]);
      K[q] = warp_bit_transpose(bits, lane);                // lane r: kept word of replicate r
      C[q] = running;
      running += __popc(K[q]);
    }
    if (rep < R) {
      uint4* dst = reinterpret_cast<uint4*>(my_table + g0);
      dst[0] = make_uint4(K[0], C[0], K[1], C[1]);
      dst[1] = make_uint4(K[2], C[2], K[3], C[3]);
    }
  }
  // counts are relative to the warp's segment; rank_kernel adds the segment bases from seg_tot
  if (rep < R) seg_tot[((int64_t)v * R + rep) * kCountWarps + warp] = running;
}

// Dynamic shared memory: mask words [round_up(W,4)] | table uint2[Gs] (only when kSmemTable).
template <bool kSmemTable>
__global__ void __launch_bounds__(kRankThreads, 1)
rank_kernel(const __grid_constant__ BlobList blobs, int n_vectors, int n, const uint32_t* __restrict__ masks,
            int W, int R, uint2* __restrict__ tables, uint32_t Gs, const uint32_t* __restrict__ seg_tot,
            int32_t* __restrict__ dev, int64_t dev_stride, double* __restrict__ ss_out,
            int64_t* __restrict__ nkept_out) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* s_mask = smem;
  uint2* s_table = reinterpret_cast<uint2*>(smem + ((W + 3) & ~3));
  __shared__ double s_red[kRankWarps];
  __shared__ uint32_t s_base[kCountWarps + 1];

  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t N = (uint32_t)triu_len(n);
  const uint32_t G = (N + 31u) / 32u;
  const uint32_t seg_groups = ((((G + 31u) / 32u) + kCountWarps - 1) / kCountWarps) * 32u;  // count_kernel's segments
  const uint32_t sbits = run_sbits(N);
  const uint32_t smask = (1u << sbits) - 1u;
  const uint32_t esc = (1u << (32u - sbits)) - 1u;

  const int64_t items = (int64_t)n_vectors * R;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int v = (int)(item / R), r = (int)(item % R);
    PresortView blob = presort_view(const_cast<void*>(blobs.p[v]), n);
    uint2* __restrict__ g_table = tables + item * Gs;
    for (int w = threadIdx.x; w < W; w += kRankThreads) s_mask[w] = masks[(int64_t)r * W + w];
    if (threadIdx.x == 0) {                                  // exclusive prefix of the segment totals
      uint32_t run = 0;
      for (int w = 0; w < kCountWarps; ++w) { s_base[w] = run; run += seg_tot[item * kCountWarps + w]; }
      s_base[kCountWarps] = run;
    }
    __syncthreads();
    const uint32_t Ns = s_base[kCountWarps];
    // bring the table in (shared memory) or finish it in place (global), adding the segment bases;
    // entries >= G become the sentinel {0, N_r} so that C(N) needs no special case
    {
      uint2* dst = kSmemTable ? s_table : g_table;
      for (uint32_t g = threadIdx.x; g < Gs; g += kRankThreads) {
        uint2 e = make_uint2(0u, Ns);
        if (g < G) { e = g_table[g]; e.y += s_base[g / seg_groups]; }
        dst[g] = e;
      }
    }
    __syncthreads();
    const uint2* __restrict__ table = kSmemTable ? s_table : g_table;

    int32_t* __restrict__ out = dev + item * dev_stride;
    const uint32_t* __restrict__ runs = blob.runs;
    long long acc = 0;
    double accd = 0.0;
    int since_flush = 0;
    // |dev| <= N, so one step adds < 4 N^2; fold into fp64 before 2^62 can be reached
    const int flush_every = (int)min(1024ull, (1ull << 62) / (4ull * (unsigned long long)N * N + 1ull) + 1ull);
    for (uint32_t p0 = threadIdx.x * 4u; p0 < (uint32_t)dev_stride; p0 += kRankThreads * 4u) {
      int32_t d4[4] = {0, 0, 0, 0};
      if (p0 < N) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(runs + p0));
        const uint32_t rw[4] = {q.x, q.y, q.z, q.w};
        uint32_t i = triu_row(p0, (uint32_t)n);
        uint32_t j = p0 - i * (2u * n - i - 1u) / 2u + i + 1u;
        uint32_t ki = (s_mask[i >> 5] >> (i & 31u)) & 1u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t s = rw[e] & smask;
          const uint32_t f = rw[e] >> sbits;
          uint32_t end = s + f + 1u;
          if (f == esc) {                                   // long run: find its end from the head tables
            const uint32_t g = s >> 5, l = s & 31u;
            const uint32_t above = l == 31u ? 0u : (__ldg(blob.head_bits + g) >> (l + 1u));
            end = above ? s + (uint32_t)__ffs(above) : __ldg(blob.next_head + g + 1);
          }
          const uint2 ts = table[s >> 5];
          const uint2 te = table[end >> 5];
          const uint32_t lo = ts.y + __popc(ts.x & low_mask(s & 31u));
          const uint32_t hi = te.y + __popc(te.x & low_mask(end & 31u));
          const uint32_t kept = ki & (s_mask[j >> 5] >> (j & 31u)) & (p0 + e < N ? 1u : 0u);
          const int32_t d = kept ? (int32_t)(lo + hi) - (int32_t)Ns : 0;
          d4[e] = d;
          acc += (long long)d * (long long)d;
          if (++j == (uint32_t)n) { ++i; j = i + 1u; ki = (s_mask[min(i, (uint32_t)n - 1u) >> 5] >> (i & 31u)) & 1u; }
        }
        if (++since_flush >= flush_every) { since_flush = 0; accd += (double)acc; acc = 0; }
      }
      *reinterpret_cast<int4*>(out + p0) = make_int4(d4[0], d4[1], d4[2], d4[3]);
    }
    accd += (double)acc;
    accd = warp_sum(accd);
    if (lane == 0) s_red[warp] = accd;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < kRankWarps; ++w) tot += s_red[w];
      if (blob.hdr->has_nan) tot = __longlong_as_double(0x7ff8000000000000LL);
      ss_out[item] = tot;
      nkept_out[item] = (int64_t)Ns;
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
  // |dev| <= N (= dev_stride up to padding): one step adds < 4 N^2 per accumulator
  const int flush_every = (int)min(256ull, (1ull << 62) / (4ull * (unsigned long long)dev_stride * dev_stride + 1ull) + 1ull);
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
    if (++since_flush >= flush_every) {
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
static inline bool rank_table_in_smem(int n) { return (size_t)table_groups(n) * 8 <= kRankSmemTableMax; }
static inline size_t rank_smem_bytes(int n, bool smem_table) {
  const size_t mask_bytes = (size_t)round_up(vr_mask_words(n), 4) * 4;
  return mask_bytes + (smem_table ? (size_t)table_groups(n) * 8 : 0);
}
// scratch of one launch over nV vectors x R replicates: transposed masks + count tables
static inline size_t rank_scratch_bytes(int n, int nV, int64_t R) {
  return arena_bytes((size_t)((R + 31) / 32) * n, 4) + arena_bytes((size_t)nV * R * table_groups(n), 8) +
         arena_bytes((size_t)nV * R * kCountWarps, 4);
}

static int launch_rank(const BlobList& blobs, int nV, int n, const uint32_t* masks, int R, int32_t* dev,
                       double* ss, int64_t* nkept, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
  const int W = vr_mask_words(n);
  const uint32_t Gs = table_groups(n);
  const int batches = (R + 31) / 32;
  Arena a(scratch, scratch_bytes);
  uint32_t* maskT = a.take<uint32_t>((size_t)batches * n);
  uint2* tables = a.take<uint2>((size_t)nV * R * Gs);
  uint32_t* seg_tot = a.take<uint32_t>((size_t)nV * R * kCountWarps);
  if (!a.ok) { set_error("rank scratch too small"); return VR_ERR_WORKSPACE; }
  ProfScope prof(F_RANK, stream);
  {
    dim3 grid((n + 255) / 256, batches);
    transpose_masks_kernel<<<grid, 256, 0, stream>>>(masks, R, n, W, maskT);
    VR_CHECK_LAUNCH();
  }
  {
    static bool attr_set = false;
    const size_t smem = (size_t)n * 4;
    if (!attr_set && smem > 40 * 1024) {
      VR_CHECK_CUDA(cudaFuncSetAttribute(count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_set = true;
    }
    dim3 grid(batches, nV);
    count_kernel<<<grid, kCountThreads, smem, stream>>>(blobs, n, maskT, R, tables, Gs, seg_tot);
    VR_CHECK_LAUNCH();
  }
  const bool in_smem = rank_table_in_smem(n);
  const size_t smem = rank_smem_bytes(n, in_smem);
  const int64_t items = (int64_t)nV * R;
  const int grid = (int)(items < kNumSMs ? items : kNumSMs);
  if (in_smem) {
    static bool attr_set = false;
    if (!attr_set) {
      VR_CHECK_CUDA(cudaFuncSetAttribute(rank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kRankSmemTableMax + 8192)));
      attr_set = true;
    }
    rank_kernel<true><<<grid, kRankThreads, smem, stream>>>(blobs, nV, n, masks, W, R, tables, Gs, seg_tot, dev,
                                                           vr_dev_stride(n), ss, nkept);
  } else {
    rank_kernel<false><<<grid, kRankThreads, smem, stream>>>(blobs, nV, n, masks, W, R, tables, Gs, seg_tot, dev,
                                                            vr_dev_stride(n), ss, nkept);
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

extern "C" size_t vr_subsample_rank_workspace_bytes(int n, int R) { return (n >= 2 && R >= 1) ? rank_scratch_bytes(n, 1, R) + 256 : 256; }

extern "C" int vr_subsample_rank(const void* presorted, int n, const uint32_t* masks, int R, int32_t* dev,
                                 double* ss, int64_t* nkept, void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(presorted && masks && dev && ss && nkept, "vr_subsample_rank: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && R >= 1, "vr_subsample_rank: bad sizes n=%d R=%d", n, R);
  if (!ws || ws_bytes < vr_subsample_rank_workspace_bytes(n, R)) {
    set_error("vr_subsample_rank: workspace %zu < %zu", ws_bytes, vr_subsample_rank_workspace_bytes(n, R));
    return VR_ERR_WORKSPACE;
  }
  BlobList bl{};
  bl.p[0] = presorted;
  return launch_rank(bl, 1, n, masks, R, dev, ss, nkept, ws, ws_bytes, stream);
}

namespace {
struct SpearmanPlan {
  size_t fixed;     // scratch
  size_t per_rep;   // dev + ss + nkept for one replicate over all vectors
};
SpearmanPlan spearman_plan(int n, int nV) {
  SpearmanPlan p;
  p.fixed = arena_bytes((size_t)n * 2100, 4) + 8192;   // transposed masks of up to 65535 replicates + slack
  p.per_rep = (size_t)nV * ((size_t)vr_dev_stride(n) * 4 + (size_t)table_groups(n) * 8 + 8 + 8 + 4 * kCountWarps);
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
  const size_t scratch_bytes = rank_scratch_bytes(n, nV, chunk) + 256;
  void* scratch = a.take<char>(scratch_bytes);
  double* ss = a.take<double>((size_t)nV * chunk);
  int64_t* nkept = a.take<int64_t>((size_t)nV * chunk);
  int32_t* dev = a.take<int32_t>((size_t)nV * chunk * vr_dev_stride(n));
  if (!a.ok) { set_error("vr_bootstrap_spearman: internal workspace plan overflow"); return VR_ERR_WORKSPACE; }
  const int W = vr_mask_words(n);
  const int tiles = ((nA + kDotTile - 1) / kDotTile) * ((nB + kDotTile - 1) / kDotTile);
  for (int64_t r0 = 0; r0 < R; r0 += chunk) {
    const int rc = (int)(R - r0 < chunk ? R - r0 : chunk);
    int st = launch_rank(bl, nV, n, masks + (size_t)r0 * W, rc, dev, ss, nkept, scratch, scratch_bytes, stream);
    if (st != VR_OK) return st;
    dim3 grid(tiles, rc);
    ProfScope prof(F_DOT, stream);
    dot_kernel<<<grid, kDotThreads, 0, stream>>>(dev, vr_dev_stride(n), ss, nA, nB, rc, (int)r0, R, out);
    VR_CHECK_LAUNCH();
  }
  return VR_OK;
}
