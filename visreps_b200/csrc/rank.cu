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
//   The side of the comparison with fewer vectors is materialised (dev vectors in HBM); the items of
//   the other side multiply their dev values with those vectors on the fly (fused) and emit rho.
#include <cstdlib>
#include <type_traits>
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

// pairbits[r][w]: bit b = canonical pair 32w+b has both stimuli in mask r (0 beyond N).
// One thread per word; the row is decoded once and then walked.
__global__ void __launch_bounds__(256)
pair_mask_kernel(const uint32_t* __restrict__ masks, int R, int n, int W, uint32_t words,
                 uint32_t* __restrict__ pairbits);

__device__ __forceinline__ uint32_t low_mask(uint32_t bit) { return (1u << bit) - 1u; }  // bit in [0,31]
// x << k with PTX semantics: k >= 32 gives 0 (C++ leaves that undefined).
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t k) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(k));
  return r;
}

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

__global__ void __launch_bounds__(256)
pair_mask_kernel(const uint32_t* __restrict__ masks, int R, int n, int W, uint32_t words,
                 uint32_t* __restrict__ pairbits) {
  extern __shared__ uint32_t s_m[];
  const int r = blockIdx.y;
  for (int w = threadIdx.x; w < W; w += 256) s_m[w] = masks[(int64_t)r * W + w];
  __syncthreads();
  const uint32_t N = (uint32_t)triu_len(n);
  for (uint32_t w = blockIdx.x * 256u + threadIdx.x; w < words; w += gridDim.x * 256u) {
    uint32_t p = w * 32u, bits = 0;
    if (p < N) {
      uint32_t i = triu_row(p, (uint32_t)n);
      uint32_t j = p - i * (2u * n - i - 1u) / 2u + i + 1u;
      uint32_t ki = (s_m[i >> 5] >> (i & 31u)) & 1u;
      for (uint32_t b = 0; b < 32u && p + b < N; ++b) {
        bits |= (ki & (s_m[j >> 5] >> (j & 31u)) & 1u) << b;
        if (++j == (uint32_t)n) { ++i; j = i + 1u; ki = (s_m[min(i, (uint32_t)n - 1u) >> 5] >> (i & 31u)) & 1u; }
      }
    }
    pairbits[(int64_t)r * words + w] = bits;
  }
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

// ---------------------------------------------------------------------------------------------
// rank_kernel: one CTA per (vector, replicate) item, persistent over items.
//   kFused == false ("materialise"): writes dev (int32, canonical order, 0 for dropped pairs) and ss.
//   kFused == true : the item's dev values never leave registers; they are multiplied on the fly with
//                    the <= kFuseMax materialised vectors of the same replicate (coalesced int4 reads)
//                    and the finished Spearman coefficients are written.  Items are ordered
//                    replicate-major so the CTAs sharing one replicate's materialised vectors run
//                    together and hit L2.
//   Acc = long long (exact) while N^3 fits 2^62, else double.
// Dynamic shared memory: mask words [round_up(W,4)] | table uint2[Gs] (only when kSmemTable).
// ---------------------------------------------------------------------------------------------
constexpr int kFuseMax = 8;

struct RankArgs {
  BlobList blobs;            // the vectors ranked by this launch
  int n_vectors;
  int n;
  const uint32_t* pairbits;  // [R][dev_stride/32]: kept bits of the canonical pairs
  int R;
  uint2* tables;             // [table_index(v)][R][Gs]
  int table_base;            // index of vector 0 of this launch in the table / seg_tot arrays
  uint32_t Gs;
  const uint32_t* seg_tot;   // [table index][R][kCountWarps]
  int64_t dev_stride;
  // materialise
  int32_t* dev;              // [n_vectors][R][dev_stride]
  double* ss;                // [n_vectors][R]
  int64_t* nkept;            // [n_vectors][R]  (nullable)
  // fused
  const int32_t* mat_dev;    // [n_mat][R][dev_stride]
  const double* mat_ss;      // [n_mat][R]
  int n_mat;
  double* out;               // (nA, nB, Rtot)
  int out_nB;
  int64_t out_Rtot;
  int out_r0;
  int mat_is_A;              // 1: materialised vectors are the model side (rows of out)
  int mat_off;               // index of materialised vector 0 on its side
  int fused_off;             // index of fused vector 0 on its side
};

template <bool kSmemTable, bool kFused, typename Acc>
__global__ void __launch_bounds__(kRankThreads, 1)
rank_kernel(const __grid_constant__ RankArgs A) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint2* s_table = reinterpret_cast<uint2*>(smem);
  __shared__ double s_red[kRankWarps][kFuseMax + 1];
  __shared__ uint32_t s_base[kCountWarps + 1];

  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int n = A.n, R = A.R;
  const uint32_t N = (uint32_t)triu_len(n);
  const uint32_t G = (N + 31u) / 32u;
  const uint32_t Gs = A.Gs;
  const uint32_t seg_groups = ((((G + 31u) / 32u) + kCountWarps - 1) / kCountWarps) * 32u;  // count_kernel's segments
  const uint32_t sbits = run_sbits(N);
  const uint32_t smask = (1u << sbits) - 1u;
  const uint32_t esc = (1u << (32u - sbits)) - 1u;
  const int64_t dev_stride = A.dev_stride;

  const int64_t items = (int64_t)A.n_vectors * R;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    // materialise: vector-major (CTAs share the vector's `runs` stream); fused: replicate-major
    const int v = kFused ? (int)(item % A.n_vectors) : (int)(item / R);
    const int r = kFused ? (int)(item / A.n_vectors) : (int)(item % R);
    PresortView blob = presort_view(const_cast<void*>(A.blobs.p[v]), n);
    const int64_t tix = (int64_t)(A.table_base + v) * R + r;
    uint2* __restrict__ g_table = A.tables + tix * Gs;
    if (threadIdx.x == 0) {                                  // exclusive prefix of the segment totals
      uint32_t run = 0;
      for (int w = 0; w < kCountWarps; ++w) { s_base[w] = run; run += A.seg_tot[tix * kCountWarps + w]; }
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

    int32_t* __restrict__ out_dev = kFused ? nullptr : A.dev + ((int64_t)v * R + r) * dev_stride;
    const int32_t* __restrict__ mat = kFused ? A.mat_dev + (int64_t)r * dev_stride : nullptr;
    const int64_t mat_step = (int64_t)R * dev_stride;        // between materialised vectors
    const uint32_t* __restrict__ runs = blob.runs;
    const uint32_t* __restrict__ keptw = A.pairbits + (int64_t)r * (dev_stride >> 5);
    Acc acc_ss = 0;
    Acc acc[kFuseMax];
#pragma unroll
    for (int a = 0; a < kFuseMax; ++a) acc[a] = 0;

    // software pipeline: the run words and kept bits of the NEXT step are in flight while this one is ranked
    const uint32_t esc_word = esc << sbits;                  // a run word >= esc_word carries the escape code
    uint32_t p0 = threadIdx.x * 4u;
    uint4 q_cur = make_uint4(0u, 0u, 0u, 0u);
    uint32_t kw_cur = 0u;
    if (p0 < N) { q_cur = __ldg(reinterpret_cast<const uint4*>(runs + p0)); kw_cur = __ldg(keptw + (p0 >> 5)); }
    for (; p0 < (uint32_t)dev_stride; p0 += kRankThreads * 4u) {
      int32_t d4[4] = {0, 0, 0, 0};
      int4 m4[4];
      if (kFused && p0 < N) {   // first batch of materialised vectors: issued before the ranking work hides their latency
#pragma unroll
        for (int a = 0; a < 4; ++a) m4[a] = __ldg(reinterpret_cast<const int4*>(mat + min(a, A.n_mat - 1) * mat_step + p0));
      }
      const uint4 q = q_cur;
      const uint32_t kw = kw_cur >> (p0 & 31u);              // 4 kept bits (p0 is a multiple of 4)
      {
        const uint32_t pn = p0 + kRankThreads * 4u;
        if (pn < N) { q_cur = __ldg(reinterpret_cast<const uint4*>(runs + pn)); kw_cur = __ldg(keptw + (pn >> 5)); }
      }
      if (p0 < N) {
        const uint32_t rw[4] = {q.x, q.y, q.z, q.w};
        uint32_t s4[4], e4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { s4[e] = rw[e] & smask; e4[e] = s4[e] + (rw[e] >> sbits) + 1u; }
        if (max(max(rw[0], rw[1]), max(rw[2], rw[3])) >= esc_word) {   // rare: a run longer than the packed field
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (rw[e] >= esc_word) {
              const uint32_t g = s4[e] >> 5, l = s4[e] & 31u;
              const uint32_t above = l == 31u ? 0u : (__ldg(blob.head_bits + g) >> (l + 1u));
              e4[e] = above ? s4[e] + (uint32_t)__ffs(above) : __ldg(blob.next_head + g + 1);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t gs = s4[e] >> 5, ge = e4[e] >> 5;
          const uint2 ts = table[gs];
          uint2 te = ts;
          if (ge != gs) te = table[ge];                      // ~5 % of pairs: the run leaves the 32-group
          // popc of the low (x & 31) bits: shifting left by 32 - (x & 31) drops the others (PTX shl clamps at 32)
          const uint32_t lo = ts.y + __popc(shl_clamp(ts.x, 32u - (s4[e] & 31u)));
          const uint32_t hi = te.y + __popc(shl_clamp(te.x, 32u - (e4[e] & 31u)));
          const int32_t d = ((kw >> e) & 1u) ? (int32_t)(lo + hi) - (int32_t)Ns : 0;
          d4[e] = d;
          acc_ss += (Acc)d * (Acc)d;
        }
      }
      if (kFused) {
        if (p0 < N) {
#pragma unroll
          for (int a = 0; a < 4; ++a)
            if (a < A.n_mat)
              acc[a] += (Acc)d4[0] * (Acc)m4[a].x + (Acc)d4[1] * (Acc)m4[a].y + (Acc)d4[2] * (Acc)m4[a].z + (Acc)d4[3] * (Acc)m4[a].w;
          if (A.n_mat > 4) {
#pragma unroll
            for (int a = 0; a < 4; ++a) m4[a] = __ldg(reinterpret_cast<const int4*>(mat + min(a + 4, A.n_mat - 1) * mat_step + p0));
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (a + 4 < A.n_mat)
                acc[a + 4] += (Acc)d4[0] * (Acc)m4[a].x + (Acc)d4[1] * (Acc)m4[a].y + (Acc)d4[2] * (Acc)m4[a].z + (Acc)d4[3] * (Acc)m4[a].w;
          }
        }
      } else {
        *reinterpret_cast<int4*>(out_dev + p0) = make_int4(d4[0], d4[1], d4[2], d4[3]);
      }
    }

    // block reduction: ss (slot kFuseMax) and, when fused, the n_mat dot products
    {
      double v_ss = warp_sum((double)acc_ss);
      if (lane == 0) s_red[warp][kFuseMax] = v_ss;
      if (kFused) {
#pragma unroll
        for (int a = 0; a < kFuseMax; ++a) {
          if (a < A.n_mat) {
            double t;
            if constexpr (std::is_integral<Acc>::value) t = (double)warp_sum_ll((long long)acc[a]);   // exact
            else t = warp_sum((double)acc[a]);
            if (lane == 0) s_red[warp][a] = t;
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x <= kFuseMax) {
      const int slot = threadIdx.x;
      double tot = 0.0;
      if (slot == kFuseMax || (kFused && slot < A.n_mat))
        for (int w = 0; w < kRankWarps; ++w) tot += s_red[w][slot];
      const double nan = __longlong_as_double(0x7ff8000000000000LL);
      // every participating lane needs ss of this item: recompute it locally from slot kFuseMax
      double ss_item = 0.0;
      for (int w = 0; w < kRankWarps; ++w) ss_item += s_red[w][kFuseMax];
      if (blob.hdr->has_nan) ss_item = nan;
      if (!kFused) {
        if (slot == kFuseMax) {
          A.ss[(int64_t)v * R + r] = ss_item;
          if (A.nkept) A.nkept[(int64_t)v * R + r] = (int64_t)Ns;
        }
      } else if (slot < A.n_mat) {
        const double ss_m = A.mat_ss[(int64_t)slot * R + r];
        double rho = tot / (sqrt(ss_m) * sqrt(ss_item));
        if (!(ss_m > 0.0) || !(ss_item > 0.0)) rho = nan;
        const int ia = A.mat_is_A ? A.mat_off + slot : A.fused_off + v;
        const int ib = A.mat_is_A ? A.fused_off + v : A.mat_off + slot;
        A.out[((int64_t)ia * A.out_nB + ib) * A.out_Rtot + A.out_r0 + r] = rho;
      }
    }
    __syncthreads();  // table, s_mask and s_red are reused by the next item
  }
}

// out[a][b][r] = sum_p devA[a][r][p] * devB[b][r][p] / sqrt(ssA[a][r] * ssB[b][r])
constexpr int kDotThreads = 256;
constexpr int kDotTile = 4;
__global__ void __launch_bounds__(kDotThreads)
dot_kernel(const int32_t* __restrict__ dev, int64_t dev_stride, const double* __restrict__ ss, int nA, int nB,
           int R, int r0, int Rtot, double* __restrict__ out, int first_is_model, int out_nB) {
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
      // the first nA vectors of `dev` are the model side when first_is_model, else the neural side
      const int ia = first_is_model ? a : b, ib = first_is_model ? b : a;
      out[((int64_t)ia * out_nB + ib) * Rtot + r0 + r] = rho;
    }
  }
}


static inline uint32_t table_groups(int n) { return (uint32_t)round_up(presort_groups(n) + 1, 32) + 32u; }
static inline bool rank_table_in_smem(int n) { return (size_t)table_groups(n) * 8 <= kRankSmemTableMax; }
static inline size_t rank_smem_bytes(int n, bool smem_table) {
  return smem_table ? (size_t)table_groups(n) * 8 : 0;
}
// scratch shared by the launches of one replicate chunk: transposed masks + count tables + segment totals
static inline size_t count_scratch_bytes(int n, int nV, int64_t R) {
  return arena_bytes((size_t)((R + 31) / 32) * n, 4) + arena_bytes((size_t)nV * R * table_groups(n), 8) +
         arena_bytes((size_t)nV * R * kCountWarps, 4) + arena_bytes((size_t)R * (vr_dev_stride(n) / 32), 4);
}
struct CountScratch {
  uint32_t* pairbits;
  uint32_t* maskT;
  uint2* tables;
  uint32_t* seg_tot;
};

// kept words + counts of every (vector, replicate) of the chunk
static int launch_count(const BlobList& blobs, int nV, int n, const uint32_t* masks, int R, void* scratch,
                        size_t scratch_bytes, CountScratch* cs, cudaStream_t stream) {
  const int W = vr_mask_words(n);
  const uint32_t Gs = table_groups(n);
  const int batches = (R + 31) / 32;
  Arena a(scratch, scratch_bytes);
  cs->maskT = a.take<uint32_t>((size_t)batches * n);
  cs->tables = a.take<uint2>((size_t)nV * R * Gs);
  cs->seg_tot = a.take<uint32_t>((size_t)nV * R * kCountWarps);
  const uint32_t words = (uint32_t)(vr_dev_stride(n) / 32);
  cs->pairbits = a.take<uint32_t>((size_t)R * words);
  if (!a.ok) { set_error("rank scratch too small"); return VR_ERR_WORKSPACE; }
  ProfScope prof(F_KEPT, stream);
  {
    dim3 grid((words + 255) / 256 < 64 ? (words + 255) / 256 : 64, R);
    pair_mask_kernel<<<grid, 256, (size_t)W * 4, stream>>>(masks, R, n, W, words, cs->pairbits);
    VR_CHECK_LAUNCH();
  }
  {
    dim3 grid((n + 255) / 256, batches);
    transpose_masks_kernel<<<grid, 256, 0, stream>>>(masks, R, n, W, cs->maskT);
    VR_CHECK_LAUNCH();
  }
  static bool attr_set = false;
  const size_t smem = (size_t)n * 4;
  if (!attr_set && smem > 40 * 1024) {
    VR_CHECK_CUDA(cudaFuncSetAttribute(count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  dim3 grid(batches, nV);
  count_kernel<<<grid, kCountThreads, smem, stream>>>(blobs, n, cs->maskT, R, cs->tables, Gs, cs->seg_tot);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

template <bool kSmem, bool kFused, typename Acc>
static int launch_rank_variant(const RankArgs& args, int grid, size_t smem, cudaStream_t stream) {
  static bool attr_set = false;
  if (kSmem && !attr_set) {
    VR_CHECK_CUDA(cudaFuncSetAttribute(rank_kernel<kSmem, kFused, Acc>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kRankSmemTableMax + 8192)));
    attr_set = true;
  }
  rank_kernel<kSmem, kFused, Acc><<<grid, kRankThreads, smem, stream>>>(args);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

static int launch_rank(const RankArgs& args, bool fused, cudaStream_t stream) {
  const int n = args.n;
  const bool in_smem = rank_table_in_smem(n);
  const size_t smem = rank_smem_bytes(n, in_smem);
  const int64_t items = (int64_t)args.n_vectors * args.R;
  const int grid = (int)(items < kNumSMs ? items : kNumSMs);
  const bool exact = triu_len(n) <= (1ll << 20);            // N^3 < 2^60: integer accumulation cannot overflow
  ProfScope prof(fused ? F_RANK_FUSED : F_RANK, stream);
  if (in_smem) {
    if (fused) return exact ? launch_rank_variant<true, true, long long>(args, grid, smem, stream)
                            : launch_rank_variant<true, true, double>(args, grid, smem, stream);
    return exact ? launch_rank_variant<true, false, long long>(args, grid, smem, stream)
                 : launch_rank_variant<true, false, double>(args, grid, smem, stream);
  }
  if (fused) return exact ? launch_rank_variant<false, true, long long>(args, grid, smem, stream)
                          : launch_rank_variant<false, true, double>(args, grid, smem, stream);
  return exact ? launch_rank_variant<false, false, long long>(args, grid, smem, stream)
               : launch_rank_variant<false, false, double>(args, grid, smem, stream);
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

extern "C" size_t vr_subsample_rank_workspace_bytes(int n, int R) {
  return (n >= 2 && R >= 1) ? count_scratch_bytes(n, 1, R) + 256 : 256;
}

extern "C" int vr_subsample_rank(const void* presorted, int n, const uint32_t* masks, int R, int32_t* dev,
                                 double* ss, int64_t* nkept, void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(presorted && masks && dev && ss && nkept, "vr_subsample_rank: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && R >= 1, "vr_subsample_rank: bad sizes n=%d R=%d", n, R);
  if (!ws || ws_bytes < vr_subsample_rank_workspace_bytes(n, R)) {
    set_error("vr_subsample_rank: workspace %zu < %zu", ws_bytes, vr_subsample_rank_workspace_bytes(n, R));
    return VR_ERR_WORKSPACE;
  }
  RankArgs args{};
  args.blobs.p[0] = presorted;
  CountScratch cs;
  int st = launch_count(args.blobs, 1, n, masks, R, ws, ws_bytes, &cs, stream);
  if (st != VR_OK) return st;
  args.n_vectors = 1; args.n = n; args.pairbits = cs.pairbits; args.R = R;
  args.tables = cs.tables; args.table_base = 0; args.Gs = table_groups(n); args.seg_tot = cs.seg_tot;
  args.dev_stride = vr_dev_stride(n); args.dev = dev; args.ss = ss; args.nkept = nkept;
  return launch_rank(args, false, stream);
}

namespace {
// VR_SPEARMAN_FUSED=1: the larger side multiplies on the fly (no dev vectors for it, no dot kernel).
// Default: both sides are materialised and dot_kernel combines them.
inline bool fused_mode() {
  static const bool f = [] { const char* e = getenv("VR_SPEARMAN_FUSED"); return e && e[0] == '1'; }();
  return f;
}
struct SpearmanPlan {
  size_t fixed;     // transposed masks + slack
  size_t per_rep;   // tables + segment totals of every vector, dev + ss of one materialised group
};
SpearmanPlan spearman_plan(int n, int nV, int n_mat) {
  SpearmanPlan p;
  p.fixed = arena_bytes((size_t)n * 2100, 4) + 8192;   // transposed masks of up to 65535 replicates + slack
  p.per_rep = (size_t)nV * ((size_t)table_groups(n) * 8 + 4 * kCountWarps) +
              (size_t)(fused_mode() ? n_mat : nV) * ((size_t)vr_dev_stride(n) * 4 + 8) + (size_t)vr_dev_stride(n) / 8;
  return p;
}
constexpr size_t kSoftCap = (size_t)40 << 30;  // default upper bound on one call's workspace
inline int mat_group(int nA, int nB) { int m = nA <= nB ? nA : nB; return m < kFuseMax ? m : kFuseMax; }
}  // namespace

extern "C" size_t vr_bootstrap_spearman_workspace_bytes(int n, int nA, int nB, int R) {
  if (n < 2 || nA < 1 || nB < 1 || R < 1) return 4096;
  SpearmanPlan p = spearman_plan(n, nA + nB, mat_group(nA, nB));
  size_t reps = (size_t)R;
  size_t cap = kSoftCap / p.per_rep;
  if (cap < 1) cap = 1;
  if (reps > cap) reps = cap;
  return p.fixed + reps * p.per_rep + 8 * 256;
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
  // The side with fewer vectors is materialised (in groups of <= kFuseMax); the other side is fused.
  const bool mat_is_A = nA <= nB;
  const int nM = mat_is_A ? nA : nB, nF = mat_is_A ? nB : nA;
  const int group = mat_group(nA, nB);
  SpearmanPlan plan = spearman_plan(n, nV, group);
  if (!ws || ws_bytes < plan.fixed + plan.per_rep + 8 * 256) {
    set_error("vr_bootstrap_spearman: workspace %zu < minimum %zu", ws_bytes, plan.fixed + plan.per_rep + 8 * 256);
    return VR_ERR_WORKSPACE;
  }
  BlobList all{};   // materialised side first, fused side after
  for (int i = 0; i < nM; ++i) { const void* p = mat_is_A ? A[i] : B[i]; VR_REQUIRE(p, "null blob"); all.p[i] = p; }
  for (int i = 0; i < nF; ++i) { const void* p = mat_is_A ? B[i] : A[i]; VR_REQUIRE(p, "null blob"); all.p[nM + i] = p; }
  int64_t chunk = (int64_t)((ws_bytes - plan.fixed - 8 * 256) / plan.per_rep);
  if (chunk > R) chunk = R;
  if (chunk > 65535) chunk = 65535;
  Arena a(ws, ws_bytes);
  const size_t scratch_bytes = count_scratch_bytes(n, nV, chunk) + 256;
  void* scratch = a.take<char>(scratch_bytes);
  const int n_dev = fused_mode() ? group : nV;
  double* ss = a.take<double>((size_t)n_dev * chunk);
  int32_t* dev = a.take<int32_t>((size_t)n_dev * chunk * vr_dev_stride(n));
  if (!a.ok) { set_error("vr_bootstrap_spearman: internal workspace plan overflow"); return VR_ERR_WORKSPACE; }
  const int W = vr_mask_words(n);
  for (int64_t r0 = 0; r0 < R; r0 += chunk) {
    const int rc = (int)(R - r0 < chunk ? R - r0 : chunk);
    const uint32_t* mk = masks + (size_t)r0 * W;
    CountScratch cs;
    int st = launch_count(all, nV, n, mk, rc, scratch, scratch_bytes, &cs, stream);
    if (st != VR_OK) return st;
    if (!fused_mode()) {
      // materialise every vector (model side first, as laid out in `all` when mat_is_A, else neural first)
      RankArgs mat{};
      mat.n = n; mat.pairbits = cs.pairbits; mat.R = rc; mat.tables = cs.tables; mat.Gs = table_groups(n);
      mat.seg_tot = cs.seg_tot; mat.dev_stride = vr_dev_stride(n);
      for (int i = 0; i < nV; ++i) mat.blobs.p[i] = all.p[i];
      mat.n_vectors = nV; mat.table_base = 0; mat.dev = dev; mat.ss = ss; mat.nkept = nullptr;
      st = launch_rank(mat, false, stream);
      if (st != VR_OK) return st;
      // dot_kernel expects the model vectors first: when the neural side was laid out first, swap the roles
      // and transpose on output by calling it with (nM = rows) and writing through a strided view.
      const int tiles = ((nM + kDotTile - 1) / kDotTile) * ((nF + kDotTile - 1) / kDotTile);
      dim3 grid(tiles, rc);
      ProfScope prof(F_DOT, stream);
      dot_kernel<<<grid, kDotThreads, 0, stream>>>(dev, vr_dev_stride(n), ss, nM, nF, rc, (int)r0, R, out,
                                                   mat_is_A ? 1 : 0, nB);
      VR_CHECK_LAUNCH();
      continue;
    }
    for (int m0 = 0; m0 < nM; m0 += group) {
      const int mg = nM - m0 < group ? nM - m0 : group;
      RankArgs base{};
      base.n = n; base.pairbits = cs.pairbits; base.R = rc; base.tables = cs.tables; base.Gs = table_groups(n);
      base.seg_tot = cs.seg_tot; base.dev_stride = vr_dev_stride(n);
      RankArgs mat = base;          // materialise vectors [m0, m0 + mg)
      for (int i = 0; i < mg; ++i) mat.blobs.p[i] = all.p[m0 + i];
      mat.n_vectors = mg; mat.table_base = m0; mat.dev = dev; mat.ss = ss; mat.nkept = nullptr;
      st = launch_rank(mat, false, stream);
      if (st != VR_OK) return st;
      RankArgs fu = base;           // every fused vector against this group
      for (int i = 0; i < nF; ++i) fu.blobs.p[i] = all.p[nM + i];
      fu.n_vectors = nF; fu.table_base = nM; fu.mat_dev = dev; fu.mat_ss = ss; fu.n_mat = mg;
      fu.out = out; fu.out_nB = nB; fu.out_Rtot = R; fu.out_r0 = (int)r0;
      fu.mat_is_A = mat_is_A ? 1 : 0; fu.mat_off = m0; fu.fused_off = 0;
      st = launch_rank(fu, true, stream);
      if (st != VR_OK) return st;
    }
  }
  return VR_OK;
}
