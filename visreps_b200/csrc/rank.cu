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
//   word (coalesced) gives C(s) and C(e) from one table entry (a predicated second read when the run
//   ends in the next 32-group) + two popc, without a data-dependent branch.
//   Reads and the dev write are coalesced; the only random accesses hit the table.
//   The side of the comparison with fewer vectors is materialised (dev vectors in HBM); the items of
//   the other side multiply their dev values with those vectors on the fly (fused) and emit rho.
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "pairmask.cuh"

namespace vr {

constexpr int kRankThreadsMat = 1024;   // materialise: 64 registers per thread
constexpr int kRankThreadsFused = 768;  // fused: 85 registers per thread, all 8 materialised loads in flight
                                        // (512 threads x two groups: 11.2 ms, 1024 threads x two load batches: 11.0 ms, 896 threads: 12.9 ms, this: 9.9 ms)
constexpr int kCountThreads = 1024;  // 32 warp segments per (vector, 32-replicate batch): 8192 warps fill the machine at NSD shape
constexpr int kCountWarps = kCountThreads / 32;
constexpr int kCountCtasMax = 16;       // CTAs sharing one (vector, 32-replicate batch) when there are few of them
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
// acc += d . m  over 4 int32 pairs, exact: four 32x32->64 multiply-adds (IMAD.WIDE).
__device__ __forceinline__ long long mad_wide(int a, int b, long long c) {
  long long r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
  return r;
}
__device__ __forceinline__ long long mac4(long long acc, const int32_t (&d)[4], const int4& m) {
  return mad_wide(d[3], m.w, mad_wide(d[2], m.z, mad_wide(d[1], m.y, mad_wide(d[0], m.x, acc))));
}
// fp64 accumulator (very long vectors): the four products are summed exactly in int64 (< 2^60), then folded in
__device__ __forceinline__ double mac4(double acc, const int32_t (&d)[4], const int4& m) {
  return acc + (double)mac4(0ll, d, m);
}
// acc += d * d: one IMAD.WIDE for the exact accumulator
__device__ __forceinline__ long long square_acc(long long acc, int d) { return mad_wide(d, d, acc); }
__device__ __forceinline__ double square_acc(double acc, int d) { return acc + (double)d * (double)d; }
// x << k with PTX semantics: k >= 32 gives 0 (C++ leaves that undefined).
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t k) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(k));
  return r;
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

// Count tables: {kept word, kept count before the group} per 32 sorted positions, Gs entries per (vector,
// replicate); a vector's replicates are padded to a multiple of 32 (table_entry_offset).
//   linear      [replicate][Gs]: what rank_kernel reads in place when the table does not fit shared memory;
//   interleaved [replicate / 32][Gs / 2][replicate % 32][2]: a lane of count_kernel owns one replicate, so
//               the 32 lanes' 16-byte pieces of a group pair are one contiguous 512-byte store (the linear
//               layout scatters them over 32 tables: 64 wavefronts per four groups, the kernel's whole cost);
//               rank_kernel straightens it out while copying the table into shared memory.
__host__ __device__ inline int64_t table_vector_stride(int64_t R, uint32_t Gs) { return (R + 31) / 32 * 32 * (int64_t)Gs; }
template <bool kInterleaved>
__device__ __forceinline__ int64_t table_entry_offset(int r, uint32_t g, uint32_t Gs) {
  if (!kInterleaved) return (int64_t)r * Gs + g;
  return (int64_t)(r >> 5) * 32 * Gs + (int64_t)(g >> 1) * 64 + (r & 31) * 2 + (g & 1u);
}

// Dynamic shared memory: maskT of the batch, n words.
template <bool kInterleaved>
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
  const uint32_t n_seg = kCountWarps * gridDim.z;           // one segment per warp; gridDim.z CTAs share a (vector, batch)
  const uint32_t seg = blockIdx.z * kCountWarps + warp;
  const uint32_t sg_per_warp = (SG + n_seg - 1) / n_seg;
  const uint32_t g_lo = min(G, seg * sg_per_warp * 32u);
  const uint32_t g_hi = min(G, g_lo + sg_per_warp * 32u);
  PresortView blob = presort_view(const_cast<void*>(blobs.p[v]), n);
  const uint32_t* __restrict__ stream = blob.stream;
  for (int i = threadIdx.x; i < n; i += kCountThreads) s_maskT[i] = maskT[(int64_t)batch * n + i];
  __syncthreads();

  const int rep = batch * 32 + (int)lane;                 // the replicate this lane accumulates
  uint2* __restrict__ vec_table = tables + (int64_t)v * table_vector_stride(R, Gs);
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
    if (kInterleaved || rep < R) {     // interleaved: the padding replicates' slots exist and keep the store whole
      *reinterpret_cast<uint4*>(vec_table + table_entry_offset<kInterleaved>(rep, g0, Gs)) = make_uint4(K[0], C[0], K[1], C[1]);
      *reinterpret_cast<uint4*>(vec_table + table_entry_offset<kInterleaved>(rep, g0 + 2u, Gs)) = make_uint4(K[2], C[2], K[3], C[3]);
    }
  }
  // counts are relative to the warp's segment; rank_kernel adds the segment bases from seg_tot
  if (rep < R) seg_tot[((int64_t)v * R + rep) * n_seg + seg] = running;
}

// ---------------------------------------------------------------------------------------------
// rank_kernel: one CTA per (vector, replicate) item, persistent over items.
//   kFused == false ("materialise"): writes dev (int32, canonical order, 0 for dropped pairs) and ss.
//   kFused == true : the item's dev values never leave registers; they are multiplied on the fly with
//                    the <= kFuseMax materialised vectors of the same replicate (coalesced int4 reads)
//                    and the finished Spearman coefficients are written.  Items are ordered
//                    replicate-major so the CTAs sharing one replicate's materialised vectors run
//                    together and hit L2.
//   Acc = long long (exact) while N^3 fits 2^60, else double.
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
  double* out;               // out[f * out_stride_f + m * out_stride_m + r]
  int64_t out_stride_f;
  int64_t out_stride_m;
  // global-table path (table too large for shared memory): an item is swept by `slices` CTAs at a time
  int slices;                // canonical range cut into this many slices of slice_len pairs
  uint32_t slice_len;
  unsigned long long* partial;   // [n_vectors][R][slices][kFuseMax + 1] raw 64-bit partial sums (int64 or fp64)
  uint32_t* ns;              // [n_vectors][R] kept pairs, written by table_finish_kernel
  int n_seg;                 // count segments per (vector, replicate): kCountWarps x CTAs per batch
};

// Global-table path, step 1: add the segment bases to every table in place, write the sentinel entries and the
// number of kept pairs.  grid (chunks of 1024 entries, vectors * R).
__global__ void __launch_bounds__(256)
table_finish_kernel(uint2* __restrict__ tables, int table_base, const uint32_t* __restrict__ seg_tot, int n_seg, int R, uint32_t G,
                    uint32_t Gs, uint32_t seg_groups, uint32_t* __restrict__ ns, int64_t item0) {
  __shared__ uint32_t s_base[kCountWarps * kCountCtasMax + 1];
  const int64_t item = item0 + blockIdx.y;
  const int v = (int)(item / R), r = (int)(item % R);
  const int64_t tix = (int64_t)(table_base + v) * R + r;
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int w = 0; w < n_seg; ++w) { s_base[w] = run; run += seg_tot[tix * n_seg + w]; }
    s_base[n_seg] = run;
    if (blockIdx.x == 0) ns[(int64_t)v * R + r] = run;
  }
  __syncthreads();
  uint2* __restrict__ tab = tables + (int64_t)(table_base + v) * table_vector_stride(R, Gs) + (int64_t)r * Gs;
  const uint32_t Ns = s_base[n_seg];
  for (uint32_t g = blockIdx.x * 1024u + threadIdx.x; g < min(Gs, (blockIdx.x + 1u) * 1024u); g += 256u) {
    uint2 e = make_uint2(0u, Ns);
    if (g < G) { e = tab[g]; e.y += s_base[g / seg_groups]; }
    tab[g] = e;
  }
}

template <bool kSmemTable, bool kFused, typename Acc, bool kCompact, bool kUntied, int kNMat>
__global__ void __launch_bounds__(kFused ? kRankThreadsFused : kRankThreadsMat, 1)
rank_kernel(const __grid_constant__ RankArgs A) {
  static_assert(kNMat >= 1 && kNMat <= kFuseMax, "materialised vectors multiplied per pass");
  static_assert(kCompact || !kUntied, "the tie-free sweep reads compact run words");
  constexpr int kRankThreads = kFused ? kRankThreadsFused : kRankThreadsMat;
  constexpr int kRankWarps = kRankThreads / 32;
  extern __shared__ __align__(16) uint32_t smem[];
  uint2* s_table = reinterpret_cast<uint2*>(smem);
  using Red = std::conditional_t<std::is_integral<Acc>::value, long long, double>;   // integer sums stay exact to the end
  __shared__ Red s_red[kRankWarps][kFuseMax + 1];
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

  // This instantiation takes the vectors of its kind: tie-free ones (header n_runs == N, decided on the device by
  // the presort) go to the kUntied sweep, the others to the general one; the host launches both.
  __shared__ int s_vec[kMaxVectors];
  __shared__ int s_nvec;
  if (threadIdx.x == 0) {
    int c = 0;
    for (int v = 0; v < A.n_vectors; ++v) {
      const bool tie_free = kCompact && __ldg(&static_cast<const PresortHeader*>(A.blobs.p[v])->n_runs) == N;
      if (tie_free == kUntied) s_vec[c++] = v;
    }
    s_nvec = c;
  }
  __syncthreads();
  const int n_vec = s_nvec;

  // Shared-memory table: one CTA per item.  Global table: an item is cut into A.slices slices and consecutive CTAs
  // take consecutive slices, so that the CTAs in flight share a few tables (L2 / L1 resident) instead of 148.
  const int slices = kSmemTable ? 1 : A.slices;
  const int64_t items = (int64_t)n_vec * R;
  const int64_t units = items * slices;
  for (int64_t unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int64_t item = unit / slices;
    const int slice = (int)(unit % slices);
    // materialise: vector-major (CTAs share the vector's `runs` stream); fused: replicate-major
    const int v = s_vec[kFused ? (int)(item % n_vec) : (int)(item / R)];
    const int r = kFused ? (int)(item / n_vec) : (int)(item % R);
    const uint32_t lo = kSmemTable ? 0u : (uint32_t)slice * A.slice_len;
    const uint32_t hi = kSmemTable ? (uint32_t)dev_stride : (uint32_t)min((int64_t)dev_stride, (int64_t)lo + A.slice_len);
    PresortView blob = presort_view(const_cast<void*>(A.blobs.p[v]), n);
    const int64_t tix = (int64_t)(A.table_base + v) * R + r;
    uint2* __restrict__ vec_table = A.tables + (int64_t)(A.table_base + v) * table_vector_stride(R, Gs);
    uint2* __restrict__ g_table = vec_table + (int64_t)r * Gs;     // linear layout (table read in place)
    uint32_t Ns;
    if constexpr (kSmemTable) {
      if (threadIdx.x == 0) {                                // exclusive prefix of the segment totals
        uint32_t run = 0;
        for (int w = 0; w < kCountWarps; ++w) { s_base[w] = run; run += A.seg_tot[tix * kCountWarps + w]; }
        s_base[kCountWarps] = run;
      }
      __syncthreads();
      Ns = s_base[kCountWarps];
      // bring the table into shared memory, adding the segment bases; entries >= G become the sentinel
      // {0, N_r} so that C(N) needs no special case
      for (uint32_t g = threadIdx.x; g < Gs; g += kRankThreads) {
        uint2 e = make_uint2(0u, Ns);
        if (g < G) { e = vec_table[table_entry_offset<true>(r, g, Gs)]; e.y += s_base[g / seg_groups]; }
        s_table[g] = e;
      }
      __syncthreads();
    } else {
      Ns = __ldg(A.ns + (int64_t)v * R + r);                 // the table was finished in place by table_finish_kernel
    }
    const uint2* __restrict__ table = kSmemTable ? s_table : g_table;

    int32_t* __restrict__ out_dev = kFused ? nullptr : A.dev + ((int64_t)v * R + r) * dev_stride;
    const int32_t* __restrict__ mat = kFused ? A.mat_dev + (int64_t)r * dev_stride : nullptr;
    // materialised vectors as int4 streams; 32-bit offsets (host guarantees n_mat * R * dev_stride < 2^34)
    const int4* __restrict__ mat4 = reinterpret_cast<const int4*>(mat);
    const uint32_t mat_step4 = (uint32_t)(((int64_t)R * dev_stride) >> 2);
    uint32_t moff[kNMat];                                    // vectors past n_mat alias the last one (results unused)
#pragma unroll
    for (int a = 0; a < kNMat; ++a) moff[a] = kFused ? (uint32_t)min(a, A.n_mat - 1) * mat_step4 + threadIdx.x : 0u;
    const uint32_t* __restrict__ runs = blob.runs;
    const uint32_t* __restrict__ keptw = A.pairbits + (int64_t)r * (dev_stride >> 5);
    // sum of squares in fp64: exact while a thread's share stays below 2^53 (it does whenever Acc is integral:
    // <= 1365 pairs of |d| <= 2.1e6), and its I2F + DFMA run on the XU / fp64 pipes the sweep leaves idle
    double acc_ss = 0.0;
    Acc acc[kNMat];
#pragma unroll
    for (int a = 0; a < kNMat; ++a) acc[a] = 0;

    // A vector without ties has run length 1 everywhere, so C(s) + C(e) = 2 C(s) + 1 for every kept pair:
    // one table entry, one shift, one popc -- and its sum of squares is the closed form N_r (N_r^2 - 1) / 3.
    const char* __restrict__ table_bytes = reinterpret_cast<const char*>(table);

    const uint32_t esc_word = esc << sbits;                  // a run word >= esc_word carries the escape code
    // ranks of four consecutive canonical pairs: q = their run words, kw = their kept bits (low four)
    auto rank4 = [&](const uint4 q, const uint32_t kw, int32_t (&d4)[4]) {
      const uint32_t rw[4] = {q.x, q.y, q.z, q.w};
      if (kCompact && kUntied) {
        const int32_t bias = 1 - (int32_t)Ns;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t w = rw[e];
          const uint2 ts = *reinterpret_cast<const uint2*>(table_bytes + (w & kRunOffMask));
          const int32_t c = (int32_t)(ts.y + __popc(shl_clamp(ts.x, w >> kRunShloShift)));
          // fused: a dropped pair's materialised partner is 0, so its (meaningless) value needs no masking
          d4[e] = (kFused || ((kw >> e) & 1u)) ? 2 * c + bias : 0;
        }
      } else if (kCompact) {
        uint32_t sum[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t w = rw[e];
          const char* entry = table_bytes + (w & kRunOffMask);
          const uint2 ts = *reinterpret_cast<const uint2*>(entry);
          uint2 te = ts;
          if (w & kRunCrossBit) te = *reinterpret_cast<const uint2*>(entry + 8);     // the run ends inside the next group
          sum[e] = ts.y + __popc(shl_clamp(ts.x, w >> kRunShloShift)) +
                   te.y + __popc(shl_clamp(te.x, (w >> kRunSheShift) & 63u));
        }
        // a long run (shlo == 0) among the four?  Rare (more than 32 equal values): its end comes from the head tables
        if (min(min(rw[0], rw[1]), min(rw[2], rw[3])) < (1u << kRunShloShift)) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t w = rw[e];
            if ((w >> kRunShloShift) == 0u) {
              const uint32_t gs = (w & kRunOffMask) >> 3, ls = ((w >> kRunSheShift) & 63u) - 1u;
              const uint32_t above = ls == 31u ? 0u : (__ldg(blob.head_bits + gs) >> (ls + 1u));
              const uint32_t endpos = above ? gs * 32u + ls + (uint32_t)__ffs(above) : __ldg(blob.next_head + gs + 1);
              const uint2 ts = table[gs], tl = table[endpos >> 5];
              sum[e] = ts.y + __popc(shl_clamp(ts.x, 32u - ls)) + tl.y + __popc(shl_clamp(tl.x, 32u - (endpos & 31u)));
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int32_t d = ((kw >> e) & 1u) ? (int32_t)sum[e] - (int32_t)Ns : 0;
          d4[e] = d;
          acc_ss = fma((double)d, (double)d, acc_ss);
        }
      } else {
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
          acc_ss = fma((double)d, (double)d, acc_ss);
        }
      }
    };

    // software pipeline: the run words and kept bits of the NEXT step are in flight while this one is ranked
    if constexpr (kFused) {
      uint32_t p0 = lo + threadIdx.x * 4u;
      const uint32_t n_hi = min(N, hi);
      uint4 q_cur = make_uint4(0u, 0u, 0u, 0u);
      uint32_t kw_cur = 0u;
      if (p0 < n_hi) { q_cur = __ldg(reinterpret_cast<const uint4*>(runs + p0)); kw_cur = __ldg(keptw + (p0 >> 5)); }
      for (; p0 < hi; p0 += kRankThreads * 4u) {
        int32_t d4[4] = {0, 0, 0, 0};
        const uint4 q = q_cur;
        const uint32_t kw = kw_cur >> (p0 & 31u);              // 4 kept bits (p0 is a multiple of 4)
        const bool any_kept = p0 < N && (kw & 0xFu) != 0u;      // a dropped stimulus' row: all four pairs are out
        int4 m4[kNMat];
        if (any_kept) {   // all materialised vectors are requested before the ranking work, which hides their latency
#pragma unroll
          for (int a = 0; a < kNMat; ++a) m4[a] = __ldg(mat4 + (moff[a] + ((p0 >> 2) - threadIdx.x)));
        }
        {
          const uint32_t pn = p0 + kRankThreads * 4u;
          if (pn < n_hi) { q_cur = __ldg(reinterpret_cast<const uint4*>(runs + pn)); kw_cur = __ldg(keptw + (pn >> 5)); }
        }
        if (any_kept) {
          rank4(q, kw, d4);
#pragma unroll
          for (int a = 0; a < kNMat; ++a) acc[a] = mac4(acc[a], d4, m4[a]);
        }
      }
    } else {
      // each warp takes 256 consecutive pairs per step as two coalesced groups of 128: eight independent
      // table lookups per thread are in flight, and both groups usually lie in the same RDM row, so a
      // dropped stimulus' row still skips the work
      constexpr uint32_t kStep = kRankThreads * 8u;
      uint32_t p0 = lo + (threadIdx.x >> 5) * 256u + lane * 4u;
      const uint32_t n_hi = min(N, hi);
      constexpr uint32_t kPadWord = kCompact ? 32u << kRunShloShift : 0u;     // what the presort stores beyond N
      uint4 qa_cur = make_uint4(kPadWord, kPadWord, kPadWord, kPadWord), qb_cur = qa_cur;
      uint32_t ka_cur = 0u, kb_cur = 0u;
      if (p0 < n_hi) { qa_cur = __ldg(reinterpret_cast<const uint4*>(runs + p0)); ka_cur = __ldg(keptw + (p0 >> 5)); }
      if (p0 + 128u < n_hi) { qb_cur = __ldg(reinterpret_cast<const uint4*>(runs + p0 + 128u)); kb_cur = __ldg(keptw + ((p0 + 128u) >> 5)); }
      for (; p0 < hi; p0 += kStep) {
        const uint32_t p1 = p0 + 128u;
        int32_t da[4] = {0, 0, 0, 0}, db[4] = {0, 0, 0, 0};
        const uint4 qa = qa_cur, qb = qb_cur;
        const uint32_t ka = p0 < n_hi ? (ka_cur >> (p0 & 31u)) & 0xFu : 0u;
        const uint32_t kb = p1 < n_hi ? (kb_cur >> (p1 & 31u)) & 0xFu : 0u;
        {
          const uint32_t pn = p0 + kStep;
          if (pn < n_hi) { qa_cur = __ldg(reinterpret_cast<const uint4*>(runs + pn)); ka_cur = __ldg(keptw + (pn >> 5)); }
          if (pn + 128u < n_hi) { qb_cur = __ldg(reinterpret_cast<const uint4*>(runs + pn + 128u)); kb_cur = __ldg(keptw + ((pn + 128u) >> 5)); }
        }
        if ((ka | kb) != 0u) {         // run words beyond N are padding words: valid to rank, never kept
          rank4(qa, ka, da);
          rank4(qb, kb, db);
        }
        *reinterpret_cast<int4*>(out_dev + p0) = make_int4(da[0], da[1], da[2], da[3]);
        if (p1 < hi) *reinterpret_cast<int4*>(out_dev + p1) = make_int4(db[0], db[1], db[2], db[3]);
      }
    }

    // block reduction: ss (slot kFuseMax) and, when fused, the n_mat dot products
    {
      auto wsum = [](Acc x) -> Red {
        if constexpr (std::is_integral<Acc>::value) return warp_sum_ll((long long)x);
        else return warp_sum((double)x);
      };
      const Red v_ss = wsum((Acc)acc_ss);
      if (lane == 0) s_red[warp][kFuseMax] = v_ss;
      if (kFused) {
#pragma unroll
        for (int a = 0; a < kNMat; ++a) {
          if (a < A.n_mat) {
            const Red t = wsum(acc[a]);
            if (lane == 0) s_red[warp][a] = t;
          }
        }
      }
    }
    __syncthreads();
    if (!kSmemTable) {               // a slice only contributes partial sums; rank_finalize_kernel adds them up in order
      if (threadIdx.x <= kFuseMax) {
        const int slot = threadIdx.x;
        Red tot_r = 0;
        if (slot == kFuseMax || (kFused && slot < A.n_mat))
          for (int w = 0; w < kRankWarps; ++w) tot_r += s_red[w][slot];
        unsigned long long raw;
        if constexpr (std::is_integral<Acc>::value) raw = (unsigned long long)tot_r;
        else raw = (unsigned long long)__double_as_longlong((double)tot_r);
        A.partial[((((int64_t)v * R + r) * slices) + slice) * (kFuseMax + 1) + slot] = raw;
      }
    } else if (threadIdx.x <= kFuseMax) {
      const int slot = threadIdx.x;
      Red tot_r = 0, ss_r = 0;       // every participating lane needs ss of this item as well
      if (slot == kFuseMax || (kFused && slot < A.n_mat))
        for (int w = 0; w < kRankWarps; ++w) tot_r += s_red[w][slot];
      for (int w = 0; w < kRankWarps; ++w) ss_r += s_red[w][kFuseMax];
      const double tot = (double)tot_r;
      double ss_item = (double)ss_r;
      const double nan = __longlong_as_double(0x7ff8000000000000LL);
      if (kUntied)                   // ranks are a permutation of 1..N_r: sum of (2 rank - N_r - 1)^2, exact in int64
        ss_item = (double)(((unsigned long long)Ns - 1ull) * Ns * ((unsigned long long)Ns + 1ull) / 3ull);   // < 2^64 for N_r < 2.6e6
      if (blob.hdr->has_nan &&                         // NaN only for a replicate that keeps a NaN entry (rsa.py:123-125)
          kept_nans(blob.hdr, N, Ns, [&](uint32_t x) { const uint2 t = table[x >> 5]; return t.y + __popc(t.x & low_mask(x & 31u)); }) != 0u)
        ss_item = nan;
      if (!kFused) {
        if (slot == kFuseMax) {
          A.ss[(int64_t)v * R + r] = ss_item;
          if (A.nkept) A.nkept[(int64_t)v * R + r] = (int64_t)Ns;
        }
      } else if (slot < A.n_mat) {
        const double ss_m = A.mat_ss[(int64_t)slot * R + r];
        double rho = tot / (sqrt(ss_m) * sqrt(ss_item));
        if (!(ss_m > 0.0) || !(ss_item > 0.0)) rho = nan;
        A.out[(int64_t)v * A.out_stride_f + (int64_t)slot * A.out_stride_m + r] = rho;
      }
    }
    __syncthreads();  // table, s_mask and s_red are reused by the next item
  }
}

// Global-table path, last step: one warp per item adds the slices' partial sums in slice order and writes what the
// single-CTA path writes at the end of an item.  grid = items of the launch (all vectors, tie-free or not).
template <bool kFused, bool kIntegral>
__global__ void __launch_bounds__(32)
rank_finalize_kernel(const __grid_constant__ RankArgs A) {
  const int v = blockIdx.x / A.R, r = blockIdx.x % A.R;
  const int slot = threadIdx.x;
  const uint32_t N = (uint32_t)triu_len(A.n);
  const PresortHeader* hdr = static_cast<const PresortHeader*>(A.blobs.p[v]);
  double mine = 0.0;
  if (slot <= kFuseMax) {
    const unsigned long long* p = A.partial + (((int64_t)v * A.R + r) * A.slices) * (kFuseMax + 1) + slot;
    if (kIntegral) {
      long long t = 0;
      for (int s = 0; s < A.slices; ++s) t += (long long)p[(int64_t)s * (kFuseMax + 1)];
      mine = (double)t;
    } else {
      for (int s = 0; s < A.slices; ++s) mine += __longlong_as_double((long long)p[(int64_t)s * (kFuseMax + 1)]);
    }
  }
  double ss_item = __shfl_sync(0xffffffffu, mine, kFuseMax);
  const uint32_t Ns = A.ns[(int64_t)v * A.R + r];
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if (run_compact(N) && hdr->n_runs == N)      // tie-free vector: the sweep did not accumulate squares (closed form)
    ss_item = (double)(((unsigned long long)Ns - 1ull) * Ns * ((unsigned long long)Ns + 1ull) / 3ull);
  if (hdr->has_nan) {
    const uint2* __restrict__ tab = A.tables + (int64_t)(A.table_base + v) * table_vector_stride(A.R, A.Gs) + (int64_t)r * A.Gs;   // finished in place
    if (kept_nans(hdr, N, Ns, [&](uint32_t x) { const uint2 t = tab[x >> 5]; return t.y + __popc(t.x & low_mask(x & 31u)); }) != 0u) ss_item = nan;
  }
  if (!kFused) {
    if (slot == kFuseMax) {
      A.ss[(int64_t)v * A.R + r] = ss_item;
      if (A.nkept) A.nkept[(int64_t)v * A.R + r] = (int64_t)Ns;
    }
  } else if (slot < A.n_mat) {
    const double ss_m = A.mat_ss[(int64_t)slot * A.R + r];
    double rho = mine / (sqrt(ss_m) * sqrt(ss_item));
    if (!(ss_m > 0.0) || !(ss_item > 0.0)) rho = nan;
    A.out[(int64_t)v * A.out_stride_f + (int64_t)slot * A.out_stride_m + r] = rho;
  }
}

static inline uint32_t table_groups(int n) { return (uint32_t)round_up(presort_groups(n) + 1, 32) + 32u; }
static inline bool rank_table_in_smem(int n) { return (size_t)table_groups(n) * 8 <= kRankSmemTableMax; }
static inline size_t rank_smem_bytes(int n, bool smem_table) {
  return smem_table ? (size_t)table_groups(n) * 8 : 0;
}

// ---- rank context: everything derived from the masks alone, shared by all rank calls of a job ----
//   pairbits [R][dev_stride/32]  kept bits of the canonical pairs        (rank_kernel)
//   maskT    [ceil(R/32)][n]     transposed stimulus masks               (count_kernel)
struct RankContext {
  uint32_t* pairbits;
  uint32_t* maskT;
};
static inline size_t rank_context_bytes(int n, int64_t R) {
  return 256 + arena_bytes((size_t)R * (vr_dev_stride(n) / 32), 4) + arena_bytes((size_t)((R + 31) / 32) * n, 4);
}
static inline RankContext rank_context_view(void* ctx, int n, int64_t R) {
  char* p = static_cast<char*>(ctx) + 256;
  RankContext c;
  c.pairbits = reinterpret_cast<uint32_t*>(p);
  p += arena_bytes((size_t)R * (vr_dev_stride(n) / 32), 4);
  c.maskT = reinterpret_cast<uint32_t*>(p);
  return c;
}
static int build_rank_context(const uint32_t* masks, int n, int R, void* ctx, cudaStream_t stream) {
  const int W = vr_mask_words(n);
  RankContext c = rank_context_view(ctx, n, R);
  const uint32_t words = (uint32_t)(vr_dev_stride(n) / 32);
  ProfScope prof(F_KEPT, stream);
  for (int r0 = 0; r0 < R; r0 += 65535) {   // grid.y limit
    const int rc = R - r0 < 65535 ? R - r0 : 65535;
    dim3 grid((words + 255) / 256 < 64 ? (words + 255) / 256 : 64, rc);
    pair_mask_kernel<<<grid, 256, (size_t)(W + 1) * 4, stream>>>(masks + (size_t)r0 * W, rc, n, W, words,
                                                           c.pairbits + (size_t)r0 * words);
    VR_CHECK_LAUNCH();
  }
  const int batches = (R + 31) / 32;
  dim3 grid((n + 255) / 256, batches);
  transpose_masks_kernel<<<grid, 256, 0, stream>>>(masks, R, n, W, c.maskT);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

// count tables + segment totals of nV vectors x R replicates
// Global-table path: slices of the canonical range swept by different CTAs.  A slice is a multiple of both sweeps'
// step (768 * 4 and 1024 * 8 pairs) and long enough (>= ~100k pairs) that the per-slice reduction stays cheap; at most
// one slice per SM, so that a table of the stress shape (12.5 MB) is shared by the whole machine.
struct SlicePlan { int slices; uint32_t slice_len; };
static inline SlicePlan rank_slice_plan(int n) {
  const int64_t stride = vr_dev_stride(n);
  if (rank_table_in_smem(n)) return {1, (uint32_t)stride};
  const int64_t unit = 24576;                                // lcm(768 * 4, 1024 * 8)
  int64_t target = stride / (unit * 4);
  if (target < 1) target = 1;
  if (target > kNumSMs) target = kNumSMs;
  const int64_t len = round_up((stride + target - 1) / target, unit);
  return {(int)((stride + len - 1) / len), (uint32_t)len};
}
// Large tables with few (vector, batch) pairs -- the stress shape with a handful of masks -- would leave the count
// to one or two CTAs; the group range is then cut over up to kCountCtasMax CTAs (more segments per table).
static inline int count_ctas(int n, int nV, int64_t R) {
  if (rank_table_in_smem(n)) return 1;
  const int64_t ctas = (int64_t)nV * ((R + 31) / 32);
  int64_t z = kNumSMs / (ctas > 0 ? ctas : 1);
  if (z < 1) z = 1;
  if (z > kCountCtasMax) z = kCountCtasMax;
  return (int)z;
}
static inline size_t tables_bytes(int n, int nV, int64_t R) {
  size_t b = arena_bytes((size_t)nV * table_vector_stride(R, table_groups(n)), 8) +
             arena_bytes((size_t)nV * R * kCountWarps * kCountCtasMax, 4);
  if (!rank_table_in_smem(n))
    b += arena_bytes((size_t)nV * R * rank_slice_plan(n).slices * (kFuseMax + 1), 8) + arena_bytes((size_t)nV * R, 4);
  return b;
}
static int launch_count(const BlobList& blobs, int nV, int n, const RankContext& ctx, int R, void* ws, size_t ws_bytes,
                        uint2** tables, uint32_t** seg_tot, unsigned long long** partial, uint32_t** ns, cudaStream_t stream) {
  const uint32_t Gs = table_groups(n);
  Arena a(ws, ws_bytes);
  *tables = a.take<uint2>((size_t)nV * table_vector_stride(R, Gs));
  *seg_tot = a.take<uint32_t>((size_t)nV * R * kCountWarps * kCountCtasMax);
  *partial = nullptr;
  *ns = nullptr;
  if (!rank_table_in_smem(n)) {
    *partial = a.take<unsigned long long>((size_t)nV * R * rank_slice_plan(n).slices * (kFuseMax + 1));
    *ns = a.take<uint32_t>((size_t)nV * R);
  }
  if (!a.ok) { set_error("rank workspace too small for the count tables"); return VR_ERR_WORKSPACE; }
  static unsigned long long attr_done = 0;
  const size_t smem = (size_t)n * 4;
  static unsigned long long attr_done_i = 0;
  const bool interleaved = rank_table_in_smem(n);          // rank_kernel copies the table to shared memory
  if (smem > 40 * 1024) {
    if (interleaved) VR_CHECK_CUDA(ensure_dynamic_smem(count_kernel<true>, 160 * 1024, &attr_done_i));
    else VR_CHECK_CUDA(ensure_dynamic_smem(count_kernel<false>, 160 * 1024, &attr_done));
  }
  ProfScope prof(F_KEPT, stream);
  dim3 grid((R + 31) / 32, nV, count_ctas(n, nV, R));
  if (interleaved) count_kernel<true><<<grid, kCountThreads, smem, stream>>>(blobs, n, ctx.maskT, R, *tables, Gs, *seg_tot);
  else count_kernel<false><<<grid, kCountThreads, smem, stream>>>(blobs, n, ctx.maskT, R, *tables, Gs, *seg_tot);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

template <bool kSmem, bool kFused, typename Acc, bool kCompact, bool kUntied, int kNMat>
static int launch_rank_inst(const RankArgs& args, int grid, size_t smem, cudaStream_t stream) {
  static unsigned long long attr_done = 0;
  auto kernel = rank_kernel<kSmem, kFused, Acc, kCompact, kUntied, kNMat>;
  if (kSmem) VR_CHECK_CUDA(ensure_dynamic_smem(kernel, (int)(kRankSmemTableMax + 8192), &attr_done));
  kernel<<<grid, kFused ? kRankThreadsFused : kRankThreadsMat, smem, stream>>>(args);
  VR_CHECK_LAUNCH();
  return VR_OK;
}
// The fused sweep is instantiated for 1, 4 and 8 materialised vectors: a single-pair bootstrap or a job with one
// neural RDM (THINGS behaviour) would otherwise load and multiply seven aliases of its only partner.
template <bool kSmem, bool kFused, typename Acc, bool kCompact, bool kUntied>
static int launch_rank_kind(const RankArgs& args, int grid, size_t smem, cudaStream_t stream) {
  if (kFused && args.n_mat <= 1) return launch_rank_inst<kSmem, kFused, Acc, kCompact, kUntied, kFused ? 1 : kFuseMax>(args, grid, smem, stream);
  if (kFused && args.n_mat <= 4) return launch_rank_inst<kSmem, kFused, Acc, kCompact, kUntied, kFused ? 4 : kFuseMax>(args, grid, smem, stream);
  return launch_rank_inst<kSmem, kFused, Acc, kCompact, kUntied, kFuseMax>(args, grid, smem, stream);
}
// Whether a vector has ties is known only on the device (presort header), so both sweeps are launched and each
// takes the vectors of its kind; the one with nothing to do returns after reading the headers.
template <bool kSmem, bool kFused, typename Acc, bool kCompact>
static int launch_rank_variant(const RankArgs& args, int grid, size_t smem, cudaStream_t stream) {
  if (kCompact) {
    int st = launch_rank_kind<kSmem, kFused, Acc, kCompact, kCompact>(args, grid, smem, stream);
    if (st != VR_OK) return st;
  }
  return launch_rank_kind<kSmem, kFused, Acc, kCompact, false>(args, grid, smem, stream);
}

template <bool kFused>
static int launch_rank_side(const RankArgs& args, cudaStream_t stream) {
  const int n = args.n;
  const int64_t N = triu_len(n);
  const bool in_smem = rank_table_in_smem(n);
  const size_t smem = rank_smem_bytes(n, in_smem);
  const int64_t items = (int64_t)args.n_vectors * args.R;
  const int64_t units = items * args.slices;
  const int grid = (int)(units < kNumSMs ? units : kNumSMs);
  if (!in_smem) {                                            // global tables: finish them in place first
    const uint32_t G = presort_groups(n);
    const uint32_t seg_groups = ((((G + 31u) / 32u) + args.n_seg - 1) / args.n_seg) * 32u;
    for (int64_t i0 = 0; i0 < items; i0 += 65535) {
      const int cnt = (int)(items - i0 < 65535 ? items - i0 : 65535);
      dim3 fg((args.Gs + 1023u) / 1024u, cnt);
      table_finish_kernel<<<fg, 256, 0, stream>>>(args.tables, args.table_base, args.seg_tot, args.n_seg, args.R, G, args.Gs, seg_groups, args.ns, i0);
      VR_CHECK_LAUNCH();
    }
  }
  // Integer accumulation is exact to the end while N^3 < 2^60.  Beyond that the sums are folded into fp64 every
  // four pairs; int64 would still fit up to N ~ 3e6, but measured at n = 1854 the fp64 variant is 7 % faster
  // (its conversions and adds run on the otherwise idle fp64 / XU pipes instead of the ALU pipe).
  const bool exact = N <= (1ll << 20);
  const bool compact = run_compact((uint32_t)N);
  if (in_smem) return launch_rank_variant<true, kFused, long long, true>(args, grid, smem, stream);   // N <= ~816k
  int st;
  if (exact) st = launch_rank_variant<false, kFused, long long, true>(args, grid, smem, stream);
  else if (compact) st = launch_rank_variant<false, kFused, double, true>(args, grid, smem, stream);
  else st = launch_rank_variant<false, kFused, double, false>(args, grid, smem, stream);
  if (st != VR_OK) return st;
  if (exact) rank_finalize_kernel<kFused, true><<<(unsigned)items, 32, 0, stream>>>(args);
  else rank_finalize_kernel<kFused, false><<<(unsigned)items, 32, 0, stream>>>(args);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

static int launch_rank(const RankArgs& args, bool fused, cudaStream_t stream) {
  ProfScope prof(fused ? F_RANK_FUSED : F_RANK, stream);
  return fused ? launch_rank_side<true>(args, stream) : launch_rank_side<false>(args, stream);
}

// count + rank of nV vectors: materialise (dev/ss) or fused against nM materialised vectors
static int rank_stage(const void* const* blobs_host, int nV, int n, const RankContext& ctx, int R, bool fused,
                      int32_t* dev, double* ss, int64_t* nkept, const int32_t* mat_dev, const double* mat_ss, int nM,
                      double* out, int64_t stride_f, int64_t stride_m, void* ws, size_t ws_bytes, cudaStream_t stream) {
  RankArgs args{};
  for (int i = 0; i < nV; ++i) {
    if (!blobs_host[i]) { set_error("null presorted blob"); return VR_ERR_INVALID; }
    args.blobs.p[i] = blobs_host[i];
  }
  uint2* tables = nullptr;
  uint32_t* seg_tot = nullptr;
  int st = launch_count(args.blobs, nV, n, ctx, R, ws, ws_bytes, &tables, &seg_tot, &args.partial, &args.ns, stream);
  if (st != VR_OK) return st;
  const SlicePlan sp = rank_slice_plan(n);
  args.slices = sp.slices; args.slice_len = sp.slice_len;
  args.n_seg = kCountWarps * count_ctas(n, nV, R);
  args.n_vectors = nV; args.n = n; args.pairbits = ctx.pairbits; args.R = R; args.tables = tables; args.table_base = 0;
  args.Gs = table_groups(n); args.seg_tot = seg_tot; args.dev_stride = vr_dev_stride(n);
  args.dev = dev; args.ss = ss; args.nkept = nkept;
  args.mat_dev = mat_dev; args.mat_ss = mat_ss; args.n_mat = nM;
  args.out = out; args.out_stride_f = stride_f; args.out_stride_m = stride_m;
  return launch_rank(args, fused, stream);
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

// Host-built masks (vr_host_draw_subsample_masks) -> device, read by the SMs straight from pinned host
// memory: 4*R*W bytes, never queued behind bulk transfers on the copy engines.
__global__ void fetch_masks_kernel(const uint32_t* __restrict__ src, int64_t count, uint32_t* __restrict__ dst) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) dst[t] = src[t];
}

extern "C" int vr_fetch_masks(const uint32_t* host_masks, int R, int n, uint32_t* masks, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(masks && (host_masks || R == 0), "vr_fetch_masks: null pointer");
  VR_REQUIRE(R >= 0 && n >= 1 && n <= VR_MAX_STIMULI, "vr_fetch_masks: bad sizes R=%d n=%d", R, n);
  const int64_t count = (int64_t)R * vr_mask_words(n);
  if (count == 0) return VR_OK;
  ProfScope prof(F_MASKS, stream);
  fetch_masks_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(host_masks, count, masks);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

extern "C" int64_t vr_dev_stride(int n) { return n >= 2 ? round_up(triu_len(n), 32) : 32; }

// ---- staged API ---------------------------------------------------------------------------------
extern "C" size_t vr_rank_context_bytes(int n, int R) { return (n >= 2 && R >= 1) ? rank_context_bytes(n, R) : 256; }

extern "C" int vr_rank_context_build(const uint32_t* masks, int n, int R, void* ctx, size_t ctx_bytes, vr_stream_t stream_) {
  VR_REQUIRE(masks && ctx, "vr_rank_context_build: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && R >= 1, "vr_rank_context_build: bad sizes n=%d R=%d", n, R);
  if (ctx_bytes < rank_context_bytes(n, R)) { set_error("vr_rank_context_build: context %zu < %zu", ctx_bytes, rank_context_bytes(n, R)); return VR_ERR_WORKSPACE; }
  return build_rank_context(masks, n, R, ctx, static_cast<cudaStream_t>(stream_));
}

extern "C" size_t vr_rank_workspace_bytes(int n, int nV, int R) { return (n >= 2 && nV >= 1 && R >= 1) ? tables_bytes(n, nV, R) + 256 : 256; }

extern "C" int vr_rank_materialise(const void* const* blobs, int nV, int n, const void* ctx, int R, int32_t* dev, double* ss,
                                   int64_t* nkept, void* ws, size_t ws_bytes, vr_stream_t stream_) {
  VR_REQUIRE(blobs && ctx && dev && ss, "vr_rank_materialise: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && R >= 1 && nV >= 1 && nV <= kMaxVectors, "vr_rank_materialise: bad sizes n=%d nV=%d R=%d", n, nV, R);
  if (!ws || ws_bytes < vr_rank_workspace_bytes(n, nV, R)) { set_error("vr_rank_materialise: workspace %zu < %zu", ws_bytes, vr_rank_workspace_bytes(n, nV, R)); return VR_ERR_WORKSPACE; }
  RankContext c = rank_context_view(const_cast<void*>(ctx), n, R);
  return rank_stage(blobs, nV, n, c, R, false, dev, ss, nkept, nullptr, nullptr, 0, nullptr, 0, 0, ws, ws_bytes,
                    static_cast<cudaStream_t>(stream_));
}

extern "C" int vr_rank_fused(const void* const* blobs, int nF, int n, const void* ctx, int R, const int32_t* mat_dev,
                             const double* mat_ss, int nM, double* out, int64_t out_stride_fused, int64_t out_stride_mat,
                             void* ws, size_t ws_bytes, vr_stream_t stream_) {
  VR_REQUIRE(blobs && ctx && mat_dev && mat_ss && out, "vr_rank_fused: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && R >= 1 && nF >= 1 && nF <= kMaxVectors, "vr_rank_fused: bad sizes n=%d nF=%d R=%d", n, nF, R);
  VR_REQUIRE(nM >= 1 && nM <= kFuseMax, "vr_rank_fused: nM=%d outside [1, %d]", nM, kFuseMax);
  VR_REQUIRE((int64_t)nM * R * vr_dev_stride(n) < (1ll << 34), "vr_rank_fused: nM * R * stride = %lld exceeds the 2^34-element addressing window; use fewer replicates per call", (long long)nM * R * vr_dev_stride(n));
  if (!ws || ws_bytes < vr_rank_workspace_bytes(n, nF, R)) { set_error("vr_rank_fused: workspace %zu < %zu", ws_bytes, vr_rank_workspace_bytes(n, nF, R)); return VR_ERR_WORKSPACE; }
  RankContext c = rank_context_view(const_cast<void*>(ctx), n, R);
  return rank_stage(blobs, nF, n, c, R, true, nullptr, nullptr, nullptr, mat_dev, mat_ss, nM, out, out_stride_fused,
                    out_stride_mat, ws, ws_bytes, static_cast<cudaStream_t>(stream_));
}

extern "C" int vr_rank_fuse_max(void) { return kFuseMax; }

// ---- single-vector convenience (context + materialise) -------------------------------------------
extern "C" size_t vr_subsample_rank_workspace_bytes(int n, int R) {
  return (n >= 2 && R >= 1) ? rank_context_bytes(n, R) + tables_bytes(n, 1, R) + 512 : 256;
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
  Arena a(ws, ws_bytes);
  void* ctx = a.take<char>(rank_context_bytes(n, R));
  void* rest = a.take<char>(tables_bytes(n, 1, R) + 256);
  int st = build_rank_context(masks, n, R, ctx, stream);
  if (st != VR_OK) return st;
  RankContext c = rank_context_view(ctx, n, R);
  const void* blobs[1] = {presorted};
  return rank_stage(blobs, 1, n, c, R, false, dev, ss, nkept, nullptr, nullptr, 0, nullptr, 0, 0, rest, tables_bytes(n, 1, R) + 256, stream);
}

// ---- one-call grouped bootstrap ---------------------------------------------------------------------
namespace {
struct SpearmanPlan {
  size_t fixed;
  size_t per_rep;   // context + tables of the larger stage + dev/ss of one materialised group, per replicate
};
inline int mat_group(int nA, int nB) { int m = nA <= nB ? nA : nB; return m < kFuseMax ? m : kFuseMax; }
SpearmanPlan spearman_plan(int n, int nA, int nB) {
  const int group = mat_group(nA, nB);
  const int nF = nA <= nB ? nB : nA;
  const int stage = group > nF ? group : nF;
  SpearmanPlan p;
  p.fixed = 16 * 256 + (size_t)n * 4 * 2 + 4096 + (size_t)32 * stage * table_groups(n) * 8;   // tables pad replicates to 32
  p.per_rep = (size_t)vr_dev_stride(n) / 8 + (size_t)n / 8 + 8 +                      // context
              (size_t)stage * ((size_t)table_groups(n) * 8 + 4 * kCountWarps * kCountCtasMax +
                               (rank_table_in_smem(n) ? 0 : (size_t)rank_slice_plan(n).slices * (kFuseMax + 1) * 8 + 4)) +   // count tables (+ slice partials)
              (size_t)group * ((size_t)vr_dev_stride(n) * 4 + 8);                    // dev + ss
  return p;
}
constexpr size_t kSoftCap = (size_t)40 << 30;  // default upper bound on one call's workspace
}  // namespace

extern "C" size_t vr_bootstrap_spearman_workspace_bytes(int n, int nA, int nB, int R) {
  if (n < 2 || nA < 1 || nB < 1 || R < 1) return 4096;
  SpearmanPlan p = spearman_plan(n, nA, nB);
  size_t reps = (size_t)R;
  size_t cap = kSoftCap / p.per_rep;
  if (cap < 32) cap = 32;
  if (reps > cap) reps = cap;
  return p.fixed + reps * p.per_rep;
}

extern "C" int vr_bootstrap_spearman(const void* const* A, int nA, const void* const* B, int nB, int n,
                                     const uint32_t* masks, int R, double* out, void* ws, size_t ws_bytes,
                                     vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(A && B && masks && out, "vr_bootstrap_spearman: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI, "vr_bootstrap_spearman: n=%d outside [2, %d]", n, VR_MAX_STIMULI);
  VR_REQUIRE(nA >= 1 && nB >= 1 && nA + nB <= kMaxVectors, "vr_bootstrap_spearman: nA=%d nB=%d (max %d in total)", nA, nB, kMaxVectors);
  VR_REQUIRE(R >= 1, "vr_bootstrap_spearman: R=%d", R);
  // The side with fewer vectors is materialised (in groups of <= kFuseMax); the other side is fused.
  const bool mat_is_A = nA <= nB;
  const int nM = mat_is_A ? nA : nB, nF = mat_is_A ? nB : nA;
  const void* const* M = mat_is_A ? A : B;
  const void* const* F = mat_is_A ? B : A;
  const int group = mat_group(nA, nB);
  SpearmanPlan plan = spearman_plan(n, nA, nB);
  const size_t min_ws = plan.fixed + (size_t)(R < 32 ? R : 32) * plan.per_rep;
  if (!ws || ws_bytes < min_ws) {
    set_error("vr_bootstrap_spearman: workspace %zu < minimum %zu", ws_bytes, min_ws);
    return VR_ERR_WORKSPACE;
  }
  int64_t chunk = (int64_t)((ws_bytes - plan.fixed) / plan.per_rep);
  if (chunk > R) chunk = R;
  if (chunk > 65535) chunk = 65535;
  {                                                         // 32-bit int4 offsets into the materialised group
    const int64_t cap = ((1ll << 34) - 1) / ((int64_t)group * vr_dev_stride(n));
    if (chunk > cap) chunk = cap;
  }
  if (chunk < R) chunk = chunk / 32 * 32;                   // keep replicate batches of 32 whole
  if (chunk < 1) { set_error("vr_bootstrap_spearman: n too large for one replicate chunk"); return VR_ERR_UNSUPPORTED; }
  const int stage = group > nF ? group : nF;
  Arena a(ws, ws_bytes);
  void* ctx = a.take<char>(rank_context_bytes(n, chunk));
  const size_t tb = tables_bytes(n, stage, chunk) + 256;
  void* tws = a.take<char>(tb);
  double* ss = a.take<double>((size_t)group * chunk);
  int32_t* dev = a.take<int32_t>((size_t)group * chunk * vr_dev_stride(n));
  if (!a.ok) { set_error("vr_bootstrap_spearman: internal workspace plan overflow"); return VR_ERR_WORKSPACE; }
  const int W = vr_mask_words(n);
  // out is (nA, nB, R): strides of the fused / materialised index
  const int64_t stride_f = mat_is_A ? (int64_t)R : (int64_t)nB * R;
  const int64_t stride_m = mat_is_A ? (int64_t)nB * R : (int64_t)R;
  for (int64_t r0 = 0; r0 < R; r0 += chunk) {
    const int rc = (int)(R - r0 < chunk ? R - r0 : chunk);
    int st = build_rank_context(masks + (size_t)r0 * W, n, rc, ctx, stream);
    if (st != VR_OK) return st;
    RankContext c = rank_context_view(ctx, n, rc);
    for (int m0 = 0; m0 < nM; m0 += group) {
      const int mg = nM - m0 < group ? nM - m0 : group;
      st = rank_stage(M + m0, mg, n, c, rc, false, dev, ss, nullptr, nullptr, nullptr, 0, nullptr, 0, 0, tws, tb, stream);
      if (st != VR_OK) return st;
      st = rank_stage(F, nF, n, c, rc, true, nullptr, nullptr, nullptr, dev, ss, mg, out + m0 * stride_m + r0, stride_f,
                      stride_m, tws, tb, stream);
      if (st != VR_OK) return st;
    }
  }
  return VR_OK;
}
