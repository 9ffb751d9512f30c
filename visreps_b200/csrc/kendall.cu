// Kendall tau-a between two RDMs' upper triangles, for any number of subsample masks.
//
// Replaces _kendall_tau_a (visreps/analysis/rsa.py:22-40), which asks SciPy for tau-b and converts it:
//     tau_a = (C - D) / n0,  n0 = M(M-1)/2 pairs of the M kept RDM entries,
// NaN when M < 2 or when either vector is constant (SciPy's tau-b is NaN there, rsa.py:29-30, :39).
// With tx, ty = pairs tied in x / in y, txy = pairs tied in both, D = strictly discordant pairs:
//     C - D = n0 - tx - ty + txy - 2 D.
//
// Once per RDM pair:   entries are sorted by (x, y) with two stable radix sorts (by y, then by x) that
//                      carry the canonical pair index.
//                      Run starts (lower bounds in the sorted keys) for equal x, equal (x, y), equal y.
// Per batch of <= 64 replicates (grid.y = replicate, ~15 launches per batch whatever its size):
//                      the (x, y) order of the kept entries is the stable compaction of the full order
//                      (same fact the Spearman path uses), so no replicate sorts:
//                      kept tables {kept word, kept count before} of both orders -> tie sums (a kept entry adds
//                      the kept entries before it in its run) -> D = inversions of the compacted y sequence,
//                      counted by a bottom-up merge sort: 2048-entry tiles in shared memory, then global
//                      levels in which every element finds its merged position by binary search in the
//                      sibling run.
#include "common.cuh"
#include "sort.cuh"
#include "pairmask.cuh"

namespace vr {

__global__ void __launch_bounds__(256)
kendall_keys_kernel(const float* __restrict__ A, const float* __restrict__ B, int n, int64_t ld,
                    uint32_t* __restrict__ xkey, uint32_t* __restrict__ ykey, uint32_t* __restrict__ idx,
                    uint32_t* __restrict__ nan_flag) {
  const uint32_t i = blockIdx.y;
  const uint32_t j = blockIdx.x * 256u + threadIdx.x;
  if (j >= (uint32_t)n || j <= i) return;
  const float a = A[(int64_t)i * ld + j], b = B[(int64_t)i * ld + j];
  if (a != a || b != b) atomicOr(nan_flag, 1u);
  const uint32_t p = triu_pos(i, j, (uint32_t)n);
  xkey[p] = float_key(a);
  ykey[p] = float_key(b);
  idx[p] = p;
}

__global__ void __launch_bounds__(256)
gather_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ index, uint32_t N, uint32_t* __restrict__ dst) {
  const uint32_t k = blockIdx.x * 256u + threadIdx.x;
  if (k < N) dst[k] = src[index[k]];
}

// Start of the run of equal keys that holds sorted position k: the keys are sorted, so it is a lower bound
// (lexicographic on (k0, k1) when k1 is given).  Replicate-independent, once per RDM pair.
__global__ void __launch_bounds__(256)
run_start_kernel(const uint32_t* __restrict__ k0, const uint32_t* __restrict__ k1, uint32_t N, uint32_t* __restrict__ start) {
  const uint32_t k = blockIdx.x * 256u + threadIdx.x;
  if (k >= N) return;
  const uint32_t a = k0[k], b = k1 ? k1[k] : 0u;
  uint32_t lo = 0, hi = k;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t am = k0[mid];
    const bool less = am < a || (k1 && am == a && k1[mid] < b);
    if (less) lo = mid + 1; else hi = mid;
  }
  start[k] = lo;
}

// ---- per batch of replicates -------------------------------------------------------------------------
// Kept tables of a sorted order for every replicate of the batch: entry g = {kept word of positions
// [32g, 32g+32), kept count before the group}, built in three steps (group words + chunk totals, scan of
// the chunk totals, add the bases).  grid (chunks, replicates, 2 orders); a chunk is 256 groups.
constexpr uint32_t kChunkGroups = 256;

struct KendallOrders {
  const uint32_t* order[2];    // pair index at each sorted position: [0] (x, y) order, [1] y order
};

__global__ void __launch_bounds__(256)
kept_words_kernel(const KendallOrders ord, const uint32_t* __restrict__ pairbits, uint32_t words, uint32_t N, uint32_t G,
                  uint2* __restrict__ tables, uint32_t* __restrict__ chunk_tot, uint32_t nchunks) {
  __shared__ uint32_t s_warp[8];
  const int r = blockIdx.y, o = blockIdx.z;
  const uint32_t* __restrict__ order = ord.order[o];
  const uint32_t* __restrict__ pb = pairbits + (size_t)r * words;
  uint2* __restrict__ tab = tables + ((size_t)o * gridDim.y + r) * G;
  const unsigned lane = lane_id();
  const uint32_t g_base = blockIdx.x * kChunkGroups + (threadIdx.x >> 5) * 32u;     // this warp's 32 groups
  uint32_t mine = 0, tot = 0;
  for (uint32_t t = 0; t < 32u; ++t) {
    const uint32_t k = (g_base + t) * 32u + lane;
    bool kept = false;
    if (k < N) { const uint32_t p = order[k]; kept = (pb[p >> 5] >> (p & 31u)) & 1u; }
    const uint32_t word = __ballot_sync(0xffffffffu, kept);
    if (t == lane) mine = word;
    tot += __popc(word);
  }
  // exclusive prefix of the 32 group counts of this warp, relative to the warp's start
  uint32_t c = __popc(mine), incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (unsigned)d) incl += x; }
  if (g_base + lane < G) tab[g_base + lane] = make_uint2(mine, incl - c);
  if (lane == 0) s_warp[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int w = 0; w < 8; ++w) run += s_warp[w];
    chunk_tot[((size_t)o * gridDim.y + r) * nchunks + blockIdx.x] = run;
  }
}

// One CTA per (replicate, order): exclusive scan of the chunk totals in place; the grand total is the number of kept entries.
__global__ void __launch_bounds__(1024)
chunk_scan_kernel(uint32_t* __restrict__ chunk_tot, uint32_t nchunks, uint32_t* __restrict__ counts) {
  __shared__ uint32_t part[1024];
  uint32_t* c = chunk_tot + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * nchunks;
  const uint32_t per = (nchunks + 1023u) / 1024u;
  const uint32_t lo = min(nchunks, threadIdx.x * per), hi = min(nchunks, lo + per);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; ++i) sum += c[i];
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t loc[32], run = 0;
    for (int t = 0; t < 32; ++t) { loc[t] = part[threadIdx.x * 32 + t]; run += loc[t]; }
    uint32_t incl = run;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d); if (threadIdx.x >= (unsigned)d) incl += x; }
    uint32_t base = incl - run;
    for (int t = 0; t < 32; ++t) { part[threadIdx.x * 32 + t] = base; base += loc[t]; }
    if (threadIdx.x == 31) counts[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = incl;
  }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (uint32_t i = lo; i < hi; ++i) { const uint32_t t = c[i]; c[i] = run; run += t; }
}

// Adds the chunk and warp bases to the table counts; for the (x, y) order it also writes the kept y keys,
// compacted, which is the sequence whose inversions are counted.
__global__ void __launch_bounds__(256)
kept_finish_kernel(uint2* __restrict__ tables, const uint32_t* __restrict__ chunk_base, uint32_t nchunks, uint32_t G, uint32_t N,
                   const uint32_t* __restrict__ yseq, uint32_t* __restrict__ cy, int64_t cy_stride) {
  __shared__ uint32_t s_warp[9];
  const int r = blockIdx.y, o = blockIdx.z;
  uint2* __restrict__ tab = tables + ((size_t)o * gridDim.y + r) * G;
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t g = blockIdx.x * kChunkGroups + threadIdx.x;
  uint2 e = g < G ? tab[g] : make_uint2(0u, 0u);
  // total of each warp's 32 groups = last lane's count + popc of its word
  const uint32_t wtot = __shfl_sync(0xffffffffu, e.y + __popc(e.x), 31);
  if (lane == 0) s_warp[warp + 1] = wtot;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = chunk_base[((size_t)o * gridDim.y + r) * nchunks + blockIdx.x];
    for (int w = 0; w < 8; ++w) { const uint32_t t = s_warp[w + 1]; s_warp[w] = run; run += t; }
  }
  __syncthreads();
  e.y += s_warp[warp];
  if (g < G) tab[g] = e;
  if (o == 0 && g < G && cy) {
    uint32_t* __restrict__ out = cy + (int64_t)r * cy_stride;
    uint32_t w = e.x, pos = e.y;
    while (w) { const int b = __ffs(w) - 1; w &= w - 1u; out[pos++] = yseq[g * 32u + b]; }
  }
}

__device__ __forceinline__ uint32_t kept_before(const uint2* __restrict__ tab, uint32_t x) {
  const uint2 e = tab[x >> 5];
  return e.y + __popc(e.x & ((1u << (x & 31u)) - 1u));
}

// Tied pairs among the kept entries: sum over runs of c(c-1)/2 = sum over kept entries of the number of kept
// entries before them in their run.  acc[r]: [0] tx  [1] ty  [2] txy.
__global__ void __launch_bounds__(256)
tie_sums_kernel(const uint2* __restrict__ tables, uint32_t G, uint32_t N, const uint32_t* __restrict__ rs_x,
                const uint32_t* __restrict__ rs_xy, const uint32_t* __restrict__ rs_y, unsigned long long* __restrict__ acc) {
  const int r = blockIdx.y;
  const uint2* __restrict__ txy_tab = tables + (size_t)r * G;
  const uint2* __restrict__ ty_tab = tables + ((size_t)gridDim.y + r) * G;
  unsigned long long tx = 0, ty = 0, txy = 0;
  for (uint32_t k = blockIdx.x * 256u + threadIdx.x; k < N; k += gridDim.x * 256u) {
    const uint2 e = txy_tab[k >> 5];
    if ((e.x >> (k & 31u)) & 1u) {
      const uint32_t here = e.y + __popc(e.x & ((1u << (k & 31u)) - 1u));
      tx += here - kept_before(txy_tab, rs_x[k]);
      txy += here - kept_before(txy_tab, rs_xy[k]);
    }
    const uint2 f = ty_tab[k >> 5];
    if ((f.x >> (k & 31u)) & 1u)
      ty += f.y + __popc(f.x & ((1u << (k & 31u)) - 1u)) - kept_before(ty_tab, rs_y[k]);
  }
  tx = (unsigned long long)warp_sum_ll((long long)tx);
  ty = (unsigned long long)warp_sum_ll((long long)ty);
  txy = (unsigned long long)warp_sum_ll((long long)txy);
  if (lane_id() == 0) {
    if (tx) atomicAdd(acc + (size_t)r * 4 + 0, tx);
    if (ty) atomicAdd(acc + (size_t)r * 4 + 1, ty);
    if (txy) atomicAdd(acc + (size_t)r * 4 + 2, txy);
  }
}

// One merge step for element k of a sequence of sorted runs of length L: left-run elements go before equal
// right-run elements; a right-run element counts the left-run elements strictly greater than itself.
template <typename Src>
__device__ __forceinline__ uint32_t merge_slot(const Src& in, uint32_t k, uint32_t M, uint32_t L, uint32_t v, uint32_t& greater) {
  const uint32_t base = k / (2u * L) * (2u * L);
  const uint32_t a1 = min(M, base + L), b1 = min(M, base + 2u * L);
  if (k < a1) {
    uint32_t lo = a1, hi = b1;                      // lower_bound in the right run
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (in[mid] < v) lo = mid + 1; else hi = mid; }
    greater = 0;
    return k + (lo - a1);
  }
  uint32_t lo = base, hi = a1;                      // upper_bound in the left run
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (in[mid] <= v) lo = mid + 1; else hi = mid; }
  greater = a1 - lo;
  return base + (k - a1) + (lo - base);
}

// Levels 1 .. kTile/2 of the bottom-up merge inside shared memory: every CTA sorts one tile of the kept y
// sequence and counts the inversions inside it.
constexpr uint32_t kTile = 2048;

__global__ void __launch_bounds__(256)
tile_sort_kernel(const uint32_t* __restrict__ cy, uint32_t* __restrict__ out, int64_t stride, const uint32_t* __restrict__ counts,
                 unsigned long long* __restrict__ acc) {
  __shared__ uint32_t buf[2][kTile];
  const int r = blockIdx.y;
  const uint32_t M = counts[r];
  const uint32_t base = blockIdx.x * kTile;
  if (base >= M) return;
  const uint32_t len = min(kTile, M - base);
  const uint32_t* __restrict__ in = cy + (int64_t)r * stride + base;
  for (uint32_t k = threadIdx.x; k < kTile; k += 256u) buf[0][k] = k < len ? in[k] : 0xFFFFFFFFu;
  __syncthreads();
  unsigned long long inv = 0;
  int src = 0;
  for (uint32_t L = 1; L < kTile; L <<= 1) {
    const uint32_t* s = buf[src];
    uint32_t* d = buf[src ^ 1];
    for (uint32_t k = threadIdx.x; k < kTile; k += 256u) {
      const uint32_t v = s[k];
      uint32_t greater;
      const uint32_t slot = merge_slot(s, k, kTile, L, v, greater);
      d[slot] = v;
      inv += greater;
    }
    __syncthreads();
    src ^= 1;
  }
  uint32_t* __restrict__ o = out + (int64_t)r * stride + base;
  for (uint32_t k = threadIdx.x; k < len; k += 256u) o[k] = buf[src][k];
  inv = (unsigned long long)warp_sum_ll((long long)inv);
  if (lane_id() == 0 && inv) atomicAdd(acc + (size_t)r * 4 + 3, inv);
}

// One level of the merge in global memory (runs of length L >= kTile), all replicates of the batch.
__global__ void __launch_bounds__(256)
merge_level_kernel(const uint32_t* __restrict__ in_all, uint32_t* __restrict__ out_all, int64_t stride,
                   const uint32_t* __restrict__ counts, uint32_t L, unsigned long long* __restrict__ acc) {
  const int r = blockIdx.y;
  const uint32_t M = counts[r];
  const uint32_t* __restrict__ in = in_all + (int64_t)r * stride;
  uint32_t* __restrict__ out = out_all + (int64_t)r * stride;
  unsigned long long local = 0;
  for (uint32_t k = blockIdx.x * 256u + threadIdx.x; k < M; k += gridDim.x * 256u) {
    const uint32_t v = in[k];
    uint32_t greater;
    const uint32_t slot = merge_slot(in, k, M, L, v, greater);
    out[slot] = v;
    local += greater;
  }
  local = (unsigned long long)warp_sum_ll((long long)local);
  if (lane_id() == 0 && local) atomicAdd(acc + (size_t)r * 4 + 3, local);
}

// ---- complement mode -------------------------------------------------------------------------------------------
// A 90 % subsample keeps 81 % of the entries: instead of counting ties and inversions among the ~404 k kept entries of every
// replicate, count them among the ~95 k DROPPED ones and use, for each relation Q in {tied in x, tied in y, tied in both,
// discordant}:      Q(kept) = Q(all) - sum over dropped entries e of q(e) + Q(dropped)
// where q(e) = number of entries in relation Q with e (replicate independent, once per RDM pair): a pair with exactly one dropped
// member is removed once, a pair with two dropped members twice and added back once.  The sum over the dropped entries is taken
// through the dropped STIMULI (as the Pearson path does):  sum_{i in D} rowsum_q(i) - sum_{i<j in D} q({i, j}).
// The merge sort that counts inversions then runs on a sequence 4x shorter (83 % of the time of the direct path).

// q[order[k]] = size of the run of equal keys holding sorted position k, minus one (lexicographic on (k0, k1) when k1 is given).
__global__ void __launch_bounds__(256)
run_ties_kernel(const uint32_t* __restrict__ k0, const uint32_t* __restrict__ k1, const uint32_t* __restrict__ start,
                const uint32_t* __restrict__ order, uint32_t N, uint32_t* __restrict__ q) {
  const uint32_t k = blockIdx.x * 256u + threadIdx.x;
  if (k >= N) return;
  const uint32_t a = k0[k], b = k1 ? k1[k] : 0u;
  uint32_t lo = k + 1u, hi = N;                       // first position whose key is greater
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t am = k0[mid];
    const bool greater = am > a || (k1 && am == a && k1[mid] > b);
    if (greater) hi = mid; else lo = mid + 1;
  }
  q[order[k]] = lo - start[k] - 1u;
}

// Bottom-up merge sort of (value, entry) pairs that adds, for every element and level, the elements of the sibling run it is
// inverted with: disc[entry] = number of entries discordant with it.  Same tiling as the counting kernels below.
__global__ void __launch_bounds__(256)
tile_sort_entries_kernel(const uint32_t* __restrict__ vin, const uint32_t* __restrict__ ein, uint32_t N,
                         uint32_t* __restrict__ vout, uint32_t* __restrict__ eout, uint32_t* __restrict__ disc) {
  constexpr uint32_t kT = 2048;
  __shared__ uint32_t val[2][kT];
  __shared__ uint32_t ent[2][kT];
  const uint32_t base = blockIdx.x * kT;
  if (base >= N) return;
  const uint32_t len = min(kT, N - base);
  for (uint32_t k = threadIdx.x; k < kT; k += 256u) {
    val[0][k] = k < len ? vin[base + k] : 0xFFFFFFFFu;
    ent[0][k] = k < len ? ein[base + k] : 0xFFFFFFFFu;
  }
  __syncthreads();
  int src = 0;
  // every element has one owner per level and levels are separated by __syncthreads(): plain read-modify-writes of disc[] do not race
  for (uint32_t L = 1; L < kT; L <<= 1) {
    const uint32_t* sv = val[src];
    const uint32_t* se = ent[src];
    uint32_t* dv = val[src ^ 1];
    uint32_t* de = ent[src ^ 1];
    for (uint32_t k = threadIdx.x; k < kT; k += 256u) {
      const uint32_t v = sv[k], e = se[k];
      const uint32_t blk = k / (2u * L) * (2u * L);
      const uint32_t a1 = min(kT, blk + L), b1 = min(kT, blk + 2u * L);
      uint32_t slot, inv;
      if (k < a1) {                                                          // left run: right elements strictly smaller
        uint32_t lo = a1, hi = b1;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sv[mid] < v) lo = mid + 1; else hi = mid; }
        inv = lo - a1;
        slot = k + (lo - a1);
      } else {                                                               // right run: left elements strictly greater
        uint32_t lo = blk, hi = a1;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sv[mid] <= v) lo = mid + 1; else hi = mid; }
        inv = a1 - lo;
        slot = blk + (k - a1) + (lo - blk);
      }
      dv[slot] = v;
      de[slot] = e;
      if (inv && e != 0xFFFFFFFFu) disc[e] += inv;                           // padding never counts: it is the largest value and sits last
    }
    __syncthreads();
    src ^= 1;
  }
  for (uint32_t k = threadIdx.x; k < len; k += 256u) { vout[base + k] = val[src][k]; eout[base + k] = ent[src][k]; }
}
__global__ void __launch_bounds__(256)
merge_level_entries_kernel(const uint32_t* __restrict__ vin, const uint32_t* __restrict__ ein, uint32_t N, uint32_t L,
                           uint32_t* __restrict__ vout, uint32_t* __restrict__ eout, uint32_t* __restrict__ disc) {
  for (uint32_t k = blockIdx.x * 256u + threadIdx.x; k < N; k += gridDim.x * 256u) {
    const uint32_t v = vin[k], e = ein[k];
    const uint32_t blk = k / (2u * L) * (2u * L);
    const uint32_t a1 = min(N, blk + L), b1 = min(N, blk + 2u * L);
    uint32_t slot, inv;
    if (k < a1) {
      uint32_t lo = a1, hi = b1;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (vin[mid] < v) lo = mid + 1; else hi = mid; }
      inv = lo - a1;
      slot = k + (lo - a1);
    } else {
      uint32_t lo = blk, hi = a1;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (vin[mid] <= v) lo = mid + 1; else hi = mid; }
      inv = a1 - lo;
      slot = blk + (k - a1) + (lo - blk);
    }
    vout[slot] = v;
    eout[slot] = e;
    if (inv) disc[e] += inv;
  }
}

// rows[q][i] = sum over b != i of q({i, b}).  grid (n, 4), 256 threads.
struct EntryCounts { const uint32_t* q[4]; };      // per canonical entry: tied in x, tied in y, tied in both, discordant
__global__ void __launch_bounds__(256)
entry_row_sums_kernel(const EntryCounts Q, int n, unsigned long long* __restrict__ rows) {
  __shared__ unsigned long long red[8];
  const uint32_t i = blockIdx.x;
  const uint32_t* __restrict__ q = Q.q[blockIdx.y];
  unsigned long long acc = 0;
  for (uint32_t b = threadIdx.x; b < (uint32_t)n; b += 256u)
    if (b != i) acc += q[b < i ? triu_pos(b, i, (uint32_t)n) : triu_pos(i, b, (uint32_t)n)];
  acc = (unsigned long long)warp_sum_ll((long long)acc);
  if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    rows[(size_t)blockIdx.y * n + i] = t;
  }
}
// totals[q] = sum_i rows[q][i] = 4 Q(all).  grid 4, 256 threads.
__global__ void __launch_bounds__(256)
entry_totals_kernel(const unsigned long long* __restrict__ rows, int n, unsigned long long* __restrict__ totals) {
  __shared__ unsigned long long red[8];
  unsigned long long acc = 0;
  for (int i = threadIdx.x; i < n; i += 256) acc += rows[(size_t)blockIdx.x * n + i];
  acc = (unsigned long long)warp_sum_ll((long long)acc);
  if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    totals[blockIdx.x] = t;
  }
}
// pairbits <- the DROPPED canonical pairs (positions >= N stay 0)
__global__ void complement_bits_kernel(uint32_t* __restrict__ pairbits, uint32_t words, uint32_t N, int Rb) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words) return;
  uint32_t valid = 0xFFFFFFFFu;
  if (w * 32u + 32u > N) valid = w * 32u >= N ? 0u : ((1u << (N - w * 32u)) - 1u);
  for (int r = 0; r < Rb; ++r) pairbits[(size_t)r * words + w] = ~pairbits[(size_t)r * words + w] & valid;
}
// sub[r][q] = sum over the dropped entries of replicate r of q(e) = sum_{i in D} rows[q][i] - sum_{i<j in D} q({i, j}).
// grid (Rb), 256 threads; `drop` is scratch of Rb x n stimulus indices.
__global__ void __launch_bounds__(256)
dropped_entry_sums_kernel(const uint32_t* __restrict__ masks, int W, int n, const EntryCounts Q, const unsigned long long* __restrict__ rows,
                          uint32_t* __restrict__ drop, unsigned long long* __restrict__ sub) {
  __shared__ int s_cnt;
  __shared__ unsigned long long red[4][8];
  const int r = blockIdx.x;
  const uint32_t* __restrict__ mask = masks + (size_t)r * W;
  uint32_t* __restrict__ list = drop + (size_t)r * n;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 256) {                                      // dropped stimuli, ascending inside every 256-block
    const int i = i0 + (int)threadIdx.x;
    const bool d = i < n && !((mask[i >> 5] >> (i & 31)) & 1u);
    const unsigned bal = __ballot_sync(0xffffffffu, d);
    int base = 0;
    if (lane_id() == 0 && bal) base = atomicAdd(&s_cnt, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (d) list[base + __popc(bal & lanemask_lt())] = (uint32_t)i;
    __syncthreads();
  }
  const uint32_t nd = (uint32_t)s_cnt;
  long long acc[4] = {0, 0, 0, 0};
  for (uint32_t t = threadIdx.x; t < nd; t += 256u) {
    const uint32_t i = list[t];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] += (long long)rows[(size_t)q * n + i];
  }
  const uint32_t npairs = nd * (nd - 1u) / 2u;                                // nd <= 32767: fits
  for (uint32_t t = threadIdx.x; t < npairs && nd >= 2u; t += 256u) {
    const uint32_t a = triu_row(t, nd);
    const uint32_t b = t - a * (2u * nd - a - 1u) / 2u + a + 1u;
    uint32_t i = list[a], j = list[b];
    if (i > j) { const uint32_t x = i; i = j; j = x; }
    const uint32_t p = triu_pos(i, j, (uint32_t)n);
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] -= (long long)Q.q[q][p];
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long long v = warp_sum_ll(acc[q]);
    if (lane_id() == 0) red[q][threadIdx.x >> 5] = (unsigned long long)v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    sub[(size_t)r * 4 + threadIdx.x] = t;
  }
}
// ---- tree mode ---------------------------------------------------------------------------------------------------
// The bottom-up merge sort of the y keys in (x, y) order has the same shape for every replicate: the sorted order of a tree
// node's kept members is a subsequence of the node's full sorted order.  So the tree is built ONCE per RDM pair -- for every
// level the entry (as its stimulus pair) at each merged position and one bit telling which child it came from -- and a replicate
// never sorts: its discordant pairs at a node are the pairs (right-child member, left-child member), both kept, with the right
// one first in the node's merged order (a left member goes before an equal right member, so ties in y never count).
// Replicates ride in the lanes (lane = replicate, 32 per batch): the kept bits of 32 merged positions are one AND of two
// transposed-mask words per position and a 32 x 32 bit transpose; per 32 positions and replicate the count is
//     (kept right members so far in the node) * popc(L) + cross(R, L),      cross(R, L) = sum_{i<j} R_i L_j,
// cross = a byte-prefix dot product (dp4a) + four lookups of a 256 x 256 table.  Inside the base blocks of 32 positions the
// discordant partners of every position are a precomputed 32-bit mask.  A warp takes a segment of 2048 positions of one level;
// nodes longer than a segment are stitched from the segments' (right, left) totals by a second small kernel.
constexpr uint32_t kTreeSegChunks = 64;                         // 32-position chunks per warp segment
constexpr uint32_t kTiesSegChunks = 8;                          // the tie passes need no stitching: short segments, many warps
constexpr int kTreeBatch = 256;                                 // replicates per pass: eight lane batches
constexpr int kTreeMinReplicates = 192;                         // below this the once-per-pair tree costs more than it saves
constexpr int kTreeMaxStimuli = 16384;                          // transposed masks of a batch live in shared memory

__device__ __forceinline__ uint32_t tree_pack_pair(uint32_t p, uint32_t n) {
  const uint32_t i = triu_row(p, n);
  const uint32_t j = p - i * (2u * n - i - 1u) / 2u + i + 1u;
  return (i << 16) | j;
}
__device__ __forceinline__ uint32_t tree_bit_transpose32(uint32_t w, unsigned lane) {      // row = lane, column = bit
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t x = __shfl_xor_sync(0xffffffffu, w, s);
    w = (lane & s) ? ((w & ~m) | ((x >> s) & m)) : ((w & m) | ((x << s) & ~m));
  }
  return w;
}

// One warp per block of 32 positions of the (x, y) order: the discordant-partner mask of every position (earlier positions of
// the block with a strictly greater y), the block sorted by y (stable) as the input of the first merge level, and the
// stimulus pair of every position.
__global__ void __launch_bounds__(256)
tree_base_kernel(const uint32_t* __restrict__ yseq, const uint32_t* __restrict__ order_xy, uint32_t N, uint32_t G, uint32_t n,
                 uint32_t* __restrict__ gt, uint32_t* __restrict__ vals, uint32_t* __restrict__ pos, uint32_t* __restrict__ ij0) {
  const uint32_t g = blockIdx.x * 8u + (threadIdx.x >> 5);
  if (g >= G) return;
  const unsigned lane = lane_id();
  const uint32_t k = g * 32u + lane;
  const uint32_t y = k < N ? yseq[k] : 0xFFFFFFFFu;
  uint32_t mask = 0, rank = 0;
#pragma unroll
  for (int a = 0; a < 32; ++a) {
    const uint32_t ya = __shfl_sync(0xffffffffu, y, a);
    if ((unsigned)a < lane && ya > y) mask |= 1u << a;
    rank += (ya < y || (ya == y && (unsigned)a < lane)) ? 1u : 0u;
  }
  gt[k] = mask;
  vals[g * 32u + rank] = y;                                      // padding (the largest key, the last indices) stays behind N
  pos[g * 32u + rank] = k;
  ij0[k] = k < N ? tree_pack_pair(order_xy[k], n) : 0u;
}

// One merge level over the whole sequence (runs of length L -> 2L), keeping what the replicates need: the stimulus pair at every
// merged position and the child bit.  `side` must be zero on entry.
__global__ void __launch_bounds__(256)
tree_merge_kernel(const uint32_t* __restrict__ vin, const uint32_t* __restrict__ pin, uint32_t N, uint32_t L,
                  uint32_t* __restrict__ vout, uint32_t* __restrict__ pout, const uint32_t* __restrict__ ij0,
                  uint32_t* __restrict__ ij, uint32_t* __restrict__ side) {
  for (uint32_t k = blockIdx.x * 256u + threadIdx.x; k < N; k += gridDim.x * 256u) {
    const uint32_t v = vin[k], p = pin[k];
    const uint32_t blk = k / (2u * L) * (2u * L);
    const uint32_t a1 = min(N, blk + L), b1 = min(N, blk + 2u * L);
    uint32_t slot;
    if (k < a1) {
      uint32_t lo = a1, hi = b1;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (vin[mid] < v) lo = mid + 1; else hi = mid; }
      slot = k + (lo - a1);
    } else {
      uint32_t lo = blk, hi = a1;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (vin[mid] <= v) lo = mid + 1; else hi = mid; }
      slot = blk + (k - a1) + (lo - blk);
      atomicOr(side + (slot >> 5), 1u << (slot & 31u));
    }
    vout[slot] = v;
    pout[slot] = p;
    ij[slot] = ij0[p];
  }
}

// cross[(r << 8) | l] = number of (i < j) with bit i of r and bit j of l set.
__global__ void tree_cross_table_kernel(uint8_t* __restrict__ cross) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 65536u) return;
  const uint32_t r = idx >> 8, l = idx & 255u;
  uint32_t c = 0;
  for (uint32_t j = 1; j < 8u; ++j) if ((l >> j) & 1u) c += __popc(r & ((1u << j) - 1u));
  cross[idx] = (uint8_t)c;
}

// counts[r] = entries kept by replicate r = pairs of its kept stimuli.  One warp per replicate.
__global__ void __launch_bounds__(256)
tree_kept_kernel(const uint32_t* __restrict__ masks, int W, int R, uint32_t* __restrict__ counts) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  long long m = 0;
  for (int w = lane_id(); w < W; w += 32) m += __popc(masks[(size_t)r * W + w]);
  m = warp_sum_ll(m);
  if (lane_id() == 0) counts[r] = (uint32_t)(m * (m - 1) / 2);
}

// mT[b][i]: bit r = stimulus i kept by replicate 32 b + r of the batch.  grid (ceil(n / 256), batches of 32).
__global__ void __launch_bounds__(256)
tree_masks_kernel(const uint32_t* __restrict__ masks, int W, int n, int R, uint32_t* __restrict__ mT) {
  const int i = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (i >= n) return;
  uint32_t w = 0;
  for (int r = 0; r < 32 && b * 32 + r < R; ++r) w |= ((masks[(size_t)(b * 32 + r) * W + (i >> 5)] >> (i & 31)) & 1u) << r;
  mT[(size_t)b * n + i] = w;
}

struct TreeArgs {
  const uint32_t* ij;          // [levels + 1][Np]: (i << 16 | j) of the entry at every position; [0] = the (x, y) order itself
  const uint32_t* side;        // [levels][G]: bit = the position's entry came from the right child
  const uint32_t* gt;          // [Np]
  const uint8_t* cross;        // [65536]
  const uint32_t* mT;          // [batches][n]
  uint32_t* segtot;            // [batches][levels][nseg][2][32]: kept right / left members of a segment (levels with long nodes)
  unsigned long long* acc;     // [rb][4], [3] = discordant pairs
  uint32_t N, Np, G, nseg;
  int levels, n, rb;
};
__device__ __forceinline__ uint32_t byte_counts(uint32_t x) {      // popcount of every byte
  x = x - ((x >> 1) & 0x55555555u);
  x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
  return (x + (x >> 4)) & 0x0F0F0F0Fu;
}

// grid (ceil((levels + 1) * nseg / 16), batches), 512 threads (16 warps share one 64 KB cross table: three CTAs per SM);
// dynamic shared memory: the cross table and the batch's masks.
constexpr int kTreeCountWarps = 16;
__global__ void __launch_bounds__(kTreeCountWarps * 32)
tree_count_kernel(const __grid_constant__ TreeArgs A) {
  extern __shared__ __align__(16) unsigned char t_smem[];
  __shared__ unsigned long long s_red[kTreeCountWarps][32];
  uint8_t* s_cross = t_smem;
  uint32_t* s_mT = reinterpret_cast<uint32_t*>(t_smem + 65536);
  const int batch = blockIdx.y;
  for (uint32_t q = threadIdx.x; q < 4096u; q += kTreeCountWarps * 32u)
    reinterpret_cast<uint4*>(s_cross)[q] = __ldg(reinterpret_cast<const uint4*>(A.cross) + q);
  for (int i = threadIdx.x; i < A.n; i += kTreeCountWarps * 32) s_mT[i] = A.mT[(size_t)batch * A.n + i];
  __syncthreads();
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t seg = blockIdx.x * (uint32_t)kTreeCountWarps + warp;
  const uint32_t lv = seg / A.nseg, s = seg % A.nseg;
  uint32_t inv = 0;
  if (lv <= (uint32_t)A.levels) {
    const uint32_t* __restrict__ ijL = A.ij + (size_t)lv * A.Np;
    const uint32_t ch0 = s * kTreeSegChunks;
    auto kept_bits = [&](uint32_t ch) {          // lane r: bit p = position 32 ch + p kept by replicate r
      const uint32_t k = ch * 32u + lane;
      uint32_t w = 0;
      if (k < A.N) { const uint32_t ij = __ldg(ijL + k); w = s_mT[ij >> 16] & s_mT[ij & 0xFFFFu]; }
      return tree_bit_transpose32(w, lane);
    };
    if (lv == 0) {
      for (uint32_t c = 0; c < kTreeSegChunks && ch0 + c < A.G; ++c) {
        const uint32_t kw = kept_bits(ch0 + c);
        const uint32_t g = __ldg(A.gt + (size_t)(ch0 + c) * 32u + lane);
#pragma unroll
        for (int l = 1; l < 32; ++l) {
          const uint32_t gl = __shfl_sync(0xffffffffu, g, l);
          if ((kw >> l) & 1u) inv += __popc(gl & kw);
        }
      }
    } else {
      const uint32_t t = lv - 1u;
      const uint32_t cpn = 2u << t;                                               // chunks per node: 2 L / 32
      const bool big = cpn > kTreeSegChunks;
      const uint32_t* __restrict__ sideL = A.side + (size_t)t * A.G;
      uint32_t carry = 0, ltot = 0;
      for (uint32_t c = 0; c < kTreeSegChunks && ch0 + c < A.G; ++c) {
        if (!big && (c & (cpn - 1u)) == 0u) carry = 0;
        const uint32_t kw = kept_bits(ch0 + c);
        const uint32_t sd = __ldg(sideL + ch0 + c);
        const uint32_t R = kw & sd, Lm = kw & ~sd;
        const uint32_t cr = byte_counts(R), cl = byte_counts(Lm);
        const uint32_t pl = (cl * 0x01010101u) >> 24;
        uint32_t x = __dp4a(cr * 0x01010100u, cl, 0u);                           // right members in lower bytes x left members of a byte
#pragma unroll
        for (int q = 0; q < 4; ++q) x += s_cross[(((R >> (8 * q)) & 255u) << 8) | ((Lm >> (8 * q)) & 255u)];
        inv += carry * pl + x;
        carry += (cr * 0x01010101u) >> 24;
        ltot += pl;
      }
      if (big) {
        uint32_t* o = A.segtot + ((((size_t)batch * A.levels + t) * A.nseg + s) * 2) * 32;
        o[lane] = carry;
        o[32 + lane] = ltot;
      }
    }
  }
  s_red[warp][lane] = inv;
  __syncthreads();
  if (warp == 0) {
    unsigned long long tot = 0;
#pragma unroll
    for (int w = 0; w < kTreeCountWarps; ++w) tot += s_red[w][lane];
    const int r = batch * 32 + (int)lane;
    if (tot && r < A.rb) atomicAdd(A.acc + (size_t)r * 4 + 3, tot);
  }
}

// Nodes longer than a segment: kept right members of the node's earlier segments x kept left members of each segment.
// grid (levels, batches), 256 threads, lanes = replicates.  A level with eight or more nodes gives every warp whole nodes;
// the few long nodes at the top are cut into eight consecutive parts, one per warp, joined through shared memory.
__global__ void __launch_bounds__(256)
tree_stitch_kernel(const __grid_constant__ TreeArgs A) {
  __shared__ unsigned long long s_r[8][32], s_l[8][32];
  const uint32_t t = blockIdx.x;
  const int batch = blockIdx.y;
  const uint32_t cpn = 2u << t;
  if (cpn <= kTreeSegChunks) return;
  const uint32_t spn = cpn / kTreeSegChunks;                                        // segments per node
  const uint32_t nodes = (A.nseg + spn - 1u) / spn;
  const unsigned lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t* __restrict__ tot = A.segtot + (((size_t)batch * A.levels + t) * A.nseg) * 64;
  unsigned long long inv = 0;
  if (nodes >= 8u) {
    for (uint32_t node = warp; node < nodes; node += 8u) {
      unsigned long long before = 0;
      const uint32_t s1 = min(A.nseg, (node + 1u) * spn);
      for (uint32_t s = node * spn; s < s1; ++s) {
        inv += before * tot[(size_t)s * 64 + 32 + lane];
        before += tot[(size_t)s * 64 + lane];
      }
    }
  } else {
    for (uint32_t node = 0; node < nodes; ++node) {
      const uint32_t s0 = node * spn, s1 = min(A.nseg, s0 + spn);
      const uint32_t per = (s1 - s0 + 7u) / 8u;
      const uint32_t lo = min(s1, s0 + warp * per), hi = min(s1, lo + per);
      unsigned long long before = 0, left = 0;
      for (uint32_t s = lo; s < hi; ++s) {
        const unsigned long long l = tot[(size_t)s * 64 + 32 + lane];
        inv += before * l;
        left += l;
        before += tot[(size_t)s * 64 + lane];
      }
      __syncthreads();
      s_r[warp][lane] = before;
      s_l[warp][lane] = left;
      __syncthreads();
      unsigned long long earlier = 0;
      for (uint32_t w = 0; w < warp; ++w) earlier += s_r[w][lane];
      inv += earlier * left;
    }
  }
  const int r = batch * 32 + (int)lane;
  if (inv && r < A.rb) atomicAdd(A.acc + (size_t)r * 4 + 3, inv);
}

// Run heads of a sorted order as bits: position k starts a run when its run start is k itself.
__global__ void __launch_bounds__(256)
tree_heads_kernel(const uint32_t* __restrict__ run_start, uint32_t N, uint32_t G, uint32_t* __restrict__ heads) {
  const uint32_t k = blockIdx.x * 256u + threadIdx.x;
  const bool head = k >= N || run_start[k] == k;                       // padding: runs of one
  const uint32_t w = __ballot_sync(0xffffffffu, head);
  if ((threadIdx.x & 31u) == 0u && (k >> 5) < G) heads[k >> 5] = w;
}
__global__ void __launch_bounds__(256)
tree_pairs_kernel(const uint32_t* __restrict__ order, uint32_t N, uint32_t Np, uint32_t n, uint32_t* __restrict__ ij) {
  const uint32_t k = blockIdx.x * 256u + threadIdx.x;
  if (k < Np) ij[k] = k < N ? tree_pack_pair(order[k], n) : 0u;
}

// Tied pairs among the kept entries, lanes = replicates: sum over the runs of equal keys of t (t - 1) / 2, t = members kept.
// Pass 0 walks the (x, y) order once for the runs of equal x and of equal (x, y), pass 1 the y order for the runs of equal y.
// A warp takes a segment of 256 positions and every run that STARTS inside it (following a run that leaves the segment to its
// end); runs of one member cost nothing, chunks made of them are skipped without looking at the masks.
struct TiesArgs {
  const uint32_t* ij[2];       // stimulus pairs along the (x, y) order and the y order
  const uint32_t* heads[3];    // run-head bits, G words: equal x, equal (x, y), equal y
  const uint32_t* mT;
  unsigned long long* acc;     // [rb][4]: [0] tx [1] ty [2] txy
  uint32_t N, G, nseg;
  int n, rb;
};
struct RunState {
  uint32_t open = 0;           // kept members so far of the run that is open at the end of the previous chunk
  bool mine = false;           // ... and whether that run started inside this warp's segment
  unsigned long long ties = 0;
};
// One chunk of 32 positions: h = its run heads, kw = this replicate's kept bits, inside = the chunk lies in the warp's segment.
__device__ __forceinline__ void run_step(RunState& st, uint32_t h, uint32_t kw, bool inside) {
  if (!inside && !st.mine) return;
  const uint32_t first = h ? (uint32_t)__ffs(h) - 1u : 32u;                   // positions before the first head continue the open run
  st.open += __popc(kw & (first >= 32u ? 0xFFFFFFFFu : ((1u << first) - 1u)));
  if (h == 0u) return;                                                         // the run goes on
  if (st.mine) st.ties += (unsigned long long)st.open * (st.open - 1u) / 2u;
  st.mine = false;
  if (!inside) return;                                                         // runs that start here belong to the next segment
  // runs that start in this chunk: all but the last end here too; only heads whose next position is not a head start a
  // run of two or more members
  const uint32_t last = 31u - (uint32_t)__clz(h);
  uint32_t multi = h & ~(h >> 1) & ~(1u << last);
  while (multi) {
    const uint32_t b = (uint32_t)__ffs(multi) - 1u;
    multi &= multi - 1u;
    const uint32_t e = (uint32_t)__ffs(h & ~((2u << b) - 1u)) - 1u;            // the next head: b is not the last one
    const uint32_t t = __popc(kw & ((1u << e) - (1u << b)));
    st.ties += t * (t - 1u) / 2u;
  }
  st.open = __popc(kw & ~((1u << last) - 1u));
  st.mine = true;
}
// grid (ceil(2 nseg / 8), batches), 256 threads; dynamic shared memory: the batch's transposed masks
__global__ void __launch_bounds__(256)
tree_ties_kernel(const __grid_constant__ TiesArgs A) {
  extern __shared__ __align__(16) uint32_t tie_mT[];
  const int batch = blockIdx.y;
  for (int i = threadIdx.x; i < A.n; i += 256) tie_mT[i] = A.mT[(size_t)batch * A.n + i];
  __syncthreads();
  const unsigned lane = lane_id();
  const uint32_t seg = blockIdx.x * 8u + (threadIdx.x >> 5);
  const uint32_t pass = seg / A.nseg, s = seg % A.nseg;
  if (pass >= 2u) return;
  const uint32_t* __restrict__ ijL = A.ij[pass];
  const uint32_t* __restrict__ H0 = A.heads[pass == 0u ? 0 : 2];                // the coarser runs of the pass
  const uint32_t* __restrict__ H1 = A.heads[1];                                 // pass 0 only: equal (x, y), a refinement of equal x
  const uint32_t ch0 = s * kTiesSegChunks, ch1 = min(A.G, ch0 + kTiesSegChunks);
  RunState coarse, fine;
  for (uint32_t ch = ch0; ch < A.G; ++ch) {
    const bool inside = ch < ch1;
    if (!inside && !coarse.mine && !fine.mine) break;                           // past the segment: only to finish runs of ours
    const uint32_t h0 = __ldg(H0 + ch);
    // 32 runs of one member, and the next position starts a run too: nothing to count, no kept bits needed
    const bool trivial = h0 == 0xFFFFFFFFu && (ch + 1u >= A.G || (__ldg(H0 + ch + 1u) & 1u));
    uint32_t kw = 0;
    if (!trivial) {
      const uint32_t k = ch * 32u + lane;
      uint32_t w = 0;
      if (k < A.N) { const uint32_t ij = __ldg(ijL + k); w = tie_mT[ij >> 16] & tie_mT[ij & 0xFFFFu]; }
      kw = tree_bit_transpose32(w, lane);
    }
    run_step(coarse, h0, kw, inside);
    if (pass == 0u) run_step(fine, trivial ? 0xFFFFFFFFu : __ldg(H1 + ch), kw, inside);
  }
  if (coarse.mine) coarse.ties += (unsigned long long)coarse.open * (coarse.open - 1u) / 2u;     // the order's last run
  if (fine.mine) fine.ties += (unsigned long long)fine.open * (fine.open - 1u) / 2u;
  const int r = batch * 32 + (int)lane;
  if (r < A.rb) {
    if (coarse.ties) atomicAdd(A.acc + (size_t)r * 4 + (pass == 0u ? 0 : 1), coarse.ties);
    if (fine.ties) atomicAdd(A.acc + (size_t)r * 4 + 2, fine.ties);
  }
}

static int tree_levels(int64_t N) { int l = 0; for (int64_t L = 32; L < N; L <<= 1) ++l; return l; }

// tau-a from the complement: acc[r] / counts[r] describe the DROPPED entries of replicate r.
__global__ void kendall_finish_complement_kernel(const unsigned long long* __restrict__ acc, const uint32_t* __restrict__ counts,
                                                 const unsigned long long* __restrict__ totals, const unsigned long long* __restrict__ sub,
                                                 unsigned long long N, const uint32_t* __restrict__ nan_flag, int Rb, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Rb) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const unsigned long long M = N - counts[r];
  if (M < 2 || *nan_flag) { out[r] = nan; return; }
  const unsigned long long n0 = M * (M - 1) / 2;
  unsigned long long q[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = totals[k] / 4ull - sub[(size_t)r * 4 + k] + acc[(size_t)r * 4 + k];
  if (q[0] == n0 || q[1] == n0) { out[r] = nan; return; }
  const long long cd = (long long)n0 - (long long)q[0] - (long long)q[1] + (long long)q[2] - 2ll * (long long)q[3];
  out[r] = (double)cd / (double)n0;
}

// acc[r]: [0] tx  [1] ty  [2] txy  [3] D ; counts[r] = kept entries
__global__ void kendall_finish_kernel(const unsigned long long* __restrict__ acc, const uint32_t* __restrict__ counts,
                                      const uint32_t* __restrict__ nan_flag, int Rb, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Rb) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const unsigned long long M = counts[r];
  if (M < 2 || *nan_flag) { out[r] = nan; return; }
  const unsigned long long n0 = M * (M - 1) / 2;
  const unsigned long long tx = acc[r * 4 + 0], ty = acc[r * 4 + 1], txy = acc[r * 4 + 2], D = acc[r * 4 + 3];
  if (tx == n0 || ty == n0) { out[r] = nan; return; }          // a constant vector: tau-b undefined (rsa.py:29-30)
  const long long cd = (long long)n0 - (long long)tx - (long long)ty + (long long)txy - 2ll * (long long)D;
  out[r] = (double)cd / (double)n0;
}

// workspace: what one RDM pair needs once, and what every replicate of a batch adds
static size_t kendall_fixed_bytes(int n) {
  const size_t Np = (size_t)round_up(triu_len(n), 32);
  // xkey_c, ykey_c, 4 sort buffers, order_y, ysorted, order_xy, xseq, yseq, 3 run-start arrays
  // + complement mode: four per-entry count arrays, 4 n row sums, totals
  return 18 * arena_bytes(Np, 4) + arena_bytes(sort_count_elems(triu_len(n), 1), 4) + arena_bytes((size_t)4 * n, 8) + 4096;
}
static size_t kendall_per_replicate_bytes(int n) {
  const size_t Np = (size_t)round_up(triu_len(n), 32);
  const size_t G = Np / 32, nchunks = (G + kChunkGroups - 1) / kChunkGroups;
  // two y sequences (ping-pong), two kept tables, pair bits, two chunk-total rows, accumulators and counts
  return 2 * Np * 4 + 2 * G * 8 + G * 4 + 2 * nchunks * 4 + 4 * 8 + 2 * 4 + (size_t)n * 4 + 4 * 8;      // + dropped list, sums over the dropped entries
}
constexpr size_t kKendallSlack = 16384;       // 256-byte alignment of the per-batch arrays
constexpr int kKendallMaxBatch = 64;
// tree mode: stimulus pairs of every level (+ the base order), child bits, base-block masks, the cross table, the transposed
// masks and the segment totals of the eight 32-replicate batches of a pass, its accumulators and kept-entry counts
static size_t kendall_tree_bytes(int n) {
  const int64_t N = triu_len(n);
  const size_t Np = (size_t)round_up(N, 32), G = Np / 32, nseg = (G + kTreeSegChunks - 1) / kTreeSegChunks;
  const size_t levels = (size_t)tree_levels(N);
  return (levels + 1) * arena_bytes(Np, 4) + arena_bytes(levels * G + 1, 4) + arena_bytes(Np, 4) + arena_bytes(65536, 1) +
         arena_bytes((size_t)(kTreeBatch / 32) * n, 4) + arena_bytes((size_t)(kTreeBatch / 32) * (levels + 1) * nseg * 64, 4) +
         arena_bytes(Np, 4) + 3 * arena_bytes(G, 4) +               // stimulus pairs of the y order, run-head bits of the tie passes
         arena_bytes((size_t)kTreeBatch * 4, 8) + arena_bytes(kTreeBatch, 4) + 4096;
}
static bool kendall_tree_wanted(int n, int R) {
  const int opt = option_kendall_tree();
  return opt != 0 && n <= kTreeMaxStimuli && (opt == 2 || R >= kTreeMinReplicates);
}

}  // namespace vr

using namespace vr;

extern "C" size_t vr_kendall_workspace_bytes(int n) {
  if (n < 2) return 256;
  return kendall_fixed_bytes(n) + kendall_per_replicate_bytes(n) + kKendallSlack;
}
extern "C" size_t vr_kendall_batch_workspace_bytes(int n, int R) {
  if (n < 2 || R < 1) return 256;
  if (kendall_tree_wanted(n, R)) return vr_kendall_workspace_bytes(n) + kendall_tree_bytes(n);      // replicates need no arrays of their own
  const int rb = R < kKendallMaxBatch ? R : kKendallMaxBatch;
  return kendall_fixed_bytes(n) + (size_t)rb * kendall_per_replicate_bytes(n) + kKendallSlack;
}

extern "C" int vr_kendall_tau_a(const float* rdmA, const float* rdmB, int n, int64_t ld, const uint32_t* masks, int R,
                                double* out, void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(rdmA && rdmB && masks && out, "vr_kendall_tau_a: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && ld >= n && R >= 1, "vr_kendall_tau_a: bad sizes n=%d ld=%lld R=%d", n, (long long)ld, R);
  if (!ws || ws_bytes < vr_kendall_workspace_bytes(n)) {
    set_error("vr_kendall_tau_a: workspace %zu < %zu", ws_bytes, vr_kendall_workspace_bytes(n));
    return VR_ERR_WORKSPACE;
  }
  const int64_t N = triu_len(n);
  const size_t Np = (size_t)round_up(N, 32);
  const uint32_t G = (uint32_t)(Np / 32), words = G;
  const uint32_t nchunks = (G + kChunkGroups - 1) / kChunkGroups;
  const int W = vr_mask_words(n);
  // tree mode when it pays and the workspace has room for the tree next to at least one replicate
  const bool tree = kendall_tree_wanted(n, R) &&
                    ws_bytes >= vr_kendall_workspace_bytes(n) + kendall_tree_bytes(n);
  // as many replicates per batch as the workspace holds (every batch is ~15 launches whatever its size)
  int Rb = (int)((ws_bytes - kendall_fixed_bytes(n) - kKendallSlack - (tree ? kendall_tree_bytes(n) : 0)) / kendall_per_replicate_bytes(n));
  if (Rb > kKendallMaxBatch) Rb = kKendallMaxBatch;
  if (Rb > R) Rb = R;
  if (Rb < 1 || tree) Rb = 1;
  Arena a(ws, ws_bytes);
  uint32_t* xkey_c = a.take<uint32_t>(Np);
  uint32_t* ykey_c = a.take<uint32_t>(Np);
  uint32_t* k0 = a.take<uint32_t>(Np);
  uint32_t* v0 = a.take<uint32_t>(Np);
  uint32_t* k1 = a.take<uint32_t>(Np);
  uint32_t* v1 = a.take<uint32_t>(Np);
  uint32_t* order_y = a.take<uint32_t>(Np);     // pair indices in y order
  uint32_t* ysorted = a.take<uint32_t>(Np);     // y keys in y order
  uint32_t* order_xy = a.take<uint32_t>(Np);    // pair indices in (x, y) order
  uint32_t* xseq = a.take<uint32_t>(Np);        // x keys in (x, y) order
  uint32_t* yseq = a.take<uint32_t>(Np);        // y keys in (x, y) order
  uint32_t* rs_x = a.take<uint32_t>(Np);        // run starts: equal x / equal (x, y) in (x, y) order, equal y in y order
  uint32_t* rs_xy = a.take<uint32_t>(Np);
  uint32_t* rs_y = a.take<uint32_t>(Np);
  uint32_t* counts_sort = a.take<uint32_t>(sort_count_elems(N, 1));
  uint32_t* cy = a.take<uint32_t>((size_t)Rb * Np);       // kept y keys in (x, y) order, per replicate
  uint32_t* cy2 = a.take<uint32_t>((size_t)Rb * Np);      // merge ping-pong
  uint2* tables = a.take<uint2>((size_t)2 * Rb * G);
  uint32_t* pairbits = a.take<uint32_t>((size_t)Rb * words);
  uint32_t* chunk_tot = a.take<uint32_t>((size_t)2 * Rb * nchunks);
  unsigned long long* acc = a.take<unsigned long long>((size_t)Rb * 4);
  uint32_t* counts = a.take<uint32_t>((size_t)2 * Rb);    // kept entries: [0][r] from the (x, y) order, [1][r] from the y order
  uint32_t* nan_flag = a.take<uint32_t>(8);
  const bool complement = !tree && option_kendall_complement() != 0;
  uint32_t* qcnt[4] = {nullptr, nullptr, nullptr, nullptr};
  unsigned long long* rows = nullptr;
  unsigned long long* totals = nullptr;
  uint32_t* drop = nullptr;
  unsigned long long* sub = nullptr;
  for (int k = 0; k < 4; ++k) qcnt[k] = a.take<uint32_t>(Np);
  rows = a.take<unsigned long long>((size_t)4 * n);
  totals = a.take<unsigned long long>(8);
  drop = a.take<uint32_t>((size_t)Rb * n);
  sub = a.take<unsigned long long>((size_t)Rb * 4);
  TreeArgs T{};
  uint32_t *t_ij = nullptr, *t_side = nullptr, *t_gt = nullptr, *t_mT = nullptr, *t_ijy = nullptr, *t_heads[3] = {nullptr, nullptr, nullptr};
  uint8_t* t_cross = nullptr;
  unsigned long long* t_acc = nullptr;
  uint32_t* t_counts = nullptr;
  if (tree) {
    T.levels = tree_levels(N);
    T.N = (uint32_t)N; T.Np = (uint32_t)Np; T.G = G; T.nseg = (G + kTreeSegChunks - 1) / kTreeSegChunks; T.n = n;
    t_ij = a.take<uint32_t>((size_t)(T.levels + 1) * Np);
    t_side = a.take<uint32_t>((size_t)T.levels * G + 1);
    t_gt = a.take<uint32_t>(Np);
    t_cross = a.take<uint8_t>(65536);
    t_mT = a.take<uint32_t>((size_t)(kTreeBatch / 32) * n);
    T.segtot = a.take<uint32_t>((size_t)(kTreeBatch / 32) * (T.levels + 1) * T.nseg * 64);
    t_acc = a.take<unsigned long long>((size_t)kTreeBatch * 4);
    t_counts = a.take<uint32_t>(kTreeBatch);
    T.ij = t_ij; T.side = t_side; T.gt = t_gt; T.cross = t_cross; T.mT = t_mT; T.acc = t_acc;
    t_ijy = a.take<uint32_t>(Np);
    for (int k = 0; k < 3; ++k) t_heads[k] = a.take<uint32_t>(G);
  }
  if (!a.ok) { set_error("vr_kendall_tau_a: internal workspace plan overflow"); return VR_ERR_WORKSPACE; }
  ProfScope prof(F_KENDALL, stream);
  const unsigned nblk = (unsigned)((N + 255) / 256);

  VR_CHECK_CUDA(cudaMemsetAsync(nan_flag, 0, 8 * sizeof(uint32_t), stream));
  dim3 grid((n + 255) / 256, n);
  kendall_keys_kernel<<<grid, 256, 0, stream>>>(rdmA, rdmB, n, ld, xkey_c, ykey_c, v0, nan_flag);
  VR_CHECK_LAUNCH();
  // sort by y, carrying the pair index
  VR_CHECK_CUDA(cudaMemcpyAsync(k0, ykey_c, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
  int rc = radix_sort_pairs(k0, v0, k1, v1, N, 0, 1, counts_sort, stream);
  if (rc != VR_OK) return rc;
  VR_CHECK_CUDA(cudaMemcpyAsync(ysorted, k0, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
  VR_CHECK_CUDA(cudaMemcpyAsync(order_y, v0, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
  // stable sort by x of the y-ordered sequence -> (x, y) order
  gather_kernel<<<nblk, 256, 0, stream>>>(xkey_c, order_y, (uint32_t)N, k0);
  VR_CHECK_LAUNCH();
  rc = radix_sort_pairs(k0, v0, k1, v1, N, 0, 1, counts_sort, stream);
  if (rc != VR_OK) return rc;
  VR_CHECK_CUDA(cudaMemcpyAsync(xseq, k0, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
  VR_CHECK_CUDA(cudaMemcpyAsync(order_xy, v0, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
  gather_kernel<<<nblk, 256, 0, stream>>>(ykey_c, order_xy, (uint32_t)N, yseq);
  VR_CHECK_LAUNCH();
  run_start_kernel<<<nblk, 256, 0, stream>>>(xseq, nullptr, (uint32_t)N, rs_x);
  VR_CHECK_LAUNCH();
  run_start_kernel<<<nblk, 256, 0, stream>>>(xseq, yseq, (uint32_t)N, rs_xy);
  VR_CHECK_LAUNCH();
  run_start_kernel<<<nblk, 256, 0, stream>>>(ysorted, nullptr, (uint32_t)N, rs_y);
  VR_CHECK_LAUNCH();

  EntryCounts Q{};
  if (complement) {
    // per-entry counts, once per RDM pair: run sizes in the three orders, and the discordant partners of every entry from a merge
    // sort of the y keys in (x, y) order that carries the entry index (k0 / v0 / k1 / v1 are free again)
    run_ties_kernel<<<nblk, 256, 0, stream>>>(xseq, nullptr, rs_x, order_xy, (uint32_t)N, qcnt[0]);
    VR_CHECK_LAUNCH();
    run_ties_kernel<<<nblk, 256, 0, stream>>>(ysorted, nullptr, rs_y, order_y, (uint32_t)N, qcnt[1]);
    VR_CHECK_LAUNCH();
    run_ties_kernel<<<nblk, 256, 0, stream>>>(xseq, yseq, rs_xy, order_xy, (uint32_t)N, qcnt[2]);
    VR_CHECK_LAUNCH();
    VR_CHECK_CUDA(cudaMemsetAsync(qcnt[3], 0, (size_t)N * 4, stream));
    const unsigned tiles_all = (unsigned)((N + 2047) / 2048);
    tile_sort_entries_kernel<<<tiles_all, 256, 0, stream>>>(yseq, order_xy, (uint32_t)N, k0, v0, qcnt[3]);
    VR_CHECK_LAUNCH();
    uint32_t *sv = k0, *se = v0, *dv = k1, *de = v1;
    const unsigned eg = nblk < 4u * kNumSMs ? nblk : 4u * kNumSMs;
    for (uint64_t L = 2048; L < (uint64_t)N; L <<= 1) {
      merge_level_entries_kernel<<<eg, 256, 0, stream>>>(sv, se, (uint32_t)N, (uint32_t)L, dv, de, qcnt[3]);
      VR_CHECK_LAUNCH();
      uint32_t* t = sv; sv = dv; dv = t;
      t = se; se = de; de = t;
    }
    for (int k = 0; k < 4; ++k) Q.q[k] = qcnt[k];
    entry_row_sums_kernel<<<dim3(n, 4), 256, 0, stream>>>(Q, n, rows);
    VR_CHECK_LAUNCH();
    entry_totals_kernel<<<4, 256, 0, stream>>>(rows, n, totals);
    VR_CHECK_LAUNCH();
  }
  const size_t tree_smem = 65536 + (size_t)n * 4;
  if (tree) {
    // the merge tree of the y keys in (x, y) order, once per RDM pair (k0 / v0 / k1 / v1 are free again)
    static unsigned long long attr_done = 0;
    static unsigned long long attr_ties = 0;
    VR_CHECK_CUDA(ensure_dynamic_smem(tree_count_kernel, 65536 + kTreeMaxStimuli * 4, &attr_done));
    VR_CHECK_CUDA(ensure_dynamic_smem(tree_ties_kernel, kTreeMaxStimuli * 4, &attr_ties));
    tree_cross_table_kernel<<<256, 256, 0, stream>>>(t_cross);
    VR_CHECK_LAUNCH();
    const unsigned pblk = (unsigned)((Np + 255) / 256);
    tree_heads_kernel<<<pblk, 256, 0, stream>>>(rs_x, (uint32_t)N, G, t_heads[0]);
    VR_CHECK_LAUNCH();
    tree_heads_kernel<<<pblk, 256, 0, stream>>>(rs_xy, (uint32_t)N, G, t_heads[1]);
    VR_CHECK_LAUNCH();
    tree_heads_kernel<<<pblk, 256, 0, stream>>>(rs_y, (uint32_t)N, G, t_heads[2]);
    VR_CHECK_LAUNCH();
    tree_pairs_kernel<<<pblk, 256, 0, stream>>>(order_y, (uint32_t)N, (uint32_t)Np, (uint32_t)n, t_ijy);
    VR_CHECK_LAUNCH();
    if (T.levels) VR_CHECK_CUDA(cudaMemsetAsync(t_side, 0, (size_t)T.levels * G * 4, stream));
    tree_base_kernel<<<(G + 7) / 8, 256, 0, stream>>>(yseq, order_xy, (uint32_t)N, G, (uint32_t)n, t_gt, k0, v0, t_ij);
    VR_CHECK_LAUNCH();
    uint32_t *sv = k0, *sp = v0, *dv = k1, *dp = v1;
    const unsigned eg = nblk < 8u * kNumSMs ? nblk : 8u * kNumSMs;
    int t = 0;
    for (uint64_t L = 32; L < (uint64_t)N; L <<= 1, ++t) {
      tree_merge_kernel<<<eg, 256, 0, stream>>>(sv, sp, (uint32_t)N, (uint32_t)L, dv, dp, t_ij, t_ij + (size_t)(t + 1) * Np, t_side + (size_t)t * G);
      VR_CHECK_LAUNCH();
      uint32_t* x = sv; sv = dv; dv = x;
      x = sp; sp = dp; dp = x;
    }
    // replicates: no sort, no tables -- masks, ties, discordant pairs, tau
    for (int r0 = 0; r0 < R; r0 += kTreeBatch) {
      const int rb = R - r0 < kTreeBatch ? R - r0 : kTreeBatch;
      const unsigned nb32 = (unsigned)((rb + 31) / 32);
      T.rb = rb;
      VR_CHECK_CUDA(cudaMemsetAsync(t_acc, 0, (size_t)rb * 4 * sizeof(unsigned long long), stream));
      tree_kept_kernel<<<(rb + 7) / 8, 256, 0, stream>>>(masks + (size_t)r0 * W, W, rb, t_counts);
      VR_CHECK_LAUNCH();
      tree_masks_kernel<<<dim3((n + 255) / 256, nb32), 256, 0, stream>>>(masks + (size_t)r0 * W, W, n, rb, t_mT);
      VR_CHECK_LAUNCH();
      TiesArgs ta{};
      ta.ij[0] = t_ij; ta.ij[1] = t_ijy;
      for (int k = 0; k < 3; ++k) ta.heads[k] = t_heads[k];
      ta.mT = t_mT; ta.acc = t_acc; ta.N = (uint32_t)N; ta.G = G; ta.nseg = (G + kTiesSegChunks - 1) / kTiesSegChunks; ta.n = n; ta.rb = rb;
      tree_ties_kernel<<<dim3((2u * ta.nseg + 7u) / 8u, nb32), 256, (size_t)n * 4, stream>>>(ta);
      VR_CHECK_LAUNCH();
      const unsigned segs = (unsigned)(T.levels + 1) * T.nseg;
      tree_count_kernel<<<dim3((segs + kTreeCountWarps - 1) / kTreeCountWarps, nb32), kTreeCountWarps * 32, tree_smem, stream>>>(T);
      VR_CHECK_LAUNCH();
      if (T.levels) {
        tree_stitch_kernel<<<dim3((unsigned)T.levels, nb32), 256, 0, stream>>>(T);
        VR_CHECK_LAUNCH();
      }
      kendall_finish_kernel<<<(rb + 63) / 64, 64, 0, stream>>>(t_acc, t_counts, nan_flag, rb, out + r0);
      VR_CHECK_LAUNCH();
    }
    return VR_OK;
  }
  KendallOrders ord;
  ord.order[0] = order_xy;
  ord.order[1] = order_y;
  const unsigned mgrid = nblk < 2u * kNumSMs ? nblk : 2u * kNumSMs;
  const unsigned tiles = (unsigned)((N + kTile - 1) / kTile);
  for (int r0 = 0; r0 < R; r0 += Rb) {
    const int rb = R - r0 < Rb ? R - r0 : Rb;
    VR_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)rb * 4 * sizeof(unsigned long long), stream));
    dim3 pg((words + 255) / 256 < 64 ? (words + 255) / 256 : 64, rb);
    pair_mask_kernel<<<pg, 256, (size_t)(W + 1) * 4, stream>>>(masks + (size_t)r0 * W, rb, n, W, words, pairbits);
    VR_CHECK_LAUNCH();
    if (complement) {
      complement_bits_kernel<<<(words + 255) / 256, 256, 0, stream>>>(pairbits, words, (uint32_t)N, rb);
      VR_CHECK_LAUNCH();
      dropped_entry_sums_kernel<<<rb, 256, 0, stream>>>(masks + (size_t)r0 * W, W, n, Q, rows, drop, sub);
      VR_CHECK_LAUNCH();
    }
    kept_words_kernel<<<dim3(nchunks, rb, 2), 256, 0, stream>>>(ord, pairbits, words, (uint32_t)N, G, tables, chunk_tot, nchunks);
    VR_CHECK_LAUNCH();
    chunk_scan_kernel<<<dim3(rb, 2), 1024, 0, stream>>>(chunk_tot, nchunks, counts);
    VR_CHECK_LAUNCH();
    kept_finish_kernel<<<dim3(nchunks, rb, 2), 256, 0, stream>>>(tables, chunk_tot, nchunks, G, (uint32_t)N, yseq, cy, (int64_t)Np);
    VR_CHECK_LAUNCH();
    tie_sums_kernel<<<dim3(mgrid, rb), 256, 0, stream>>>(tables, G, (uint32_t)N, rs_x, rs_xy, rs_y, acc);
    VR_CHECK_LAUNCH();
    tile_sort_kernel<<<dim3(tiles, rb), 256, 0, stream>>>(cy, cy2, (int64_t)Np, counts, acc);
    VR_CHECK_LAUNCH();
    uint32_t* src = cy2;
    uint32_t* dst = cy;
    for (uint64_t L = kTile; L < (uint64_t)N; L <<= 1) {      // M <= N: levels beyond M only copy
      merge_level_kernel<<<dim3(mgrid, rb), 256, 0, stream>>>(src, dst, (int64_t)Np, counts, (uint32_t)L, acc);
      VR_CHECK_LAUNCH();
      uint32_t* t = src; src = dst; dst = t;
    }
    if (complement) kendall_finish_complement_kernel<<<(rb + 63) / 64, 64, 0, stream>>>(acc, counts, totals, sub, (unsigned long long)N, nan_flag, rb, out + r0);
    else kendall_finish_kernel<<<(rb + 63) / 64, 64, 0, stream>>>(acc, counts, nan_flag, rb, out + r0);
    VR_CHECK_LAUNCH();
  }
  return VR_OK;
}
