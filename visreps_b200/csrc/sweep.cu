// Grouped Spearman bootstrap, "one pair x 32 replicates per warp" formulation.
//
// Replaces, for nA model RDMs x nB neural RDMs x R subsample masks, the loop of visreps/evals.py:361-369 (and its
// copies at :512-520, rsa.py:248-256): `rdm[idx][:, idx]` x2 -> triu gather (rsa.py:111-113) ->
// scipy.stats.spearmanr == rankdata('average') x2 + np.corrcoef (rsa.py:122).
//
// As in rank.cu the ranks of a subsample come from the presorted order of the FULL vector: with C_r(x) = number of
// pairs kept by replicate r at sorted positions < x and [s, e) the run of equal values holding pair p,
//     dev_r(p) = 2 rank - (N_r + 1) = C_r(s) + C_r(e) - N_r.
// What changes is who does what.  A warp takes ONE canonical pair and the 32 replicates of a batch (lane = replicate):
//   * count_rows_kernel lays the count tables out as rows [batch][vector][32-group g][lane] of {kept word, T} with
//     T = 2 C_r(32 g) - N_r + 2^31, so the row of a pair's group is one coalesced 256-byte read (L2 resident: a
//     batch block is 64 MB at the NSD shape) and lane r gets its replicate's entry without any bank conflict;
//   * the run word of the pair (presort, replicate independent) is the same for all lanes: no divergence.  A pair in a
//     run of length one -- most pairs -- needs one popc:  dev + 2^31 = T + 2 popc(kept & below_s) + 1;
//   * dev + 2^31 is the low word of the double 2^52 + 2^31 + dev, so one DADD makes it an exact fp64 value and the
//     nA x nB dot products of the replicate are accumulated per lane with DFMA (exact: every partial sum stays below
//     2^53 within a warp's slice; slices are added as int64).  nA x nB <= 8 x 8 accumulators live in registers, the
//     table rows of the NEXT pair are in flight while the current pair is multiplied.
//   * sum(dev^2) never needs a pass of its own:  sum = (N_r^3 - N_r - sum over runs (t^3 - t)) / 3  with t the kept
//     members of a run; tie_sums_kernel visits only the runs of two or more equal values.
// Work sharing is the same as before: all pairs of a job see the same masks (evals.py:356), so the 16 rankings of an
// 8 x 8 job are looked up once per (pair, replicate) and feed 64 products.
#include <type_traits>
#include "common.cuh"
#include "pairmask.cuh"

namespace vr {

constexpr int kSweepWarps = 8;
constexpr int kSweepThreads = kSweepWarps * 32;
constexpr int kSweepMax = 8;                 // vectors per side in one sweep launch
constexpr int kCountWarpsRows = 32;          // warps (segments) per CTA of the count kernels
constexpr uint32_t kDevBias = 0x80000000u;   // dev + 2^31 is kept in unsigned registers
constexpr uint32_t kRunSingleBit = 1u;       // compact run word, bit 0: the run has one member (e = s + 1)

// Exact mode: int64 totals (|sum| <= N^3 / 3, t^3 of a run) and a warp share of at least 32 pairs whose fp64 partial sums
// stay below 2^53.  Beyond that (N > ~2 M pairs) everything is summed in fp64, in a fixed order.
static inline bool sweep_exact(uint32_t N) {
  return (double)N * (double)N * (double)N < 9.0e18 && (double)N * (double)N * 32.0 < 9007199254740992.0;
}

__device__ __forceinline__ uint32_t shl_clamp32(uint32_t x, uint32_t k) {   // PTX shl: k >= 32 gives 0
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(k));
  return r;
}
__device__ __forceinline__ uint2 ldg_row(const char* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
// exact (dev + 2^31 as uint32) -> fp64 dev with one DADD
__device__ __forceinline__ double biased_to_double(uint32_t v) {
  return __hiloint2double(0x43300000, (int)v) - 4503601774854144.0;          // 2^52 + 2^31
}

// ------------------------------------------------------------------------------------------------------------
// Job layout (device, caller's memory):
//   maskT  [batches][n]            bit r of word i: stimulus i kept by replicate 32 b + r
//   nk     [batches * 32]          N_r = m (m - 1) / 2, m = stimuli kept by replicate r   (0 for the padding replicates)
//   tie    [slots][batches * 32]   sum over runs of (t^3 - t), int64
//   seg    [slots][batches][segments][32]   kept pairs per count segment (scratch of the table build)
//   rows   [batches][slots][Gs][32] uint2
// ------------------------------------------------------------------------------------------------------------
struct SweepJob {
  uint32_t* maskT;
  uint32_t* nk;
  long long* tie;
  uint32_t* seg;
  uint2* rows;
  int n, R, batches, slots, segments;
  uint32_t N, G, Gs;
  size_t slot_stride;     // bytes between the rows of two vectors
  size_t batch_stride;    // bytes between the rows of two batches
  size_t group_stride;    // bytes between the rows of two consecutive 32-groups
  int batch_inner;        // layout [slot][g][batch][lane] instead of [batch][slot][g][lane]
};
// Row {kept, T} x 32 lanes of (batch, vector slot, 32-group g).
__host__ __device__ inline uint2* job_row(const SweepJob& J, size_t batch, size_t slot, size_t g) {
  return reinterpret_cast<uint2*>(reinterpret_cast<char*>(J.rows) + batch * J.batch_stride + slot * J.slot_stride + g * J.group_stride);
}
// Two vectors (one pair of RDMs) and several batches -- the stress shape: 50 M pairs, tables of 400 MB per (vector, batch), far
// beyond L2 -- are laid out batch-innermost: the rows of all batches for one (vector, group) are contiguous, so the sweep, which
// then walks the batches of a pair in its inner loop, reads one 256 x batches byte burst per lookup instead of `batches` row
// reads in DRAM pages hundreds of MB apart.
static inline bool sweep_batch_inner(int slots, int64_t batches) { return slots == 2 && batches >= 2 && batches <= 8; }
// Count segments per (vector, batch): 32 warps x z CTAs, z chosen so that a job with few (vector, batch) pairs -- the
// stress shape with a handful of masks -- still fills the machine.
constexpr int kCountCtasRows = 16;
static inline int sweep_segments(int n, int vectors, int64_t batches) {      // of one table build (vectors added together)
  int64_t z = (2 * kNumSMs + vectors * batches - 1) / (vectors * batches);
  if (z < 1) z = 1;
  if (z > kCountCtasRows) z = kCountCtasRows;
  const int64_t zmin = ((int64_t)presort_groups(n) + 32 * 100000 - 1) / (32 * 100000);   // the bit-plane counters hold < 2^17 groups
  if (z < zmin) z = zmin;
  return (int)(z * kCountWarpsRows);
}
static inline int sweep_segments_max(int n) {
  const int64_t zmin = ((int64_t)presort_groups(n) + 32 * 100000 - 1) / (32 * 100000);
  return (int)((zmin > kCountCtasRows ? zmin : kCountCtasRows) * kCountWarpsRows);
}
static inline size_t sweep_job_bytes(int n, int slots, int64_t R) {
  const int64_t batches = (R + 31) / 32;
  const size_t Gs = (size_t)presort_groups(n) + 1;
  return 256 + arena_bytes((size_t)batches * n, 4) + arena_bytes((size_t)batches * 32, 4) + arena_bytes((size_t)slots * batches * 32, 8) +
         arena_bytes((size_t)slots * batches * sweep_segments_max(n) * 32, 4) + arena_bytes((size_t)batches * slots * Gs * 32, 8);
}
static inline SweepJob sweep_job_view(void* mem, int n, int slots, int R) {
  SweepJob j;
  j.n = n; j.R = R; j.batches = (R + 31) / 32; j.slots = slots; j.segments = sweep_segments_max(n);
  j.N = (uint32_t)triu_len(n); j.G = presort_groups(n); j.Gs = j.G + 1;
  char* p = static_cast<char*>(mem) + 256;
  j.maskT = reinterpret_cast<uint32_t*>(p); p += arena_bytes((size_t)j.batches * n, 4);
  j.nk = reinterpret_cast<uint32_t*>(p); p += arena_bytes((size_t)j.batches * 32, 4);
  j.tie = reinterpret_cast<long long*>(p); p += arena_bytes((size_t)slots * j.batches * 32, 8);
  j.seg = reinterpret_cast<uint32_t*>(p); p += arena_bytes((size_t)slots * j.batches * j.segments * 32, 4);
  j.rows = reinterpret_cast<uint2*>(p);
  j.batch_inner = sweep_batch_inner(slots, j.batches) ? 1 : 0;
  if (j.batch_inner) {
    j.batch_stride = 256;
    j.group_stride = (size_t)256 * j.batches;
    j.slot_stride = (size_t)j.Gs * j.group_stride;
  } else {
    j.group_stride = 256;
    j.slot_stride = (size_t)j.Gs * 256;
    j.batch_stride = j.slot_stride * slots;
  }
  return j;
}

// maskT and N_r from the (R, W) masks.  grid (ceil(n / 256), batches)
__global__ void sweep_masks_kernel(const uint32_t* __restrict__ masks, int R, int n, int W, uint32_t* __restrict__ maskT) {
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
__global__ void sweep_nk_kernel(const uint32_t* __restrict__ masks, int R, int n, int W, int total, uint32_t* __restrict__ nk) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= total) return;
  uint32_t m = 0;
  if (r < R)
    for (int w = 0; w < W; ++w) {
      uint32_t x = masks[(int64_t)r * W + w];
      if (w == W - 1 && (n & 31)) x &= (1u << (n & 31)) - 1u;
      m += __popc(x);
    }
  nk[r] = m * (m - 1u) / 2u;                        // m <= 32767: fits
}

// ---- count tables ----------------------------------------------------------------------------------------------
// Segment s of a vector = groups [s * seg_groups, (s + 1) * seg_groups); one warp per (segment, vector, batch).
// Phase 1 (count_totals_kernel): kept pairs of the segment for each of the batch's 32 replicates.  Lane l reads sorted
// position 32 g + l; its 32-bit word holds the pair's kept bit for all replicates, so the per-replicate totals are the
// column sums of the segment's bit matrix: carry-save adders fold eight words into bit planes (ones/twos/fours), higher
// planes ripple once per eight words, and the planes are transposed once at the end.
constexpr int kPlanes = 14;     // weights 8 .. 65536: a lane sees at most 2^17 - 1 groups of a segment (n = 32767: 2^24 groups / 32 segments ... / z)

__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t u = a ^ b;
  h = (a & b) | (u & c);
  l = u ^ c;
}
__device__ __forceinline__ uint32_t kept_bits_of(const uint32_t* __restrict__ stream, uint32_t G, uint32_t g, unsigned lane,
                                                 const uint32_t* __restrict__ s_maskT) {
  if (g >= G) return 0u;
  const uint32_t wd = __ldg(stream + (size_t)g * 32u + lane);
  return (wd & kPadBit) ? 0u : (s_maskT[pair_i(wd)] & s_maskT[pair_j(wd)]);
}
// Transpose a 32x32 bit matrix held one row per lane (row = lane, column = bit).
__device__ __forceinline__ uint32_t bit_transpose32(uint32_t w, unsigned lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t x = __shfl_xor_sync(0xffffffffu, w, s);
    w = (lane & s) ? ((w & ~m) | ((x >> s) & m)) : ((w & m) | ((x << s) & ~m));
  }
  return w;
}

struct BlobPtrs { const void* p[kSweepMax * 2]; };

// grid (segments / 32, count, batches), 1024 threads; dynamic shared memory: n words (maskT of the batch)
__global__ void __launch_bounds__(kCountWarpsRows * 32)
count_totals_kernel(const __grid_constant__ BlobPtrs blobs, int slot0, const __grid_constant__ SweepJob J, uint32_t seg_groups) {   // J.segments: row pitch of `seg`
  extern __shared__ __align__(16) uint32_t s_maskT[];
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int v = blockIdx.y, batch = blockIdx.z;
  const uint32_t seg = blockIdx.x * kCountWarpsRows + warp;
  for (int i = threadIdx.x; i < J.n; i += blockDim.x) s_maskT[i] = J.maskT[(size_t)batch * J.n + i];
  __syncthreads();
  PresortView blob = presort_view(const_cast<void*>(blobs.p[v]), J.n);
  const uint32_t g_lo = min(J.G, seg * seg_groups), g_hi = min(J.G, g_lo + seg_groups);
  uint32_t plane[kPlanes];
#pragma unroll
  for (int k = 0; k < kPlanes; ++k) plane[k] = 0u;
  uint32_t ones = 0, twos = 0, fours = 0;
  for (uint32_t g0 = g_lo; g0 < g_hi; g0 += 8u) {
    uint32_t w[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) w[q] = (g0 + q < g_hi) ? kept_bits_of(blob.stream, J.G, g0 + q, lane, s_maskT) : 0u;
    uint32_t twoA, twoB, fourA, fourB, eight;
    csa(twoA, ones, ones, w[0], w[1]);
    csa(twoB, ones, ones, w[2], w[3]);
    csa(fourA, twos, twos, twoA, twoB);
    csa(twoA, ones, ones, w[4], w[5]);
    csa(twoB, ones, ones, w[6], w[7]);
    csa(fourB, twos, twos, twoA, twoB);
    csa(eight, fours, fours, fourA, fourB);
    // ripple `eight` (weight 8) into the planes: plane[k] has weight 8 << k
    uint32_t carry = eight;
#pragma unroll
    for (int k = 0; k < kPlanes; ++k) {
      const uint32_t t = plane[k] & carry;
      plane[k] ^= carry;
      carry = t;
    }
  }
  // per-replicate totals: lane r sums, over the 32 lanes' planes, bit r of each plane
  uint32_t total = __popc(bit_transpose32(ones, lane)) + 2u * __popc(bit_transpose32(twos, lane)) + 4u * __popc(bit_transpose32(fours, lane));
#pragma unroll
  for (int k = 0; k < kPlanes; ++k) total += (8u << k) * __popc(bit_transpose32(plane[k], lane));
  J.seg[(((size_t)(slot0 + v) * J.batches + batch) * J.segments + seg) * 32 + lane] = total;
}

// Phase 2: the rows.  Same grid.  T = 2 C(32 g) - N_r + 2^31.
__global__ void __launch_bounds__(kCountWarpsRows * 32)
count_rows_kernel(const __grid_constant__ BlobPtrs blobs, int slot0, const __grid_constant__ SweepJob J, uint32_t seg_groups) {
  extern __shared__ __align__(16) uint32_t s_maskT[];
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int v = blockIdx.y, batch = blockIdx.z;
  const uint32_t seg = blockIdx.x * kCountWarpsRows + warp;
  for (int i = threadIdx.x; i < J.n; i += blockDim.x) s_maskT[i] = J.maskT[(size_t)batch * J.n + i];
  __syncthreads();
  PresortView blob = presort_view(const_cast<void*>(blobs.p[v]), J.n);
  const uint32_t g_lo = min(J.G, seg * seg_groups), g_hi = min(J.G, g_lo + seg_groups);
  const uint32_t* segs = J.seg + ((size_t)(slot0 + v) * J.batches + batch) * J.segments * 32;
  uint32_t running = 0;
  for (uint32_t s = 0; s < seg; ++s) running += segs[(size_t)s * 32 + lane];
  const uint32_t nr = J.nk[batch * 32 + lane];
  const uint32_t bias = kDevBias - nr;
  for (uint32_t g0 = g_lo; g0 < g_hi; g0 += 4u) {
    uint32_t bits[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) bits[q] = (g0 + q < g_hi) ? kept_bits_of(blob.stream, J.G, g0 + q, lane, s_maskT) : 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t K = bit_transpose32(bits[q], lane);          // lane r: kept word of replicate r for the group
      if (g0 + q < g_hi) job_row(J, batch, slot0 + v, g0 + q)[lane] = make_uint2(K, 2u * running + bias);
      running += __popc(K);
    }
  }
  // the sentinel row behind the last group: C = N_r (runs that end at N exactly on a group boundary read it)
  if (g_hi == J.G && g_lo < g_hi) job_row(J, batch, slot0 + v, J.G)[lane] = make_uint2(0u, 2u * running + bias);
  if (J.G == 0 && seg == 0) job_row(J, batch, slot0 + v, 0)[lane] = make_uint2(0u, bias);
}

// ---- tie sums ---------------------------------------------------------------------------------------------------
// sum over the runs of two or more equal values of t^3 - t, t = members kept by the replicate.  One warp per 32-group,
// lane = replicate.  grid (chunks, count, batches), 256 threads.
template <bool kExact>
__global__ void __launch_bounds__(256)
tie_sums_kernel(const __grid_constant__ BlobPtrs blobs, int slot0, const __grid_constant__ SweepJob J) {
  using Acc = std::conditional_t<kExact, long long, double>;
  const unsigned lane = lane_id();
  const int v = blockIdx.y, batch = blockIdx.z;
  PresortView blob = presort_view(const_cast<void*>(blobs.p[v]), J.n);
  Acc acc = 0;
  const uint32_t warps = gridDim.x * 8u;
  for (uint32_t g = blockIdx.x * 8u + (threadIdx.x >> 5); g < J.G; g += warps) {
    const uint32_t H = __ldg(blob.head_bits + g);
    // is the first position of the next group a head?  (position N counts as one)
    const uint32_t first_next = (g + 1u < J.G) ? (__ldg(blob.head_bits + g + 1u) & 1u) : 1u;
    uint32_t valid = 0xFFFFFFFFu;
    if (g == J.G - 1u && (J.N & 31u)) valid = (1u << (J.N & 31u)) - 1u;    // positions < N (bit N % 32 of H marks the end)
    const uint32_t next_is_head = (H >> 1) | (first_next << 31);
    uint32_t M = H & ~next_is_head & valid;                                // heads of runs with two or more members
    if (M == 0u) continue;
    const uint2 row = job_row(J, batch, slot0 + v, g)[lane];
    while (M) {
      const uint32_t s = __ffs(M) - 1u;
      M &= M - 1u;
      const uint32_t above = s == 31u ? 0u : ((H >> (s + 1u)) << (s + 1u));
      uint32_t t;
      if (above) {                                                         // the run ends inside this group
        const uint32_t e = __ffs(above) - 1u;
        t = __popc(row.x & ((e == 32u ? 0xFFFFFFFFu : ((1u << e) - 1u)) & ~((1u << s) - 1u)));
      } else {
        const uint32_t e = __ldg(blob.next_head + g + 1u);                // global position of the run's end
        const uint2 re = job_row(J, batch, slot0 + v, e >> 5)[lane];
        const uint32_t ce2 = re.y + 2u * __popc(shl_clamp32(re.x, 32u - (e & 31u)));     // 2 C(e) + const
        const uint32_t cs2 = row.y + 2u * __popc(shl_clamp32(row.x, 32u - s));           // 2 C(s) + const
        t = (ce2 - cs2) >> 1;
      }
      const Acc tt = (Acc)t;
      acc += tt * tt * tt - tt;
    }
  }
  __shared__ Acc red[8][32];
  red[threadIdx.x >> 5][lane] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    Acc t = 0;
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    long long* dst = J.tie + ((size_t)(slot0 + v) * J.batches + batch) * 32 + lane;
    if (t != 0) {
      if constexpr (kExact) atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)t);
      else atomicAdd(reinterpret_cast<double*>(dst), t);
    }
  }
}

// ---- the sweep ---------------------------------------------------------------------------------------------------
struct SweepVec {
  const uint32_t* runs;        // canonical run words of the vector (presorted blob)
  const uint32_t* head_bits;   // long runs only
  const uint32_t* next_head;
  uint32_t row_base;           // byte offset of the vector's rows inside a batch block (fits 32 bits, checked on the host)
};
struct SweepArgs {
  SweepVec vec[2 * kSweepMax];   // A vectors first, then B
  int n;
  uint32_t N, sbits;
  const char* rows;              // job rows
  size_t batch_stride;
  const uint32_t* maskT;
  int batches, splits;
  void* partial;                 // [batch][split][NA * NB][32]  int64 (exact mode) or fp64
  int exact;
};

// What a pair needs from a vector's table, decoded once from its run word (replicate independent, so the same for all 32
// lanes): the byte offset of the row of the run's first group, the mask of the positions of that group below s, and the mask
// of the positions below e.  With the row {kept, T = 2 C(32 g) - N_r + 2^31}:
//     dev + 2^31 = T + popc(kept & mask_s) + popc(kept & mask_e)
// A run of ONE value has e = s + 1, so C(e) = C(s) + 1 for a kept pair: mask_e = 0 marks it and saves the second popc.
// Runs that end in a later group (a few per cent) set `slow` and are recomputed from the run word after the fast pass.
struct PairDesc { uint32_t off, mask_s, mask_e; bool slow; };
template <bool kCompact>
__device__ __forceinline__ PairDesc decode_run(uint32_t w, uint32_t row_base, uint32_t sbits) {
  PairDesc d;
  if (kCompact) {
    const uint32_t shlo = w >> kRunShloShift, she = (w >> kRunSheShift) & 63u;
    d.off = row_base + ((w & kRunOffMask) << 5);                                // + gs * 256
    d.slow = (w & kRunCrossBit) || shlo == 0u;
    d.mask_s = shlo >= 32u ? 0u : (0xFFFFFFFFu >> shlo);                        // positions below s: low 32 - shlo bits
    d.mask_e = (w & kRunSingleBit) ? 0u : (0xFFFFFFFFu >> she);                 // she = 0: the whole group
  } else {
    const uint32_t smask = (1u << sbits) - 1u, esc = (1u << (32u - sbits)) - 1u;
    const uint32_t s = w & smask, len1 = w >> sbits, ls = s & 31u;
    d.off = row_base + ((s >> 5) << 8);
    d.mask_s = (1u << ls) - 1u;
    const uint32_t le = ls + len1 + 1u;                                         // e relative to the group start
    d.slow = len1 == esc || le > 32u;
    d.mask_e = len1 == 0u ? 0u : (le >= 32u ? 0xFFFFFFFFu : ((1u << le) - 1u));
  }
  return d;
}
// The slow cases: dev + 2^31 from the run word, loading whatever other row the run's end lies in.
template <bool kCompact>
__device__ __noinline__ uint32_t dev_slow(uint32_t w, const uint2 e0, const char* base, const SweepVec& V, uint32_t sbits) {
  uint32_t ls, e;
  if (kCompact) {
    const uint32_t gs = (w & kRunOffMask) >> 3;
    if ((w >> kRunShloShift) != 0u) {                                           // the run ends inside the next group
      ls = 32u - (w >> kRunShloShift);
      e = (gs + 1u) * 32u + ((32u - ((w >> kRunSheShift) & 63u)) & 31u);
    } else {                                                                    // longer than a group: head tables
      ls = ((w >> kRunSheShift) & 63u) - 1u;
      const uint32_t above = ls == 31u ? 0u : (__ldg(V.head_bits + gs) >> (ls + 1u));
      e = above ? gs * 32u + ls + (uint32_t)__ffs(above) : __ldg(V.next_head + gs + 1);
    }
  } else {
    const uint32_t smask = (1u << sbits) - 1u, esc = (1u << (32u - sbits)) - 1u;
    const uint32_t s = w & smask, len1 = w >> sbits;
    ls = s & 31u;
    e = s + len1 + 1u;
    if (len1 == esc) {
      const uint32_t g = s >> 5;
      const uint32_t above = ls == 31u ? 0u : (__ldg(V.head_bits + g) >> (ls + 1u));
      e = above ? s + (uint32_t)__ffs(above) : __ldg(V.next_head + g + 1);
    }
  }
  const uint2 e1 = ldg_row(base + ((size_t)(e >> 5) << 8));
  return e0.y + ((e1.y - e0.y) >> 1) + __popc(shl_clamp32(e0.x, 32u - ls)) + __popc(shl_clamp32(e1.x, 32u - (e & 31u)));
}

// Dynamic shared memory per warp: two blocks of 32 pairs each --
//   offs  [2][NQ][32][4]   row offsets, four vectors per quad            (issue:   NQ broadcast LDS.128 per pair)
//   masks [2][NH][32][4]   {mask_s, mask_e} of two vectors per quad      (consume: NH broadcast LDS.128 per pair)
//   meta  [2][32][2]       kept word of the pair (bit r: kept by replicate r), bit mask of the vectors on the slow path
template <int NVEC> struct SweepSmem {
  static constexpr int NQ = (NVEC + 3) / 4, NH = (NVEC + 1) / 2;
  static constexpr size_t kWarpBytes = (size_t)2 * (NQ + NH) * 32 * 16 + 2 * 32 * 8;
  static constexpr size_t kBytes = kWarpBytes * kSweepWarps;
};
// Measured and dropped (round 2, NSD shape, 8 x 8 tile at 10.0 ms): bulk-copy (cp.async.bulk, 256 B per row) fetches into a
// shared-memory ring -- 19 ms, the copy engine takes ~20 cycles per 256-byte request; per-lane cp.async into a ring -- 9.7 vs
// 8.4 ms in the microbenchmark (tools/experiments/sweep_bench.cu), the extra LDS traffic costs more than the deeper prefetch
// gains; two warps per pair exchanging dev values through shared memory (32 accumulators, 16 warps per SM) -- 13 ms, the
// per-pair barrier serialises the two warps and no more rows are in flight; 4 x 8 / 4 x 4 tiles -- 26 / 15 ms (more lookups).
// Table rows are fetched into registers kDepth - 1 pairs ahead of the pair being multiplied.  The 8 x 8 tile has room for one
// pair ahead (64 accumulators = 128 registers); smaller tiles look further ahead, which is what hides the L2 latency when a
// pair needs only a few rows (a single matrix swept against eight: 9 rows per pair).
__host__ __device__ constexpr int sweep_depth(int nvec) { return nvec > 12 ? 2 : nvec > 8 ? 3 : 4; }
__host__ __device__ constexpr int sweep_ctas_per_sm(int na, int nb) { return na * nb <= 16 ? 2 : 1; }

// grid (batches * splits), kSweepThreads.  A CTA takes the canonical range [c_lo, c_hi) of one batch; each warp a
// contiguous share of it, in blocks of 32 pairs whose run words are decoded once (lane = pair) into shared memory.
template <int NA, int NB, bool kCompact>
__global__ void __launch_bounds__(kSweepThreads, sweep_ctas_per_sm(NA, NB))
sweep_kernel(const __grid_constant__ SweepArgs A) {
  constexpr int NVEC = NA + NB;
  using SM = SweepSmem<NVEC>;
  constexpr int NQ = SM::NQ, NH = SM::NH, D = sweep_depth(NVEC);
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  __shared__ unsigned long long red[kSweepWarps][32];
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  unsigned char* wbase = sweep_smem + (size_t)warp * SM::kWarpBytes;
  uint4 (*offs)[NQ][32] = reinterpret_cast<uint4 (*)[NQ][32]>(wbase);                                   // [2][NQ][32]
  uint4 (*masks)[NH][32] = reinterpret_cast<uint4 (*)[NH][32]>(wbase + (size_t)2 * NQ * 32 * 16);          // [2][NH][32]
  uint2 (*meta)[32] = reinterpret_cast<uint2 (*)[32]>(wbase + (size_t)2 * (NQ + NH) * 32 * 16);            // [2][32]

  const int b = blockIdx.x / A.splits, sp = blockIdx.x % A.splits;
  const uint32_t per = (A.N + A.splits - 1) / A.splits;
  const uint32_t c_lo = min(A.N, sp * per), c_hi = min(A.N, c_lo + per);
  const uint32_t wper = ((c_hi - c_lo) + kSweepWarps - 1) / kSweepWarps;
  const uint32_t p_lo = min(c_hi, c_lo + warp * wper), p_hi = min(c_hi, p_lo + wper);
  const int count = (int)(p_hi - p_lo);
  const char* const bbase = A.rows + (size_t)b * A.batch_stride + lane * 8;     // + 32-bit offsets (host: a batch block < 4 GB)
  const uint32_t* mt = A.maskT + (size_t)b * A.n;
  double acc[NA][NB];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int c = 0; c < NB; ++c) acc[a][c] = 0.0;

  auto vecref = [&](int v) -> const SweepVec& { return A.vec[v < NA ? v : kSweepMax + (v - NA)]; };
  auto stage_block = [&](int blk) {              // pairs p_lo + 32 blk .. + 31 -> buffer blk & 1; lane = pair
    const uint32_t p = min(p_lo + (uint32_t)blk * 32u + lane, A.N - 1u);
    uint32_t off[NQ * 4], ms[NH * 2], me[NH * 2];
    uint32_t slow = 0u;
#pragma unroll
    for (int v = 0; v < NQ * 4; ++v) off[v] = 0u;
#pragma unroll
    for (int v = 0; v < NH * 2; ++v) { ms[v] = 0u; me[v] = 0u; }
#pragma unroll
    for (int v = 0; v < NVEC; ++v) {
      const PairDesc d = decode_run<kCompact>(__ldg(vecref(v).runs + p), vecref(v).row_base, A.sbits);
      off[v] = d.off; ms[v] = d.mask_s; me[v] = d.mask_e;
      slow |= d.slow ? (1u << v) : 0u;
    }
    const uint32_t i = triu_row(p, (uint32_t)A.n);
    const uint32_t j = p - i * (2u * A.n - i - 1u) / 2u + i + 1u;
    const uint32_t kw = __ldg(mt + i) & __ldg(mt + j);
    const int buf = blk & 1;
#pragma unroll
    for (int q = 0; q < NQ; ++q) offs[buf][q][lane] = make_uint4(off[4 * q], off[4 * q + 1], off[4 * q + 2], off[4 * q + 3]);
#pragma unroll
    for (int h = 0; h < NH; ++h) masks[buf][h][lane] = make_uint4(ms[2 * h], me[2 * h], ms[2 * h + 1], me[2 * h + 1]);
    meta[buf][lane] = make_uint2(kw, slow);
    __syncwarp();
  };
  auto issue = [&](int i, uint2 (&e)[NVEC]) {    // table rows of the warp's i-th pair
    const int buf = (i >> 5) & 1, k = i & 31;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const uint4 o = offs[buf][q][k];
      const uint32_t oo[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (4 * q + t < NVEC) e[4 * q + t] = ldg_row(bbase + (size_t)oo[t]);
    }
  };
  // A pair takes two steps that use different pipes: its ranks (shared-memory descriptors, popc on the XU pipe, integer adds)
  // and its products (fp64 pipe).  For the smaller tiles the products of pair i - 1 are written next to the ranks of pair i, in
  // one basic block, so that the scheduler can interleave the two instruction streams inside a warp (THINGS shape, 8 x 1:
  // 23.9 -> 21.5 ms).  The 8 x 8 tile has no registers to spare for the extra live ranks (254, and 13.5 instead of 12.7 ms
  // at the NSD shape with the interleaved order), so it keeps ranks -> products per pair.
  constexpr bool kPipelined = NA * NB < 64;
  auto ranks = [&](int i, const uint2 (&e)[NVEC], uint32_t (&dev)[NVEC], uint2& mk) {
    const int buf = (i >> 5) & 1, k = i & 31;
    mk = meta[buf][k];
    const bool keep = (mk.x >> lane) & 1u;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const uint4 m = masks[buf][h][k];
      const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int v = 2 * h + t;
        if (v < NVEC) {
          const uint32_t ps = __popc(e[v].x & mm[2 * t]);
          uint32_t pe = ps + 1u;                                                // a run of one value: C(e) = C(s) + 1
          if (mm[2 * t + 1] != 0u) pe = __popc(e[v].x & mm[2 * t + 1]);
          const uint32_t d = e[v].y + ps + pe;
          dev[v] = (v < NA && !keep) ? kDevBias : d;                            // a dropped pair contributes nothing
        }
      }
    }
  };
  auto ranks_slow = [&](int i, const uint2 (&e)[NVEC], uint32_t (&dev)[NVEC], const uint2 mk) {   // rare: runs that end in a later group
    const uint32_t p = p_lo + (uint32_t)i;
    const bool keep = (mk.x >> lane) & 1u;
#pragma unroll 1
    for (uint32_t todo = mk.y; todo; todo &= todo - 1u) {
      const int v = __ffs(todo) - 1;
      uint2 ev = e[0];
#pragma unroll
      for (int u = 0; u < NVEC; ++u) if (u == v) ev = e[u];
      uint32_t dv = dev_slow<kCompact>(__ldg(vecref(v).runs + p), ev, bbase + vecref(v).row_base, vecref(v), A.sbits);
      if (v < NA && !keep) dv = kDevBias;
#pragma unroll
      for (int u = 0; u < NVEC; ++u) if (u == v) dev[u] = dv;
    }
  };
  auto products = [&](const uint32_t (&dev)[NVEC]) {
    double dB[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) dB[c] = biased_to_double(dev[NA + c]);
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const double dA = biased_to_double(dev[a]);
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[a][c] = fma(dA, dB[c], acc[a][c]);
    }
  };

  if (count > 0) {
    stage_block(0);
    if (count > 32) stage_block(1);
    uint2 e[D][NVEC];                                                           // register ring, compile-time indices only
    uint32_t dcur[NVEC];                                                        // ranks of the previous pair (none yet: all products 0)
#pragma unroll
    for (int v = 0; v < NVEC; ++v) dcur[v] = kDevBias;
#pragma unroll
    for (int d = 0; d < D - 1; ++d)
      if (d < count) issue(d, e[d]);
    for (int i0 = 0; i0 < count; i0 += D) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int i = i0 + d;
        if (i < count) {
          if (i + D - 1 < count) issue(i + D - 1, e[(d + D - 1) % D]);
          uint2 mk;
          if constexpr (kPipelined) {
            uint32_t dnext[NVEC];
            ranks(i, e[d], dnext, mk);
            products(dcur);
            if (mk.y != 0u) ranks_slow(i, e[d], dnext, mk);
#pragma unroll
            for (int v = 0; v < NVEC; ++v) dcur[v] = dnext[v];
          } else {
            ranks(i, e[d], dcur, mk);
            if (mk.y != 0u) ranks_slow(i, e[d], dcur, mk);
            products(dcur);
          }
          // block (i >> 5) has been read for the last time after its pair 31 (the issue side is at most D - 1 pairs into the
          // next block): its buffer takes block blk + 2
          if ((i & 31) == 31 && ((i >> 5) + 2) * 32 < count) stage_block((i >> 5) + 2);
        }
      }
    }
    if constexpr (kPipelined) products(dcur);
  }
  // CTA reduction.  Exact mode: a warp's sums are integers below 2^53 (the host sized the slices for it); they are added
  // as int64.  Otherwise fp64 in warp order.
  unsigned long long* out = static_cast<unsigned long long*>(A.partial) + ((size_t)b * A.splits + sp) * (NA * NB) * 32;
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      red[warp][lane] = A.exact ? (unsigned long long)(long long)acc[a][c] : (unsigned long long)__double_as_longlong(acc[a][c]);
      __syncthreads();
      if (warp == 0) {
        unsigned long long r;
        if (A.exact) {
          long long t = 0;
#pragma unroll
          for (int w = 0; w < kSweepWarps; ++w) t += (long long)red[w][lane];
          r = (unsigned long long)t;
        } else {
          double t = 0.0;
#pragma unroll
          for (int w = 0; w < kSweepWarps; ++w) t += __longlong_as_double((long long)red[w][lane]);
          r = (unsigned long long)__double_as_longlong(t);
        }
        out[(a * NB + c) * 32 + lane] = r;
      }
      __syncthreads();
    }
}

// ---- one pair of vectors, batch-innermost rows (sweep_batch_inner) -----------------------------------------------------
// A warp takes a canonical pair and walks ALL batches of the job for it: the pair's run words are decoded once, and the rows of the
// batches for one (vector, group) are one contiguous burst.  grid (splits), kSweepThreads; each warp a contiguous share of pairs.
constexpr int kPairBatches = 8;
struct SweepPairArgs {
  SweepVec vec[2];
  int slot[2];
  int n;
  uint32_t N, sbits;
  SweepJob job;
  int splits;
  void* partial;                 // [batch][split][1][32]
  int exact;
};
// The run [s, e) of a pair from its run word (the head tables only for runs the word cannot hold).
template <bool kCompact>
__device__ __forceinline__ void run_bounds(uint32_t w, const SweepVec& V, uint32_t sbits, uint32_t& s, uint32_t& e) {
  if (kCompact) {
    const uint32_t gs = (w & kRunOffMask) >> 3, shlo = w >> kRunShloShift, she = (w >> kRunSheShift) & 63u;
    if (shlo != 0u) {
      s = gs * 32u + (32u - shlo);
      if (w & kRunSingleBit) e = s + 1u;
      else if (w & kRunCrossBit) e = (gs + 1u) * 32u + (32u - she);
      else e = gs * 32u + (32u - she);                                          // she = 0: the group's end
    } else {
      const uint32_t ls = she - 1u;
      s = gs * 32u + ls;
      const uint32_t above = ls == 31u ? 0u : (__ldg(V.head_bits + gs) >> (ls + 1u));
      e = above ? s + (uint32_t)__ffs(above) : __ldg(V.next_head + gs + 1);
    }
  } else {
    const uint32_t smask = (1u << sbits) - 1u, esc = (1u << (32u - sbits)) - 1u;
    const uint32_t len1 = w >> sbits;
    s = w & smask;
    e = s + len1 + 1u;
    if (len1 == esc) {
      const uint32_t g = s >> 5, ls = s & 31u;
      const uint32_t above = ls == 31u ? 0u : (__ldg(V.head_bits + g) >> (ls + 1u));
      e = above ? s + (uint32_t)__ffs(above) : __ldg(V.next_head + g + 1);
    }
  }
}

// Per pair and batch: dev + 2^31 = T(gs) + (T(ge) - T(gs)) / 2 + popc(kept(gs) below s) + popc(kept(ge) below e), ge the group
// of position e - 1 -- one formula for every run shape; the two rows coincide unless the run leaves its first group.
template <bool kCompact>
__global__ void __launch_bounds__(kSweepThreads, 2)
sweep_pair_kernel(const __grid_constant__ SweepPairArgs A) {
  __shared__ uint32_t kws[kSweepWarps][kPairBatches][32];                       // kept words of the block's pairs, per batch
  __shared__ unsigned long long red[kSweepWarps][32];
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const SweepJob& J = A.job;
  const int nb = J.batches;                                                   // <= kPairBatches
  const int sp = blockIdx.x;
  const uint32_t per = (A.N + A.splits - 1) / A.splits;
  const uint32_t c_lo = min(A.N, sp * per), c_hi = min(A.N, c_lo + per);
  const uint32_t wper = ((c_hi - c_lo) + kSweepWarps - 1) / kSweepWarps;
  const uint32_t p_lo = min(c_hi, c_lo + warp * wper), p_hi = min(c_hi, p_lo + wper);
  const char* const rows0 = reinterpret_cast<const char*>(J.rows) + (size_t)A.slot[0] * J.slot_stride + lane * 8;
  const char* const rows1 = reinterpret_cast<const char*>(J.rows) + (size_t)A.slot[1] * J.slot_stride + lane * 8;
  double acc[kPairBatches];
#pragma unroll
  for (int bb = 0; bb < kPairBatches; ++bb) acc[bb] = 0.0;

  for (uint32_t p0 = p_lo; p0 < p_hi; p0 += 32u) {
    const uint32_t cnt = min(32u, p_hi - p0);
    const uint32_t p = min(p0 + lane, A.N - 1u);
    // lane = pair: run bounds of both vectors, kept words of every batch
    uint32_t g_l[4], m_l[4];                                                    // {gs, ge, below-s mask, below-e mask} x {A, B}
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      uint32_t s_, e_;
      run_bounds<kCompact>(__ldg(A.vec[v].runs + p), A.vec[v], A.sbits, s_, e_);
      const uint32_t ge = (e_ - 1u) >> 5, le = e_ - ge * 32u;                   // le in 1..32
      g_l[2 * v] = s_ >> 5; g_l[2 * v + 1] = ge;
      m_l[2 * v] = (1u << (s_ & 31u)) - 1u;
      m_l[2 * v + 1] = le >= 32u ? 0xFFFFFFFFu : (1u << le) - 1u;
    }
    {
      const uint32_t i = triu_row(p, (uint32_t)A.n);
      const uint32_t j = p - i * (2u * A.n - i - 1u) / 2u + i + 1u;
      __syncwarp();
#pragma unroll
      for (int bb = 0; bb < kPairBatches; ++bb)
        if (bb < nb) kws[warp][bb][lane] = __ldg(J.maskT + (size_t)bb * A.n + i) & __ldg(J.maskT + (size_t)bb * A.n + j);
      __syncwarp();
    }
#pragma unroll 1
    for (uint32_t k = 0; k < cnt; ++k) {
      uint32_t g[4], m[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { g[q] = __shfl_sync(0xffffffffu, g_l[q], k); m[q] = __shfl_sync(0xffffffffu, m_l[q], k); }
      const char* r0s = rows0 + (size_t)g[0] * J.group_stride;
      const char* r1s = rows1 + (size_t)g[2] * J.group_stride;
      uint2 a0[kPairBatches], a1[kPairBatches], b0[kPairBatches], b1[kPairBatches];
#pragma unroll
      for (int bb = 0; bb < kPairBatches; ++bb)
        if (bb < nb) { a0[bb] = ldg_row(r0s + (size_t)bb * 256); b0[bb] = ldg_row(r1s + (size_t)bb * 256); }
      // every load is issued before anything waits on one (the copies below would otherwise serialise the two vectors)
      if (g[1] != g[0]) {
        const char* r0e = rows0 + (size_t)g[1] * J.group_stride;
#pragma unroll
        for (int bb = 0; bb < kPairBatches; ++bb) if (bb < nb) a1[bb] = ldg_row(r0e + (size_t)bb * 256);
      }
      if (g[3] != g[2]) {
        const char* r1e = rows1 + (size_t)g[3] * J.group_stride;
#pragma unroll
        for (int bb = 0; bb < kPairBatches; ++bb) if (bb < nb) b1[bb] = ldg_row(r1e + (size_t)bb * 256);
      }
      if (g[1] == g[0]) {
#pragma unroll
        for (int bb = 0; bb < kPairBatches; ++bb) a1[bb] = a0[bb];
      }
      if (g[3] == g[2]) {
#pragma unroll
        for (int bb = 0; bb < kPairBatches; ++bb) b1[bb] = b0[bb];
      }
#pragma unroll
      for (int bb = 0; bb < kPairBatches; ++bb)
        if (bb < nb) {
          const bool keep = (kws[warp][bb][k] >> lane) & 1u;
          const uint32_t dA = a0[bb].y + ((a1[bb].y - a0[bb].y) >> 1) + __popc(a0[bb].x & m[0]) + __popc(a1[bb].x & m[1]);
          const uint32_t dB = b0[bb].y + ((b1[bb].y - b0[bb].y) >> 1) + __popc(b0[bb].x & m[2]) + __popc(b1[bb].x & m[3]);
          acc[bb] = fma(biased_to_double(keep ? dA : kDevBias), biased_to_double(dB), acc[bb]);
        }
    }
  }
  unsigned long long* out = static_cast<unsigned long long*>(A.partial);
#pragma unroll 1
  for (int bb = 0; bb < nb; ++bb) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < kPairBatches; ++q) if (q == bb) v = acc[q];
    red[warp][lane] = A.exact ? (unsigned long long)(long long)v : (unsigned long long)__double_as_longlong(v);
    __syncthreads();
    if (warp == 0) {
      unsigned long long r;
      if (A.exact) {
        long long t = 0;
#pragma unroll
        for (int w = 0; w < kSweepWarps; ++w) t += (long long)red[w][lane];
        r = (unsigned long long)t;
      } else {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kSweepWarps; ++w) t += __longlong_as_double((long long)red[w][lane]);
        r = (unsigned long long)__double_as_longlong(t);
      }
      out[((size_t)bb * A.splits + sp) * 32 + lane] = r;
    }
    __syncthreads();
  }
}

// rho for every (a, c, replicate): slice sums added in slice order (int64 when exact), then
// sum(dev_a dev_c) / (sqrt(ss_a) sqrt(ss_c)).  grid (batches, NA * NB), 32 threads.
struct FinalizeArgs {
  const unsigned long long* partial;
  int splits, nA, nB, R, batches;
  const uint32_t* nk;
  const long long* tie;            // [slots][batches * 32]
  int slotA[kSweepMax], slotB[kSweepMax];
  const void* blobA[kSweepMax];
  const void* blobB[kSweepMax];
  double* out;
  int64_t strideA, strideB;
  int exact;
  int validA, validB;              // vectors that are not padding aliases
  SweepJob job;                    // row addressing (per-replicate NaN test)
  uint32_t N;
};
__global__ void __launch_bounds__(32)
sweep_finalize_kernel(const __grid_constant__ FinalizeArgs F) {
  const int b = blockIdx.x, pair = blockIdx.y, lane = threadIdx.x;
  const int a = pair / F.nB, c = pair % F.nB;
  const int r = b * 32 + lane;
  if (r >= F.R || a >= F.validA || c >= F.validB) return;
  const unsigned long long* p = F.partial + ((size_t)b * F.splits) * (F.nA * F.nB) * 32 + (size_t)pair * 32 + lane;
  double tot;
  if (F.exact) {
    long long t = 0;
    for (int s = 0; s < F.splits; ++s) t += (long long)p[(size_t)s * (F.nA * F.nB) * 32];
    tot = (double)t;
  } else {
    tot = 0.0;
    for (int s = 0; s < F.splits; ++s) tot += __longlong_as_double((long long)p[(size_t)s * (F.nA * F.nB) * 32]);
  }
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  auto ss_of = [&](int slot, const void* blob) -> double {
    const double nr = (double)F.nk[r];
    const long long tie = F.tie[((size_t)slot * F.batches + b) * 32 + lane];      // int64, or fp64 bits when not exact
    double ss;
    if (F.exact) {
      const unsigned long long n_ = F.nk[r];
      ss = (double)(((n_ * n_ * n_ - n_) - (unsigned long long)tie) / 3ull);               // exact integer
    } else {
      ss = (nr * nr * nr - nr - __longlong_as_double(tie)) / 3.0;
    }
    const PresortHeader* hdr = static_cast<const PresortHeader*>(blob);
    if (hdr->has_nan) {                              // NaN only for a replicate that keeps a NaN entry (rsa.py:123-125)
      const uint32_t n_kept = F.nk[r];
      const uint32_t kn = kept_nans(hdr, F.N, n_kept, [&](uint32_t x) {
        const uint2 e = job_row(F.job, b, slot, x >> 5)[lane];
        return ((e.y - kDevBias + n_kept) >> 1) + (uint32_t)__popc(e.x & ((1u << (x & 31u)) - 1u)); });
      if (kn != 0u) ss = nan;
    }
    return ss;
  };
  const double ssa = ss_of(F.slotA[a], F.blobA[a]), ssc = ss_of(F.slotB[c], F.blobB[c]);
  double rho = tot / (sqrt(ssc) * sqrt(ssa));
  if (!(ssa > 0.0) || !(ssc > 0.0)) rho = nan;
  F.out[(int64_t)a * F.strideA + (int64_t)c * F.strideB + r] = rho;
}

// ---- host side -----------------------------------------------------------------------------------------------------
static inline int sweep_splits(uint32_t N, int batches, bool exact, int ctas_per_sm = 1) {
  // CTAs are numbered batch-major and one CTA is resident per SM, so with a multiple of the SM count per batch the CTAs in
  // flight work on ONE batch block at a time (64 MB at the NSD shape: L2 resident; with 37 splits four blocks were in flight
  // and 77 % of the rows came from DRAM).  A warp's share of pairs must keep its fp64 partial sums exact (N^2 * share < 2^53).
  const int64_t wave = (int64_t)kNumSMs * ctas_per_sm;
  int64_t splits = wave;
  if (exact) {
    const double limit = 9007199254740992.0 / ((double)N * (double)N);
    const int64_t share = limit > 1e9 ? 1000000000 : (int64_t)limit;
    const int64_t need = ((int64_t)N + share * kSweepWarps - 1) / (share * kSweepWarps);
    while (splits < need) splits += wave;
  }
  const int64_t cap = ((int64_t)N + 32 * kSweepWarps - 1) / (32 * kSweepWarps);      // at least a block of pairs per warp
  if (splits > cap) splits = cap > 0 ? cap : 1;
  (void)batches;
  return (int)splits;
}
template <int NA, int NB>
static int launch_sweep(const SweepArgs& args, bool compact, cudaStream_t stream) {
  dim3 grid((unsigned)(args.batches * args.splits));
  constexpr size_t smem = SweepSmem<NA + NB>::kBytes;
  static unsigned long long attr_c = 0, attr_g = 0;
  if (compact) {
    if (smem > 40 * 1024) VR_CHECK_CUDA(ensure_dynamic_smem(sweep_kernel<NA, NB, true>, (int)smem, &attr_c));
    sweep_kernel<NA, NB, true><<<grid, kSweepThreads, smem, stream>>>(args);
  } else {
    if (smem > 40 * 1024) VR_CHECK_CUDA(ensure_dynamic_smem(sweep_kernel<NA, NB, false>, (int)smem, &attr_g));
    sweep_kernel<NA, NB, false><<<grid, kSweepThreads, smem, stream>>>(args);
  }
  VR_CHECK_LAUNCH();
  return VR_OK;
}
// Instantiated for 1, 2, 4 and 8 vectors per side; other counts are padded with aliases of the last vector (their
// products are never read).
static inline int sweep_pad(int k) { return k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : 8; }
template <int NA>
static int launch_sweep_b(const SweepArgs& args, int nB, bool compact, cudaStream_t stream) {
  switch (sweep_pad(nB)) {
    case 1: return launch_sweep<NA, 1>(args, compact, stream);
    case 2: return launch_sweep<NA, 2>(args, compact, stream);
    case 4: return launch_sweep<NA, 4>(args, compact, stream);
    default: return launch_sweep<NA, 8>(args, compact, stream);
  }
}
static int launch_sweep_ab(const SweepArgs& args, int nA, int nB, bool compact, cudaStream_t stream) {
  switch (sweep_pad(nA)) {
    case 1: return launch_sweep_b<1>(args, nB, compact, stream);
    case 2: return launch_sweep_b<2>(args, nB, compact, stream);
    case 4: return launch_sweep_b<4>(args, nB, compact, stream);
    default: return launch_sweep_b<8>(args, nB, compact, stream);
  }
}

}  // namespace vr

using namespace vr;

extern "C" size_t vr_spearman_job_bytes(int n, int slots, int R) {
  return (n >= 2 && slots >= 1 && R >= 1) ? sweep_job_bytes(n, slots, R) : 256;
}

extern "C" int vr_spearman_job_init(const uint32_t* masks, int n, int R, int slots, void* job, size_t job_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(masks && job, "vr_spearman_job_init: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && R >= 1 && slots >= 1, "vr_spearman_job_init: bad sizes n=%d R=%d slots=%d", n, R, slots);
  if (job_bytes < sweep_job_bytes(n, slots, R)) { set_error("vr_spearman_job_init: job buffer %zu < %zu", job_bytes, sweep_job_bytes(n, slots, R)); return VR_ERR_WORKSPACE; }
  SweepJob J = sweep_job_view(job, n, slots, R);
  VR_REQUIRE(J.batch_inner || (size_t)slots * J.slot_stride < ((size_t)1 << 32), "vr_spearman_job_init: %d vectors x %u groups exceed the 4 GB batch block", slots, J.Gs);
  const int W = vr_mask_words(n);
  ProfScope prof(F_KEPT, stream);
  sweep_masks_kernel<<<dim3((n + 255) / 256, J.batches), 256, 0, stream>>>(masks, R, n, W, J.maskT);
  VR_CHECK_LAUNCH();
  sweep_nk_kernel<<<(J.batches * 32 + 255) / 256, 256, 0, stream>>>(masks, R, n, W, J.batches * 32, J.nk);
  VR_CHECK_LAUNCH();
  VR_CHECK_CUDA(cudaMemsetAsync(J.tie, 0, (size_t)slots * J.batches * 32 * 8, stream));
  return VR_OK;
}

extern "C" int vr_spearman_job_add(void* job, int n, int R, int slots, const void* const* blobs_host, int count, int slot0,
                                   vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(job && blobs_host, "vr_spearman_job_add: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && R >= 1 && count >= 1 && slot0 >= 0 && slot0 + count <= slots,
             "vr_spearman_job_add: bad sizes n=%d R=%d slots=%d slot0=%d count=%d", n, R, slots, slot0, count);
  SweepJob J = sweep_job_view(job, n, slots, R);
  const uint32_t blocks = (J.G + 31u) / 32u;
  static unsigned long long attr_a = 0, attr_b = 0;
  const size_t smem = (size_t)n * 4;
  if (smem > 40 * 1024) {
    VR_CHECK_CUDA(ensure_dynamic_smem(count_totals_kernel, 160 * 1024, &attr_a));
    VR_CHECK_CUDA(ensure_dynamic_smem(count_rows_kernel, 160 * 1024, &attr_b));
  }
  for (int c0 = 0; c0 < count; c0 += 2 * kSweepMax) {
    const int cc = count - c0 < 2 * kSweepMax ? count - c0 : 2 * kSweepMax;
    BlobPtrs bp{};
    for (int k = 0; k < cc; ++k) {
      if (!blobs_host[c0 + k]) { set_error("vr_spearman_job_add: null presorted blob"); return VR_ERR_INVALID; }
      bp.p[k] = blobs_host[c0 + k];
    }
    {
      ProfScope prof(F_KEPT, stream);
      const int segments = sweep_segments(n, cc, J.batches);                 // <= J.segments (the pitch of the scratch)
      const uint32_t seg_groups = ((blocks + segments - 1) / segments) * 32u;
      dim3 grid(segments / kCountWarpsRows, cc, J.batches);
      count_totals_kernel<<<grid, kCountWarpsRows * 32, smem, stream>>>(bp, slot0 + c0, J, seg_groups);
      VR_CHECK_LAUNCH();
      count_rows_kernel<<<grid, kCountWarpsRows * 32, smem, stream>>>(bp, slot0 + c0, J, seg_groups);
      VR_CHECK_LAUNCH();
    }
    {
      ProfScope prof(F_RANK, stream);
      uint32_t chunks = (J.G + 63u) / 64u;                           // 8 warps x 8 groups per CTA at least
      if (chunks > 64u) chunks = 64u;
      if (chunks < 1u) chunks = 1u;
      if (sweep_exact(J.N)) tie_sums_kernel<true><<<dim3(chunks, cc, J.batches), 256, 0, stream>>>(bp, slot0 + c0, J);
      else tie_sums_kernel<false><<<dim3(chunks, cc, J.batches), 256, 0, stream>>>(bp, slot0 + c0, J);
      VR_CHECK_LAUNCH();
    }
  }
  return VR_OK;
}

extern "C" size_t vr_spearman_sweep_workspace_bytes(int n, int nA, int nB, int R) {
  if (n < 2 || nA < 1 || nB < 1 || R < 1) return 256;
  const uint32_t N = (uint32_t)triu_len(n);
  const int batches = (R + 31) / 32;
  const int splits = sweep_splits(N, batches, sweep_exact(N), 2);
  return arena_bytes((size_t)batches * splits * kSweepMax * kSweepMax * 32, 8) + 256;
}

extern "C" int vr_spearman_job_sweep(void* job, int n, int R, int slots, const void* const* blobsA_host, const int* slotsA_host, int nA,
                                     const void* const* blobsB_host, const int* slotsB_host, int nB, double* out,
                                     int64_t strideA, int64_t strideB, void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(job && blobsA_host && slotsA_host && blobsB_host && slotsB_host && out, "vr_spearman_job_sweep: null pointer");
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI && R >= 1 && nA >= 1 && nB >= 1, "vr_spearman_job_sweep: bad sizes n=%d R=%d nA=%d nB=%d", n, R, nA, nB);
  if (!ws || ws_bytes < vr_spearman_sweep_workspace_bytes(n, nA, nB, R)) {
    set_error("vr_spearman_job_sweep: workspace %zu < %zu", ws_bytes, vr_spearman_sweep_workspace_bytes(n, nA, nB, R));
    return VR_ERR_WORKSPACE;
  }
  SweepJob J = sweep_job_view(job, n, slots, R);
  const bool exact = sweep_exact(J.N);
  const bool compact = run_compact(J.N);
  const int tile_a = option_sweep_tile_a(), tile_b = option_sweep_tile_b();
  for (int k = 0; k < nA; ++k) VR_REQUIRE(slotsA_host[k] >= 0 && slotsA_host[k] < slots && blobsA_host[k], "vr_spearman_job_sweep: bad A slot");
  for (int k = 0; k < nB; ++k) VR_REQUIRE(slotsB_host[k] >= 0 && slotsB_host[k] < slots && blobsB_host[k], "vr_spearman_job_sweep: bad B slot");
  void* partial = ws;
  ProfScope prof(F_RANK_FUSED, stream);
  if (J.batch_inner) {
    // (slots == 2: every sweep of this job is one vector against one vector)
    VR_REQUIRE(J.batches <= kPairBatches, "vr_spearman_job_sweep: internal: %d batches in a batch-inner job", J.batches);
    for (int a = 0; a < nA; ++a)
      for (int c = 0; c < nB; ++c) {
        const int splits = sweep_splits(J.N, J.batches, exact, 2);
        SweepPairArgs pa{};
        const void* bl[2] = {blobsA_host[a], blobsB_host[c]};
        const int sl[2] = {slotsA_host[a], slotsB_host[c]};
        for (int k = 0; k < 2; ++k) {
          PresortView pv = presort_view(const_cast<void*>(bl[k]), n);
          pa.vec[k].runs = pv.runs; pa.vec[k].head_bits = pv.head_bits; pa.vec[k].next_head = pv.next_head; pa.vec[k].row_base = 0u;
          pa.slot[k] = sl[k];
        }
        pa.n = n; pa.N = J.N; pa.sbits = run_sbits(J.N); pa.job = J; pa.splits = splits; pa.partial = partial; pa.exact = exact ? 1 : 0;
        if (compact) sweep_pair_kernel<true><<<splits, kSweepThreads, 0, stream>>>(pa);
        else sweep_pair_kernel<false><<<splits, kSweepThreads, 0, stream>>>(pa);
        VR_CHECK_LAUNCH();
        FinalizeArgs fin{};
        fin.partial = static_cast<const unsigned long long*>(ws); fin.splits = splits; fin.nA = 1; fin.nB = 1; fin.R = R; fin.batches = J.batches;
        fin.nk = J.nk; fin.tie = J.tie; fin.slotA[0] = sl[0]; fin.slotB[0] = sl[1]; fin.blobA[0] = bl[0]; fin.blobB[0] = bl[1];
        fin.out = out + (int64_t)a * strideA + (int64_t)c * strideB; fin.strideA = strideA; fin.strideB = strideB; fin.exact = exact ? 1 : 0;
        fin.validA = 1; fin.validB = 1; fin.job = J; fin.N = J.N;
        sweep_finalize_kernel<<<dim3(J.batches, 1), 32, 0, stream>>>(fin);
        VR_CHECK_LAUNCH();
      }
    return VR_OK;
  }
  for (int a0 = 0; a0 < nA; a0 += tile_a) {
    const int na = nA - a0 < tile_a ? nA - a0 : tile_a;
    for (int b0 = 0; b0 < nB; b0 += tile_b) {
      const int nb = nB - b0 < tile_b ? nB - b0 : tile_b;
      const int splits = sweep_splits(J.N, J.batches, exact, sweep_ctas_per_sm(sweep_pad(na), sweep_pad(nb)));
      SweepArgs args{};
      FinalizeArgs fin{};
      auto fill = [&](SweepVec& v, const void* blob, int slot) {
        PresortView pv = presort_view(const_cast<void*>(blob), n);
        v.runs = pv.runs; v.head_bits = pv.head_bits; v.next_head = pv.next_head;
        v.row_base = (uint32_t)((size_t)slot * J.slot_stride);
      };
      const int pa = sweep_pad(na), pb = sweep_pad(nb);
      for (int k = 0; k < pa; ++k) { const int q = a0 + (k < na ? k : na - 1); fill(args.vec[k], blobsA_host[q], slotsA_host[q]); fin.slotA[k] = slotsA_host[q]; fin.blobA[k] = blobsA_host[q]; }
      for (int k = 0; k < pb; ++k) { const int q = b0 + (k < nb ? k : nb - 1); fill(args.vec[kSweepMax + k], blobsB_host[q], slotsB_host[q]); fin.slotB[k] = slotsB_host[q]; fin.blobB[k] = blobsB_host[q]; }
      args.n = n; args.N = J.N; args.sbits = run_sbits(J.N); args.rows = reinterpret_cast<const char*>(J.rows);
      args.batch_stride = J.batch_stride; args.maskT = J.maskT; args.batches = J.batches; args.splits = splits; args.partial = partial;
      args.exact = exact ? 1 : 0;
      int rc = launch_sweep_ab(args, na, nb, compact, stream);
      if (rc != VR_OK) return rc;
      fin.partial = static_cast<const unsigned long long*>(ws); fin.splits = splits; fin.nA = pa; fin.nB = pb; fin.R = R; fin.batches = J.batches; fin.nk = J.nk; fin.tie = J.tie;
      fin.out = out + (int64_t)a0 * strideA + (int64_t)b0 * strideB; fin.strideA = strideA; fin.strideB = strideB; fin.exact = exact ? 1 : 0;
      fin.validA = na; fin.validB = nb; fin.job = J; fin.N = J.N;
      sweep_finalize_kernel<<<dim3(J.batches, pa * pb), 32, 0, stream>>>(fin);
      VR_CHECK_LAUNCH();
    }
  }
  return VR_OK;
}
