// Stable LSD radix sort of (uint32 key, uint32 payload) pairs, batched over independent
// segments.  Hand-written for sm_100a; no library sort is used anywhere in this path.
//
// Per 8-bit pass:  tile_count  ->  scan_counts  ->  tile_scatter
//   - a tile is kSortTile consecutive elements handled by one CTA; warp w owns a contiguous
//     slice of the tile and visits it in order, so ranking inside a digit is stable;
//   - equal digits inside a 32-element row are resolved with __match_any_sync + popc;
//   - per-(segment, digit, tile) counts are scanned digit-major, which yields each tile's
//     global base per digit;
//   - the scatter stages the tile in shared memory in digit order and writes it out from there, so
//     every digit's run leaves as consecutive addresses.
// HBM traffic per pass: 2 reads + 1 write of the 8-byte pairs (24 B/element).
#pragma once
#include "common.cuh"

namespace vr {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;                          // rows of 32 per warp
constexpr int kSortTile = kSortThreads * kSortItems;    // 4096 elements per CTA
constexpr int kRadix = 256;

inline int sort_num_tiles(int64_t len) { return (int)((len + kSortTile - 1) / kSortTile); }
// uint32 counts: [segments][256][tiles]
inline size_t sort_count_elems(int64_t len, int segments) {
  return (size_t)segments * kRadix * sort_num_tiles(len);
}

static __global__ void __launch_bounds__(kSortThreads)
tile_count_kernel(const uint32_t* __restrict__ keys, int64_t len, int64_t seg_stride, int shift,
                  uint32_t* __restrict__ counts, int tiles) {
  __shared__ uint32_t hist[kRadix];
  const int tile = blockIdx.x, seg = blockIdx.y;
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t* k = keys + (int64_t)seg * seg_stride;
  const int64_t base = (int64_t)tile * kSortTile;
#pragma unroll 4
  for (int it = 0; it < kSortItems; ++it) {
    int64_t idx = base + (int64_t)it * kSortThreads + threadIdx.x;
    if (idx < len) atomicAdd(&hist[(k[idx] >> shift) & 0xFFu], 1u);
  }
  __syncthreads();
  counts[((size_t)seg * kRadix + threadIdx.x) * tiles + tile] = hist[threadIdx.x];
}

// Exclusive scan, in place, of each segment's [256][tiles] count table (digit-major).
static __global__ void __launch_bounds__(1024)
scan_counts_kernel(uint32_t* __restrict__ counts, int tiles) {
  __shared__ uint32_t warp_tot[32];
  uint32_t* c = counts + (size_t)blockIdx.x * kRadix * tiles;
  const int64_t M = (int64_t)kRadix * tiles;
  const int64_t chunk = (M + 1023) / 1024;
  const int64_t lo = (int64_t)threadIdx.x * chunk;
  const int64_t hi = lo + chunk < M ? lo + chunk : M;
  uint32_t sum = 0;
  for (int64_t i = lo; i < hi; ++i) sum += c[i];
  // block exclusive scan of `sum`
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane_id() >= (unsigned)o) incl += t;
  }
  if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = warp_tot[threadIdx.x], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane_id() >= (unsigned)o) wi += t;
    }
    warp_tot[threadIdx.x] = wi - w;
  }
  __syncthreads();
  uint32_t run = warp_tot[threadIdx.x >> 5] + incl - sum;
  for (int64_t i = lo; i < hi; ++i) {
    uint32_t t = c[i];
    c[i] = run;
    run += t;
  }
}

static __global__ void __launch_bounds__(kSortThreads)
tile_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                    int64_t len, int64_t seg_stride, int shift,
                    const uint32_t* __restrict__ bases, int tiles) {
  __shared__ uint32_t cnt[kSortWarps][kRadix];     // per-warp digit counts -> tile-local start of (warp, digit)
  __shared__ uint32_t dstart[kRadix];              // tile-local start of each digit
  __shared__ uint32_t gbase[kRadix];               // global position of the tile's first element of each digit
  __shared__ uint32_t skey[kSortTile];             // the tile, locally sorted by digit (stable)
  __shared__ uint32_t sval[kSortTile];
  __shared__ uint32_t wtot[kSortWarps];
  const int tile = blockIdx.x, seg = blockIdx.y;
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();

  const int64_t seg_off = (int64_t)seg * seg_stride;
  const uint32_t* kin = keys_in + seg_off;
  const uint32_t* vin = vals_in + seg_off;
  const int64_t tile_base = (int64_t)tile * kSortTile;
  const int64_t base = tile_base + (int64_t)warp * (kSortItems * 32);

  uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    int64_t idx = base + r * 32 + lane;
    bool ok = idx < len;
    key[r] = ok ? kin[idx] : 0xFFFFFFFFu;
    val[r] = ok ? vin[idx] : 0u;
  }
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    int64_t idx = base + r * 32 + lane;
    bool ok = idx < len;
    uint32_t digit = ok ? ((key[r] >> shift) & 0xFFu) : 0x100u;
    uint32_t peers = __match_any_sync(0xffffffffu, digit);
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (ok && (int)lane == leader) {
      old = cnt[warp][digit];
      cnt[warp][digit] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & lanemask_lt());
    __syncwarp();
  }
  __syncthreads();
  // per digit (one thread each): the tile's count, and the warps' counts turned into offsets inside the digit
  uint32_t dcount;
  {
    const int digit = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      uint32_t t = cnt[w][digit];
      cnt[w][digit] = run;
      run += t;
    }
    dcount = run;
    gbase[digit] = bases[((size_t)seg * kRadix + digit) * tiles + tile];
  }
  // exclusive scan of the 256 digit counts -> dstart
  {
    uint32_t incl = dcount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) before += w < warp ? wtot[w] : 0u;
    dstart[threadIdx.x] = before + incl - dcount;
  }
  __syncthreads();
  // stage the tile in digit order, then write it out: consecutive threads write consecutive addresses inside every
  // digit's run instead of 32 scattered 8-byte pieces per warp
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    int64_t idx = base + r * 32 + lane;
    if (idx < len) {
      const uint32_t d = (key[r] >> shift) & 0xFFu;
      const uint32_t p = dstart[d] + cnt[warp][d] + rank[r];
      skey[p] = key[r];
      sval[p] = val[r];
    }
  }
  __syncthreads();
  uint32_t* kout = keys_out + seg_off;
  uint32_t* vout = vals_out + seg_off;
  const int64_t remaining = len - tile_base;
  const uint32_t valid = (uint32_t)(remaining < kSortTile ? remaining : kSortTile);
  for (uint32_t p = threadIdx.x; p < valid; p += kSortThreads) {
    const uint32_t k = skey[p];
    const uint32_t d = (k >> shift) & 0xFFu;
    const uint32_t pos = gbase[d] + (p - dstart[d]);
    kout[pos] = k;
    vout[pos] = sval[p];
  }
}

// Sorts `segments` independent runs of `len` pairs (segment s starts at s*seg_stride).
// keys0/vals0 hold the input; the sorted result ends in keys0/vals0 again (4 passes).
// `counts` needs sort_count_elems(len, segments) uint32.
inline int radix_sort_pairs(uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1,
                            int64_t len, int64_t seg_stride, int segments, uint32_t* counts,
                            cudaStream_t stream) {
  if (len <= 0 || segments <= 0) return VR_OK;
  const int tiles = sort_num_tiles(len);
  dim3 grid(tiles, segments);
  uint32_t *ki = keys0, *vi = vals0, *ko = keys1, *vo = vals1;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    tile_count_kernel<<<grid, kSortThreads, 0, stream>>>(ki, len, seg_stride, shift, counts, tiles);
    VR_CHECK_LAUNCH();
    scan_counts_kernel<<<segments, 1024, 0, stream>>>(counts, tiles);
    VR_CHECK_LAUNCH();
    tile_scatter_kernel<<<grid, kSortThreads, 0, stream>>>(ki, vi, ko, vo, len, seg_stride, shift,
                                                           counts, tiles);
    VR_CHECK_LAUNCH();
    uint32_t* t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  return VR_OK;
}

}  // namespace vr
