// Shared host/device helpers for the visreps_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/visreps_b200.h"

namespace vr {

// ---- error plumbing ------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(unsigned n = 1);
// Tuning options (vr_set_option; initial values from VR_SYRK_PAIR / VR_SYRK_WAVES, for A/B measurements and tests).
int option_syrk_pair();
int option_syrk_waves();
int option_sweep_tile_a();   // vectors per side in one Spearman sweep launch (1..8)
int option_sweep_tile_b();
int option_kendall_tree();         // Kendall bootstrap: the once-per-pair merge tree (sort-free replicates): 0 off, 1 when it pays, 2 always
int option_kendall_complement();   // Kendall bootstrap: count among the dropped entries (1, default) or the kept ones (0)

// ---- built-in stage profiler (CUDA events on the launching stream) ---------------------
// Off by default.  Between vr_profile_start() and vr_profile_stop() every ProfScope records a
// start/stop event pair; stop() synchronises them and returns per-family device time and
// kernel-launch counts.  bench.py uses it to time the dominant kernel inside the timed region.
enum Family {
  F_MASKS = 0, F_PRESORT, F_KEPT, F_RANK, F_RANK_FUSED, F_PEARSON, F_KENDALL, F_RDM_CENTER, F_RDM_SYRK, F_RDM_SYRK_PAIR, F_RDM_EPILOGUE, F_ROWRANK, F_COLCORR,
  F_NUM_FAMILIES
};
struct ProfScope {
  int fam;
  cudaStream_t stream;
  int slot;
  ProfScope(int family, cudaStream_t s);
  ~ProfScope();
};

#define VR_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      vr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                    __LINE__);                                                           \
      return VR_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define VR_CHECK_LAUNCH()                                                                \
  do {                                                                                   \
    vr::count_launch();                                                                  \
    VR_CHECK_CUDA(cudaGetLastError());                                                   \
  } while (0)

#define VR_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      vr::set_error(__VA_ARGS__);                                                        \
      return VR_ERR_INVALID;                                                             \
    }                                                                                    \
  } while (0)

// Opt-in dynamic shared memory is a per-device function attribute: `done` is a per-kernel bit set of the
// devices already configured (a process may drive several GPUs).
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, unsigned long long* done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && ((*done >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev < 64) *done |= 1ull << dev;
  return e;
}

// ---- sizes ---------------------------------------------------------------------------
__host__ __device__ inline int64_t triu_len(int64_t n) { return n * (n - 1) / 2; }
__host__ __device__ inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
constexpr int kNumSMs = 148;  // B200

// Bump allocator over the caller's workspace; every block is 256-byte aligned.
struct Arena {
  char* base;
  size_t cap;
  size_t off = 0;
  bool ok = true;
  Arena(void* p, size_t bytes) : base(static_cast<char*>(p)), cap(bytes) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
    if (off + bytes > cap) { ok = false; off += bytes; return nullptr; }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};
inline size_t arena_bytes(size_t count, size_t elem) { return (count * elem + 255) / 256 * 256; }

// ---- packed (i, j) pair word of the presorted stream -----------------------------------
//  bits  0..14  column j
//  bits 15..29  row    i   (i < j)
//  bit  30      head: first element of a run of equal values in sorted order
//  bit  31      pad : position >= N in the last 32-group (never kept, never written)
constexpr uint32_t kHeadBit = 1u << 30;
constexpr uint32_t kPadBit = 1u << 31;
__host__ __device__ inline uint32_t pack_pair(uint32_t i, uint32_t j) { return (i << 15) | j; }
__host__ __device__ inline uint32_t pair_i(uint32_t w) { return (w >> 15) & 0x7FFFu; }
__host__ __device__ inline uint32_t pair_j(uint32_t w) { return w & 0x7FFFu; }
// Row-major position of (i<j) in the strict upper triangle of an n x n matrix.
__host__ __device__ inline uint32_t triu_pos(uint32_t i, uint32_t j, uint32_t n) {
  return i * (2u * n - i - 1u) / 2u + (j - i - 1u);
}

// ---- presorted blob layout --------------------------------------------------------------
//  [0, 256)          header
//  stream            uint32[G*32]   packed pairs in ascending value order (+pad words)
//  runs              uint32[G*32]   per canonical pair p: the run [s, e) of equal values holding it,
//                                   packed as s | min(e-s-1, ESC) << sbits   (see run_sbits)
//  head_bits         uint32[G]      bit l of word g: sorted position 32g+l starts a run
//                                   (position N, if it falls inside the last group, is set too)
//  prev_head         uint32[G]      last head position <= 32g+31
//  next_head         uint32[G+1]    first head position >= 32g   (N if none; [G] = N)
struct PresortHeader {
  uint32_t magic;
  uint32_t n;
  uint32_t N;        // n(n-1)/2
  uint32_t G;        // ceil(N/32)
  uint32_t has_nan;  // written on the device
  uint32_t n_runs;   // number of distinct values (device)
  uint32_t nan_lo;   // NaNs with the sign bit set: they sort to positions [0, nan_lo)            (device)
  uint32_t nan_hi;   // NaNs without it: positions [N - nan_hi, N).  A replicate's scores are NaN only if it KEEPS one of them
};
// Kept entries among the NaN positions of a presorted vector, given C(x) = kept entries at sorted positions < x.
template <typename CountBefore>
__host__ __device__ inline uint32_t kept_nans(const PresortHeader* hdr, uint32_t N, uint32_t n_kept, CountBefore count_before) {
  uint32_t k = 0;
  if (hdr->nan_lo) k += count_before(hdr->nan_lo);
  if (hdr->nan_hi) k += n_kept - count_before(N - hdr->nan_hi);
  return k;
}
constexpr uint32_t kPresortMagic = 0x56525053u;  // 'VRPS'
struct PresortView {
  PresortHeader* hdr;
  uint32_t* stream;
  uint32_t* runs;
  uint32_t* head_bits;
  uint32_t* prev_head;
  uint32_t* next_head;
};
// Bits that hold a sorted position (< N); the remaining high bits of a `runs` word hold the run
// length minus one, saturating at the escape code (all ones): longer runs are resolved from the
// head tables.
__host__ __device__ inline uint32_t run_sbits(uint32_t N) {
  uint32_t b = 1;
  while (b < 32 && (1ull << b) < (unsigned long long)N + 1ull) ++b;
  return b;                 // n <= VR_MAX_STIMULI gives N < 2^29, so at least 3 length bits remain
}
// Compact run word, used while the table index fits 16 bits (N <= 2,097,120 pairs, n <= 2048):
//   [ 000 ][ gs : 16 ][ cross : 1 ][ she : 6 ][ shlo : 6 ]
// for the run [s, e) of equal values that holds the pair:  gs = s >> 5, shlo = 32 - (s & 31), and the end e is
// described relative to the same table entry: cross = 0 when e lies in group gs (she = 32 - (e & 31)) or is
// exactly the next group boundary (she = 0: every bit of the group counts); cross = 1 when e lies inside
// group gs + 1.  With table entries {kept word, kept count before the group},
//     C(s) = cnt[gs] + popc(kept[gs] << shlo),      C(e) = cnt[gs + cross] + popc(kept[gs + cross] << she)
// with no data-dependent branch.  The low field is the BYTE offset of the 8-byte entry and shlo sits in the
// top bits, so each is one instruction away.  A run that ends beyond group gs + 1 (longer than 32) is marked
// shlo = 0 with she = (s & 31) + 1 and resolved from the head tables.  A vector without ties (header
// n_runs == N) has e = s + 1 everywhere: C(s) + C(e) = 2 C(s) + 1 for a kept pair.
__host__ __device__ inline uint32_t run_gbits(uint32_t N) {
  const uint32_t G = (N + 31u) / 32u;
  uint32_t b = 1;
  while ((1u << b) < G + 1u) ++b;
  return b;
}
__host__ __device__ inline bool run_compact(uint32_t N) { return run_gbits(N) <= 16u; }
constexpr uint32_t kRunOffMask = 0x7FFF8u;     // gs << 3
constexpr uint32_t kRunCrossBit = 1u << 19;
constexpr uint32_t kRunSheShift = 20u;
constexpr uint32_t kRunShloShift = 26u;
__host__ __device__ inline uint32_t presort_groups(int n) { return (uint32_t)((triu_len(n) + 31) / 32); }
inline size_t presort_blob_bytes(int n) {
  size_t G = presort_groups(n);
  return 256 + 2 * arena_bytes(G * 32, 4) + 2 * arena_bytes(G, 4) + arena_bytes(G + 1, 4);
}
__host__ __device__ inline PresortView presort_view(void* blob, int n) {
  size_t G = presort_groups(n);
  char* p = static_cast<char*>(blob);
  PresortView v;
  v.hdr = reinterpret_cast<PresortHeader*>(p);
  p += 256;
  v.stream = reinterpret_cast<uint32_t*>(p);
  p += (G * 32 * 4 + 255) / 256 * 256;
  v.runs = reinterpret_cast<uint32_t*>(p);
  p += (G * 32 * 4 + 255) / 256 * 256;
  v.head_bits = reinterpret_cast<uint32_t*>(p);
  p += (G * 4 + 255) / 256 * 256;
  v.prev_head = reinterpret_cast<uint32_t*>(p);
  p += (G * 4 + 255) / 256 * 256;
  v.next_head = reinterpret_cast<uint32_t*>(p);
  return v;
}

// ---- device helpers ------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint32_t lanemask_le() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint32_t lanemask_gt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
  return m;
}
// Order-preserving map fp32 -> uint32 (ascending), with -0.0 folded onto +0.0.
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (u == 0x80000000u) u = 0u;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace vr
