// Host-side replay of numpy's legacy RandomState stream for the bootstrap index draws (no CUDA here;
// compiled by g++ so the tempering loop can be multiversioned for AVX2).
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <vector>
#include "../../include/visreps_b200.h"

namespace vr { void set_error(const char* fmt, ...); }
#define VR_REQUIRE(cond, ...) do { if (!(cond)) { vr::set_error(__VA_ARGS__); return VR_ERR_INVALID; } } while (0)

// Exact restatement of numpy's legacy stream for `RandomState.choice(n, m, replace=False)`, which is
// `permutation(n)[:m]`: a Fisher-Yates shuffle from the top, j = random_interval(i) drawn by masked
// rejection from 32-bit MT19937 outputs.  The reference draws one such subsample per bootstrap
// replicate (visreps/evals.py:362-364, visreps/analysis/rsa.py:249-251); doing the B draws here
// instead of in a Python loop takes the host off the critical path.  State in / state out, so a
// RandomState shared with other consumers (rsa.py:169,176) stays in step.
namespace {
struct MT { uint32_t* key; int pos; };
inline void mt_refill(MT& g) {
  constexpr int N = 624, M = 397;
  constexpr uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
  uint32_t* k = g.key;
  int i = 0;
  for (; i < N - M; ++i) { uint32_t y = (k[i] & UP) | (k[i + 1] & LO); k[i] = k[i + M] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
  for (; i < N - 1; ++i) { uint32_t y = (k[i] & UP) | (k[i + 1] & LO); k[i] = k[i + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
  uint32_t y = (k[N - 1] & UP) | (k[0] & LO);
  k[N - 1] = k[M - 1] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  g.pos = 0;
}
__attribute__((target_clones("avx2", "default")))
void temper_range(const uint32_t* key, uint32_t* out, int from) {
  for (int k = from; k < 624; ++k) {
    uint32_t y = key[k];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    out[k] = y;
  }
}
inline uint32_t mt_next32(MT& g) {
  if (g.pos == 624) mt_refill(g);
  uint32_t y = g.key[g.pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
}  // namespace

extern "C" int vr_host_draw_subsamples(uint32_t* mt_key, int* mt_pos, int n, int m, int B, int32_t* out) {
  VR_REQUIRE(mt_key && mt_pos && (out || (int64_t)B * m == 0), "vr_host_draw_subsamples: null pointer");
  VR_REQUIRE(n >= 1 && m >= 0 && m <= n && B >= 0 && *mt_pos >= 0 && *mt_pos <= 624,
             "vr_host_draw_subsamples: bad arguments n=%d m=%d B=%d pos=%d", n, m, B, *mt_pos);
  MT g{mt_key, *mt_pos};
  // Outputs are tempered a whole state block at a time (the loop vectorises); `pos` walks the block
  // exactly as numpy's generator does, so the state handed back is numpy's.
  uint32_t tempered[624];
  auto temper_block = [&](int from) { temper_range(g.key, tempered, from); };
  int pos = g.pos;
  temper_block(pos);
  std::vector<int32_t> arr((size_t)n);
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < n; ++i) arr[i] = i;
    for (uint32_t i = (uint32_t)n - 1; i >= 1; --i) {
      const uint32_t mask = 0xffffffffu >> __builtin_clz(i);     // smallest 2^k - 1 >= i
      uint32_t j;
      do {
        if (pos == 624) { mt_refill(g); temper_block(0); pos = 0; }
        j = tempered[pos++] & mask;
      } while (j > i);
      const int32_t t = arr[i]; arr[i] = arr[j]; arr[j] = t;
    }
    for (int i = 0; i < m; ++i) out[(size_t)b * m + i] = arr[i];
  }
  *mt_pos = pos;
  return VR_OK;
}

// Same stream, but only what the device needs: the n-bit membership mask of each draw.  The kept set of
// `permutation(n)[:m]` is the complement of the shuffle's last n-m slots, and those are final after the
// first n-m swap steps (i = n-1 .. m); the remaining m-1 steps only have to consume the generator
// exactly as numpy does, which needs no array at all: per 32-bit output, `i -= ((w & mask) <= i)`.
// Replicates outside [b_begin, b_end) (another rank's share) are consumed without any swaps.
// masks: HOST ((b_end - b_begin), vr_mask_words(n)) uint32.
extern "C" int vr_host_draw_subsample_masks(uint32_t* mt_key, int* mt_pos, int n, int m, int B, int b_begin, int b_end,
                                            uint32_t* masks) {
  VR_REQUIRE(mt_key && mt_pos, "vr_host_draw_subsample_masks: null pointer");
  VR_REQUIRE(n >= 1 && m >= 0 && m <= n && B >= 0 && *mt_pos >= 0 && *mt_pos <= 624 && b_begin >= 0 && b_begin <= b_end && b_end <= B,
             "vr_host_draw_subsample_masks: bad arguments n=%d m=%d B=%d range=[%d,%d) pos=%d", n, m, B, b_begin, b_end, *mt_pos);
  VR_REQUIRE(masks || b_begin == b_end, "vr_host_draw_subsample_masks: null mask buffer");
  MT g{mt_key, *mt_pos};
  uint32_t tempered[624];
  int pos = g.pos;
  temper_range(g.key, tempered, pos);
  const int W = (n + 31) / 32;
  std::vector<int32_t> arr((size_t)n);
  for (int b = 0; b < B; ++b) {
    const bool wanted = b >= b_begin && b < b_end && m > 0;
    uint32_t i = (uint32_t)n - 1;
    if (wanted) {
      for (int k = 0; k < n; ++k) arr[k] = k;
      for (; i >= (uint32_t)m && i >= 1; --i) {                    // the slots that end up dropped
        const uint32_t mask = 0xffffffffu >> __builtin_clz(i);
        uint32_t j;
        do {
          if (pos == 624) { mt_refill(g); temper_range(g.key, tempered, 0); pos = 0; }
          j = tempered[pos++] & mask;
        } while (j > i);
        const int32_t t = arr[i]; arr[i] = arr[j]; arr[j] = t;
      }
    }
    while (i >= 1) {                                               // consume-only steps, one mask range at a time
      const uint32_t mask = 0xffffffffu >> __builtin_clz(i);
      const uint32_t lo = (mask >> 1) + 1;                         // mask is constant for i in [lo, mask]
      while (i >= lo) {
        if (pos == 624) { mt_refill(g); temper_range(g.key, tempered, 0); pos = 0; }
        int room = 624 - pos;
        const uint32_t* w = tempered + pos;
        int k = 0;
        for (; k < room && i >= lo; ++k) i -= ((w[k] & mask) <= i) ? 1u : 0u;
        pos += k;
      }
    }
    if (b >= b_begin && b < b_end) {
      uint32_t* row = masks + (size_t)(b - b_begin) * W;
      if (m == 0) { for (int k = 0; k < W; ++k) row[k] = 0u; continue; }
      for (int k = 0; k < W; ++k) row[k] = 0xffffffffu;
      if (n % 32) row[W - 1] = (1u << (n % 32)) - 1u;
      for (int k = m; k < n; ++k) row[arr[k] >> 5] &= ~(1u << (arr[k] & 31));
    }
  }
  *mt_pos = pos;
  return VR_OK;
}
