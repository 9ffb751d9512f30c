// Canonical-order kept bits of the RDM pairs for a set of subsample masks (shared by the Spearman
// rank context and the Kendall path).
#pragma once
#include "common.cuh"

namespace vr {

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

// pairbits[r][w]: bit b = canonical pair 32w+b has both stimuli in mask r (0 beyond N).
// One thread per word.  Canonical pairs run along the columns of a row, so a word is a bit field of the mask
// itself -- columns j .. j+31 of row i, gated by stimulus i -- continued in the next row(s) when the row ends.
static __global__ void __launch_bounds__(256)
pair_mask_kernel(const uint32_t* __restrict__ masks, int R, int n, int W, uint32_t words,
                 uint32_t* __restrict__ pairbits) {
  extern __shared__ uint32_t s_m[];                 // W + 1 words: the mask and a zero word behind it
  const int r = blockIdx.y;
  for (int w = threadIdx.x; w <= W; w += 256) s_m[w] = w < W ? masks[(int64_t)r * W + w] : 0u;
  __syncthreads();
  const uint32_t N = (uint32_t)triu_len(n);
  for (uint32_t w = blockIdx.x * 256u + threadIdx.x; w < words; w += gridDim.x * 256u) {
    const uint32_t p = w * 32u;
    uint32_t bits = 0;
    if (p < N) {
      uint32_t i = triu_row(p, (uint32_t)n);
      uint32_t j = p - i * (2u * n - i - 1u) / 2u + i + 1u;
      uint32_t filled = 0;
      while (filled < 32u && i + 1u < (uint32_t)n) {
        const uint32_t len = min(32u - filled, (uint32_t)n - j);                 // pairs of row i that fit
        if ((s_m[i >> 5] >> (i & 31u)) & 1u) {
          const uint32_t field = __funnelshift_r(s_m[j >> 5], s_m[(j >> 5) + 1u], j & 31u);   // mask bits j .. j+31
          bits |= (len == 32u ? field : (field & ((1u << len) - 1u))) << filled;
        }
        filled += len;
        ++i;
        j = i + 1u;
      }
    }
    pairbits[(int64_t)r * words + w] = bits;
  }
}


}  // namespace vr
