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
// Per replicate:       the (x, y) order of the kept entries is the stable compaction of the full order
//                      (same fact the Spearman path uses), so no replicate sorts:
//                      compact -> tie sums (run lengths) -> D = inversions of the y sequence, counted by a
//                      bottom-up merge sort in which every element finds its merged position by binary
//                      search in the sibling run (O(M log^2 M) steps, fully parallel, ~20 launches).
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

// kept flag of sorted position k (pair index order[k]) for one replicate, and block-wise counts for the scan.
// Single-CTA exclusive scan + stable compaction of up to two value arrays.  N <= ~2^27.
__global__ void __launch_bounds__(1024)
compact_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ pairbits, uint32_t N,
               const uint32_t* __restrict__ v0, const uint32_t* __restrict__ v1, uint32_t* __restrict__ o0,
               uint32_t* __restrict__ o1, uint32_t* __restrict__ count_out) {
  __shared__ uint32_t part[1024];
  const uint32_t chunk = (N + 1023u) / 1024u;
  const uint32_t lo = threadIdx.x * chunk, hi = min(N, lo + chunk);
  uint32_t c = 0;
  for (uint32_t k = lo; k < hi; ++k) { const uint32_t p = order[k]; c += (pairbits[p >> 5] >> (p & 31u)) & 1u; }
  part[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t loc[32], run = 0;
    for (int t = 0; t < 32; ++t) { loc[t] = part[threadIdx.x * 32 + t]; run += loc[t]; }
    uint32_t incl = run;
    for (int o = 1; o < 32; o <<= 1) { uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (threadIdx.x >= (unsigned)o) incl += x; }
    uint32_t base = incl - run;
    for (int t = 0; t < 32; ++t) { part[threadIdx.x * 32 + t] = base; base += loc[t]; }
    if (threadIdx.x == 31) *count_out = incl;
  }
  __syncthreads();
  uint32_t w = part[threadIdx.x];
  for (uint32_t k = lo; k < hi; ++k) {
    const uint32_t p = order[k];
    if ((pairbits[p >> 5] >> (p & 31u)) & 1u) {
      o0[w] = v0[k];
      if (v1) o1[w] = v1[k];
      ++w;
    }
  }
}

// sum over runs of equal (k0[, k1]) of c(c-1)/2; the thread at a run head walks its run.
__global__ void __launch_bounds__(256)
tie_pairs_kernel(const uint32_t* __restrict__ k0, const uint32_t* __restrict__ k1, const uint32_t* __restrict__ count,
                 unsigned long long* __restrict__ out) {
  const uint32_t M = *count;
  unsigned long long local = 0;
  for (uint32_t k = blockIdx.x * 256u + threadIdx.x; k < M; k += gridDim.x * 256u) {
    const bool head = k == 0 || k0[k] != k0[k - 1] || (k1 && k1[k] != k1[k - 1]);
    if (head) {
      uint32_t e = k + 1;
      while (e < M && k0[e] == k0[k] && (!k1 || k1[e] == k1[k])) ++e;
      const unsigned long long c = e - k;
      local += c * (c - 1) / 2;
    }
  }
  local = (unsigned long long)warp_sum_ll((long long)local);
  if (lane_id() == 0 && local) atomicAdd(out, local);
}

// One level of the bottom-up merge: runs of length L are merged pairwise; elements of the right run count the
// left-run elements strictly greater than themselves (discordant pairs).
__global__ void __launch_bounds__(256)
merge_level_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ count,
                   uint32_t L, unsigned long long* __restrict__ inversions) {
  const uint32_t M = *count;
  unsigned long long local = 0;
  for (uint32_t k = blockIdx.x * 256u + threadIdx.x; k < M; k += gridDim.x * 256u) {
    const uint32_t base = k / (2u * L) * (2u * L);
    const uint32_t a1 = min(M, base + L), b1 = min(M, base + 2u * L);
    const uint32_t v = in[k];
    if (k < a1) {                                   // left run: goes before equal right elements
      uint32_t lo = a1, hi = b1;                    // lower_bound in the right run
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (in[mid] < v) lo = mid + 1; else hi = mid; }
      out[k + (lo - a1)] = v;
    } else {                                        // right run
      uint32_t lo = base, hi = a1;                  // upper_bound in the left run
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (in[mid] <= v) lo = mid + 1; else hi = mid; }
      out[base + (k - a1) + (lo - base)] = v;
      local += a1 - lo;                             // left elements > v
    }
  }
  local = (unsigned long long)warp_sum_ll((long long)local);
  if (lane_id() == 0 && local) atomicAdd(inversions, local);
}

// acc: [0] tx  [1] ty  [2] txy  [3] D ; counts: [0] M (x order)  [1] M (y order)
__global__ void kendall_finish_kernel(const unsigned long long* __restrict__ acc, const uint32_t* __restrict__ counts,
                                      const uint32_t* __restrict__ nan_flag, double* __restrict__ out) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const unsigned long long M = counts[0];
  if (M < 2 || *nan_flag) { *out = nan; return; }
  const unsigned long long n0 = M * (M - 1) / 2;
  const unsigned long long tx = acc[0], ty = acc[1], txy = acc[2], D = acc[3];
  if (tx == n0 || ty == n0) { *out = nan; return; }            // a constant vector: tau-b undefined (rsa.py:29-30)
  const long long cd = (long long)n0 - (long long)tx - (long long)ty + (long long)txy - 2ll * (long long)D;
  *out = (double)cd / (double)n0;
}

__global__ void zero_u64_kernel(unsigned long long* p, int count) {
  if ((int)threadIdx.x < count) p[threadIdx.x] = 0ull;
}

}  // namespace vr

using namespace vr;

extern "C" size_t vr_kendall_workspace_bytes(int n) {
  if (n < 2) return 256;
  const size_t Np = (size_t)round_up(triu_len(n), 32);
  // xkey_c, ykey_c, idx, 4 sort buffers, order_y, order_xy, xseq, yseq, ysorted, 4 compact/merge buffers, pairbits
  return 16 * arena_bytes(Np, 4) + arena_bytes(sort_count_elems(triu_len(n), 1), 4) + arena_bytes(Np / 32, 4) + 4096;
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
  const int W = vr_mask_words(n);
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
  uint32_t* cx = a.take<uint32_t>(Np);          // compacted x keys
  uint32_t* cy = a.take<uint32_t>(Np);          // compacted y keys (x, y order)
  uint32_t* cy2 = a.take<uint32_t>(Np);         // merge ping-pong
  uint32_t* cys = a.take<uint32_t>(Np);         // compacted y keys (y order)
  uint32_t* counts_sort = a.take<uint32_t>(sort_count_elems(N, 1));
  uint32_t* pairbits = a.take<uint32_t>(Np / 32);
  unsigned long long* acc = a.take<unsigned long long>(8);
  uint32_t* small = a.take<uint32_t>(8);        // [0] M (x order) [1] M (y order) [2] nan flag
  if (!a.ok) { set_error("vr_kendall_tau_a: internal workspace plan overflow"); return VR_ERR_WORKSPACE; }
  ProfScope prof(F_KENDALL, stream);
  const unsigned nblk = (unsigned)((N + 255) / 256);

  VR_CHECK_CUDA(cudaMemsetAsync(small, 0, 8 * sizeof(uint32_t), stream));
  dim3 grid((n + 255) / 256, n);
  kendall_keys_kernel<<<grid, 256, 0, stream>>>(rdmA, rdmB, n, ld, xkey_c, ykey_c, v0, small + 2);
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

  const uint32_t words = (uint32_t)(Np / 32);
  const int mgrid = (int)(nblk < 4 * kNumSMs ? nblk : 4 * kNumSMs);
  for (int r = 0; r < R; ++r) {
    dim3 pg((words + 255) / 256 < 64 ? (words + 255) / 256 : 64, 1);
    pair_mask_kernel<<<pg, 256, (size_t)W * 4, stream>>>(masks + (size_t)r * W, 1, n, W, words, pairbits);
    VR_CHECK_LAUNCH();
    zero_u64_kernel<<<1, 32, 0, stream>>>(acc, 8);
    VR_CHECK_LAUNCH();
    compact_kernel<<<1, 1024, 0, stream>>>(order_xy, pairbits, (uint32_t)N, xseq, yseq, cx, cy, small + 0);
    VR_CHECK_LAUNCH();
    compact_kernel<<<1, 1024, 0, stream>>>(order_y, pairbits, (uint32_t)N, ysorted, nullptr, cys, nullptr, small + 1);
    VR_CHECK_LAUNCH();
    tie_pairs_kernel<<<mgrid, 256, 0, stream>>>(cx, nullptr, small + 0, acc + 0);
    VR_CHECK_LAUNCH();
    tie_pairs_kernel<<<mgrid, 256, 0, stream>>>(cys, nullptr, small + 1, acc + 1);
    VR_CHECK_LAUNCH();
    tie_pairs_kernel<<<mgrid, 256, 0, stream>>>(cx, cy, small + 0, acc + 2);
    VR_CHECK_LAUNCH();
    uint32_t* src = cy;
    uint32_t* dst = cy2;
    for (uint64_t L = 1; L < (uint64_t)N; L <<= 1) {     // M <= N: levels beyond M are no-ops that only copy
      merge_level_kernel<<<mgrid, 256, 0, stream>>>(src, dst, small + 0, (uint32_t)L, acc + 3);
      VR_CHECK_LAUNCH();
      uint32_t* t = src; src = dst; dst = t;
    }
    kendall_finish_kernel<<<1, 1, 0, stream>>>(acc, small, small + 2, out + r);
    VR_CHECK_LAUNCH();
  }
  return VR_OK;
}
