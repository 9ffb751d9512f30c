// Presort of an RDM's strict upper triangle, and row-wise ordinal ranks.
//
// vr_presort_triu_f32 replaces, for every later comparison of the same RDM, the
// `triu_indices` gather (visreps/analysis/rsa.py:111-113) and the stable argsort inside
// scipy.stats.rankdata that scipy.stats.spearmanr runs per call (rsa.py:122).
// vr_rank_rows_ordinal_f32 replaces _rank (rsa.py:50-52).
#include "common.cuh"
#include "sort.cuh"

namespace vr {

constexpr int kPresortBatch = 64;
struct PresortBatch {
  const float* rdm[kPresortBatch];
  void* blob[kPresortBatch];
};

// ---- upper triangle -> (sortable key, packed pair) ------------------------------------
// grid (ceil(n/256), n, batch); keys/vals of matrix b start at b * seg_stride.
__global__ void __launch_bounds__(256)
triu_keys_kernel(const __grid_constant__ PresortBatch batch, int n, int64_t ld, uint32_t* __restrict__ keys,
                 uint32_t* __restrict__ vals, int64_t seg_stride) {
  const uint32_t i = blockIdx.y;
  const uint32_t j = blockIdx.x * 256u + threadIdx.x;
  const int b = blockIdx.z;
  PresortHeader* hdr = reinterpret_cast<PresortHeader*>(batch.blob[b]);
  if (i == 0 && j == 0) {
    hdr->magic = kPresortMagic;
    hdr->n = (uint32_t)n;
    hdr->N = (uint32_t)triu_len(n);
    hdr->G = presort_groups(n);
  }
  if (j >= (uint32_t)n || j <= i) return;
  const float v = batch.rdm[b][(int64_t)i * ld + j];
  if (v != v) {
    atomicOr(&hdr->has_nan, 1u);
    atomicAdd((__float_as_uint(v) & 0x80000000u) ? &hdr->nan_lo : &hdr->nan_hi, 1u);
  }
  const uint32_t p = triu_pos(i, j, (uint32_t)n);
  keys[(int64_t)b * seg_stride + p] = float_key(v);
  vals[(int64_t)b * seg_stride + p] = pack_pair(i, j);
}

// Clears the device-written header fields (has_nan, n_runs) before triu_keys_kernel sets them.
__global__ void clear_headers_kernel(const __grid_constant__ PresortBatch batch, int count) {
  const int b = threadIdx.x;
  if (b < count) {
    PresortHeader* hdr = reinterpret_cast<PresortHeader*>(batch.blob[b]);
    hdr->has_nan = 0;
    hdr->n_runs = 0;
    hdr->nan_lo = 0;
    hdr->nan_hi = 0;
  }
}

// One warp per 32-group of the sorted order: head flags, packed stream, per-group
// candidates for the prev/next head tables.  grid (ceil(G/8), batch).
__global__ void __launch_bounds__(256)
heads_kernel(const uint32_t* __restrict__ keys_all, const uint32_t* __restrict__ vals_all, int64_t seg_stride,
             uint32_t N, uint32_t G, uint32_t n, const __grid_constant__ PresortBatch batch) {
  const uint32_t g = blockIdx.x * 8u + (threadIdx.x >> 5);
  if (g >= G) return;
  const int b = blockIdx.y;
  const uint32_t* __restrict__ keys = keys_all + (int64_t)b * seg_stride;
  const uint32_t* __restrict__ vals = vals_all + (int64_t)b * seg_stride;
  PresortView out = presort_view(batch.blob[b], (int)n);
  const unsigned lane = lane_id();
  const uint32_t k = g * 32u + lane;
  const bool valid = k < N;
  const uint32_t key = valid ? keys[k] : 0u;
  const uint32_t prev = (valid && k > 0) ? keys[k - 1] : 0u;
  const bool head = valid && (k == 0 || key != prev);
  const bool pad_head = (k == N);  // sentinel that closes the last run
  uint32_t word;
  if (valid) word = vals[k] | (head ? kHeadBit : 0u);
  else word = kPadBit | (pad_head ? kHeadBit : 0u);
  out.stream[k] = word;
  const uint32_t H = __ballot_sync(0xffffffffu, head);
  const uint32_t Hn = __ballot_sync(0xffffffffu, head || pad_head);
  if (lane == 0) {
    out.head_bits[g] = Hn;
    out.prev_head[g] = H ? g * 32u + (31u - __clz(H)) : 0u;
    out.next_head[g] = Hn ? g * 32u + (__ffs(Hn) - 1u) : 0xFFFFFFFFu;
    if (H) atomicAdd(&out.hdr->n_runs, (uint32_t)__popc(H));
  }
}

// prev_head: inclusive running max (forward).  next_head: inclusive running min (backward),
// seeded with next_head[G] = N.  One CTA per matrix, each thread owns a contiguous chunk.
__global__ void __launch_bounds__(1024)
head_tables_kernel(const __grid_constant__ PresortBatch batch, uint32_t N, uint32_t G, uint32_t n) {
  __shared__ uint32_t part[1024];
  PresortView view = presort_view(batch.blob[blockIdx.x], (int)n);
  uint32_t* __restrict__ prev_head = view.prev_head;
  uint32_t* __restrict__ next_head = view.next_head;
  const uint32_t chunk = (G + 1023u) / 1024u;
  const uint32_t lo = threadIdx.x * chunk;
  const uint32_t hi = lo + chunk < G ? lo + chunk : G;
  // forward max
  uint32_t m = 0;
  for (uint32_t g = lo; g < hi; ++g) m = max(m, prev_head[g]);
  part[threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.x < 32) {   // exclusive max-scan of the 1024 chunk maxima, 32 per lane
    uint32_t run = 0, loc[32];
    for (int t = 0; t < 32; ++t) { loc[t] = part[threadIdx.x * 32 + t]; run = max(run, loc[t]); }
    uint32_t incl = run;
    for (int o = 1; o < 32; o <<= 1) { uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (threadIdx.x >= (unsigned)o) incl = max(incl, x); }
    uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
    run = threadIdx.x == 0 ? 0u : excl;
    for (int t = 0; t < 32; ++t) { part[threadIdx.x * 32 + t] = run; run = max(run, loc[t]); }
  }
  __syncthreads();
  m = part[threadIdx.x];
  for (uint32_t g = lo; g < hi; ++g) { m = max(m, prev_head[g]); prev_head[g] = m; }
  __syncthreads();
  // backward min
  m = 0xFFFFFFFFu;
  for (uint32_t g = lo; g < hi; ++g) m = min(m, next_head[g]);
  part[threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.x < 32) {   // exclusive suffix-min-scan, seeded with N
    const int L = 31 - (int)threadIdx.x;            // lane 0 owns the LAST 32 chunks
    uint32_t run = 0xFFFFFFFFu, loc[32];
    for (int t = 0; t < 32; ++t) { loc[t] = part[L * 32 + 31 - t]; run = min(run, loc[t]); }
    uint32_t incl = run;
    for (int o = 1; o < 32; o <<= 1) { uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (threadIdx.x >= (unsigned)o) incl = min(incl, x); }
    uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
    run = threadIdx.x == 0 ? N : min(N, excl);
    for (int t = 0; t < 32; ++t) { part[L * 32 + 31 - t] = run; run = min(run, loc[t]); }
    if (threadIdx.x == 0) next_head[G] = N;
  }
  __syncthreads();
  m = part[threadIdx.x];
  for (uint32_t g = hi; g > lo; --g) { m = min(m, next_head[g - 1]); next_head[g - 1] = m; }
}

// One warp per 32-group of the sorted order, after the head tables are final: the run [s, e)
// of every element, scattered to the element's canonical (row-major triu) position.
__global__ void __launch_bounds__(256)
runs_kernel(uint32_t N, uint32_t G, uint32_t n, const __grid_constant__ PresortBatch batch) {
  const uint32_t g = blockIdx.x * 8u + (threadIdx.x >> 5);
  if (g >= G) return;
  PresortView blob = presort_view(batch.blob[blockIdx.y], (int)n);
  const unsigned lane = lane_id();
  const uint32_t k = g * 32u + lane;
  const uint32_t wd = blob.stream[k];
  const uint32_t H = blob.head_bits[g];
  if (k >= N) { blob.runs[k] = run_compact(N) ? 32u << kRunShloShift : 0u; return; }   // canonical padding (never kept, never a long run)
  const uint32_t below = H & lanemask_le();
  const uint32_t s = below ? g * 32u + (31u - __clz(below)) : blob.prev_head[g - 1];
  const uint32_t above = H & lanemask_gt();
  const uint32_t e = above ? g * 32u + (__ffs(above) - 1u) : blob.next_head[g + 1];
  uint32_t word;
  if (run_compact(N)) {
    const uint32_t gs = s >> 5, ls = s & 31u, ge = e >> 5, le = e & 31u;
    const uint32_t head = (gs << 3) | ((32u - ls) << kRunShloShift);
    if (e == s + 1u) word = head | 1u | (le ? (32u - le) << kRunSheShift : 0u);     // bit 0: a run of one (sweep.cu); she as below
    else if (ge == gs) word = head | ((32u - le) << kRunSheShift);                  // le > ls: she in 1..31
    else if (ge == gs + 1u && le == 0u) word = head;                                // she = 0: the whole group
    else if (ge == gs + 1u) word = head | kRunCrossBit | ((32u - le) << kRunSheShift);
    else word = (gs << 3) | ((ls + 1u) << kRunSheShift);                            // shlo = 0: long run
  } else {
    const uint32_t sbits = run_sbits(N);
    const uint32_t esc = (1u << (32u - sbits)) - 1u;
    word = s | (min(e - s - 1u, esc) << sbits);
  }
  blob.runs[triu_pos(pair_i(wd), pair_j(wd), n)] = word;
}

// ---- row-wise ordinal ranks ------------------------------------------------------------
__global__ void __launch_bounds__(256)
row_keys_kernel(const float* __restrict__ X, int64_t d, int64_t ldx, uint32_t* __restrict__ keys,
                uint32_t* __restrict__ vals) {
  const int64_t row = blockIdx.y;
  for (int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x; c < d; c += (int64_t)gridDim.x * 256) {
    keys[row * d + c] = float_key(X[row * ldx + c]);
    vals[row * d + c] = (uint32_t)c;
  }
}
__global__ void __launch_bounds__(256)
row_rank_scatter_kernel(const uint32_t* __restrict__ vals, int64_t d, float* __restrict__ out) {
  const int64_t row = blockIdx.y;
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < d; k += (int64_t)gridDim.x * 256)
    out[row * d + vals[row * d + k]] = (float)k;
}

}  // namespace vr

using namespace vr;

extern "C" size_t vr_presort_bytes(int n) { return n >= 2 ? presort_blob_bytes(n) : 256; }

static size_t presort_ws(int n, int count) {
  if (n < 2 || count < 1) return 256;
  const int64_t N = triu_len(n);
  return 4 * arena_bytes((size_t)count * round_up(N, 32), 4) + arena_bytes(sort_count_elems(N, count), 4) + 256;
}
extern "C" size_t vr_presort_workspace_bytes(int n) { return presort_ws(n, 1); }
extern "C" size_t vr_presort_batch_workspace_bytes(int n, int count) { return presort_ws(n, count); }

extern "C" int vr_presort_triu_batch_f32(const float* const* rdms, int count, int n, int64_t ld, void* const* blobs,
                                         void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(rdms && blobs, "vr_presort_triu_batch_f32: null pointer");
  VR_REQUIRE(count >= 1 && count <= kPresortBatch, "vr_presort_triu_batch_f32: count=%d outside [1, %d]", count, kPresortBatch);
  VR_REQUIRE(n >= 2 && n <= VR_MAX_STIMULI, "vr_presort_triu_f32: n=%d outside [2, %d]", n, VR_MAX_STIMULI);
  VR_REQUIRE(ld >= n, "vr_presort_triu_f32: ld=%lld < n=%d", (long long)ld, n);
  if (ws_bytes < presort_ws(n, count) || !ws) {
    set_error("vr_presort_triu_f32: workspace %zu < %zu", ws_bytes, presort_ws(n, count));
    return VR_ERR_WORKSPACE;
  }
  PresortBatch batch{};
  for (int b = 0; b < count; ++b) {
    VR_REQUIRE(rdms[b] && blobs[b], "vr_presort_triu_f32: null pointer in batch");
    batch.rdm[b] = rdms[b];
    batch.blob[b] = blobs[b];
  }
  const int64_t N = triu_len(n);
  const int64_t Npad = round_up(N, 32);
  const uint32_t G = presort_groups(n);
  Arena a(ws, ws_bytes);
  uint32_t* k0 = a.take<uint32_t>((size_t)count * Npad);
  uint32_t* v0 = a.take<uint32_t>((size_t)count * Npad);
  uint32_t* k1 = a.take<uint32_t>((size_t)count * Npad);
  uint32_t* v1 = a.take<uint32_t>((size_t)count * Npad);
  uint32_t* counts = a.take<uint32_t>(sort_count_elems(N, count));
  ProfScope prof(F_PRESORT, stream);

  clear_headers_kernel<<<1, kPresortBatch, 0, stream>>>(batch, count);
  VR_CHECK_LAUNCH();
  dim3 grid((n + 255) / 256, n, count);
  triu_keys_kernel<<<grid, 256, 0, stream>>>(batch, n, ld, k0, v0, Npad);
  VR_CHECK_LAUNCH();
  int rc = radix_sort_pairs(k0, v0, k1, v1, N, Npad, count, counts, stream);
  if (rc != VR_OK) return rc;
  dim3 hg((G + 7) / 8, count);
  heads_kernel<<<hg, 256, 0, stream>>>(k0, v0, Npad, (uint32_t)N, G, (uint32_t)n, batch);
  VR_CHECK_LAUNCH();
  head_tables_kernel<<<count, 1024, 0, stream>>>(batch, (uint32_t)N, G, (uint32_t)n);
  VR_CHECK_LAUNCH();
  runs_kernel<<<hg, 256, 0, stream>>>((uint32_t)N, G, (uint32_t)n, batch);
  VR_CHECK_LAUNCH();
  return VR_OK;
}

extern "C" int vr_presort_triu_f32(const float* rdm, int n, int64_t ld, void* presorted, void* ws,
                                   size_t ws_bytes, vr_stream_t stream_) {
  VR_REQUIRE(rdm && presorted, "vr_presort_triu_f32: null pointer");
  const float* r[1] = {rdm};
  void* b[1] = {presorted};
  return vr_presort_triu_batch_f32(r, 1, n, ld, b, ws, ws_bytes, stream_);
}

extern "C" size_t vr_rank_rows_workspace_bytes(int n, int64_t d) {
  if (n < 1 || d < 1) return 256;
  return 4 * arena_bytes((size_t)n * d, 4) + arena_bytes(sort_count_elems(d, n), 4) + 256;
}

extern "C" int vr_rank_rows_ordinal_f32(const float* X, int n, int64_t d, int64_t ldx, float* out,
                                        void* ws, size_t ws_bytes, vr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VR_REQUIRE(X && out, "vr_rank_rows_ordinal_f32: null pointer");
  VR_REQUIRE(n >= 1 && n <= 65535 && d >= 1 && ldx >= d, "vr_rank_rows_ordinal_f32: bad shape n=%d d=%lld ldx=%lld",
             n, (long long)d, (long long)ldx);
  VR_REQUIRE((int64_t)n * d < (1ll << 40), "vr_rank_rows_ordinal_f32: too large");
  if (ws_bytes < vr_rank_rows_workspace_bytes(n, d) || !ws) {
    set_error("vr_rank_rows_ordinal_f32: workspace %zu < %zu", ws_bytes, vr_rank_rows_workspace_bytes(n, d));
    return VR_ERR_WORKSPACE;
  }
  Arena a(ws, ws_bytes);
  const size_t tot = (size_t)n * d;
  uint32_t* k0 = a.take<uint32_t>(tot);
  uint32_t* v0 = a.take<uint32_t>(tot);
  uint32_t* k1 = a.take<uint32_t>(tot);
  uint32_t* v1 = a.take<uint32_t>(tot);
  uint32_t* counts = a.take<uint32_t>(sort_count_elems(d, n));
  const int gx = (int)((d + 255) / 256 < 1024 ? (d + 255) / 256 : 1024);
  dim3 grid(gx, n);
  ProfScope prof(F_ROWRANK, stream);
  row_keys_kernel<<<grid, 256, 0, stream>>>(X, d, ldx, k0, v0);
  VR_CHECK_LAUNCH();
  int rc = radix_sort_pairs(k0, v0, k1, v1, d, d, n, counts, stream);
  if (rc != VR_OK) return rc;
  row_rank_scatter_kernel<<<grid, 256, 0, stream>>>(v0, d, out);
  VR_CHECK_LAUNCH();
  return VR_OK;
}
