// Experiment (round 2): "lanes = replicates" Spearman sweep.  One warp handles ONE canonical pair for 32 replicates at a
// time: per vector a coalesced 256-byte table row {kept word, count} (L2-resident) gives the pair's rank in all 32
// replicates with one popc, and the 8 x 8 dot products are accumulated per lane.  Synthetic tie-free data; checks a few
// (a, b, r) cells on the host.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o sweep_bench sweep_bench.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NA = 8, NB = 8, NV = NA + NB;

// tables[v][batch][g][lane] = {kept word of replicate batch*32+lane over sorted positions 32g..32g+31, kept count before 32g}
__global__ void build_tables(const uint32_t* __restrict__ sorted_pairs /*[v][N] (i<<16|j) in sorted order*/, int N, int G,
                             const uint32_t* __restrict__ maskT /*[batch][n]*/, int n, int batches, uint2* __restrict__ tables,
                             uint32_t* __restrict__ nkept /*[v][batch][32]*/) {
  const int v = blockIdx.y, b = blockIdx.x, lane = threadIdx.x;
  const uint32_t* sp = sorted_pairs + (size_t)v * N;
  const uint32_t* mt = maskT + (size_t)b * n;
  uint2* tab = tables + ((size_t)b * NV + v) * G * 32;
  uint32_t run = 0;
  for (int g = 0; g < G; ++g) {
    uint32_t bits = 0;
    const int x = g * 32 + lane;
    if (x < N) { const uint32_t w = sp[x]; bits = mt[w >> 16] & mt[w & 0xFFFF]; }
    // transpose: lane r collects bit r of every lane's word
    uint32_t K = 0;
    for (int k = 0; k < 32; ++k) K |= ((__shfl_sync(0xffffffffu, bits, k) >> lane) & 1u) << k;
    tab[(size_t)g * 32 + lane] = make_uint2(K, run);
    run += __popc(K);
  }
  if (v == 0) nkept[(size_t)b * 32 + lane] = run;
}

struct Desc { uint32_t row_off; uint32_t lowmask; };   // byte offset of the table row inside the batch block (v * G * 256 + g * 256), mask of positions below s in the group

__device__ __forceinline__ void mac(unsigned long long& acc, uint32_t a, uint32_t b) {
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
}
__device__ __forceinline__ uint2 ldg64(const char* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// tables: [batch][v][G][32] uint2.  desc: [N][NV] Desc.
__device__ __forceinline__ double u2d(uint32_t v) {     // exact uint32 -> fp64 with one DADD (no I2F)
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}

// kFlavour 0: 64-bit integer multiply-accumulate (IMAD.WIDE + IADD3 + IMAD.X);  1: fp64 DFMA (exact within a warp's slice)
template <int kWarps, int kFlavour>
__global__ void __launch_bounds__(kWarps * 32, 1)
sweep_kernel(const Desc* __restrict__ desc, int N, int G, const uint2* __restrict__ tables, int batches,
             const uint32_t* __restrict__ nkept, const uint32_t* __restrict__ pairs_canon /*[N] i<<16|j*/,
             const uint32_t* __restrict__ maskT, int n, int splits, unsigned long long* __restrict__ out /*[NA][NB][batches*32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x / splits, s = blockIdx.x % splits;
  const int per = (N + splits - 1) / splits;
  const int p_lo = s * per, p_hi = min(N, p_lo + per);
  const char* base = reinterpret_cast<const char*>(tables) + (size_t)b * ((size_t)NV * G * 256) + lane * 8;
  const int32_t bias = 1 - (int32_t)nkept[(size_t)b * 32 + lane];
  const uint32_t* mt = maskT + (size_t)b * n;
  unsigned long long acc[NA][NB];
  double accd[NA][NB];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int c = 0; c < NB; ++c) { acc[a][c] = 0; accd[a][c] = 0.0; }

  int p = p_lo + warp;
  uint4 dn[NV / 2];                                     // descriptors of the next pair (uniform loads)
  uint32_t kwn = 0;
  auto load_desc = [&](int q) {
    const uint4* dp = reinterpret_cast<const uint4*>(desc + (size_t)q * NV);
#pragma unroll
    for (int k = 0; k < NV / 2; ++k) dn[k] = __ldg(dp + k);
    const uint32_t pw = __ldg(pairs_canon + q);
    kwn = __ldg(mt + (pw >> 16)) & __ldg(mt + (pw & 0xFFFF));
  };
  if (p < p_hi) load_desc(p);
  for (; p < p_hi; p += kWarps) {
    uint2 e[NV];
    uint32_t lm[NV];
#pragma unroll
    for (int k = 0; k < NV / 2; ++k) {
      e[2 * k] = ldg64(base + dn[k].x);     lm[2 * k] = dn[k].y;
      e[2 * k + 1] = ldg64(base + dn[k].z); lm[2 * k + 1] = dn[k].w;
    }
    const int32_t keep = -(int32_t)((kwn >> lane) & 1u);           // all ones for a kept (pair, replicate)
    if (p + kWarps < p_hi) load_desc(p + kWarps);
    uint32_t dev[NV];                                     // 2 * rank - 1 = C(s) + C(e) (tie-free: 2 C(s) + 1), always > 0
#pragma unroll
    for (int v = 0; v < NV; ++v) dev[v] = 2u * (e[v].y + __popc(e[v].x & lm[v])) + 1u;
#pragma unroll
    for (int a = 0; a < NA; ++a) dev[a] &= (uint32_t)keep;
    if (kFlavour == 0) {
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int c = 0; c < NB; ++c) mac(acc[a][c], dev[a], dev[NA + c]);
    } else {
      double dd[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) dd[v] = u2d(dev[v]);
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int c = 0; c < NB; ++c) accd[a][c] = fma(dd[a], dd[NA + c], accd[a][c]);
    }
  }
  if (kFlavour == 1) {     // a warp's slice: < 2^53, exact
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[a][c] = (unsigned long long)accd[a][c];
  }
  // sum over kept pairs of (2 rA - 1)(2 rB - 1) -> centred: subtract N_r^3  (sum of 2 r - 1 over kept pairs is N_r^2)
  {
    const unsigned long long nr = (unsigned long long)(1 - bias);
    const unsigned long long share = (warp == 0 && s == 0) ? nr * nr * nr : 0ull;   // subtracted once per (batch, lane)
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[a][c] -= share;
  }
  // CTA reduction through shared memory, then one atomic per (a, c, lane)
  __shared__ unsigned long long red[kWarps][32];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      red[warp][lane] = acc[a][c];
      __syncthreads();
      if (warp == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kWarps; ++w) t += red[w][lane];
        atomicAdd(out + ((size_t)(a * NB + c) * batches + b) * 32 + lane, (unsigned long long)t);
      }
      __syncthreads();
    }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000;
  const int R = argc > 2 ? atoi(argv[2]) : 1024;
  const int N = n * (n - 1) / 2, G = (N + 31) / 32, batches = (R + 31) / 32;
  printf("n=%d N=%d R=%d batches=%d NV=%d\n", n, N, R, batches, NV);
  std::mt19937 rng(1);
  // canonical pairs
  std::vector<uint32_t> canon(N);
  { int p = 0; for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) canon[p++] = ((uint32_t)i << 16) | j; }
  // per vector: random sorted order
  std::vector<uint32_t> sorted_pairs((size_t)NV * N);
  std::vector<Desc> desc((size_t)N * NV);
  std::vector<std::vector<uint32_t>> pos(NV, std::vector<uint32_t>(N));
  for (int v = 0; v < NV; ++v) {
    std::vector<uint32_t> perm(N);
    std::iota(perm.begin(), perm.end(), 0u);
    std::shuffle(perm.begin(), perm.end(), rng);       // perm[x] = canonical pair at sorted position x
    for (int x = 0; x < N; ++x) { sorted_pairs[(size_t)v * N + x] = canon[perm[x]]; pos[v][perm[x]] = x; }
    for (int p = 0; p < N; ++p) { const uint32_t x = pos[v][p]; desc[(size_t)p * NV + v] = Desc{((uint32_t)v * G + (x >> 5)) * 256u, (1u << (x & 31)) - 1u}; }
  }
  // masks: each replicate drops n/10 stimuli; maskT[b][i] bit r
  std::vector<uint32_t> maskT((size_t)batches * n, 0u);
  std::vector<std::vector<uint8_t>> kept(R, std::vector<uint8_t>(n, 1));
  for (int r = 0; r < R; ++r) {
    std::vector<int> idx(n); std::iota(idx.begin(), idx.end(), 0); std::shuffle(idx.begin(), idx.end(), rng);
    for (int k = 0; k < n / 10; ++k) kept[r][idx[k]] = 0;
    for (int i = 0; i < n; ++i) if (kept[r][i]) maskT[(size_t)(r / 32) * n + i] |= 1u << (r & 31);
  }
  uint32_t *d_sorted, *d_maskT, *d_nkept, *d_canon; uint2* d_tables; Desc* d_desc; unsigned long long* d_out;
  CK(cudaMalloc(&d_sorted, sorted_pairs.size() * 4)); CK(cudaMemcpy(d_sorted, sorted_pairs.data(), sorted_pairs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_maskT, maskT.size() * 4)); CK(cudaMemcpy(d_maskT, maskT.data(), maskT.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_canon, (size_t)N * 4)); CK(cudaMemcpy(d_canon, canon.data(), (size_t)N * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_desc, desc.size() * 8)); CK(cudaMemcpy(d_desc, desc.data(), desc.size() * 8, cudaMemcpyHostToDevice));
  const size_t tab_bytes = (size_t)NV * batches * G * 32 * 8;
  printf("tables: %.1f MB total, %.1f MB per batch\n", tab_bytes / 1e6, tab_bytes / 1e6 / batches);
  CK(cudaMalloc(&d_tables, tab_bytes)); CK(cudaMalloc(&d_nkept, (size_t)NV * batches * 32 * 4));
  CK(cudaMalloc(&d_out, (size_t)NA * NB * batches * 32 * 8));
  build_tables<<<dim3(batches, NV), 32>>>(d_sorted, N, G, d_maskT, n, batches, d_tables, d_nkept);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](int splits, int warps, int flavour) {
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
      CK(cudaMemset(d_out, 0, (size_t)NA * NB * batches * 32 * 8));
      cudaEventRecord(e0);
#define LAUNCH(W, F) sweep_kernel<W, F><<<batches * splits, W * 32>>>(d_desc, N, G, d_tables, batches, d_nkept, d_canon, d_maskT, n, splits, d_out)
      if (warps == 8 && flavour == 0) LAUNCH(8, 0);
      else if (warps == 8) LAUNCH(8, 1);
      else if (warps == 4 && flavour == 0) LAUNCH(4, 0);
      else if (warps == 4) LAUNCH(4, 1);
      else if (flavour == 0) LAUNCH(6, 0);
      else LAUNCH(6, 1);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    printf("flavour=%d splits=%3d warps=%2d : %.3f ms  (%.2f G pair-rep-vector lookups/s, %.1f GB/s table rows)\n", flavour, splits, warps, best,
           (double)N * batches * 32 * NV / best / 1e6, (double)N * batches * NV * 256 / best / 1e6);
    return best;
  };
  if (argc > 5) run(atoi(argv[5]), atoi(argv[4]), atoi(argv[3]));
  else for (int flavour : {0, 1}) for (int warps : {4, 8}) for (int splits : {148, 296}) run(splits, warps, flavour);
  // check a few cells against a direct host computation
  std::vector<unsigned long long> out((size_t)NA * NB * batches * 32);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r : {0, 37, R - 1}) {
    std::vector<std::vector<int32_t>> devv(NV, std::vector<int32_t>(N));
    long long Nr = 0;
    for (int v = 0; v < NV; ++v) {
      uint32_t c = 0;
      std::vector<uint32_t> C(N);
      for (int x = 0; x < N; ++x) { const uint32_t w = sorted_pairs[(size_t)v * N + x]; C[x] = c; c += kept[r][w >> 16] & kept[r][w & 0xFFFF]; }
      Nr = c;
      for (int p = 0; p < N; ++p) { const uint32_t w = canon[p]; const bool k = kept[r][w >> 16] && kept[r][w & 0xFFFF]; devv[v][p] = k ? 2 * (int32_t)C[pos[v][p]] + 1 - (int32_t)c : 0; }
    }
    for (int a : {0, 7}) for (int c : {0, 5}) {
      long long t = 0;
      for (int p = 0; p < N; ++p) t += (long long)devv[a][p] * devv[NA + c][p];
      const long long got = (long long)out[((size_t)(a * NB + c) * batches + r / 32) * 32 + (r & 31)];
      if (got != t) { ++bad; printf("MISMATCH r=%d a=%d c=%d got=%lld want=%lld (Nr=%lld)\n", r, a, c, got, t, Nr); }
    }
  }
  printf(bad ? "FAILED\n" : "check ok\n");
  return bad != 0;
}
