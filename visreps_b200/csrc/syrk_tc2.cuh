// Pair-tile variant of the tcgen05 SYRK (syrk_tc.cuh): two CTAs of a cluster (one TPC) compute a 256 x 256 tile of
// Z Z^T with tcgen05.mma.cta_group::2 -- M = 256 split over the two SMs, N = 256, each CTA holding 128 rows of A and
// 128 of the 256 rows of B.  Per flop the tensor cores read half the shared-memory operand bytes of the 128 x 128
// single-CTA tile, which is what holds that kernel at ~71 % of the TF32 pipe.
//
//   CTA c of the pair (c = %cluster_ctarank):  A = rows 256*I + 128*c .. +128,  B half = rows 256*J + 128*c .. +128
//   warp 0   TMA producer (both CTAs): boxes land in the CTA's own shared memory, completion bytes are
//            signalled on the LEADER's (rank 0) full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1   leader only: one thread issues the MMAs for the pair; tcgen05.commit ... .multicast::cluster releases
//            the stage in both CTAs / hands the accumulator to both epilogues
//   warp 2   TMEM allocation (cta_group::2, 512 columns = two 256-column fp32 accumulators), both CTAs
//   warps 4-11 epilogue: warp w drains lane quarter w % 4, column half (w - 4) / 4 of the CTA's own TMEM into
//            registers per chunk and arrives on the leader's tmem_empty barrier (remote arrive from the peer)
//   Diagonal pair tiles (I == J): the B half of a CTA is its own A tile, so only A is loaded.
#pragma once
#include "syrk_tc.cuh"

namespace vr {
namespace tc2 {

using namespace vr::tc;

constexpr int kPM = 256;                        // pair tile edge
constexpr int kStages2 = 3;
constexpr int kStageBytes2 = 4 * kTileBytes;    // A_hi | A_lo | Bhalf_hi | Bhalf_lo, 16 KB each
constexpr int kThreads2 = 384;
constexpr int kTmemCols2 = 512;
constexpr size_t kSmemBytes2 = (size_t)kStages2 * kStageBytes2 + 1024;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: the leader's copy

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into this CTA's shared memory, completion on the LEADER CTA's barrier at the same offset
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 256 (pair), N = 256
constexpr uint32_t kInstrDesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kPM >> 3) << 17) | ((uint32_t)(kPM >> 4) << 24);
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDesc2), "r"(accumulate)
      : "memory");
}

// grid (2 * upper-triangular pair tiles, K splits), cluster (2, 1, 1).  partial[split][n_pad][n_pad], n_pad % 256 == 0.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
syrk_tc2_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, int nb2,
                int kb_total, int kb_per_split, float* __restrict__ partial, int n_pad) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages2];
  __shared__ __align__(8) uint64_t empty_bar[kStages2];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  int bi, bj;
  decode_tri_tile((int)(blockIdx.x >> 1), nb2, bi, bj);      // both CTAs of a pair decode the same tile
  const bool diag = bi == bj;
  const int split = blockIdx.y;
  const int kb0 = split * kb_per_split;
  const int nkb = min(kb_total, kb0 + kb_per_split) - kb0;
  const int nchunks = (nkb + kChunkKb - 1) / kChunkKb;
  const int row_a = bi * kPM + (int)cta * kBM;                // this CTA's 128 rows of A
  const int row_b = bj * kPM + (int)cta * kBM;                // this CTA's 128 of the 256 rows of B

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols2) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                         // the peer's barriers exist before anything remote happens
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = 2u * (diag ? 2u * kTileBytes : 4u * kTileBytes);   // bytes of BOTH CTAs per stage
      for (int it = 0; it < nkb; ++it) {
        const int s = it % kStages2;
        const uint32_t ph = (uint32_t)(it / kStages2) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        if (leader) mbar_expect_tx(&full_bar[s], tx);
        uint8_t* st = smem + (size_t)s * kStageBytes2;
        const int k = (kb0 + it) * kBK;
        tma_load_2d_pair(&map_hi, &full_bar[s], st, k, row_a);
        tma_load_2d_pair(&map_lo, &full_bar[s], st + kTileBytes, k, row_a);
        if (!diag) {
          tma_load_2d_pair(&map_hi, &full_bar[s], st + 2 * kTileBytes, k, row_b);
          tma_load_2d_pair(&map_lo, &full_bar[s], st + 3 * kTileBytes, k, row_b);
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      int it = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        const uint32_t use = (uint32_t)(c >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[buf], use ^ 1u);            // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * kPM);
        const int it_end = min(nkb, it + kChunkKb);
        for (int first = 1; it < it_end; ++it, first = 0) {
          const int s = it % kStages2;
          const uint32_t ph = (uint32_t)(it / kStages2) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + (size_t)s * kStageBytes2);
          const uint32_t a_lo = a_hi + kTileBytes;
          const uint32_t b_hi = diag ? a_hi : a_hi + 2 * kTileBytes;
          const uint32_t b_lo = diag ? a_lo : a_hi + 3 * kTileBytes;
#pragma unroll
          for (int j = 0; j < kBK / 8; ++j) {
            const uint32_t off = (uint32_t)j * 32u;
            mma_tf32_pair(tmem_acc, smem_desc_sw128(a_hi + off), smem_desc_sw128(b_hi + off), (first && j == 0) ? 0u : 1u);
            mma_tf32_pair(tmem_acc, smem_desc_sw128(a_hi + off), smem_desc_sw128(b_lo + off), 1u);
            mma_tf32_pair(tmem_acc, smem_desc_sw128(a_lo + off), smem_desc_sw128(b_hi + off), 1u);
          }
          tc_commit_pair(&empty_bar[s]);
        }
        tc_commit_pair(&tmem_full_bar[buf]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;                        // which 128 of the 256 accumulator columns
    const int row = q * 32 + (int)lane;
    float acc[kBN];
#pragma unroll
    for (int x = 0; x < kBN; ++x) acc[x] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      const uint32_t use = (uint32_t)(c >> 1) & 1u;
      mbar_wait(&tmem_full_bar[buf], use);
      tc_fence_after();
#pragma unroll
      for (int cg = 0; cg < kBN / 32; ++cg) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kPM + half * kBN + cg * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int x = 0; x < 32; ++x) acc[cg * 32 + x] += __uint_as_float(v[x]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[buf]);   // 8 epilogue warps x 2 CTAs -> barrier count 16
    }
    float* out = partial + (size_t)split * n_pad * n_pad + (size_t)(row_a + row) * n_pad + (size_t)bj * kPM + (size_t)half * kBN;
#pragma unroll
    for (int x = 0; x < kBN; x += 4)
      *reinterpret_cast<float4*>(out + x) = make_float4(acc[x], acc[x + 1], acc[x + 2], acc[x + 3]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                         // neither CTA leaves while the other may still touch it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols2) : "memory");
  }
}

}  // namespace tc2
}  // namespace vr
