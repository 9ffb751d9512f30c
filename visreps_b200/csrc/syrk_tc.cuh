// Symmetric rank-k update Z Z^T on the 5th-generation tensor cores (tcgen05 / TMEM / TMA),
// fp32-accurate through a 3xTF32 split.   Replaces `x @ x.T` of compute_rdm
// (visreps/analysis/rsa.py:89), which the reference runs as a CPU sgemm.
//
//   Z (centred rows) is stored as two tf32-exact planes: hi = rn_tf32(z), lo = rn_tf32(z - hi).
//   z_a z_b = hi_a hi_b + hi_a lo_b + lo_a hi_b + O(2^-22): three kind::tf32 MMAs per K step into
//   the same fp32 TMEM accumulator.  Only tiles with block-row <= block-column are computed.
//
//   CTA = one 128x128 output tile x one K split.  Inside the CTA the fp32 accumulation chain in TMEM is
//   cut every 1024 columns (two TMEM accumulators, drained into registers by the epilogue warps while the
//   next chunk is multiplied); the few K splits are summed in fp64 by rdm_epilogue_kernel.
//   warp 0   TMA producer: 128x32 fp32 boxes (128 B rows, SWIZZLE_128B) of hi/lo for the row block
//            and the column block into a 3-stage ring, mbarrier complete_tx
//   warp 1   one elected thread issues tcgen05.mma (M=128, N=128, K=8) x 4 K-steps x 3 products per
//            stage, tcgen05.commit releases the stage / signals the epilogue
//   warp 2   TMEM allocation (128 columns) and release
//   warps 4-7 epilogue: per chunk tcgen05.ld 32 lanes x 32 columns at a time -> fp32 register accumulators;
//            at the end one partial tile -> memory
//   Diagonal tiles load only the row block (B = A).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace vr {
namespace tc {

constexpr int kBM = 128;
constexpr int kBN = 128;
constexpr int kBK = 32;                       // fp32 elements per stage row = 128 bytes = one swizzle row
constexpr int kStages = 3;
constexpr int kTileBytes = kBM * kBK * 4;     // 16 KB
constexpr int kStageBytes = 4 * kTileBytes;   // A_hi | A_lo | B_hi | B_lo
constexpr int kThreads = 256;
constexpr int kTmemCols = 256;              // two 128-column fp32 accumulators
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024;   // + slack for 1024 B alignment
constexpr int kSuper = 8;                     // super-block edge, in tiles
constexpr int kChunkKb = 32;                  // 1024 columns per accumulation chain inside TMEM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M=128, N=128
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void decode_tri_tile(int t, int nb, int& bi, int& bj) {
  bi = 0;
  while (t >= nb - bi) { t -= nb - bi; ++bi; }
  bj = bi + t;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// grid (upper-triangular 128-tiles, K splits).  partial[split][n_pad][n_pad].
// A CTA walks its whole K range; the fp32 accumulation chain inside TMEM is cut every kChunkKb K-blocks
// (1024 columns): the MMA warp alternates between two 128-column TMEM accumulators while the epilogue
// warps drain the finished one into registers (fp32 adds), so chunk c+1 is multiplied while chunk c is
// read out.  Only one partial tile per CTA reaches memory.
__global__ void __launch_bounds__(kThreads, 1)
syrk_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, int nb,
               int kb_total, int kb_per_split, float* __restrict__ partial, int n_pad, int super) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31;
  // Rasterisation: blockIdx.x walks super-blocks of kSuper x kSuper tiles (upper-triangular order), so the
  // ~148 CTAs in flight touch few distinct row/column blocks per K step and share them through L2.
  // Slots below the diagonal or past the edge exit at once (uniformly, before any barrier or TMEM use).
  int bi, bj;
  {
    const int sb_count = (nb + super - 1) / super;
    int sbi, sbj;
    decode_tri_tile((int)blockIdx.x / (super * super), sb_count, sbi, sbj);
    const int within = (int)blockIdx.x % (super * super);
    bi = sbi * super + within / super;
    bj = sbj * super + within % super;
    if (bi >= nb || bj >= nb || bi > bj) return;
  }
  const bool diag = bi == bj;
  const int split = blockIdx.y;
  const int kb0 = split * kb_per_split;
  const int nkb = min(kb_total, kb0 + kb_per_split) - kb0;
  const int nchunks = (nkb + kChunkKb - 1) / kChunkKb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = diag ? 2u * kTileBytes : 4u * kTileBytes;
      for (int it = 0; it < nkb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (uint32_t)(it / kStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], tx);
        uint8_t* st = smem + (size_t)s * kStageBytes;
        const int k = (kb0 + it) * kBK;
        tma_load_2d(&map_hi, &full_bar[s], st, k, bi * kBM);
        tma_load_2d(&map_lo, &full_bar[s], st + kTileBytes, k, bi * kBM);
        if (!diag) {
          tma_load_2d(&map_hi, &full_bar[s], st + 2 * kTileBytes, k, bj * kBN);
          tma_load_2d(&map_lo, &full_bar[s], st + 3 * kTileBytes, k, bj * kBN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int it = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        const uint32_t use = (uint32_t)(c >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[buf], use ^ 1u);        // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * kBN);
        const int it_end = min(nkb, it + kChunkKb);
        for (int first = 1; it < it_end; ++it, first = 0) {
          const int s = it % kStages;
          const uint32_t ph = (uint32_t)(it / kStages) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + (size_t)s * kStageBytes);
          const uint32_t a_lo = a_hi + kTileBytes;
          const uint32_t b_hi = diag ? a_hi : a_hi + 2 * kTileBytes;
          const uint32_t b_lo = diag ? a_lo : a_hi + 3 * kTileBytes;
#pragma unroll
          for (int j = 0; j < kBK / 8; ++j) {         // UMMA K = 8 tf32 = 32 bytes inside the 128 B swizzle row
            const uint32_t off = (uint32_t)j * 32u;
            mma_tf32(tmem_acc, smem_desc_sw128(a_hi + off), smem_desc_sw128(b_hi + off), (first && j == 0) ? 0u : 1u);
            mma_tf32(tmem_acc, smem_desc_sw128(a_hi + off), smem_desc_sw128(b_lo + off), 1u);
            mma_tf32(tmem_acc, smem_desc_sw128(a_lo + off), smem_desc_sw128(b_hi + off), 1u);
          }
          tc_commit(&empty_bar[s]);                      // stage reusable once these MMAs have read it
        }
        tc_commit(&tmem_full_bar[buf]);                  // this chunk's accumulator is complete
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
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
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kBN + cg * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int x = 0; x < 32; ++x) acc[cg * 32 + x] += __uint_as_float(v[x]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);  // 4 epilogue warps -> barrier count 4
    }
    float* out = partial + (size_t)split * n_pad * n_pad + (size_t)(bi * kBM + row) * n_pad + (size_t)bj * kBN;
#pragma unroll
    for (int x = 0; x < kBN; x += 4)
      *reinterpret_cast<float4*>(out + x) = make_float4(acc[x], acc[x + 1], acc[x + 2], acc[x + 3]);
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

inline int grid_slots(int nb, int super) { const int sb = (nb + super - 1) / super; return sb * (sb + 1) / 2 * super * super; }

// Host: 2-D tensor map over a (rows = n_pad, cols = d_pad) fp32 plane, box 128 rows x 32 columns, 128 B swizzle.
inline int make_plane_map(CUtensorMap* map, const float* base, int n_pad, int64_t d_pad) {
  cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)n_pad};
  cuuint64_t strides[1] = {(cuuint64_t)d_pad * 4};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
  cuuint32_t estr[2] = {1, 1};
  // The driver entry point is resolved through the runtime so the library carries no link-time
  // dependency on libcuda (it must load on a box without a driver for the build check).
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      (void)cudaGetLastError();
      return VR_ERR_UNSUPPORTED;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (n_pad=%d d_pad=%lld)", (int)r, n_pad, (long long)d_pad);
    return VR_ERR_CUDA;
  }
  return VR_OK;
}

}  // namespace tc
}  // namespace vr
