// kaldi-cnn_b200/csrc/cnslmat/gemm_tma.cuh
//
// TMA-fed TF32 GEMM pipeline on tcgen05 / TMEM, and the "problems" that ride on it.
// Unlike gemm_tc.cuh no thread ever touches an operand element: one lane issues
// cp.async.bulk.tensor (TMA) boxes that land in shared memory already in the swizzled
// layout the UMMA descriptors describe, one lane issues tcgen05.mma, four warps drain
// the accumulator through a padded shared-memory staging tile into coalesced stores.
//
// Tensor maps are encoded with data type TFLOAT32: the TMA unit rounds FP32 to TF32
// while it copies (measured: 3e-4 max-norm error against FP64 vs 7.5e-4 when the raw
// FP32 bits are truncated by the tensor core), so no rounding pass is needed.
//
// Operand majors, both straight from row-major matrices (no transpose pass):
//   K-major   [MN rows][K cols]  one box {32 k, rows}, SWIZZLE_128B
//   MN-major  [K rows][MN cols]  MN/32 boxes {32 mn, 32 k}, TMA SWIZZLE_128B_ATOM_32B;
//             the only layout the tensor core transposes for 32-bit elements is
//             SWIZZLE_128B_BASE32B (4-k-row atoms): LBO = 4096 (next MN atom),
//             SBO = 512 (next 4 k rows)
//
// Pipeline (one kernel template, parameterised by a Problem):
//   CTA = 128 x 128 output tile, 6 warps, 2 CTAs per SM (3 x 32 KB stages each) so one
//   CTA's epilogue overlaps the other's main loop
//   warp 0    TMA producer (one lane), full / empty mbarrier ring
//   warp 1    TMEM allocator + MMA issuer (one lane), tcgen05.commit frees the stages
//   warps 2-5 epilogue: tcgen05.ld 32x32b -> staging [128][132] (the idle stages) ->
//             Problem::store (row segments, transposed [map][pos] runs, split-K partials,
//             or the fused momentum / weight-decay SGD update)
// M / N / K tails and the zero padding of the convolutions are TMA out-of-bounds
// zero fill, never memory traffic.
//
// Problems:
//   DenseProb    the three GEMMs of the affine layer (nnet2/nnet-component.cc:1216-1258,
//                nnet0/nnet-component-nnet0.cc:1133-1143)
//   ConvRowsProb convolution forward and input-gradient over a channels-last staging
//                copy (conv_tma.cuh)
//   ConvWgradProb convolution weight-gradient (conv_tma.cuh)
//
// Eligibility (host side): 16-byte aligned base and pitches for every operand
// (cuTensorMapEncodeTiled); anything else takes the software-producer kernel.

#ifndef KCNN_GEMM_TMA_CUH_
#define KCNN_GEMM_TMA_CUH_

#include <cuda.h>

#include "gemm_tc.cuh"

namespace kcnn {
namespace tma {

using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tmem_ld32;
using tc::umma_commit;
using tc::umma_tf32;

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int THREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 4;
constexpr int B_STAGE_BYTES = BN * BK * 4;
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ATOM_BYTES = BK * 128;          // one MN-major box: 32 k-rows x 128 bytes
constexpr int PITCH = BN + 4;                 // staging row pitch (floats): conflict-free v4 stores
constexpr int STAGING_BYTES = 128 * PITCH * 4;
// kStages = 3: 2 CTAs per SM (6 stages in flight per SM, one CTA's epilogue under the other's
// main loop) for grids of more than one wave; kStages = 6: 1 CTA per SM with the same bytes
// in flight, for grids that leave at most one CTA per SM anyway (M = 512 FC layers, the
// convolutions) -- with 3 stages those are latency-bound (measured 330 vs 600 TFLOP/s).
template <int kStages>
struct Ring {
  static constexpr int RING_BYTES = kStages * STAGE_BYTES;
  static constexpr int DATA_BYTES = RING_BYTES > STAGING_BYTES ? RING_BYTES : STAGING_BYTES;
  static constexpr int BAR_OFFSET = DATA_BYTES;
  static constexpr int SMEM_TOTAL = BAR_OFFSET + (2 * kStages + 1) * 8 + 16 + 1024;
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// kPair: cta_group::2 form -- the box lands in THIS CTA's shared memory, the bytes are
// credited to the mbarrier of the pair's leader CTA (`bar` is then a shared::cluster address).
#define KCNN_TMA_LOAD(NAME, DIMS, COORD_FMT, COORD_ARGS, COORD_OPS, BAR_IDX)                                   \
  template <bool kPair>                                                                                       \
  __device__ __forceinline__ void NAME(uint32_t dst, const CUtensorMap *map, COORD_ARGS, uint32_t bar) {      \
    if (kPair)                                                                                                \
      asm volatile("cp.async.bulk.tensor." DIMS ".cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes " \
                   "[%0], [%1, " COORD_FMT "], [%" BAR_IDX "];" ::"r"(dst), "l"(map), COORD_OPS, "r"(bar)      \
                   : "memory");                                                                               \
    else                                                                                                      \
      asm volatile("cp.async.bulk.tensor." DIMS ".shared::cluster.global.mbarrier::complete_tx::bytes "       \
                   "[%0], [%1, " COORD_FMT "], [%" BAR_IDX "];" ::"r"(dst), "l"(map), COORD_OPS, "r"(bar)      \
                   : "memory");                                                                               \
  }
#define KCNN_COMMA ,
KCNN_TMA_LOAD(tma_load_2d, "2d", "{%2, %3}", int c0 KCNN_COMMA int c1, "r"(c0) KCNN_COMMA "r"(c1), "4")
KCNN_TMA_LOAD(tma_load_3d, "3d", "{%2, %3, %4}", int c0 KCNN_COMMA int c1 KCNN_COMMA int c2,
              "r"(c0) KCNN_COMMA "r"(c1) KCNN_COMMA "r"(c2), "5")
KCNN_TMA_LOAD(tma_load_4d, "4d", "{%2, %3, %4, %5}", int c0 KCNN_COMMA int c1 KCNN_COMMA int c2 KCNN_COMMA int c3,
              "r"(c0) KCNN_COMMA "r"(c1) KCNN_COMMA "r"(c2) KCNN_COMMA "r"(c3), "6")
#undef KCNN_COMMA
#undef KCNN_TMA_LOAD
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major SWIZZLE_128B: 128-byte rows, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t desc_k_major(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 k-rows x 128 bytes (32 mn);
// MN atoms LBO apart, 4-k groups SBO = 512 bytes apart (one MMA = 8 k = two groups).
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(ATOM_BYTES >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Problem interface (mt / nt: 128-row / 128-column tile indices):
//   static constexpr bool kAMn, kBMn
//   void kb_range(int z, int &begin, int &end) const             K-blocks of split z
//   uint32_t tx_bytes() const                                     bytes one CTA's boxes of a stage deliver
//   void load<kPair>(kb, a_addr, b_addr, bar, &map_a, &map_b, mt, nt)   A rows of tile mt, B rows of tile nt
//   void prefetch(tid, mt, nt) const                              epilogue threads, before the accumulator is ready
//   void store<kRows>(stage, tid, mt, nt, z) const                128 epilogue threads, stage[128][PITCH];
//                                                                 kRows = rows of global loads kept in flight
//
// kPair = false: one CTA computes the 128 x 128 tile (blockIdx.x, blockIdx.y).
// kPair = true:  a cluster of two CTAs (one TPC) computes a 256 x 256 tile with
//   tcgen05.mma.cta_group::2 (M = 256, N = 256): CTA r stages A rows of tile 2*pair + r
//   (= blockIdx.x) and the B rows of column tile 2*blockIdx.y + r -- 32 KB per K-block per
//   SM for twice the FLOPs per SM, which is what lifts these TF32 GEMMs off the ~40 B/clk
//   per-SM L2 -> shared-memory ingress limit (profiles/).  The leader CTA issues the MMAs
//   for both; its tcgen05.commit multicasts the stage release and the "accumulator ready"
//   signal to both CTAs; each CTA drains its own 128 lanes x 256 columns of TMEM.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

template <class Prob, bool kPair, int STAGES>
__global__ void __launch_bounds__(THREADS, STAGES <= 2 ? 3 : (STAGES <= 3 ? 2 : 1))
tma_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const Prob prob) {
  kcnn::pdl_trigger();
  constexpr int kCols = kPair ? 2 * BN : BN;        // TMEM columns = accumulator columns per CTA
  constexpr int BAR_OFFSET = Ring<STAGES>::BAR_OFFSET;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 8 * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8 * (2 * STAGES);
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem_gen + BAR_OFFSET + 8 * (2 * STAGES + 1));

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int mt = blockIdx.x;
  const int nt_load = kPair ? 2 * (int)blockIdx.y + (int)rank : (int)blockIdx.y;
  int kb_begin, kb_end;
  prob.kb_range((int)blockIdx.z, kb_begin, kb_end);
  const int num_kb = kb_end - kb_begin;

  if (warp == 1) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_ptr_smem)), "r"((uint32_t)kCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_ptr_smem)), "r"((uint32_t)kCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (t == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar(s), kPair ? 2 : 1);      // pair: one arrive.expect_tx per CTA's producer
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  kcnn::pdl_wait();                 // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer --
    if (lane == 0) {
      const uint32_t tx = prob.tx_bytes();
      for (int i = 0; i < num_kb; i++) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait(empty_bar(s), ((i / STAGES) - 1) & 1);
        const uint32_t a_addr = smem_base + s * STAGE_BYTES;
        // the full barrier that counts: this CTA's, or the pair leader's
        const uint32_t fb = kPair ? map_to_cta(full_bar(s), 0) : full_bar(s);
        if (kPair) mbar_expect_tx_cluster(fb, tx); else mbar_expect_tx(fb, tx);
        prob.template load<kPair>(kb_begin + i, a_addr, a_addr + A_STAGE_BYTES, fb, &map_a, &map_b, mt, nt_load);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer --
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(kPair ? 2 * BM : BM, kCols, Prob::kAMn, Prob::kBMn);
      for (int i = 0; i < num_kb; i++) {
        const int s = i % STAGES;
        mbar_wait(full_bar(s), (i / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * STAGE_BYTES;
        const uint32_t b_addr = a_addr + A_STAGE_BYTES;
        const uint64_t adesc = Prob::kAMn ? desc_mn_major(a_addr) : desc_k_major(a_addr);
        const uint64_t bdesc = Prob::kBMn ? desc_mn_major(b_addr) : desc_k_major(b_addr);
#pragma unroll
        for (int k = 0; k < BK / 8; k++) {
          // one MMA = 8 TF32 of K: +32 bytes along a K-major row, +1024 bytes (two 4-row
          // groups) in an MN-major atom; the start-address field is in 16-byte units
          const uint64_t ad = adesc + (Prob::kAMn ? 64 * k : 2 * k);
          const uint64_t bd = bdesc + (Prob::kBMn ? 64 * k : 2 * k);
          if (kPair) umma_tf32_pair(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
          else       umma_tf32(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        if (kPair) umma_commit_pair(empty_bar(s)); else umma_commit(empty_bar(s));
      }
      if (kPair) umma_commit_pair(accum_bar); else umma_commit(accum_bar);
    }
    __syncwarp();
  } else {
    // ----------------------------------------------------------------- epilogue --
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    float *stage = reinterpret_cast<float *>(smem_gen);
    const int tid = t - 64;
#pragma unroll
    for (int h = 0; h < kCols / BN; h++) prob.prefetch(tid, mt, kPair ? 2 * (int)blockIdx.y + h : (int)blockIdx.y);
    if (num_kb > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int h = 0; h < kCols / BN; h++) {
      // TMEM -> registers -> padded staging: lane = tile row, 32 columns per tcgen05.ld
#pragma unroll 1
      for (int j0 = 0; j0 < BN; j0 += 32) {
        uint32_t v[32];
        if (num_kb > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * BN + j0), v);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = 0u;
        }
        float *dst = stage + (q * 32 + lane) * PITCH + j0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4 *>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      prob.template store<8>(stage, tid, mt, kPair ? 2 * (int)blockIdx.y + h : (int)blockIdx.y, (int)blockIdx.z);
      if (h + 1 < kCols / BN) asm volatile("bar.sync 1, 128;" ::: "memory");     // staging is reused
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kPair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kCols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kCols)
                   : "memory");
  }
}

// ---- persistent variant ------------------------------------------------------------------
// For grids of several waves (the weight gradients: 1024 tiles x 16 K-blocks for FC2): one
// CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (tile row fastest, so CTAs
// running together share B tiles in L2).  The 4-stage operand ring runs straight across tile
// boundaries, the accumulator is double-buffered in TMEM (2 x 128 columns) and the staging
// tile has its own shared memory, so the epilogue of tile i -- TMEM drain, then the HBM-heavy
// fused SGD read-modify-write of the W / prev_grad tile -- overlaps the main loop of tile
// i + 1 inside the SAME CTA instead of relying on a second resident CTA.
constexpr int PSTAGES = 4;
struct PRing {
  static constexpr int RING_BYTES = PSTAGES * STAGE_BYTES;
  static constexpr int STAGING_OFFSET = RING_BYTES;
  static constexpr int BAR_OFFSET = RING_BYTES + STAGING_BYTES;
  static constexpr int SMEM_TOTAL = BAR_OFFSET + (2 * PSTAGES + 4) * 8 + 16 + 1024;
};

__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <class Prob>
__global__ void __launch_bounds__(THREADS, 1)
tma_gemm_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                           const Prob prob, int tiles_m, int tiles_n, int splits) {
  kcnn::pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + PRing::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 8 * (PSTAGES + s); };
  auto acc_full = [&](int b) { return bar_base + 8 * (2 * PSTAGES + b); };
  auto acc_empty = [&](int b) { return bar_base + 8 * (2 * PSTAGES + 2 + b); };
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem_gen + PRing::BAR_OFFSET + 8 * (2 * PSTAGES + 4));

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int total_tiles = tiles_m * tiles_n * splits;
  auto decode = [&](int tile, int &mt, int &nt, int &z) {
    mt = tile % tiles_m;
    const int r = tile / tiles_m;
    nt = r % tiles_n;
    z = r / tiles_n;
  };

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)), "r"((uint32_t)(2 * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < PSTAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  kcnn::pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer --
    if (lane == 0) {
      const uint32_t tx = prob.tx_bytes();
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt, nt, z, kb_begin, kb_end;
        decode(tile, mt, nt, z);
        prob.kb_range(z, kb_begin, kb_end);
        for (int kb = kb_begin; kb < kb_end; kb++, it++) {
          const int s = it % PSTAGES;
          if (it >= PSTAGES) mbar_wait(empty_bar(s), ((it / PSTAGES) - 1) & 1);
          const uint32_t a_addr = smem_base + s * STAGE_BYTES;
          mbar_expect_tx(full_bar(s), tx);
          prob.template load<false>(kb, a_addr, a_addr + A_STAGE_BYTES, full_bar(s), &map_a, &map_b, mt, nt);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer --
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, Prob::kAMn, Prob::kBMn);
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tcount++) {
        int mt, nt, z, kb_begin, kb_end;
        decode(tile, mt, nt, z);
        prob.kb_range(z, kb_begin, kb_end);
        const int buf = tcount & 1, use = tcount >> 1;
        if (use > 0) mbar_wait(acc_empty(buf), (use - 1) & 1);       // epilogue drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int i = 0; i < kb_end - kb_begin; i++, it++) {
          const int s = it % PSTAGES;
          mbar_wait(full_bar(s), (it / PSTAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
          const uint64_t adesc = Prob::kAMn ? desc_mn_major(a_addr) : desc_k_major(a_addr);
          const uint64_t bdesc = Prob::kBMn ? desc_mn_major(b_addr) : desc_k_major(b_addr);
#pragma unroll
          for (int k = 0; k < BK / 8; k++) {
            const uint64_t ad = adesc + (Prob::kAMn ? 64 * k : 2 * k);
            const uint64_t bd = bdesc + (Prob::kBMn ? 64 * k : 2 * k);
            umma_tf32(tmem_d, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(acc_full(buf));
      }
    }
    __syncwarp();
  } else {
    // ----------------------------------------------------------------- epilogue --
    const int q = warp & 3;
    float *stage = reinterpret_cast<float *>(smem_gen + PRing::STAGING_OFFSET);
    const int tid = t - 64;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tcount++) {
      int mt, nt, z;
      decode(tile, mt, nt, z);
      const int buf = tcount & 1, use = tcount >> 1;
      prob.prefetch(tid, mt, nt);
      mbar_wait(acc_full(buf), use & 1);
      tc_fence_after();
      asm volatile("bar.sync 1, 128;" ::: "memory");              // previous tile's staging fully consumed
#pragma unroll 1
      for (int j0 = 0; j0 < BN; j0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + j0), v);
        float *dst = stage + (q * 32 + lane) * PITCH + j0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4 *>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      tc_fence_before();
      mbar_arrive_local(acc_empty(buf));                           // TMEM buffer may be overwritten
      asm volatile("bar.sync 1, 128;" ::: "memory");
      prob.template store<16>(stage, tid, mt, nt, z);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN))
                 : "memory");
  }
}

// ------------------------------------------------------------------ DenseProb --

// What the epilogue does with the accumulator tile.
enum EpiMode {
  EPI_STORE = 0,      // out = acc (+ bias_n)
  EPI_PARTIAL = 1,    // workspace[z] = acc           (split-K)
  EPI_SGD = 2,        // prev = m prev - lr wd W + lr acc ; W += prev   (out = W, aux = prev)
};

// prev = momentum*prev ; prev += a_decay*W ; prev += a_grad*grad ; W += prev  -- the four
// roundings of the reference's Scale / AddMat / AddMat / AddMat, as sgd_momentum_kernel.
struct SgdCoef {
  float momentum, a_decay, a_grad;
};
__device__ __forceinline__ void sgd_apply(float &w, float &p, float g, const SgdCoef &c) {
  p = p * c.momentum;
  p = fmaf(c.a_decay, w, p);
  p = fmaf(c.a_grad, g, p);
  w = w + p;
}

// Row-major 128-column segment store shared by the dense and the weight-gradient
// problems: tid -> (row group, 4 columns); each warp instruction writes one 512-byte
// row segment.  row_of(m) maps a tile row to the output row; a negative result marks a padding
// row of the GEMM that has no output row (skipped).
template <int kEpi, int kRows, class RowMap>
__device__ __forceinline__ void store_rows(const float *stage, int tid, int m0, int n0, int M, int N,
                                           float *obase, int ld, const float *bias_n, float *aux,
                                           const SgdCoef &sgd, RowMap row_of, bool relu = false) {
  const int wl = tid >> 5, lane = tid & 31;
  const int n = n0 + 4 * lane;
  const bool vec_ok = (n + 3 < N) && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(obase) & 15u) == 0);
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kEpi == EPI_STORE && bias_n) {
    if (n < N) bias4.x = __ldg(bias_n + n);
    if (n + 1 < N) bias4.y = __ldg(bias_n + n + 1);
    if (n + 2 < N) bias4.z = __ldg(bias_n + n + 2);
    if (n + 3 < N) bias4.w = __ldg(bias_n + n + 3);
  }
  if (kEpi == EPI_SGD && vec_ok) {
    // nnet0/nnet-component-nnet0.cc:767-773, 1138-1142 on the tile: W and prev_grad rows are
    // read in batches of kRows rows (2 kRows independent 128-bit loads per thread in flight; the
    // lines were L2-prefetched by Problem::prefetch while the main loop ran), updated, written back.
    constexpr int RB = kRows;
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += RB) {
      float4 w[RB], pv[RB];
      size_t roff[RB];
      bool okr[RB];
#pragma unroll
      for (int i = 0; i < RB; i++) {
        const int mt = wl * 32 + r0 + i;
        const int row = m0 + mt < M ? row_of(m0 + mt) : -1;
        okr[i] = row >= 0;
        roff[i] = (size_t)(okr[i] ? row : 0) * ld + n;
        if (okr[i]) {
          w[i] = *reinterpret_cast<const float4 *>(obase + roff[i]);
          pv[i] = *reinterpret_cast<const float4 *>(aux + roff[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < RB; i++) {
        const int mt = wl * 32 + r0 + i;
        if (!okr[i]) continue;
        const float4 a = *reinterpret_cast<const float4 *>(stage + mt * PITCH + 4 * lane);
        sgd_apply(w[i].x, pv[i].x, a.x, sgd); sgd_apply(w[i].y, pv[i].y, a.y, sgd);
        sgd_apply(w[i].z, pv[i].z, a.z, sgd); sgd_apply(w[i].w, pv[i].w, a.w, sgd);
        *reinterpret_cast<float4 *>(aux + roff[i]) = pv[i];
        *reinterpret_cast<float4 *>(obase + roff[i]) = w[i];
      }
    }
    return;
  }
#pragma unroll 4
  for (int r = 0; r < 32; r++) {
    const int mt = wl * 32 + r;
    if (m0 + mt >= M) break;
    const int row = row_of(m0 + mt);
    if (row < 0) continue;
    float4 a = *reinterpret_cast<const float4 *>(stage + mt * PITCH + 4 * lane);
    const size_t roff = (size_t)row * ld + n;
    float *orow = obase + roff;
    if (kEpi == EPI_SGD) {
      float *prow = aux + roff;
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (n + j < N) {
          float w = orow[j], pv = prow[j];
          sgd_apply(w, pv, av[j], sgd);
          prow[j] = pv;
          orow[j] = w;
        }
    } else {
      a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
      if (relu) {       // RectifiedLinearComponent::Propagate fused: x > 0 ? x : 0
        a.x = a.x > 0.0f ? a.x : 0.0f; a.y = a.y > 0.0f ? a.y : 0.0f;
        a.z = a.z > 0.0f ? a.z : 0.0f; a.w = a.w > 0.0f ? a.w : 0.0f;
      }
      if (vec_ok) {
        *reinterpret_cast<float4 *>(orow) = a;
      } else {
        if (n < N) orow[0] = a.x;
        if (n + 1 < N) orow[1] = a.y;
        if (n + 2 < N) orow[2] = a.z;
        if (n + 3 < N) orow[3] = a.w;
      }
    }
  }
}

// L2 prefetch of the 128 x 128 tiles of two row-major matrices the epilogue will read
// (EPI_SGD: W and prev_grad), issued by the 128 epilogue threads before the main loop ends.
template <class RowMap>
__device__ __forceinline__ void prefetch_tile_l2(int tid, int m0, int n0, int M, int N, const float *a,
                                                 const float *b, int ld, RowMap row_of) {
  // 128 rows x 4 lines of 128 bytes per matrix; thread -> (row, line) pairs
  for (int i = tid; i < 128 * 4; i += 128) {
    const int r = i >> 2, line = i & 3;
    const int m = m0 + r, n = n0 + line * 32;
    const int row = m < M ? row_of(m) : -1;
    if (row >= 0 && n < N) {
      const size_t off = (size_t)row * ld + n;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a + off));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(b + off));
    }
  }
}

struct IdentityRow {
  __device__ __forceinline__ int operator()(int m) const { return m; }
};

template <bool kAMn_, bool kBMn_, int kEpi>
struct DenseProb {
  static constexpr bool kAMn = kAMn_, kBMn = kBMn_;
  int M, N, K;
  int kb_per_split;
  float *out;              // row-major [M][ldo]
  int ldo;
  const float *bias_n;     // per column, or nullptr
  float *workspace;        // [splits][M][N] partials
  float *aux;              // EPI_SGD: prev_grad, same shape / pitch as out
  SgdCoef sgd;
  int relu;                // EPI_STORE: max(., 0) after the bias

  __device__ __forceinline__ void kb_range(int z, int &b, int &e) const {
    const int total = (K + BK - 1) / BK;
    b = z * kb_per_split;
    e = min(total, b + kb_per_split);
    if (e < b) e = b;
  }
  __device__ __forceinline__ uint32_t tx_bytes() const { return STAGE_BYTES; }
  template <bool kPair>
  __device__ __forceinline__ void load(int kb, uint32_t a_addr, uint32_t b_addr, uint32_t bar,
                                       const CUtensorMap *ma, const CUtensorMap *mb, int mt, int nt) const {
    const int m0 = mt * BM, n0 = nt * BN, k0 = kb * BK;
    if (kAMn) {
#pragma unroll
      for (int i = 0; i < BM / 32; i++) tma_load_2d<kPair>(a_addr + i * ATOM_BYTES, ma, m0 + 32 * i, k0, bar);
    } else {
      tma_load_2d<kPair>(a_addr, ma, k0, m0, bar);
    }
    if (kBMn) {
#pragma unroll
      for (int i = 0; i < BN / 32; i++) tma_load_2d<kPair>(b_addr + i * ATOM_BYTES, mb, n0 + 32 * i, k0, bar);
    } else {
      tma_load_2d<kPair>(b_addr, mb, k0, n0, bar);
    }
  }
  __device__ __forceinline__ void prefetch(int tid, int mt, int nt) const {
    if (kEpi == EPI_SGD) prefetch_tile_l2(tid, mt * BM, nt * BN, M, N, out, aux, ldo, IdentityRow());
  }
  template <int kRows>
  __device__ __forceinline__ void store(const float *stage, int tid, int mt, int nt, int z) const {
    const int m0 = mt * BM, n0 = nt * BN;
    if (kEpi == EPI_PARTIAL)
      store_rows<EPI_STORE, kRows>(stage, tid, m0, n0, M, N, workspace + (size_t)z * M * N, N, nullptr, nullptr,
                                   sgd, IdentityRow());
    else
      store_rows<kEpi, kRows>(stage, tid, m0, n0, M, N, out, ldo, bias_n, aux, sgd, IdentityRow(), relu != 0);
  }
};

// Optional tail of the split-K reduction: the last stage of a per-column sum (the convolution's
// bias gradient from the pack kernel's partial rows), run by extra blocks of the same launch:
//   dst[c] = alpha * sum_r partial[r][c]  (+ dst[c] when accumulate)
struct ColSumTail {
  const float *partial;   // [rows][cols], nullptr = no tail
  int rows, cols;
  float *dst;
  float alpha;
  int accumulate;
};
constexpr int kColSumTailCols = 32;      // columns per tail block (256 threads = 32 columns x 8 row groups)
inline unsigned colsum_tail_blocks(int cols) { return (unsigned)((cols + kColSumTailCols - 1) / kColSumTailCols); }

// out[row_of(m)][n] = sum_z ws[z][m][n] (+ bias_n[n]), or the SGD update with that sum as the
// gradient.  N % 4 == 0, 16-byte aligned rows (checked by the launchers).
template <int kEpi, class RowMap>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float *__restrict__ ws, int splits, int M, int N, float *__restrict__ out, int ldo,
                     const float *__restrict__ bias_n, float *__restrict__ aux, SgdCoef sgd, RowMap row_of,
                     ColSumTail tail, unsigned main_blocks, int relu) {
  kcnn::pdl_prologue();
  if (blockIdx.x >= main_blocks) {
    // kColSumTailCols columns per block: 32 lanes x 8 row groups, 4 independent loads in flight
    // per thread (a single thread walking all rows of a column made this tail, not the
    // reduction, the duration of the launch: ~13 us for 128 rows); fixed summation order.
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (int)(blockIdx.x - main_blocks) * kColSumTailCols + tx;
    const bool live = tail.partial != nullptr && c < tail.cols;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    if (live) {
      const float *p = tail.partial + c;
      const size_t ld = (size_t)tail.cols;
      int r = ty;
      for (; r + 24 < tail.rows; r += 32) {
        s0 += __ldg(p + (size_t)r * ld);
        s1 += __ldg(p + (size_t)(r + 8) * ld);
        s2 += __ldg(p + (size_t)(r + 16) * ld);
        s3 += __ldg(p + (size_t)(r + 24) * ld);
      }
      for (; r < tail.rows; r += 8) s0 += __ldg(p + (size_t)r * ld);
    }
    red[ty][tx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ty == 0 && live) {
      float s = red[0][tx];
#pragma unroll
      for (int i = 1; i < 8; i++) s += red[i][tx];
      tail.dst[c] = tail.accumulate ? fmaf(tail.alpha, s, tail.dst[c]) : tail.alpha * s;
    }
    return;
  }
  const long long total4 = ((long long)M * N) >> 2;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const long long e = i << 2;
  const int m = (int)(e / N), n = (int)(e - (long long)m * N);
  float4 s = __ldg(reinterpret_cast<const float4 *>(ws + e));
  for (int z = 1; z < splits; z++) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(ws + (size_t)z * M * N + e));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const int row = row_of(m);
  if (row < 0) return;
  const size_t off = (size_t)row * ldo + n;
  if (kEpi == EPI_SGD) {
    float4 w = *reinterpret_cast<const float4 *>(out + off);
    float4 pv = *reinterpret_cast<const float4 *>(aux + off);
    sgd_apply(w.x, pv.x, s.x, sgd); sgd_apply(w.y, pv.y, s.y, sgd);
    sgd_apply(w.z, pv.z, s.z, sgd); sgd_apply(w.w, pv.w, s.w, sgd);
    *reinterpret_cast<float4 *>(aux + off) = pv;
    *reinterpret_cast<float4 *>(out + off) = w;
  } else {
    if (bias_n) {
      s.x += __ldg(bias_n + n); s.y += __ldg(bias_n + n + 1); s.z += __ldg(bias_n + n + 2); s.w += __ldg(bias_n + n + 3);
    }
    if (relu) {
      s.x = s.x > 0.0f ? s.x : 0.0f; s.y = s.y > 0.0f ? s.y : 0.0f;
      s.z = s.z > 0.0f ? s.z : 0.0f; s.w = s.w > 0.0f ? s.w : 0.0f;
    }
    *reinterpret_cast<float4 *>(out + off) = s;
  }
}

// ------------------------------------------------------------------- host side --

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn();         // kernels_gemm.cu (driver entry point, resolved once)
int tma_data_type();                     // CU_TENSOR_MAP_DATA_TYPE_* used for the operands
bool enabled();                          // KCNN_TMA=0 disables the TMA paths
bool pair_enabled();                     // KCNN_TMA_PAIR=1 opts in to the 2-CTA (cta_group::2) tiles
bool deep_ring_enabled();                // KCNN_TMA_DEEP=0 keeps 3 stages for one-wave grids
bool persistent_enabled();               // KCNN_TMA_PERSIST=0 keeps multi-wave grids on one tile per CTA

// Grow-only device scratch, one buffer per slot (kernels_gemm.cu).  Returns nullptr when it
// would have to grow while the stream is being captured; callers then take another path.
// SCRATCH_SPLITK_ROWS: partials of the convolution fprop / dgrad split; its own slot because
// inside a convolution's Backprop the dgrad GEMM runs CONCURRENTLY with the weight-gradient GEMM
// (kcnn::ForkJoin), whose partials live in SCRATCH_SPLITK.
enum ScratchSlot { SCRATCH_SPLITK = 0, SCRATCH_XCL = 1, SCRATCH_DYCL = 2, SCRATCH_BIAS = 3, SCRATCH_SPLITK_ROWS = 4,
                   SCRATCH_SLOTS = 5 };
float *scratch(int slot, size_t bytes);

// Tensor map of rank 2 to 4 over FP32 data; dims[0] is the contiguous axis, strides_bytes[i]
// is the pitch of dims[i + 1].  mn_major picks the 32-byte-atom swizzle.
inline bool encode_map(CUtensorMap *map, const float *base, int rank, const unsigned long long *dims,
                       const unsigned long long *strides_bytes, const unsigned *box, bool mn_major) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  if (!host_aligned16(base)) return false;
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; i++) {
    if (dims[i] == 0 || box[i] == 0 || box[i] > 256) return false;
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; i++) {
    if (strides_bytes[i] % 16 != 0 || strides_bytes[i] == 0) return false;
    gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(map, (CUtensorMapDataType)tma_data_type(), (cuuint32_t)rank, const_cast<float *>(base), gdim,
                  gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// A pitched row-major matrix as a GEMM operand.
struct Matrix {
  const float *base;
  int rows, cols, ld;
};

// 2-D tensor map over `m` with a {32 cols, box_rows} box.
inline bool encode_2d(CUtensorMap *map, const Matrix &m, int box_rows, bool mn_major) {
  if (m.rows <= 0 || m.cols <= 0) return false;
  unsigned long long dims[2] = {(unsigned long long)m.cols, (unsigned long long)m.rows};
  unsigned long long str[1] = {(unsigned long long)m.ld * 4};
  unsigned box[2] = {32, (unsigned)box_rows};
  return encode_map(map, m.base, 2, dims, str, box, mn_major);
}

struct Epilogue {
  int mode = EPI_STORE;
  const float *bias_n = nullptr;
  float *aux = nullptr;          // EPI_SGD: prev_grad
  SgdCoef sgd = {0.f, 0.f, 0.f};
  int relu = 0;                  // EPI_STORE: rectify after the bias
};

inline int pick_splits(long long tiles, int num_kb) {
  if (tiles >= 100) return 1;
  long long want = (2 * kNumSMs) / tiles;
  long long max_by_k = num_kb / 8;            // at least 8 K-blocks per split
  if (want > max_by_k) want = max_by_k;
  if (want > 16) want = 16;
  return (int)(want < 1 ? 1 : want);
}



// One tile per CTA; kPair launches 2-CTA clusters.
template <class Prob, bool kPair, int kStages>
void launch_variant(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, const Prob &p, dim3 grid) {
  auto kernel = tma_gemm_kernel<Prob, kPair, kStages>;
  constexpr int smem = Ring<kStages>::SMEM_TOTAL;
  static bool attr_set = false;          // one flag per instantiation
  if (!attr_set) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = true;
  }
  launch_kernel(kernel, grid, dim3(THREADS), (size_t)smem, st, kPair ? 2u : 1u, ma, mb, p);
}

template <class Prob>
void launch_persistent(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, const Prob &p, dim3 tiles) {
  auto kernel = tma_gemm_persistent_kernel<Prob>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PRing::SMEM_TOTAL);
    attr_set = true;
  }
  long long total = (long long)tiles.x * tiles.y * tiles.z;
  unsigned ctas = (unsigned)(total < kNumSMs ? total : kNumSMs);
  launch_kernel(kernel, dim3(ctas), dim3(THREADS), (size_t)PRing::SMEM_TOTAL, st, 1u, ma, mb, p, (int)tiles.x,
                (int)tiles.y, (int)tiles.z);
}

// grid = (128-row tiles, 128-column tiles, splits).  pair: 2-CTA clusters along x, 256-column
// tiles along y (grid.x rounded up to even, grid.y halved).
template <class Prob>
void launch_prob(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, const Prob &p, dim3 grid,
                 int kb_per_tile, bool pair = false) {
  if (pair) {
    launch_variant<Prob, true, 3>(st, ma, mb, p, dim3((grid.x + 1) & ~1u, (grid.y + 1) / 2, grid.z));
  } else if ((long long)grid.x * grid.y * grid.z <= kNumSMs && deep_ring_enabled()) {
    launch_variant<Prob, false, 6>(st, ma, mb, p, grid);
  } else if ((long long)grid.x * grid.y * grid.z > kNumSMs + kNumSMs / 2 && kb_per_tile <= 64 &&
             persistent_enabled()) {
    // many SHORT tiles (weight gradients): the epilogue is a large share of a tile, overlap it.
    // Long-K tiles keep one tile per CTA, 2 CTAs per SM: 6 stages in flight per SM beat the
    // persistent kernel's 4 (4096^3: 599 vs 488 TFLOP/s).
    launch_persistent<Prob>(st, ma, mb, p, grid);
  } else {
    launch_variant<Prob, false, 3>(st, ma, mb, p, grid);
  }
}

// Pair tiles pay when both extents fill a 256 x 256 tile reasonably.
inline bool use_pair(int M, int N) { return pair_enabled() && M >= 256 && N >= 192; }

// D[M x N] (row-major, pitch ldo) = A B^T with A logical [M x K], B logical [N x K].
// a / b are the matrices as stored: K-major -> [MN][K], MN-major -> [K][MN].
// Returns false (nothing launched) when an operand is not TMA-addressable.
template <bool kAMn, bool kBMn>
bool gemm(cudaStream_t st, const Matrix &a, const Matrix &b, int M, int N, int K, float *out, int ldo,
          const Epilogue &epi, bool allow_split) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  CUtensorMap ma, mb;
  if (!encode_2d(&ma, a, kAMn ? BK : BM, kAMn)) return false;
  if (!encode_2d(&mb, b, kBMn ? BK : BN, kBMn)) return false;
  const int num_kb = (K + BK - 1) / BK;
  const bool out_vec = (N & 3) == 0 && (ldo & 3) == 0 && host_aligned16(out);
  const bool pair = use_pair(M, N);
  int splits = 1;
  if (allow_split && out_vec) {
    long long ctas = pair ? (long long)(2 * ceil_div_u(M, 2 * BM)) * ceil_div_u(N, 2 * BN)
                          : (long long)ceil_div_u(M, BM) * ceil_div_u(N, BN);
    splits = pick_splits(ctas, num_kb);
  }
  int per = (num_kb + splits - 1) / splits;
  splits = (num_kb + per - 1) / per;
  float *ws = nullptr;
  if (splits > 1) {
    ws = scratch(SCRATCH_SPLITK, (size_t)splits * M * N * sizeof(float));
    if (!ws) { splits = 1; per = num_kb; }
  }
  dim3 grid(ceil_div_u(M, BM), ceil_div_u(N, BN), splits);
  auto fill = [&](auto &p) {
    p.M = M; p.N = N; p.K = K; p.kb_per_split = per;
    p.out = out; p.ldo = ldo; p.bias_n = epi.bias_n; p.workspace = ws; p.aux = epi.aux; p.sgd = epi.sgd;
    p.relu = epi.relu;
  };
  if (splits > 1) {
    DenseProb<kAMn, kBMn, EPI_PARTIAL> p; fill(p);
    launch_prob(st, ma, mb, p, grid, per, pair);
    const unsigned blocks = ceil_div_u(((long long)M * N) >> 2, 256);
    if (epi.mode == EPI_SGD)
      KCNN_LAUNCH((splitk_reduce_kernel<EPI_SGD, IdentityRow>), blocks, 256, 0, st, ws, splits, M, N, out, ldo,
                  nullptr, epi.aux, epi.sgd, IdentityRow(), ColSumTail{nullptr, 0, 0, nullptr, 0.f, 0}, blocks, 0);
    else
      KCNN_LAUNCH((splitk_reduce_kernel<EPI_STORE, IdentityRow>), blocks, 256, 0, st, ws, splits, M, N, out, ldo,
                  epi.bias_n, nullptr, epi.sgd, IdentityRow(), ColSumTail{nullptr, 0, 0, nullptr, 0.f, 0}, blocks,
                  epi.relu);
  } else if (epi.mode == EPI_SGD) {
    DenseProb<kAMn, kBMn, EPI_SGD> p; fill(p);
    launch_prob(st, ma, mb, p, grid, per, pair);
  } else {
    DenseProb<kAMn, kBMn, EPI_STORE> p; fill(p);
    launch_prob(st, ma, mb, p, grid, per, pair);
  }
  return true;
}

}  // namespace tma
}  // namespace kcnn

#endif
