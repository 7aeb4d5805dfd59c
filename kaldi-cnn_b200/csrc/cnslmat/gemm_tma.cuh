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
#include <stdlib.h>

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
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
// 128-bit load from the shared memory of a CTA of this cluster (address from mapa)
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(cluster_addr)
               : "memory");
  return v;
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

//
// kClusterK: split-K WITHOUT a workspace or a second launch.  The gridDim.z K-splits of one
// output tile form a thread-block cluster (1, 1, S): every CTA accumulates its K range in TMEM
// and drains it to its staging tile as usual; after a cluster barrier CTA z sums rows
// [z * 128/S, (z+1) * 128/S) of all S staging tiles through distributed shared memory, in split
// order 0..S-1 (the summation order of the former splitk_reduce_kernel: same bits), and runs the
// problem's epilogue -- bias / ReLU / mask, transposed store, or the momentum SGD step -- on just
// those rows.  The reduction is spread over the S CTAs and costs ~1 us (DSMEM ~20 B/clk/SM)
// instead of a 7-16 us launch that re-reads S partial tiles from L2.
template <class Prob, bool kPair, int STAGES, bool kClusterK = false>
__global__ void __launch_bounds__(THREADS, STAGES <= 2 ? 3 : (STAGES <= 3 ? 2 : 1))
tma_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const Prob prob) {
  static_assert(!(kPair && kClusterK), "cluster split-K uses one CTA per tile");
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
      if (!kClusterK) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        prob.template store<8>(stage, tid, mt, kPair ? 2 * (int)blockIdx.y + h : (int)blockIdx.y, (int)blockIdx.z,
                               0, BM);
        if (h + 1 < kCols / BN) asm volatile("bar.sync 1, 128;" ::: "memory");     // staging is reused
      }
    }
  }

  if (kClusterK) {
    cluster_sync_all();                          // every split's partial tile sits in its CTA's staging buffer
    const int S = (int)cluster_nctarank(), me = (int)cluster_ctarank();
    const int rp = (BM + S - 1) / S;
    const int rb = min(BM, me * rp), re = min(BM, rb + rp);
    if (warp >= 2) {
      float *stage = reinterpret_cast<float *>(smem_gen);
      const int wl = (t - 64) >> 5;
      for (int r = rb + wl; r < re; r += 4) {
        float *mine = stage + r * PITCH + 4 * lane;
        const uint32_t a = smem_u32(mine);
        float4 v[8];
#pragma unroll
        for (int z = 0; z < 8; z++)
          if (z < S) v[z] = ld_dsmem_v4(map_to_cta(a, (uint32_t)z));
        float4 s = v[0];
#pragma unroll
        for (int z = 1; z < 8; z++)
          if (z < S) { s.x += v[z].x; s.y += v[z].y; s.z += v[z].z; s.w += v[z].w; }
        *reinterpret_cast<float4 *>(mine) = s;    // rows [rb, re) of MY buffer are read by nobody else
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    cluster_arrive();                            // this CTA no longer reads its peers' shared memory
    if (warp >= 2)
      prob.template store<8>(reinterpret_cast<float *>(smem_gen), t - 64, mt, (int)blockIdx.y, 0, rb, re);
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
  if (kClusterK) cluster_wait();                 // my staging buffer stays alive until every peer has read it
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
      // (prefetching one tile AHEAD instead was measured: FC2 wgrad + SGD 75.7 -> 79.9 us alone -- the lines
      // compete with the current tile's read-modify-write for L2 and HBM; this placement stays)
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
      prob.template store<16>(stage, tid, mt, nt, z, 0, BM);
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
  EPI_STORE = 0,      // out = acc (+ bias_n) (ReLU / mask / dropout: RowsEpi)
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

// Element-wise work fused into a row-major EPI_STORE epilogue (all optional):
//   forward   out = relu?(acc + bias_n) ; out2 = out .* dropout-scale(seed, row, col)
//             (RectifiedLinearComponent + DropoutComponent::Propagate, upstream nnet2/nnet-component.cc:
//             799-806, 3592-3620, applied where the pre-activation is still in shared memory)
//   backward  out = mask_x > 0 ? acc : 0                      (ReLU backward of the PRODUCER, :813-827)
//             out = mask_x > 0 ? acc * mask_y / mask_x : 0    (dropout backward :3634-3636, then ReLU backward)
//   layout    perm_r > 0: column n = g * perm_r + pos of the logical matrix (the reference's [map][pos]
//             order) is written to column pos * perm_g + g (channels-last [pos][map])
struct RowsEpi {
  const float *bias_n;
  int relu;
  const float *mask_x; int ld_mx;
  const float *mask_y; int ld_my;
  float *out2; int ld_o2;
  float dp, low, high;
  const unsigned long long *seed;
  int perm_r, perm_g;
  FastDiv div_r;
  int vec;               // host: every pointer above is 16-byte aligned with pitches % 4 == 0
};
inline RowsEpi rows_epi_none() {
  RowsEpi e;
  e.bias_n = nullptr; e.relu = 0; e.mask_x = nullptr; e.ld_mx = 0; e.mask_y = nullptr; e.ld_my = 0;
  e.out2 = nullptr; e.ld_o2 = 0; e.dp = 0.f; e.low = 0.f; e.high = 0.f; e.seed = nullptr;
  e.perm_r = 0; e.perm_g = 0; e.div_r = FastDiv(1); e.vec = 1;
  return e;
}
inline bool rows_epi_vec_ok(const RowsEpi &e) {
  if (e.mask_x && (!host_aligned16(e.mask_x) || (e.ld_mx & 3))) return false;
  if (e.mask_y && (!host_aligned16(e.mask_y) || (e.ld_my & 3))) return false;
  if (e.out2 && (!host_aligned16(e.out2) || (e.ld_o2 & 3))) return false;
  return true;
}

// The dropout scale of one unit recovered from its activations, y / x with y = x * scale (x > 0).  The
// reference's backward pass divides exactly like this (out_deriv * out_value / in_value); the IEEE division
// sequence, though, is ~20 dependent instructions per element on a warp that has its scheduler to itself --
// measured (ncu, profiles/r02_gated_dgrad.md): the gated FC2 input-gradient GEMM executed 6.6 M warp
// instructions against 2.8 M ungated and took 73 us against 42.  MUFU.RCP + FMUL is within 2 ulp of it.
__device__ __forceinline__ float gate_ratio(float y, float x) { return __fdividef(y, x); }

// Row-major 128-column segment store shared by the dense, the channels-last convolution and the
// weight-gradient problems: tid -> (warp, 4 columns); each warp instruction writes one 512-byte
// row segment.  Tile rows [rb, re) are stored (the whole tile, or this CTA's slice of a cluster
// split-K reduction), warp w taking rows rb + w, rb + w + 4, ...  row_of(m) maps a tile row to
// the output row; a negative result marks a padding row of the GEMM that has no output row.
template <int kEpi, int kRows, class RowMap>
__device__ __forceinline__ void store_rows(const float *stage, int tid, int m0, int n0, int M, int N,
                                           float *obase, int ld, float *aux, const SgdCoef &sgd, RowMap row_of,
                                           int rb, int re, const RowsEpi &e) {
  const int wl = tid >> 5, lane = tid & 31;
  const int n = n0 + 4 * lane;
  if (re > M - m0) re = M - m0;
  const bool vec_ok = (n + 3 < N) && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(obase) & 15u) == 0);
  if (kEpi == EPI_SGD && vec_ok) {
    // nnet0/nnet-component-nnet0.cc:767-773, 1138-1142 on the tile: W and prev_grad rows are
    // read in batches of kRows rows (2 kRows independent 128-bit loads per thread in flight; the
    // lines were L2-prefetched by Problem::prefetch while the main loop ran), updated, written back.
    constexpr int RB = kRows;
#pragma unroll 1
    for (int r0 = rb + wl; r0 < re; r0 += 4 * RB) {
      float4 w[RB], pv[RB];
      size_t roff[RB];
      bool okr[RB];
#pragma unroll
      for (int i = 0; i < RB; i++) {
        const int mt = r0 + 4 * i;
        const int row = mt < re ? row_of(m0 + mt) : -1;
        okr[i] = row >= 0;
        roff[i] = (size_t)(okr[i] ? row : 0) * ld + n;
        if (okr[i]) {
          w[i] = *reinterpret_cast<const float4 *>(obase + roff[i]);
          pv[i] = *reinterpret_cast<const float4 *>(aux + roff[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < RB; i++) {
        const int mt = r0 + 4 * i;
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
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kEpi == EPI_STORE && e.bias_n) {
    if (n < N) bias4.x = __ldg(e.bias_n + n);
    if (n + 1 < N) bias4.y = __ldg(e.bias_n + n + 1);
    if (n + 2 < N) bias4.z = __ldg(e.bias_n + n + 2);
    if (n + 3 < N) bias4.w = __ldg(e.bias_n + n + 3);
  }
  if (kEpi == EPI_STORE && vec_ok && e.mask_x == nullptr && e.out2 == nullptr && e.perm_r == 0) {
    // the common case (bias / ReLU only) as a tight loop of independent 128-bit moves
#pragma unroll 8
    for (int mt = rb + wl; mt < re; mt += 4) {
      const int row = row_of(m0 + mt);
      if (row < 0) continue;
      float4 a = *reinterpret_cast<const float4 *>(stage + mt * PITCH + 4 * lane);
      a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
      if (e.relu) {
        a.x = a.x > 0.0f ? a.x : 0.0f; a.y = a.y > 0.0f ? a.y : 0.0f;
        a.z = a.z > 0.0f ? a.z : 0.0f; a.w = a.w > 0.0f ? a.w : 0.0f;
      }
      *reinterpret_cast<float4 *>(obase + (size_t)row * ld + n) = a;
    }
    return;
  }
  const bool evec = vec_ok && e.vec != 0;
  if (kEpi == EPI_STORE && evec && e.mask_x != nullptr && e.out2 == nullptr) {
    // Backward gates: the mask rows come from global memory.  The compiler may not move a load
    // above the previous row's store (possible aliasing), which left ONE row of loads in flight per
    // warp -- 32 dependent DRAM / L2 round trips per tile.  So: all loads of kRows rows first (2 kRows
    // independent 128-bit loads per thread), then the arithmetic and the stores.
    constexpr int RB = kRows;
    const bool has_y = e.mask_y != nullptr;
#pragma unroll 1
    for (int r0 = rb + wl; r0 < re; r0 += 4 * RB) {
      float4 x[RB], y[RB];
      int rw[RB];
#pragma unroll
      for (int i = 0; i < RB; i++) {
        const int mt = r0 + 4 * i;
        rw[i] = mt < re ? row_of(m0 + mt) : -1;
        if (rw[i] >= 0) {
          x[i] = __ldg(reinterpret_cast<const float4 *>(e.mask_x + (size_t)rw[i] * e.ld_mx + n));
          if (has_y) y[i] = __ldg(reinterpret_cast<const float4 *>(e.mask_y + (size_t)rw[i] * e.ld_my + n));
        }
      }
#pragma unroll
      for (int i = 0; i < RB; i++) {
        if (rw[i] < 0) continue;
        float4 a = *reinterpret_cast<const float4 *>(stage + (r0 + 4 * i) * PITCH + 4 * lane);
        a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
        if (has_y) {       // d * (y / x): dropout_bprop_kernel's gate, then the ReLU gate (see gate_ratio)
          a.x = x[i].x > 0.0f ? a.x * gate_ratio(y[i].x, x[i].x) : 0.0f;
          a.y = x[i].y > 0.0f ? a.y * gate_ratio(y[i].y, x[i].y) : 0.0f;
          a.z = x[i].z > 0.0f ? a.z * gate_ratio(y[i].z, x[i].z) : 0.0f;
          a.w = x[i].w > 0.0f ? a.w * gate_ratio(y[i].w, x[i].w) : 0.0f;
        } else {
          a.x = x[i].x > 0.0f ? a.x : 0.0f; a.y = x[i].y > 0.0f ? a.y : 0.0f;
          a.z = x[i].z > 0.0f ? a.z : 0.0f; a.w = x[i].w > 0.0f ? a.w : 0.0f;
        }
        if (e.perm_r > 0) {
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            uint32_t g, pos;
            e.div_r.divmod((uint32_t)(n + j), g, pos);
            obase[(size_t)rw[i] * ld + (size_t)pos * e.perm_g + g] = av[j];
          }
        } else {
          *reinterpret_cast<float4 *>(obase + (size_t)rw[i] * ld + n) = a;
        }
      }
    }
    return;
  }
  unsigned long long seed = 0;
  if (kEpi == EPI_STORE && e.out2) seed = *e.seed;
#pragma unroll 4
  for (int mt = rb + wl; mt < re; mt += 4) {
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
      continue;
    }
    a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
    if (e.relu) {       // RectifiedLinearComponent::Propagate fused: x > 0 ? x : 0
      a.x = a.x > 0.0f ? a.x : 0.0f; a.y = a.y > 0.0f ? a.y : 0.0f;
      a.z = a.z > 0.0f ? a.z : 0.0f; a.w = a.w > 0.0f ? a.w : 0.0f;
    }
    if (e.mask_x) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = make_float4(1.f, 1.f, 1.f, 1.f);
      const float *xr = e.mask_x + (size_t)row * e.ld_mx + n;
      const float *yr = e.mask_y ? e.mask_y + (size_t)row * e.ld_my + n : nullptr;
      if (evec) {
        x = __ldg(reinterpret_cast<const float4 *>(xr));
        if (yr) y = __ldg(reinterpret_cast<const float4 *>(yr));
      } else {
        if (n < N) x.x = __ldg(xr);
        if (n + 1 < N) x.y = __ldg(xr + 1);
        if (n + 2 < N) x.z = __ldg(xr + 2);
        if (n + 3 < N) x.w = __ldg(xr + 3);
        if (yr) {
          if (n < N) y.x = __ldg(yr);
          if (n + 1 < N) y.y = __ldg(yr + 1);
          if (n + 2 < N) y.z = __ldg(yr + 2);
          if (n + 3 < N) y.w = __ldg(yr + 3);
        }
      }
      if (yr) {          // d * (y / x): dropout_bprop_kernel's gate, then the ReLU gate (see gate_ratio)
        a.x = x.x > 0.0f ? a.x * gate_ratio(y.x, x.x) : 0.0f; a.y = x.y > 0.0f ? a.y * gate_ratio(y.y, x.y) : 0.0f;
        a.z = x.z > 0.0f ? a.z * gate_ratio(y.z, x.z) : 0.0f; a.w = x.w > 0.0f ? a.w * gate_ratio(y.w, x.w) : 0.0f;
      } else {
        a.x = x.x > 0.0f ? a.x : 0.0f; a.y = x.y > 0.0f ? a.y : 0.0f;
        a.z = x.z > 0.0f ? a.z : 0.0f; a.w = x.w > 0.0f ? a.w : 0.0f;
      }
    }
    if (e.perm_r > 0) {
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (n + j < N) {
          uint32_t g, pos;
          e.div_r.divmod((uint32_t)(n + j), g, pos);
          obase[(size_t)row * ld + (size_t)pos * e.perm_g + g] = av[j];
        }
      continue;
    }
    if (vec_ok) {
      *reinterpret_cast<float4 *>(orow) = a;
    } else {
      if (n < N) orow[0] = a.x;
      if (n + 1 < N) orow[1] = a.y;
      if (n + 2 < N) orow[2] = a.z;
      if (n + 3 < N) orow[3] = a.w;
    }
    if (e.out2) {
      float4 d = a;
      d.x *= dropout_scale_at(seed, (unsigned long long)row, N, n, e.dp, e.low, e.high);
      d.y *= dropout_scale_at(seed, (unsigned long long)row, N, n + 1, e.dp, e.low, e.high);
      d.z *= dropout_scale_at(seed, (unsigned long long)row, N, n + 2, e.dp, e.low, e.high);
      d.w *= dropout_scale_at(seed, (unsigned long long)row, N, n + 3, e.dp, e.low, e.high);
      float *o2 = e.out2 + (size_t)row * e.ld_o2 + n;
      if (evec) {
        *reinterpret_cast<float4 *>(o2) = d;
      } else {
        if (n < N) o2[0] = d.x;
        if (n + 1 < N) o2[1] = d.y;
        if (n + 2 < N) o2[2] = d.z;
        if (n + 3 < N) o2[3] = d.w;
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
  float *aux;              // EPI_SGD: prev_grad, same shape / pitch as out
  SgdCoef sgd;
  RowsEpi epi;             // EPI_STORE: bias, ReLU, masks, dropout, channels-last columns

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
    if (kEpi == EPI_STORE && epi.mask_x != nullptr)
      prefetch_tile_l2(tid, mt * BM, nt * BN, M, N, epi.mask_x,
                       (epi.mask_y && epi.ld_my == epi.ld_mx) ? epi.mask_y : epi.mask_x, epi.ld_mx, IdentityRow());
  }
  template <int kRows>
  __device__ __forceinline__ void store(const float *stage, int tid, int mt, int nt, int z, int rb, int re) const {
    const int m0 = mt * BM, n0 = nt * BN;
    store_rows<kEpi, kRows>(stage, tid, m0, n0, M, N, out, ldo, aux, sgd, IdentityRow(), rb, re, epi);
  }
};

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
bool cluster_k_enabled();                // KCNN_TMA_CLUSTERK=0: no split-K (small grids then leave SMs idle)

// Grow-only device scratch, one buffer per slot (kernels_gemm.cu).  Returns nullptr when it
// would have to grow while the stream is being captured; callers then take another path.
// Only the bare-component path (a Component called outside NnetMinibatchUpdater's fused step)
// uses it: channels-last staging copies and the bias-gradient partial sums of the pack kernel.
enum ScratchSlot { SCRATCH_XCL = 0, SCRATCH_DYCL = 1, SCRATCH_BIAS = 2, SCRATCH_SLOTS = 3 };
float *scratch(int slot, size_t bytes, cudaStream_t st);

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
  float *aux = nullptr;          // EPI_SGD: prev_grad
  SgdCoef sgd = {0.f, 0.f, 0.f};
  RowsEpi rows = rows_epi_none();
};

// K-splits of one output tile = CTAs of one cluster (tma_gemm_kernel<..., kClusterK>): as many as
// fit one wave of the machine, at least 8 K-blocks each, at most the portable cluster size.
constexpr int kMaxClusterK = 8;
inline int pick_splits(long long tiles, int num_kb) {
  if (!cluster_k_enabled() || tiles * 2 > kNumSMs) return 1;
  long long want = kNumSMs / tiles;
  long long max_by_k = num_kb / 8;
  if (want > max_by_k) want = max_by_k;
  if (want > kMaxClusterK) want = kMaxClusterK;
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

// Cluster split-K: grid.z K-splits per tile, one cluster (1, 1, grid.z) each, 1 CTA per SM.
template <class Prob>
void launch_cluster_k(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, const Prob &p, dim3 grid) {
  auto kernel = tma_gemm_kernel<Prob, false, 6, true>;
  constexpr int smem = Ring<6>::SMEM_TOTAL;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = true;
  }
  launch_kernel_cluster(kernel, grid, dim3(THREADS), (size_t)smem, st, dim3(1, 1, grid.z), ma, mb, p);
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
  // KCNN_TMA_PERSIST_CTAS: fewer than one CTA per SM leaves SMs to whatever runs beside this kernel (in the
  // training step: the input-gradient chain on the compute stream while this runs on the side branch)
  static int cap = -1;
  if (cap < 0) {
    const char *e = getenv("KCNN_TMA_PERSIST_CTAS");
    cap = e ? atoi(e) : kNumSMs;
    if (cap < 1 || cap > kNumSMs) cap = kNumSMs;
  }
  unsigned ctas = (unsigned)(total < cap ? total : cap);
  launch_kernel(kernel, dim3(ctas), dim3(THREADS), (size_t)PRing::SMEM_TOTAL, st, 1u, ma, mb, p, (int)tiles.x,
                (int)tiles.y, (int)tiles.z);
}

// grid = (128-row tiles, 128-column tiles, splits).  pair: 2-CTA clusters along x, 256-column
// tiles along y (grid.x rounded up to even, grid.y halved).
template <class Prob>
void launch_prob(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, const Prob &p, dim3 grid,
                 int kb_per_tile, bool pair = false) {
  if (grid.z > 1) {
    launch_cluster_k<Prob>(st, ma, mb, p, grid);
  } else if (pair) {
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
  const bool pair = use_pair(M, N);
  int splits = 1;
  if (allow_split && !pair) splits = pick_splits((long long)ceil_div_u(M, BM) * ceil_div_u(N, BN), num_kb);
  int per = (num_kb + splits - 1) / splits;
  splits = (num_kb + per - 1) / per;
  dim3 grid(ceil_div_u(M, BM), ceil_div_u(N, BN), splits);
  auto fill = [&](auto &p) {
    p.M = M; p.N = N; p.K = K; p.kb_per_split = per;
    p.out = out; p.ldo = ldo; p.aux = epi.aux; p.sgd = epi.sgd;
    p.epi = epi.rows; p.epi.vec = rows_epi_vec_ok(epi.rows) ? 1 : 0;
  };
  if (epi.mode == EPI_SGD) {
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
