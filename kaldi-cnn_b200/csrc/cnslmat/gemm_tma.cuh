// kaldi-cnn_b200/csrc/cnslmat/gemm_tma.cuh
//
// TMA-fed TF32 GEMM on tcgen05 / TMEM for operands that ARE plain pitched row-major
// matrices: the three GEMMs of the affine layer (nnet2/nnet-component.cc:1216-1258,
// nnet0/nnet-component-nnet0.cc:1133-1143) and, through the channels-last staging of
// conv_tma.cuh, the convolutions.  Unlike gemm_tc.cuh no thread ever touches an
// operand element: one lane issues cp.async.bulk.tensor (TMA) boxes that land in
// shared memory already in the SWIZZLE_128B layout the UMMA descriptors describe,
// one lane issues tcgen05.mma, four warps drain the accumulator.
//
// Both operand majors are supported without a transpose pass:
//   K-major   matrix is [MN rows][K cols]  -> one box {32 k, rows} per stage
//   MN-major  matrix is [K rows][MN cols]  -> MN/32 boxes {32 mn, 32 k} per stage,
//             TMA swizzle 128B_ATOM_32B; the UMMA descriptor uses the MN-major
//             SWIZZLE_128B_BASE32B canonical layout (4-k-row atoms, LBO = 4096, SBO = 512)
// so  fprop  Y = X W^T      is (A K-major,  B K-major)
//     dgrad  dX = dY W      is (A K-major,  B MN-major)
//     wgrad  dW = dY^T X    is (A MN-major, B MN-major)
//
// CTA = 128 x BN tile (BN = 128), 6 warps, 2 CTAs per SM (3 x 32 KB stages each) so
// one CTA's epilogue overlaps the other's main loop:
//   warp 0   TMA producer (one lane), full/empty mbarrier ring
//   warp 1   TMEM allocator + MMA issuer (one lane), tcgen05.commit frees the stages
//   warps 2-5 epilogue: tcgen05.ld 32x32b -> padded smem staging (the idle stages) ->
//            512-byte coalesced row segments to HBM, with bias, split-K partials, or
//            the fused momentum / weight-decay SGD update (wgrad of the FC layer).
// K / M / N tails are zero-filled by TMA (out-of-bounds box elements), never read.
//
// Eligibility (host side): 16-byte aligned base and row pitch for every operand
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

constexpr int BM = 128, BK = 32;
constexpr int THREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 4;
constexpr int ATOM_BYTES = BK * 128;          // one MN-major box: 32 k-rows x 128 bytes

template <int BN, int STAGES>
struct Smem {
  static constexpr int B_STAGE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGING_PITCH = BN + 4;                    // floats; conflict-free v4 stores
  static constexpr int STAGING_BYTES = 128 * STAGING_PITCH * 4;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int DATA_BYTES = RING_BYTES > STAGING_BYTES ? RING_BYTES : STAGING_BYTES;
  static constexpr int BAR_OFFSET = DATA_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;
};

// What the epilogue does with the accumulator tile.
enum EpiMode {
  EPI_STORE = 0,      // out = acc (+ bias_n)
  EPI_PARTIAL = 1,    // workspace[z] = acc           (split-K)
  EPI_SGD = 2,        // prev = m prev - lr wd W + lr acc ; W += prev   (out = W, aux = prev)
};

struct Params {
  int M, N, K;
  int k_chunk;             // K range per blockIdx.z (multiple of BK)
  float *out;              // row-major [M][ldo]
  int ldo;
  const float *bias_n;     // per column, or nullptr
  float *workspace;        // [splits][M][N] partials
  float *aux;              // EPI_SGD: prev_grad, same shape / pitch as out
  float *grad_out;         // EPI_SGD: optional copy of the raw gradient (nullptr = none)
  int ldg;
  float lr, lr_wd, momentum;
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
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
// MN-major, 32-bit elements: the only layout the tensor core transposes is
// SWIZZLE_128B_BASE32B (layout type 1; 32-byte chunks XOR-ed with the row index mod 4,
// TMA's SWIZZLE_128B_ATOM_32B): atoms of 4 k-rows x 128 bytes (32 mn); MN atoms LBO
// apart, 4-k groups SBO = 512 bytes apart (one MMA = 8 k = two groups).
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

template <int BN, int STAGES, bool kAMn, bool kBMn, int kEpi>
__global__ void __launch_bounds__(THREADS, 2)
gemm_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const Params p) {
  using S = Smem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + S::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 8 * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8 * (2 * STAGES);
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem_gen + S::BAR_OFFSET + 8 * (2 * STAGES + 1));

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_begin = blockIdx.z * p.k_chunk;
  const int k_end = min(p.K, k_begin + p.k_chunk);
  const int num_kb = (k_end - k_begin + BK - 1) / BK;

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)), "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer --
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(empty_bar(s), ((kb / STAGES) - 1) & 1);
        const uint32_t a_addr = smem_base + s * S::STAGE_BYTES;
        const uint32_t b_addr = a_addr + A_STAGE_BYTES;
        const int k0 = k_begin + kb * BK;
        mbar_expect_tx(full_bar(s), S::STAGE_BYTES);
        if (kAMn) {
#pragma unroll
          for (int i = 0; i < BM / 32; i++) tma_load_2d(a_addr + i * ATOM_BYTES, &map_a, m0 + 32 * i, k0, full_bar(s));
        } else {
          tma_load_2d(a_addr, &map_a, k0, m0, full_bar(s));
        }
        if (kBMn) {
#pragma unroll
          for (int i = 0; i < BN / 32; i++) tma_load_2d(b_addr + i * ATOM_BYTES, &map_b, n0 + 32 * i, k0, full_bar(s));
        } else {
          tma_load_2d(b_addr, &map_b, k0, n0, full_bar(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer --
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, kAMn, kBMn);
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        mbar_wait(full_bar(s), (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * S::STAGE_BYTES;
        const uint32_t b_addr = a_addr + A_STAGE_BYTES;
        const uint64_t adesc = kAMn ? desc_mn_major(a_addr) : desc_k_major(a_addr);
        const uint64_t bdesc = kBMn ? desc_mn_major(b_addr) : desc_k_major(b_addr);
#pragma unroll
        for (int k = 0; k < BK / 8; k++) {
          // one MMA = 8 TF32 of K: +32 bytes along a K-major row, +1024 bytes (one 8-row
          // group) in an MN-major atom; start-address field is in 16-byte units
          const uint64_t ad = adesc + (kAMn ? 64 * k : 2 * k);
          const uint64_t bd = bdesc + (kBMn ? 64 * k : 2 * k);
          umma_tf32(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  } else {
    // ----------------------------------------------------------------- epilogue --
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    float *stage = reinterpret_cast<float *>(smem_gen) + q * 32 * S::STAGING_PITCH;
    if (num_kb > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    // TMEM -> registers -> padded staging: lane = row, 32 columns per tcgen05.ld
#pragma unroll 1
    for (int j0 = 0; j0 < BN; j0 += 32) {
      uint32_t v[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)j0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 0u;
      }
      float *dst = stage + lane * S::STAGING_PITCH + j0;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4 *>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    __syncwarp();
    // staging -> HBM: one 512-byte row segment per warp instruction (lane = 4 columns)
    static_assert(BN == 128, "epilogue readback assumes 128 columns = 32 lanes x float4");
    const int n = n0 + 4 * lane;
    const bool n_full = n + 3 < p.N;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kEpi == EPI_STORE && p.bias_n) {
      if (n < p.N) bias4.x = __ldg(p.bias_n + n);
      if (n + 1 < p.N) bias4.y = __ldg(p.bias_n + n + 1);
      if (n + 2 < p.N) bias4.z = __ldg(p.bias_n + n + 2);
      if (n + 3 < p.N) bias4.w = __ldg(p.bias_n + n + 3);
    }
    float *obase;
    int ld;
    if (kEpi == EPI_PARTIAL) {
      obase = p.workspace + (size_t)blockIdx.z * p.M * p.N;
      ld = p.N;
    } else {
      obase = p.out;
      ld = p.ldo;
    }
    const bool vec_ok = n_full && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(obase) & 15u) == 0);
#pragma unroll 4
    for (int r = 0; r < 32; r++) {
      const int m = m0 + q * 32 + r;
      if (m >= p.M) break;
      float4 a = *reinterpret_cast<const float4 *>(stage + r * S::STAGING_PITCH + 4 * lane);
      float *orow = obase + (size_t)m * ld + n;
      if (kEpi == EPI_SGD) {
        // nnet0/nnet-component-nnet0.cc:1138-1142: prev = m prev - lr wd W + lr grad ; W += prev
        float *prow = p.aux + (size_t)m * ld + n;
        if (vec_ok) {
          float4 w = *reinterpret_cast<const float4 *>(orow);
          float4 pv = *reinterpret_cast<const float4 *>(prow);
          if (p.grad_out && ((p.ldg & 3) == 0))
            *reinterpret_cast<float4 *>(p.grad_out + (size_t)m * p.ldg + n) = a;
          pv.x = p.momentum * pv.x - p.lr_wd * w.x + p.lr * a.x;
          pv.y = p.momentum * pv.y - p.lr_wd * w.y + p.lr * a.y;
          pv.z = p.momentum * pv.z - p.lr_wd * w.z + p.lr * a.z;
          pv.w = p.momentum * pv.w - p.lr_wd * w.w + p.lr * a.w;
          w.x += pv.x; w.y += pv.y; w.z += pv.z; w.w += pv.w;
          *reinterpret_cast<float4 *>(prow) = pv;
          *reinterpret_cast<float4 *>(orow) = w;
        } else {
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (n + j < p.N) {
              float w = orow[j];
              float pv = p.momentum * prow[j] - p.lr_wd * w + p.lr * av[j];
              prow[j] = pv;
              orow[j] = w + pv;
              if (p.grad_out) p.grad_out[(size_t)m * p.ldg + n + j] = av[j];
            }
        }
      } else {
        a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
        if (vec_ok) {
          *reinterpret_cast<float4 *>(orow) = a;
        } else {
          if (n < p.N) orow[0] = a.x;
          if (n + 1 < p.N) orow[1] = a.y;
          if (n + 2 < p.N) orow[2] = a.z;
          if (n + 3 < p.N) orow[3] = a.w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN)
                 : "memory");
  }
}

// out[m][n] = sum_z ws[z][m][n] (+ bias_n[n])
__global__ void __launch_bounds__(256)
splitk_reduce_rows_kernel(const float *__restrict__ ws, int splits, int M, int N, float *__restrict__ out,
                          int ldo, const float *__restrict__ bias_n) {
  const long long total4 = ((long long)M * N) >> 2;         // N % 4 == 0 is checked by the launcher
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const long long e = i << 2;
  const int m = (int)(e / N), n = (int)(e - (long long)m * N);
  float4 s = __ldg(reinterpret_cast<const float4 *>(ws + e));
  for (int z = 1; z < splits; z++) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(ws + (size_t)z * M * N + e));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  if (bias_n) {
    s.x += __ldg(bias_n + n); s.y += __ldg(bias_n + n + 1); s.z += __ldg(bias_n + n + 2); s.w += __ldg(bias_n + n + 3);
  }
  *reinterpret_cast<float4 *>(out + (size_t)m * ldo + n) = s;
}

// ------------------------------------------------------------------- host side --

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn();         // kernels_gemm.cu (driver entry point, resolved once)
int tma_data_type();                     // CU_TENSOR_MAP_DATA_TYPE_* used for the operands

// A pitched row-major matrix as a GEMM operand.
struct Matrix {
  const float *base;
  int rows, cols, ld;
};

inline bool matrix_tma_ok(const Matrix &m) {
  return host_aligned16(m.base) && (m.ld & 3) == 0 && m.rows > 0 && m.cols > 0;
}

// 2-D tensor map over `m` with a {32 cols, box_rows} box, 128-byte swizzle.
inline bool encode_2d(CUtensorMap *map, const Matrix &m, int box_rows, bool mn_major) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)m.cols, (cuuint64_t)m.rows};
  cuuint64_t gstr[1] = {(cuuint64_t)m.ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, (CUtensorMapDataType)tma_data_type(), 2, const_cast<float *>(m.base), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct Epilogue {
  int mode = EPI_STORE;
  const float *bias_n = nullptr;
  float *aux = nullptr;          // EPI_SGD: prev_grad
  float *grad_out = nullptr;
  int ldg = 0;
  float lr = 0.f, lr_wd = 0.f, momentum = 0.f;
};

float *splitk_workspace(size_t bytes);   // kernels_gemm.cu: grow-only device scratch

inline int pick_splits(int M, int N, int K) {
  long long tiles = (long long)((M + BM - 1) / BM) * ((N + 127) / 128);
  if (tiles >= 100) return 1;
  long long want = (2 * kNumSMs) / tiles;
  long long max_by_k = K / (BK * 8);
  if (want > max_by_k) want = max_by_k;
  if (want > 16) want = 16;
  return (int)(want < 1 ? 1 : want);
}

template <bool kAMn, bool kBMn, int kEpi>
void launch_cfg(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, const Params &p, int splits) {
  constexpr int BN = 128, STAGES = 3;
  using S = Smem<BN, STAGES>;
  auto kernel = gemm_tma_kernel<BN, STAGES, kAMn, kBMn, kEpi>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    attr_set = true;
  }
  dim3 grid(ceil_div_u(p.M, BM), ceil_div_u(p.N, BN), splits);
  kernel<<<grid, THREADS, S::TOTAL, st>>>(ma, mb, p);
  count_launch();
}

// D[M x N] (row-major, pitch ldo) = A B^T with A logical [M x K], B logical [N x K].
// a / b are the matrices as stored: K-major -> [MN][K], MN-major -> [K][MN].
// Returns false (nothing launched) when an operand is not TMA-addressable.
template <bool kAMn, bool kBMn>
bool gemm(cudaStream_t st, const Matrix &a, const Matrix &b, int M, int N, int K, float *out, int ldo,
          const Epilogue &epi, bool allow_split) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  if (!matrix_tma_ok(a) || !matrix_tma_ok(b)) return false;
  CUtensorMap ma, mb;
  if (!encode_2d(&ma, a, kAMn ? BK : BM, kAMn)) return false;
  if (!encode_2d(&mb, b, kBMn ? BK : 128, kBMn)) return false;
  Params p;
  p.M = M; p.N = N; p.K = K;
  p.out = out; p.ldo = ldo; p.bias_n = epi.bias_n; p.workspace = nullptr;
  p.aux = epi.aux; p.grad_out = epi.grad_out; p.ldg = epi.ldg;
  p.lr = epi.lr; p.lr_wd = epi.lr_wd; p.momentum = epi.momentum;
  int splits = 1;
  if (allow_split && epi.mode == EPI_STORE && (N & 3) == 0 && (ldo & 3) == 0 && host_aligned16(out))
    splits = pick_splits(M, N, K);
  int chunk = (K + splits - 1) / splits;
  chunk = ((chunk + BK - 1) / BK) * BK;
  splits = (K + chunk - 1) / chunk;
  p.k_chunk = chunk;
  if (splits > 1) {
    p.workspace = splitk_workspace((size_t)splits * M * N * sizeof(float));
    if (!p.workspace) { splits = 1; p.k_chunk = ((K + BK - 1) / BK) * BK; }
  }
  if (splits > 1) {
    launch_cfg<kAMn, kBMn, EPI_PARTIAL>(st, ma, mb, p, splits);
    KCNN_LAUNCH(splitk_reduce_rows_kernel, ceil_div_u(((long long)M * N) >> 2, 256), 256, 0, st, p.workspace,
                splits, M, N, out, ldo, epi.bias_n);
  } else if (epi.mode == EPI_SGD) {
    launch_cfg<kAMn, kBMn, EPI_SGD>(st, ma, mb, p, 1);
  } else {
    launch_cfg<kAMn, kBMn, EPI_STORE>(st, ma, mb, p, 1);
  }
  return true;
}

}  // namespace tma
}  // namespace kcnn

#endif
