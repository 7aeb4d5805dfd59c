// kaldi-cnn_b200/csrc/cnslmat/gemm_tc.cuh
//
// TF32 implicit GEMM on the 5th-generation tensor cores (KCNN_MATH_TF32_TC):
// tcgen05.mma kind::tf32 issued by one thread, FP32 accumulators in TMEM, read back
// with tcgen05.ld for the epilogue.  One kernel template serves convolution forward,
// input-gradient, weight-gradient and the three affine GEMMs: the operands are
// addressed through the same separable decoders as the FP32 kernel
// (gemm_operands.cuh), so zero padding, the kernel flip, the im2col window and every
// block / row permutation of the reference are index arithmetic in the PRODUCER,
// never a buffer in HBM.
//
// CTA = 128 x BN output tile (BN = 64 | 128), 9 warps:
//   warps 0-7  producers: gather a 128 x 32 slice of A and a BN x 32 slice of B per
//              stage straight from the layer's tensors, round to TF32 (cvt.rna) and
//              store them in the canonical K-major SWIZZLE_128B layout the UMMA
//              descriptors describe; then fence.proxy.async + mbarrier arrive.
//              Loads run along whichever axis is contiguous in memory (128-bit when
//              the four K-neighbours are contiguous and aligned).
//   warp  8    one thread issues 4 x tcgen05.mma (K = 8 each) per stage and
//              tcgen05.commit's the stage back to the producers.
//   warps 0-3  epilogue after the main loop: tcgen05.ld 32x32b, add bias, scatter
//              through the output map (coalesced across lanes for the [C][W][H]
//              activations, 128-bit rows for row-major matrices), or write split-K
//              partials.
// 4-stage ring of full / empty mbarriers; gridDim.z splits K.
//
// A TMA producer is not used here on purpose: none of the conv operands satisfies
// the tensor-map constraints in the reference layout (non-inner strides of W*4 or
// H*W*4 bytes are not multiples of 16 for W = 18, 14, 6; sub-matrix views are not
// 16-byte aligned) -- see DESIGN.md "why a software producer".

#ifndef KCNN_GEMM_TC_CUH_
#define KCNN_GEMM_TC_CUH_

#include "gemm_operands.cuh"

namespace kcnn {
namespace tc {

constexpr int BM = 128, BK = 32, STAGES = 4;
constexpr int PREFETCH = 2;          // K-blocks of loads in flight ahead of the one being stored
constexpr int PRODUCER_WARPS = 8, PRODUCER_THREADS = PRODUCER_WARPS * 32;
constexpr int THREADS = PRODUCER_THREADS + 32;
constexpr int A_STAGE_BYTES = BM * BK * 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row
// groups 1024 bytes apart (SBO), version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address      bits [0,14)
  d |= (uint64_t)1 << 16;                               // LBO (unused here)  bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                     // SBO = 1024 B       bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version bits [46,48)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B       bits [61,64)
  return d;
}

// Instruction descriptor for kind::tf32: D = F32, A = B = TF32, both K-major.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <class OpA, class OpB, class Out>
struct TcGemm {
  OpA a;
  OpB b;
  Out out;
  int M, N, K;
  int k_chunk;        // K range per blockIdx.z, a multiple of BK
  float *workspace;   // split-K partials [splits][M][N]
};

template <int BN>
struct Smem {
  static constexpr int B_STAGE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  // full[STAGES], empty[STAGES], accum, tmem pointer
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

// One thread's share of a ROWS x 32 operand slice: ROWS / 32 rows x one 16-byte K-chunk.
// kFastK:  thread owns K-chunk (t & 7) of rows (t >> 3) + 32 i
// !kFastK: warp w owns K-chunk w of rows lane + 32 i (consecutive lanes -> consecutive rows)
// load_slice only ISSUES the global loads (into registers); store_slice rounds to TF32 and
// writes the swizzled stage.  Keeping them apart lets the producer keep several K-blocks
// of loads in flight (see the prefetch ring in the kernel).
template <int ROWS, bool kFastK, class Op>
__device__ __forceinline__ void load_slice(const Op &op, const Ctx (&rows)[ROWS / 32], int k0, int k_end,
                                           int t, float4 (&v)[ROWS / 32]) {
  const int chunk = kFastK ? (t & 7) : (t >> 5);
  Ctx ck[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    int k = k0 + 4 * chunk + j;
    ck[j] = op.k(k < k_end ? k : 0x7fffffff);
  }
  const bool contig = kFastK && ck[1].off == ck[0].off + 1 && ck[2].off == ck[0].off + 2 &&
                      ck[3].off == ck[0].off + 3 && ck[3].dw != kInvalidCoord;
#pragma unroll
  for (int i = 0; i < ROWS / 32; i++) {
    const Ctx &cr = rows[i];
    bool vec = false;
    const float *p = op.base + cr.off + ck[0].off;
    if (contig) {
      // all four taps inside the window?  (dw / dh of the K-neighbours may differ)
      bool ok = true;
#pragma unroll
      for (int j = 0; j < 4; j++)
        ok = ok && (unsigned)(cr.dw + ck[j].dw) < (unsigned)op.wlim &&
             (unsigned)(cr.dh + ck[j].dh) < (unsigned)op.hlim;
      vec = ok && (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
    }
    if (vec) {
      v[i] = __ldg(reinterpret_cast<const float4 *>(p));
    } else {
      v[i].x = op.load(cr, ck[0]);
      v[i].y = op.load(cr, ck[1]);
      v[i].z = op.load(cr, ck[2]);
      v[i].w = op.load(cr, ck[3]);
    }
  }
}

template <int ROWS, bool kFastK>
__device__ __forceinline__ void store_slice(const float4 (&v)[ROWS / 32], uint32_t stage_addr, int t) {
  const int chunk = kFastK ? (t & 7) : (t >> 5);
  const int rbase = kFastK ? (t >> 3) : (t & 31);
#pragma unroll
  for (int i = 0; i < ROWS / 32; i++) {
    const int r = rbase + 32 * i;
    uint32_t dst = stage_addr + r * 128 + ((chunk ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(to_tf32(v[i].x)),
                 "r"(to_tf32(v[i].y)), "r"(to_tf32(v[i].z)), "r"(to_tf32(v[i].w))
                 : "memory");
  }
}

template <int BN, bool kAFastK, bool kBFastK, bool kEpiFastN, class OpA, class OpB, class Out>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const TcGemm<OpA, OpB, Out> g) {
  kcnn::pdl_prologue();
  using S = Smem<BN>;
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
  const int k_begin = blockIdx.z * g.k_chunk;
  const int k_end = min(g.K, k_begin + g.k_chunk);
  const int num_kb = (k_end - k_begin + BK - 1) / BK;

  if (warp == PRODUCER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)), "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar(s), PRODUCER_WARPS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < PRODUCER_WARPS) {
    // ---------------------------------------------------------------- producers --
    Ctx a_rows[BM / 32], b_rows[BN / 32];
#pragma unroll
    for (int i = 0; i < BM / 32; i++)
      a_rows[i] = g.a.mn(m0 + (kAFastK ? (t >> 3) : (t & 31)) + 32 * i);
#pragma unroll
    for (int i = 0; i < BN / 32; i++)
      b_rows[i] = g.b.mn(n0 + (kBFastK ? (t >> 3) : (t & 31)) + 32 * i);
    // Prefetch ring: the loads of K-block kb + PREFETCH are issued before the stores of
    // K-block kb, so PREFETCH + 1 K-blocks (32 KB each per CTA) are in flight per SM.
    float4 fa[PREFETCH + 1][BM / 32], fb[PREFETCH + 1][BN / 32];
#pragma unroll
    for (int u = 0; u < PREFETCH; u++)
      if (u < num_kb) {
        load_slice<BM, kAFastK>(g.a, a_rows, k_begin + u * BK, k_end, t, fa[u]);
        load_slice<BN, kBFastK>(g.b, b_rows, k_begin + u * BK, k_end, t, fb[u]);
      }
    for (int kb0 = 0; kb0 < num_kb; kb0 += PREFETCH + 1) {
#pragma unroll
      for (int u = 0; u < PREFETCH + 1; u++) {
        const int kb = kb0 + u;
        if (kb < num_kb) {
          constexpr int R = PREFETCH + 1;
          const int nxt = kb + PREFETCH;
          if (nxt < num_kb) {
            load_slice<BM, kAFastK>(g.a, a_rows, k_begin + nxt * BK, k_end, t, fa[(u + PREFETCH) % R]);
            load_slice<BN, kBFastK>(g.b, b_rows, k_begin + nxt * BK, k_end, t, fb[(u + PREFETCH) % R]);
          }
          const int s = kb % STAGES;
          if (kb >= STAGES) mbar_wait(empty_bar(s), ((kb / STAGES) - 1) & 1);
          const uint32_t a_addr = smem_base + s * S::STAGE_BYTES;
          store_slice<BM, kAFastK>(fa[u], a_addr, t);
          store_slice<BN, kBFastK>(fb[u], a_addr + A_STAGE_BYTES, t);
          fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar(s));
        }
      }
    }
  } else {
    // --------------------------------------------------------------- MMA issuer --
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        mbar_wait(full_bar(s), (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * S::STAGE_BYTES;
        const uint64_t adesc = make_smem_desc(a_addr);
        const uint64_t bdesc = make_smem_desc(a_addr + A_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 8; k++)      // 8 TF32 = 32 bytes per MMA: +2 in 16-byte units
          umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(s));            // frees the stage when these MMAs retire
      }
      umma_commit(accum_bar);                 // accumulator complete
    }
    __syncwarp();
  }

  if (warp < 4) {
    // ----------------------------------------------------------------- epilogue --
    // column offsets / bias of this tile, staged in the (now idle) first stage
    int *n_off = reinterpret_cast<int *>(smem_gen);
    float *n_bias = reinterpret_cast<float *>(smem_gen + BN * 4);
    if (num_kb > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const bool split = gridDim.z > 1;
    for (int j = t; j < BN; j += 128) {
      int n = n0 + j;
      n_off[j] = n < g.N ? g.out.n(n).off : 0x7fffffff;
      n_bias[j] = (!split && g.out.bias_n && n < g.N) ? __ldg(g.out.bias_n + n) : 0.0f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int m = m0 + warp * 32 + lane;
    const bool m_ok = m < g.M;
    const int m_off = (m_ok && !split) ? g.out.m(m).off : 0;
    const float bm = (m_ok && !split && g.out.bias_m) ? __ldg(g.out.bias_m + m) : 0.0f;
    float *ws_row = split ? g.workspace + ((size_t)blockIdx.z * g.M + (m_ok ? m : 0)) * g.N : nullptr;
#pragma unroll 1
    for (int j0 = 0; j0 < BN; j0 += 32) {
      uint32_t v[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)j0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 0u;
      }
      if (!m_ok || n0 + j0 >= g.N) continue;
      if (split) {
        float *p = ws_row + n0 + j0;
        if (n0 + j0 + 32 <= g.N && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4 *>(p + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                             __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++)
            if (n0 + j0 + j < g.N) p[j] = __uint_as_float(v[j]);
        }
      } else if (kEpiFastN) {
        // row-major output: this thread owns 32 consecutive floats of row m
        float *p = g.out.base + m_off + n_off[j0];
        const bool vec = n0 + j0 + 32 <= g.N && n_off[j0 + 31] == n_off[j0] + 31 &&
                         (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
        if (vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4 *>(p + j) =
                make_float4(__uint_as_float(v[j]) + bm + n_bias[j0 + j],
                            __uint_as_float(v[j + 1]) + bm + n_bias[j0 + j + 1],
                            __uint_as_float(v[j + 2]) + bm + n_bias[j0 + j + 2],
                            __uint_as_float(v[j + 3]) + bm + n_bias[j0 + j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++)
            if (n0 + j0 + j < g.N)
              g.out.base[m_off + n_off[j0 + j]] = __uint_as_float(v[j]) + bm + n_bias[j0 + j];
        }
      } else {
        // [C][W][H] activations: consecutive lanes (rows m) are consecutive addresses
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (n0 + j0 + j < g.N)
            g.out.base[m_off + n_off[j0 + j]] = __uint_as_float(v[j]) + bm + n_bias[j0 + j];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PRODUCER_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN)
                 : "memory");
  }
}

template <class Out>
__global__ void __launch_bounds__(256)
tc_splitk_reduce_kernel(const float *__restrict__ ws, int splits, int M, int N, Out out, FastDiv div_n) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)M * N) return;
  uint32_t m, n;
  div_n.divmod((uint32_t)t, m, n);
  float s = 0.0f;
  for (int z = 0; z < splits; z++) s += __ldg(ws + (size_t)z * M * N + t);
  if (out.bias_m) s += __ldg(out.bias_m + m);
  if (out.bias_n) s += __ldg(out.bias_n + n);
  out.base[out.m((int)m).off + out.n((int)n).off] = s;
}

inline int pick_splits(int M, int N, int K, int bn) {
  long long tiles = (long long)((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  if (tiles >= 96) return 1;
  long long want = (kNumSMs + tiles - 1) / tiles;
  long long max_by_k = K / (BK * 8);           // at least 8 K-blocks per split
  if (max_by_k < 1) max_by_k = 1;
  if (want > max_by_k) want = max_by_k;
  if (want > 32) want = 32;
  return (int)(want < 1 ? 1 : want);
}

inline size_t workspace_bytes(int M, int N, int K) {
  int bn = N <= 64 ? 64 : 128;
  int s = pick_splits(M, N, K, bn);
  return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

template <int BN, bool kAFastK, bool kBFastK, bool kEpiFastN, class OpA, class OpB, class Out>
void launch_bn(cudaStream_t st, const OpA &a, const OpB &b, const Out &out, int M, int N, int K,
               int splits, float *workspace) {
  using S = Smem<BN>;
  TcGemm<OpA, OpB, Out> g;
  g.a = a; g.b = b; g.out = out;
  g.M = M; g.N = N; g.K = K;
  if (splits < 1 || workspace == nullptr) splits = 1;
  int chunk = (K + splits - 1) / splits;
  chunk = ((chunk + BK - 1) / BK) * BK;
  if (chunk < BK) chunk = BK;
  splits = K > 0 ? (K + chunk - 1) / chunk : 1;
  g.k_chunk = chunk;
  g.workspace = workspace;
  auto kernel = gemm_tc_kernel<BN, kAFastK, kBFastK, kEpiFastN, OpA, OpB, Out>;
  static bool attr_set = false;          // one flag per template instantiation
  if (!attr_set) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    attr_set = true;
  }
  dim3 grid(ceil_div_u(M, BM), ceil_div_u(N, BN), splits);
  launch_kernel(kernel, grid, dim3(THREADS), (size_t)S::TOTAL, st, 1u, g);
  if (splits > 1)
    KCNN_LAUNCH(tc_splitk_reduce_kernel<Out>, ceil_div_u((long long)M * N, 256), 256, 0, st, workspace,
                splits, M, N, out, FastDiv((uint32_t)N));
}

// Host launcher; workspace must hold workspace_bytes(M, N, K) when that is non-zero.
template <bool kAFastK, bool kBFastK, bool kEpiFastN, class OpA, class OpB, class Out>
void launch_gemm_tc(cudaStream_t st, const OpA &a, const OpB &b, const Out &out, int M, int N, int K,
                    float *workspace) {
  if (M <= 0 || N <= 0) return;
  if (N <= 64) {
    launch_bn<64, kAFastK, kBFastK, kEpiFastN>(st, a, b, out, M, N, K, pick_splits(M, N, K, 64), workspace);
  } else {
    launch_bn<128, kAFastK, kBFastK, kEpiFastN>(st, a, b, out, M, N, K, pick_splits(M, N, K, 128), workspace);
  }
}

}  // namespace tc
}  // namespace kcnn

#endif
