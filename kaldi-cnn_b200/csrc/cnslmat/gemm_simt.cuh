// kaldi-cnn_b200/csrc/cnslmat/gemm_simt.cuh
//
// FP32 implicit-GEMM on the CUDA cores (KCNN_MATH_FP32_SIMT): the 1e-5-class
// path.  C[M x N] = A[M x K] * B[K x N] where A, B and C are addressed through
// the separable decoders of gemm_operands.cuh, so convolution forward,
// input-gradient and weight-gradient and the three affine GEMMs are all
// instances of ONE kernel template; no im2col / pad / flip / transpose buffer
// exists.  Roofline: FP32 FMA peak (148 SMs x 128 lanes x 2 flop x clock).
//
// Tile 128 x 128 x 16, 256 threads, 8 x 8 accumulators per thread, operands
// staged through double-buffered shared memory with register prefetch.
// Each operand tile is fetched along whichever axis is contiguous in memory
// (kFastK: K runs fastest across the threads of a warp; otherwise M/N does).
// gridDim.z splits K (weight gradient: K = N*OH*OW is long and M x N is small);
// partial tiles go to a workspace reduced by splitk_reduce_kernel.

#ifndef KCNN_GEMM_SIMT_CUH_
#define KCNN_GEMM_SIMT_CUH_

#include "gemm_operands.cuh"

namespace kcnn {

constexpr int SBM = 128, SBN = 128, SBK = 16, STHREADS = 256;
constexpr int SPITCH = SBM + 4;

template <class OpA, class OpB, class Out, bool kAFastK, bool kBFastK, bool kEpiFastN>
struct SimtGemm {
  OpA a;
  OpB b;
  Out out;
  int M, N, K;
  int k_chunk;        // K range handled by one blockIdx.z (multiple of SBK)
  float *workspace;   // split-K partials [splits][M][N] when gridDim.z > 1
};

template <class G, class OpA, class OpB, class Out, bool kAFastK, bool kBFastK, bool kEpiFastN>
__global__ void __launch_bounds__(STHREADS)
gemm_simt_kernel(const SimtGemm<OpA, OpB, Out, kAFastK, kBFastK, kEpiFastN> g) {
  kcnn::pdl_prologue();
  __shared__ float As[2][SBK][SPITCH];
  __shared__ float Bs[2][SBK][SPITCH];

  const int t = threadIdx.x;
  const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
  const int k_begin = blockIdx.z * g.k_chunk;
  const int k_end = min(g.K, k_begin + g.k_chunk);

  // ---- loader assignment: 8 elements of A and 8 of B per thread per tile
  // fast-K : k = t % 16, mn = t / 16 + 16 i      fast-MN : mn = t % 128, k = t / 128 + 2 i
  Ctx a_fix[kAFastK ? 8 : 1], b_fix[kBFastK ? 8 : 1];
  if (kAFastK) {
#pragma unroll
    for (int i = 0; i < 8; i++) a_fix[i] = g.a.mn(m0 + (t >> 4) + 16 * i);
  } else {
    a_fix[0] = g.a.mn(m0 + (t & 127));
  }
  if (kBFastK) {
#pragma unroll
    for (int i = 0; i < 8; i++) b_fix[i] = g.b.mn(n0 + (t >> 4) + 16 * i);
  } else {
    b_fix[0] = g.b.mn(n0 + (t & 127));
  }

  float a_reg[8], b_reg[8];
  auto fetch = [&](int k0) {
    if (kAFastK) {
      int k = k0 + (t & 15);
      Ctx ck = g.a.k(k < k_end ? k : 0x7fffffff);
#pragma unroll
      for (int i = 0; i < 8; i++) a_reg[i] = g.a.load(a_fix[i], ck);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        int k = k0 + (t >> 7) + 2 * i;
        Ctx ck = g.a.k(k < k_end ? k : 0x7fffffff);
        a_reg[i] = g.a.load(a_fix[0], ck);
      }
    }
    if (kBFastK) {
      int k = k0 + (t & 15);
      Ctx ck = g.b.k(k < k_end ? k : 0x7fffffff);
#pragma unroll
      for (int i = 0; i < 8; i++) b_reg[i] = g.b.load(b_fix[i], ck);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        int k = k0 + (t >> 7) + 2 * i;
        Ctx ck = g.b.k(k < k_end ? k : 0x7fffffff);
        b_reg[i] = g.b.load(b_fix[0], ck);
      }
    }
  };
  auto stash = [&](int buf) {
    if (kAFastK) {
#pragma unroll
      for (int i = 0; i < 8; i++) As[buf][t & 15][(t >> 4) + 16 * i] = a_reg[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) As[buf][(t >> 7) + 2 * i][t & 127] = a_reg[i];
    }
    if (kBFastK) {
#pragma unroll
      for (int i = 0; i < 8; i++) Bs[buf][t & 15][(t >> 4) + 16 * i] = b_reg[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) Bs[buf][(t >> 7) + 2 * i][t & 127] = b_reg[i];
    }
  };

  // ---- compute assignment: rows {4 tm .. +3, 64 + 4 tm .. +3}, cols likewise with tn
  const int tm = kEpiFastN ? (t >> 4) : (t & 15);
  const int tn = kEpiFastN ? (t & 15) : (t >> 4);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;

  int buf = 0;
  if (k_begin < k_end) {
    fetch(k_begin);
    stash(0);
  }
  __syncthreads();
  for (int k0 = k_begin; k0 < k_end; k0 += SBK) {
    const bool more = k0 + SBK < k_end;
    if (more) fetch(k0 + SBK);
#pragma unroll
    for (int kk = 0; kk < SBK; kk++) {
      float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][4 * tm]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][64 + 4 * tm]);
      float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][4 * tn]);
      float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][64 + 4 * tn]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // ---- epilogue
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int m = m0 + (i < 4 ? 4 * tm + i : 64 + 4 * tm + (i - 4));
    if (m >= g.M) continue;
    if (split) {
      float *ws = g.workspace + ((size_t)blockIdx.z * g.M + m) * g.N;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        int n = n0 + (j < 4 ? 4 * tn + j : 64 + 4 * tn + (j - 4));
        if (n < g.N) ws[n] = acc[i][j];
      }
    } else {
      Ctx cm = g.out.m(m);
      float bm = g.out.bias_m ? __ldg(g.out.bias_m + m) : 0.0f;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        int n = n0 + (j < 4 ? 4 * tn + j : 64 + 4 * tn + (j - 4));
        if (n < g.N) {
          Ctx cn = g.out.n(n);
          float v = acc[i][j] + bm;
          if (g.out.bias_n) v += __ldg(g.out.bias_n + n);
          g.out.base[cm.off + cn.off] = v;
        }
      }
    }
  }
}

// Sums split-K partials in a fixed order (deterministic) and scatters through the
// output map.  One thread per output element, n fastest.
template <class Out>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float *__restrict__ ws, int splits, int M, int N, Out out,
                     FastDiv div_n) {
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

// Host launcher.  splits > 1 needs workspace of splits*M*N floats.
template <bool kAFastK, bool kBFastK, bool kEpiFastN, class OpA, class OpB, class Out>
void launch_gemm_simt(cudaStream_t st, const OpA &a, const OpB &b, const Out &out, int M, int N,
                      int K, int splits, float *workspace) {
  if (M <= 0 || N <= 0) return;
  using G = SimtGemm<OpA, OpB, Out, kAFastK, kBFastK, kEpiFastN>;
  G g;
  g.a = a; g.b = b; g.out = out;
  g.M = M; g.N = N; g.K = K;
  if (splits < 1) splits = 1;
  int chunk = (K + splits - 1) / splits;
  chunk = ((chunk + SBK - 1) / SBK) * SBK;
  if (chunk < SBK) chunk = SBK;
  splits = K > 0 ? (K + chunk - 1) / chunk : 1;
  g.k_chunk = chunk;
  g.workspace = workspace;
  dim3 grid(ceil_div_u(M, SBM), ceil_div_u(N, SBN), splits);
  KCNN_LAUNCH((gemm_simt_kernel<G, OpA, OpB, Out, kAFastK, kBFastK, kEpiFastN>), grid, STHREADS, 0,
              st, g);
  if (splits > 1)
    KCNN_LAUNCH(splitk_reduce_kernel<Out>, ceil_div_u((long long)M * N, 256), 256, 0, st,
                workspace, splits, M, N, out, FastDiv((uint32_t)N));
}

}  // namespace kcnn

#endif
