// kaldi-cnn_b200/csrc/cnslmat/kernels_gemm.cu
//
// Entry points of the GEMM-shaped part of the hot path: convolution forward /
// input-gradient / weight-gradient and the three affine GEMMs.  Each builds the
// operand decoders (gemm_operands.cuh) for its tensor roles and launches either
// the tcgen05 TF32 kernel (gemm_tc.cuh, KCNN_MATH_TF32_TC) or the FP32 CUDA-core
// kernel (gemm_simt.cuh, KCNN_MATH_FP32_SIMT).
//
// Index algebra (verified against the reference's CPU loops):
//   activation row n : col = h + w*H + c*H*W                      [C][W][H]
//   kernel           : row = kh + kw*KH + c*KH*KW, col = g
//   fprop  Y[n,g,ow,oh] = b[g] + sum_{c,kw,kh} Xp[n,c,ow+kw,oh+kh] K[c,kw,kh,g]
//   dgrad  dX[n,c,w,h]  = sum_{g,kw,kh} dY[n,g,w+pw-kw,h+ph-kh] K[c,kw,kh,g]
//   wgrad  dK[c,kw,kh,g]= sum_{n,ow,oh} Xp[n,c,ow+kw,oh+kh] dY[n,g,ow,oh]

#include <mutex>
#include <stdio.h>
#include <stdlib.h>

#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "conv_tma.cuh"

namespace kcnn {

namespace tma {

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library links
// the static CUDA runtime only, libcuda.so is whatever the process already has.
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      fprintf(stderr, "kaldi-cnn_b200: cuTensorMapEncodeTiled unavailable; TMA GEMMs disabled\n");
  }
  return fn;
}

// TFLOAT32: the TMA unit converts FP32 -> TF32 while it copies, so the tensor core sees
// rounded operands like the software producer's cvt.rna (KCNN_TMA_DTYPE=f32 copies raw bits,
// which the tensor core truncates).
int tma_data_type() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_TMA_DTYPE");
    v = (e && e[0] == 'f') ? (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (int)CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  }
  return v;
}

bool enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_TMA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    // Validated (same results) but measured SLOWER than the single-CTA tiles on this part
    // (4096^3: 472 vs 599 TFLOP/s, profiles/), so it is opt-in.
    const char *e = getenv("KCNN_TMA_PAIR");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

bool persistent_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_TMA_PERSIST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool cluster_k_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_TMA_CLUSTERK");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool deep_ring_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_TMA_DEEP");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// Grow-only device scratch, one buffer per (device, slot).  Grown with cudaMalloc, so the first call of a
// shape must happen outside stream capture (warm-up steps do that; while `st` is being captured a buffer
// that would have to grow is refused and the caller takes another path); outgrown buffers are kept alive
// because CUDA graphs captured earlier still point at them.  The table is per device and guarded by a
// mutex; the CONTENTS of a slot belong to whoever launched last, so the bare-component path that uses it
// is limited to one stream per device at a time (NnetMinibatchUpdater's fused step never comes here).
float *scratch(int slot, size_t bytes, cudaStream_t st) {
  constexpr int kMaxDevices = 32;
  static float *buf[kMaxDevices][SCRATCH_SLOTS] = {};
  static size_t cap[kMaxDevices][SCRATCH_SLOTS] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lock(mu);
  if (bytes > cap[dev][slot]) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return nullptr;
    }
    size_t want = bytes > 2 * cap[dev][slot] ? bytes : 2 * cap[dev][slot];
    float *p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    buf[dev][slot] = p;
    cap[dev][slot] = want;
  }
  return buf[dev][slot];
}

}  // namespace tma

using Op33 = Operand<Dec3, Dec3>;
using Op3I = Operand<Dec3, Dec3Inner>;
using Out33 = OutputMap<Dec3, Dec3>;

static void check_int32(const MatrixDim &d, const char *what) {
  if ((long long)d.rows * d.stride >= (1ll << 31)) {
    fprintf(stderr, "kaldi-cnn_b200: %s has rows*stride >= 2^31 (%d x %d); the hot path indexes "
            "with int32 like the reference\n", what, d.rows, d.stride);
    abort();
  }
}

static Op33 make_op(const float *base, const Dec3 &mn, const Dec3 &k, int wlim, int hlim) {
  Op33 o; o.base = base; o.mn = mn; o.k = k; o.wlim = wlim; o.hlim = hlim; return o;
}
static Out33 make_out(float *base, const Dec3 &m, const Dec3 &n, const float *bias_n) {
  Out33 o; o.base = base; o.m = m; o.n = n; o.bias_n = bias_n; o.bias_m = nullptr; return o;
}

// Column sums over rows and over the `inner` positions of each map:
//   out[g] = sum_{n < rows} sum_{p < inner} m[n*stride + g*inner + p]
// (db of the convolution; inner = 1 gives AddRowSumMat for the affine layer).
__global__ void __launch_bounds__(256)
colsum_maps_kernel(const float *__restrict__ m, int rows, int stride, int inner,
                   float *__restrict__ out, FastDiv div_inner) {
  kcnn::pdl_prologue();
  __shared__ float scratch[8];
  const int g = blockIdx.x;
  const float *base = m + (size_t)g * inner;
  float s = 0.0f;
  const int total = rows * inner;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    uint32_t n, p;
    div_inner.divmod((uint32_t)e, n, p);
    s += __ldg(base + (size_t)n * stride + p);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < 8 ? scratch[threadIdx.x] : 0.0f;
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[g] = s;
  }
}

// inner == 1: a block owns 32 adjacent columns and reads 128-byte row segments with 32
// row-threads, four independent accumulators each.  out[c] = alpha * sum (+ out[c] if accumulate).
__global__ void __launch_bounds__(1024)
colsum_rows_kernel(const float *__restrict__ m, int rows, int cols, int stride,
                   float *__restrict__ out, float alpha, int accumulate) {
  kcnn::pdl_prologue();
  __shared__ float part[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  if (col < cols) {
    const float *p = m + col;
    int r = ty;
    for (; r + 96 < rows; r += 128) {
      s0 += __ldg(p + (size_t)r * stride);
      s1 += __ldg(p + (size_t)(r + 32) * stride);
      s2 += __ldg(p + (size_t)(r + 64) * stride);
      s3 += __ldg(p + (size_t)(r + 96) * stride);
    }
    for (; r < rows; r += 32) s0 += __ldg(p + (size_t)r * stride);
  }
  part[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && col < cols) {
    float v = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) v += part[i][tx];
    out[col] = accumulate ? fmaf(alpha, v, out[col]) : alpha * v;
  }
}

static void launch_colsum(cudaStream_t st, const float *m, int rows, int stride, int maps,
                          int inner, float *out, float alpha = 1.0f, int accumulate = 0) {
  if (maps == 0) return;
  if (inner == 1)
    KCNN_LAUNCH(colsum_rows_kernel, ceil_div_u(maps, 32), 1024, 0, st, m, rows, maps, stride, out, alpha,
                accumulate);
  else
    KCNN_LAUNCH(colsum_maps_kernel, maps, 256, 0, st, m, rows, stride, inner, out,
                FastDiv((uint32_t)inner));
}

static int pick_splits(int M, int N, int K) {
  long long tiles = (long long)((M + SBM - 1) / SBM) * ((N + SBN - 1) / SBN);
  long long want = (2 * kNumSMs + tiles - 1) / tiles;
  long long max_by_k = K / (8 * SBK);
  if (max_by_k < 1) max_by_k = 1;
  if (want > max_by_k) want = max_by_k;
  if (want > 64) want = 64;
  return (int)(want < 1 ? 1 : want);
}

// One problem, two kernels: tcgen05 TF32 or FP32 CUDA cores.  `ws` must hold
// gemm_workspace_bytes(M, N, K) bytes when that is non-zero (split-K partials).
template <bool kAFastK, bool kBFastK, bool kEpiFastN, class OpA, class OpB, class Out>
static void launch_gemm(cudaStream_t st, int math, const OpA &a, const OpB &b, const Out &o, int M,
                        int N, int K, bool allow_split, float *ws) {
  if (math == KCNN_MATH_TF32_TC) {
    tc::launch_gemm_tc<kAFastK, kBFastK, kEpiFastN>(st, a, b, o, M, N, K, allow_split ? ws : nullptr);
  } else {
    int splits = (allow_split && ws) ? pick_splits(M, N, K) : 1;
    launch_gemm_simt<kAFastK, kBFastK, kEpiFastN>(st, a, b, o, M, N, K, splits, ws);
  }
}

static size_t gemm_workspace_bytes(int M, int N, int K) {
  size_t simt = (size_t)pick_splits(M, N, K) * M * N * sizeof(float);
  size_t tcb = tc::workspace_bytes(M, N, K);
  return simt > tcb ? simt : tcb;
}

struct ConvGeom {
  int N, H, W, C, ph, pw, KH, KW, G, OH, OW, P, ks;
};

static ConvGeom conv_geom(int N, int H, int W, int C, int ph, int pw, int KH, int KW, int G) {
  ConvGeom q;
  q.N = N; q.H = H; q.W = W; q.C = C; q.ph = ph; q.pw = pw; q.KH = KH; q.KW = KW; q.G = G;
  q.OH = H + 2 * ph - KH + 1;
  q.OW = W + 2 * pw - KW + 1;
  q.P = q.OH * q.OW;
  q.ks = KH * KW;
  return q;
}

}  // namespace kcnn

using namespace kcnn;

extern "C" {

size_t kcnn_conv2d_staging_floats(int num_rows, int H, int W, int C, int ph, int pw, int KH, int KW, int G) {
  ConvGeom q = conv_geom(num_rows, H, W, C, ph, pw, KH, KW, G);
  if (!(H == 1 && KH == 1 && ph == 0) || q.P <= 0) return 0;
  tma::ConvShape cs = {q.N, W, C, pw, KW, G, q.OW};
  return tma::conv_tma_shape_ok(cs) ? (size_t)q.N * W * C : 0;
}

void cudaF_conv2d_fprop(cudaStream_t st, int math, const float *in, MatrixDim id,
                        const float *kernel, MatrixDim kd, const float *bias, float *out,
                        MatrixDim od, int H, int W, int C, int ph, int pw, int KH, int KW, int G,
                        int concat) {
  (void)cudaF_conv2d_fprop_act(st, math, in, id, kernel, kd, bias, out, od, H, W, C, ph, pw, KH, KW, G,
                               concat, nullptr, KCNN_ACT_NONE);
}

int cudaF_conv2d_fprop_staged(cudaStream_t st, int math, const float *in, MatrixDim id,
                               const float *kernel, MatrixDim kd, const float *bias, float *out,
                               MatrixDim od, int H, int W, int C, int ph, int pw, int KH, int KW, int G,
                               int concat, float *staging) {
  return cudaF_conv2d_fprop_act(st, math, in, id, kernel, kd, bias, out, od, H, W, C, ph, pw, KH, KW, G, concat,
                                staging, KCNN_ACT_NONE);
}

int cudaF_conv2d_fprop_act(cudaStream_t st, int math, const float *in, MatrixDim id,
                            const float *kernel, MatrixDim kd, const float *bias, float *out,
                            MatrixDim od, int H, int W, int C, int ph, int pw, int KH, int KW, int G,
                            int concat, float *staging, int act) {
  ConvGeom q = conv_geom(id.rows, H, W, C, ph, pw, KH, KW, G);
  if (q.N == 0 || q.P <= 0 || G == 0) return 0;
  check_int32(id, "conv input"); check_int32(od, "conv output"); check_int32(kd, "conv kernel");
  const int M = q.N * q.P, K = q.ks * C;
  const bool relu = act == KCNN_ACT_RELU;
  if (math == KCNN_MATH_TF32_TC && concat && H == 1 && KH == 1 && ph == 0) {
    tma::ConvShape cs = {q.N, W, C, pw, KW, G, q.OW};
    if (tma::conv_fprop(st, cs, in, id.stride, kernel, kd.stride, bias, out, od.stride, staging, relu))
      return staging != nullptr ? 1 : 0;
  }
  // (any other path does not fill `staging`: the caller must only trust it when
  //  kcnn_conv2d_staging_floats() was non-zero AND the math mode is the tensor-core one.)
  if (math == KCNN_MATH_TF32_TC && concat && KH == H && H > 1 && ph == 0 && pw == 0) {
    tma::ConvFullShape fs = {q.N, H, W, C, KW, G, q.OW};
    tma::ConvOut o = {out, od.stride, false, bias, relu, nullptr};
    if (tma::conv_full_fprop(st, fs, in, id.stride, kernel, kd.stride, o)) return 0;
  }
  // the generic kernels below have no activation in their epilogue: one in-place pass after them
  struct ReluAfter {
    cudaStream_t st; float *out; MatrixDim od; bool on;
    ~ReluAfter() { if (on) cudaF_relu_fprop(st, out, od, out, od); }
  } relu_after = {st, out, od, relu};
  // A(m = (n, ow, oh), k = (c, kw, kh)) = Xpad[n, c, ow + kw, oh + kh]
  Op33 a = make_op(in,
      make_dec3(q.N, q.OW, q.OH, id.stride, H, 1, -pw * H - ph, 1, 1, -pw, -ph),
      make_dec3(C, KW, KH, H * W, H, 1, 0, 1, 1, 0, 0), W, H);
  // B(k, g) = kernel[k, g]
  Op33 b = make_op(kernel, make_linear(G, 1), make_linear(K, kd.stride), 1, 1);
  Out33 o = concat
      // out[n, g*P + pos]                      (folds _convmat_to_out)
      ? make_out(out, make_dec3(q.N, q.P, 1, od.stride, 1, 0, 0, 0, 0, 0, 0), make_linear(G, q.P), bias)
      // out[pos*N + n, g]                      (concat = false, conv2D.cc:199)
      : make_out(out, make_dec3(q.N, q.P, 1, od.stride, (long long)q.N * od.stride, 0, 0, 0, 0, 0, 0),
                 make_linear(G, 1), bias);
  const bool a_fast_k = (q.OH == 1 && KH > 1);   // time-axis layers: the (kw, kh) run is contiguous
  if (concat) {
    if (a_fast_k) launch_gemm<true, false, false>(st, math, a, b, o, M, G, K, false, nullptr);
    else          launch_gemm<false, false, false>(st, math, a, b, o, M, G, K, false, nullptr);
  } else {
    if (a_fast_k) launch_gemm<true, false, true>(st, math, a, b, o, M, G, K, false, nullptr);
    else          launch_gemm<false, false, true>(st, math, a, b, o, M, G, K, false, nullptr);
  }
  return 0;
}

void cudaF_conv2d_dgrad(cudaStream_t st, int math, const float *out_deriv, MatrixDim odd,
                        const float *kernel, MatrixDim kd, float *in_deriv, MatrixDim idd, int H,
                        int W, int C, int ph, int pw, int KH, int KW, int G) {
  ConvGeom q = conv_geom(odd.rows, H, W, C, ph, pw, KH, KW, G);
  if (q.N == 0 || C == 0) return;
  check_int32(odd, "conv out_deriv"); check_int32(idd, "conv in_deriv"); check_int32(kd, "conv kernel");
  if (math == KCNN_MATH_TF32_TC && H == 1 && KH == 1 && ph == 0) {
    tma::ConvShape cs = {q.N, W, C, pw, KW, G, q.OW};
    tma::ConvBackward b = {};
    b.out_deriv = out_deriv; b.ld_od = odd.stride; b.kernel = const_cast<float *>(kernel); b.ld_k = kd.stride;
    b.in_deriv = in_deriv; b.ld_id = idd.stride;
    if (tma::conv_backward(st, cs, b)) return;
  }
  if (math == KCNN_MATH_TF32_TC && KH == H && H > 1 && ph == 0 && pw == 0) {
    tma::ConvFullShape fs = {q.N, H, W, C, KW, G, q.OW};
    tma::ConvBackward b = {};
    b.out_deriv = out_deriv; b.ld_od = odd.stride; b.kernel = const_cast<float *>(kernel); b.ld_k = kd.stride;
    b.in_deriv = in_deriv; b.ld_id = idd.stride;
    if (tma::conv_full_backward(st, fs, b)) return;
  }
  if (q.OH == 1) {
    // Full-height kernel (every layer of egs/exp/nnet/nnet.config): out_deriv has one row
    // of positions, so kh is fixed by h (kh = h + ph) and moves from the reduction into the
    // GEMM's N axis:  dX[(n, w), (c, h)] = sum_{kw, g} dY[n, g, w + pw - kw] K[c, kw, h + ph, g].
    // (The reference's "no-flip" branch exploits the same structure with four copies,
    // nnet0/nnet-component-nnet0.cc:499-528.)  N = C*H instead of C: conv1 (C = 1, H = 40)
    // becomes a 40-wide GEMM instead of a 1-wide one.
    const int M = q.N * W, Ng = C * H, K = KW * G;
    Op3I a; a.base = out_deriv;
    a.mn = make_dec3(q.N, W, 1, odd.stride, 1, 0, pw, 1, 0, pw, 0);
    a.k = make_dec3_inner(KW, 1, G, q.OW, -1, 0, 0, -1, 0, 0, 0);
    a.wlim = q.OW; a.hlim = 1;
    Op3I b; b.base = kernel;
    b.mn = make_dec3(C, H, 1, (long long)q.ks * kd.stride, kd.stride, 0, (long long)ph * kd.stride, 0, 0, 0, 0);
    b.k = make_dec3_inner(KW, 1, G, 1, (long long)KH * kd.stride, 0, 0, 0, 0, 0, 0);
    b.wlim = 1; b.hlim = 1;
    Out33 o = make_out(in_deriv, make_dec3(q.N, W, 1, idd.stride, H, 0, 0, 0, 0, 0, 0),
                       make_dec3(C, H, 1, H * W, 1, 0, 0, 0, 0, 0, 0), nullptr);
    launch_gemm<false, true, true>(st, math, a, b, o, M, Ng, K, false, nullptr);
    return;
  }
  const int M = q.N * H * W, K = q.ks * G;
  // A(m = (n, w, h), k = (kw, kh, g)) = dY[n, g, w + pw - kw, h + ph - kh]
  // (folds PaddingZero of out_deriv and the 180-degree rotation of FlipMat)
  Op3I a; a.base = out_deriv;
  a.mn = make_dec3(q.N, W, H, odd.stride, q.OH, 1, pw * q.OH + ph, 1, 1, pw, ph);
  a.k = make_dec3_inner(KW, KH, G, q.P, -q.OH, -1, 0, -1, -1, 0, 0);
  a.wlim = q.OW; a.hlim = q.OH;
  // B(k = (kw, kh, g), c) = kernel[(c*KW + kw)*KH + kh, g]   (folds FlipMat's in/out swap)
  Op3I b; b.base = kernel;
  b.mn = make_linear(C, (long long)q.ks * kd.stride);
  b.k = make_dec3_inner(KW, KH, G, 1, (long long)KH * kd.stride, kd.stride, 0, 0, 0, 0, 0);
  b.wlim = 1; b.hlim = 1;
  // dX[n, c*H*W + w*H + h]
  Out33 o = make_out(in_deriv, make_dec3(q.N, H * W, 1, idd.stride, 1, 0, 0, 0, 0, 0, 0),
                     make_linear(C, H * W), nullptr);
  launch_gemm<false, true, false>(st, math, a, b, o, M, C, K, false, nullptr);
}

size_t kcnn_conv2d_wgrad_workspace(int num_rows, int H, int W, int C, int ph, int pw, int KH,
                                   int KW, int G) {
  ConvGeom q = conv_geom(num_rows, H, W, C, ph, pw, KH, KW, G);
  int M = q.ks * C, K = q.N * q.P;
  if (M <= 0 || G <= 0 || K <= 0) return 0;
  return gemm_workspace_bytes(M, G, K);
}

void cudaF_conv2d_wgrad(cudaStream_t st, int math, const float *in_value, MatrixDim ivd,
                        const float *out_deriv, MatrixDim odd, float *kernel_grad, MatrixDim kgd,
                        float *bias_grad, void *workspace, int H, int W, int C, int ph, int pw,
                        int KH, int KW, int G) {
  ConvGeom q = conv_geom(ivd.rows, H, W, C, ph, pw, KH, KW, G);
  if (q.N == 0 || q.P <= 0 || G == 0 || C == 0) return;
  check_int32(ivd, "conv in_value"); check_int32(odd, "conv out_deriv");
  const int M = q.ks * C, K = q.N * q.P;
  if (math == KCNN_MATH_TF32_TC && H == 1 && KH == 1 && ph == 0) {
    tma::ConvShape cs = {q.N, W, C, pw, KW, G, q.OW};
    tma::ConvBackward b = {};
    b.in_value = in_value; b.ld_iv = ivd.stride; b.out_deriv = out_deriv; b.ld_od = odd.stride;
    b.kernel_grad = kernel_grad; b.ld_kg = kgd.stride; b.want_bias = bias_grad != nullptr;
    if (tma::conv_backward(st, cs, b)) {
      if (bias_grad) launch_colsum(st, b.bias_partial, b.bias_rows, G, G, 1, bias_grad);
      return;
    }
  }
  if (math == KCNN_MATH_TF32_TC && KH == H && H > 1 && ph == 0 && pw == 0) {
    tma::ConvFullShape fs = {q.N, H, W, C, KW, G, q.OW};
    tma::ConvBackward b = {};
    b.in_value = in_value; b.ld_iv = ivd.stride; b.out_deriv = out_deriv; b.ld_od = odd.stride;
    b.kernel_grad = kernel_grad; b.ld_kg = kgd.stride; b.want_bias = bias_grad != nullptr;
    if (tma::conv_full_backward(st, fs, b)) {
      if (bias_grad) launch_colsum(st, b.bias_partial, b.bias_rows, G, G, 1, bias_grad);
      return;
    }
  }
  if (bias_grad) launch_colsum(st, out_deriv, q.N, odd.stride, G, q.P, bias_grad);
  // A(m = (c, kw, kh), k = (n, ow, oh)) = Xpad[n, c, ow + kw, oh + kh]
  // (folds PaddingZero + TpBlock; row order (c, kw, kh) folds ModPermuteRow)
  Op33 a = make_op(in_value,
      make_dec3(C, KW, KH, H * W, H, 1, -pw * H - ph, 1, 1, -pw, -ph),
      make_dec3(q.N, q.OW, q.OH, ivd.stride, H, 1, 0, 1, 1, 0, 0), W, H);
  // B(k = (n, pos), g) = dY[n, g*P + pos]      (folds TpInsideBlock)
  Op33 b = make_op(out_deriv, make_linear(G, q.P),
                   make_dec3(q.N, q.P, 1, odd.stride, 1, 0, 0, 0, 0, 0, 0), 1, 1);
  Out33 o = make_out(kernel_grad, make_linear(M, kgd.stride), make_linear(G, 1), nullptr);
  const bool a_fast_k = !(q.OH == 1 && KH > 1);
  const bool b_fast_k = q.P >= 4;
  float *ws = static_cast<float *>(workspace);
  if (a_fast_k && b_fast_k)       launch_gemm<true, true, true>(st, math, a, b, o, M, G, K, true, ws);
  else if (a_fast_k && !b_fast_k) launch_gemm<true, false, true>(st, math, a, b, o, M, G, K, true, ws);
  else if (!a_fast_k && b_fast_k) launch_gemm<false, true, true>(st, math, a, b, o, M, G, K, true, ws);
  else                            launch_gemm<false, false, true>(st, math, a, b, o, M, G, K, true, ws);
}

int cudaF_conv2d_backward(cudaStream_t st, int math, const float *in_value, MatrixDim ivd,
                          const float *out_deriv, MatrixDim odd, float *kernel, MatrixDim kd,
                          float *in_deriv, MatrixDim idd, float *kernel_grad, MatrixDim kgd,
                          float *bias_grad, float *prev_grad, MatrixDim pd, float *bias, int apply,
                          float momentum, float a_decay, float a_grad, const float *staged_input, int H,
                          int W, int C, int ph, int pw, int KH, int KW, int G) {
  ConvGeom q = conv_geom(odd.rows, H, W, C, ph, pw, KH, KW, G);
  if (q.N == 0 || q.P <= 0 || G == 0 || C == 0) return 0;
  if (math != KCNN_MATH_TF32_TC) return 0;
  const bool time_axis = H == 1 && KH == 1 && ph == 0;
  const bool full_height = KH == H && H > 1 && ph == 0 && pw == 0;
  if (!time_axis && !full_height) return 0;
  tma::ConvShape cs = {q.N, W, C, pw, KW, G, q.OW};
  tma::ConvFullShape fs = {q.N, H, W, C, KW, G, q.OW};
  tma::SgdCoef coef = {momentum, a_decay, a_grad};
  tma::ConvBackward b = {};
  b.in_value = in_value; b.ld_iv = ivd.stride; b.out_deriv = out_deriv; b.ld_od = odd.stride;
  b.kernel = kernel; b.ld_k = kd.stride; b.in_deriv = in_deriv; b.ld_id = idd.stride;
  b.want_bias = true;
  b.staged_x = time_axis ? staged_input : nullptr;
  if (apply) {
    b.prev = prev_grad; b.ld_p = pd.stride; b.sgd = &coef;
  } else {
    b.kernel_grad = kernel_grad; b.ld_kg = kgd.stride;
  }
  if (!(time_axis ? tma::conv_backward(st, cs, b) : tma::conv_full_backward(st, fs, b))) return 0;
  if (apply) launch_colsum(st, b.bias_partial, b.bias_rows, G, G, 1, bias, a_grad, 1);    // bias += lr * db
  else       launch_colsum(st, b.bias_partial, b.bias_rows, G, G, 1, bias_grad);
  return 1;
}

int cudaF_affine_wgrad_sgd(cudaStream_t st, int math, const float *in_value, MatrixDim ivd,
                           const float *out_deriv, MatrixDim odd, float *w, MatrixDim wd,
                           float *prev_grad, MatrixDim pd, float *bias, float momentum,
                           float a_decay, float a_grad) {
  const int M = odd.cols, N = ivd.cols, K = ivd.rows;
  if (M == 0 || N == 0 || K == 0) return 0;
  if (!(math == KCNN_MATH_TF32_TC && tma::enabled())) return 0;
  if (pd.stride != wd.stride || !host_aligned16(prev_grad) || !host_aligned16(w) || (wd.stride & 3)) return 0;
  tma::Epilogue epi;
  epi.mode = tma::EPI_SGD; epi.aux = prev_grad;
  epi.sgd.momentum = momentum; epi.sgd.a_decay = a_decay; epi.sgd.a_grad = a_grad;
  // bias first: it reads out_deriv only, and nothing below touches the bias
  if (!tma::gemm<true, true>(st, tma::Matrix{out_deriv, K, M, odd.stride}, tma::Matrix{in_value, K, N, ivd.stride},
                             M, N, K, w, wd.stride, epi, true))
    return 0;
  if (bias) launch_colsum(st, out_deriv, odd.rows, odd.stride, M, 1, bias, a_grad, 1);
  return 1;
}

void cudaF_sum_rows_per_map(cudaStream_t st, const float *m, MatrixDim md, int inner, float *out) {
  if (inner <= 0 || md.cols == 0) return;
  launch_colsum(st, m, md.rows, md.stride, md.cols / inner, inner, out);
}

void cudaF_affine_fprop(cudaStream_t st, int math, const float *in, MatrixDim id, const float *w,
                        MatrixDim wd, const float *bias, float *out, MatrixDim od) {
  cudaF_affine_fprop_act(st, math, in, id, w, wd, bias, out, od, KCNN_ACT_NONE);
}

void cudaF_affine_fprop_act(cudaStream_t st, int math, const float *in, MatrixDim id, const float *w,
                            MatrixDim wd, const float *bias, float *out, MatrixDim od, int act) {
  const int M = id.rows, N = wd.rows, K = wd.cols;
  if (M == 0 || N == 0) return;
  check_int32(id, "affine input"); check_int32(od, "affine output"); check_int32(wd, "affine weights");
  // out = 1 bias^T + in W^T : A(m, k) = in[m, k], B(k, n) = W[n, k]
  if (math == KCNN_MATH_TF32_TC && tma::enabled()) {
    tma::Epilogue epi; epi.rows.bias_n = bias; epi.rows.relu = act == KCNN_ACT_RELU ? 1 : 0;
    if (tma::gemm<false, false>(st, tma::Matrix{in, M, K, id.stride}, tma::Matrix{w, N, K, wd.stride}, M, N, K,
                                out, od.stride, epi, true))
      return;
  }
  Op33 a = make_op(in, make_linear(M, id.stride), make_linear(K, 1), 1, 1);
  Op33 b = make_op(w, make_linear(N, wd.stride), make_linear(K, 1), 1, 1);
  Out33 o = make_out(out, make_linear(M, od.stride), make_linear(N, 1), bias);
  launch_gemm<true, true, true>(st, math, a, b, o, M, N, K, false, nullptr);
  if (act == KCNN_ACT_RELU) cudaF_relu_fprop(st, out, od, out, od);      // generic kernels: separate pass
}

void cudaF_affine_dgrad(cudaStream_t st, int math, const float *out_deriv, MatrixDim odd,
                        const float *w, MatrixDim wd, float *in_deriv, MatrixDim idd) {
  const int M = odd.rows, N = wd.cols, K = wd.rows;
  if (M == 0 || N == 0) return;
  check_int32(odd, "affine out_deriv"); check_int32(idd, "affine in_deriv");
  // in_deriv = out_deriv W : A(m, k) = dY[m, k], B(k, n) = W[k, n]
  if (math == KCNN_MATH_TF32_TC && tma::enabled()) {
    tma::Epilogue epi;
    if (tma::gemm<false, true>(st, tma::Matrix{out_deriv, M, K, odd.stride}, tma::Matrix{w, K, N, wd.stride}, M, N, K,
                               in_deriv, idd.stride, epi, true))
      return;
  }
  Op33 a = make_op(out_deriv, make_linear(M, odd.stride), make_linear(K, 1), 1, 1);
  Op33 b = make_op(w, make_linear(N, 1), make_linear(K, wd.stride), 1, 1);
  Out33 o = make_out(in_deriv, make_linear(M, idd.stride), make_linear(N, 1), nullptr);
  launch_gemm<true, false, true>(st, math, a, b, o, M, N, K, false, nullptr);
}

void cudaF_affine_wgrad(cudaStream_t st, int math, const float *in_value, MatrixDim ivd,
                        const float *out_deriv, MatrixDim odd, float *w_grad, MatrixDim wgd,
                        float *bias_grad) {
  const int M = odd.cols, N = ivd.cols, K = ivd.rows;
  if (M == 0 || N == 0) return;
  check_int32(ivd, "affine in_value"); check_int32(odd, "affine out_deriv");
  if (bias_grad) launch_colsum(st, out_deriv, odd.rows, odd.stride, M, 1, bias_grad);
  // w_grad = out_deriv^T in_value : A(m, k) = dY[k, m], B(k, n) = X[k, n]
  if (math == KCNN_MATH_TF32_TC && tma::enabled()) {
    tma::Epilogue epi;
    if (tma::gemm<true, true>(st, tma::Matrix{out_deriv, K, M, odd.stride}, tma::Matrix{in_value, K, N, ivd.stride},
                              M, N, K, w_grad, wgd.stride, epi, true))
      return;
  }
  Op33 a = make_op(out_deriv, make_linear(M, 1), make_linear(K, odd.stride), 1, 1);
  Op33 b = make_op(in_value, make_linear(N, 1), make_linear(K, ivd.stride), 1, 1);
  Out33 o = make_out(w_grad, make_linear(M, wgd.stride), make_linear(N, 1), nullptr);
  launch_gemm<false, false, true>(st, math, a, b, o, M, N, K, false, nullptr);
}

// ---- (4) the fused training step: channels-last activations (nnet2/nnet-fused.cc) -------------

int kcnn_conv_time_shape_ok(int N, int W, int C, int pw, int KW, int G) {
  tma::ConvShape cs = {N, W, C, pw, KW, G, W + 2 * pw - KW + 1};
  return tma::conv_tma_shape_ok(cs) && tma::encode_tiled_fn() != nullptr ? 1 : 0;
}

int kcnn_conv_full_shape_ok(int N, int H, int W, int C, int KW, int G) {
  tma::ConvFullShape fs = {N, H, W, C, KW, G, W - KW + 1};
  return tma::conv_full_shape_ok(fs) && ((KW * H) & 31) == 0 && (G & 3) == 0 && tma::encode_tiled_fn() != nullptr ? 1 : 0;
}

int cudaF_conv_time_fprop_cl(cudaStream_t st, const float *x, int N, int W, int C, int pw, int KW, int G,
                             const float *kernel, MatrixDim kd, const float *bias, float *out, int out_cl,
                             int ldo, int relu) {
  if (N == 0) return 1;
  tma::ConvShape cs = {N, W, C, pw, KW, G, W + 2 * pw - KW + 1};
  tma::ConvOut o = {out, ldo, out_cl != 0, bias, relu != 0, nullptr};
  return tma::conv_rows_fprop(st, cs, x, kernel, kd.stride, o) ? 1 : 0;
}

int cudaF_conv_time_dgrad_cl(cudaStream_t st, const float *dy, int N, int W, int C, int pw, int KW, int G,
                             const float *kernel, MatrixDim kd, float *dx, const float *mask) {
  if (N == 0) return 1;
  tma::ConvShape cs = {N, W, C, pw, KW, G, W + 2 * pw - KW + 1};
  tma::ConvOut o = {dx, 0, true, nullptr, false, mask};
  return tma::conv_rows_dgrad(st, cs, dy, kernel, kd.stride, o) ? 1 : 0;
}

int cudaF_conv_time_wgrad_cl(cudaStream_t st, const float *x, const float *dy, int N, int W, int C, int pw,
                             int KW, int G, float *w, MatrixDim wd, float *prev_grad, MatrixDim pd, int apply,
                             float momentum, float a_decay, float a_grad) {
  if (N == 0) return 1;
  tma::ConvShape cs = {N, W, C, pw, KW, G, W + 2 * pw - KW + 1};
  tma::SgdCoef coef = {momentum, a_decay, a_grad};
  tma::ConvWgradOut o = {w, wd.stride, apply ? prev_grad : nullptr, pd.stride, apply ? &coef : nullptr};
  if (!host_aligned16(x) || !host_aligned16(dy)) return 0;
  return tma::conv_rows_wgrad(st, cs, x, dy, o) ? 1 : 0;
}

int cudaF_conv_full_fprop_cl(cudaStream_t st, const float *in, MatrixDim id, int H, int W, int C, int KW, int G,
                             const float *kernel, MatrixDim kd, const float *bias, float *out, int relu) {
  if (id.rows == 0) return 1;
  tma::ConvFullShape fs = {id.rows, H, W, C, KW, G, W - KW + 1};
  tma::ConvOut o = {out, 0, true, bias, relu != 0, nullptr};
  return tma::conv_full_fprop(st, fs, in, id.stride, kernel, kd.stride, o) ? 1 : 0;
}

int cudaF_conv_full_dgrad_cl(cudaStream_t st, const float *dy, int N, int H, int W, int C, int KW, int G,
                             const float *kernel, MatrixDim kd, float *in_deriv, MatrixDim idd) {
  if (N == 0) return 1;
  tma::ConvFullShape fs = {N, H, W, C, KW, G, W - KW + 1};
  return tma::conv_full_dgrad(st, fs, dy, kernel, kd.stride, in_deriv, idd.stride) ? 1 : 0;
}

int cudaF_conv_full_wgrad_cl(cudaStream_t st, const float *in, MatrixDim id, const float *dy, int H, int W, int C,
                             int KW, int G, float *w, MatrixDim wd, float *prev_grad, MatrixDim pd, int apply,
                             float momentum, float a_decay, float a_grad) {
  if (id.rows == 0) return 1;
  tma::ConvFullShape fs = {id.rows, H, W, C, KW, G, W - KW + 1};
  tma::SgdCoef coef = {momentum, a_decay, a_grad};
  tma::ConvWgradOut o = {w, wd.stride, apply ? prev_grad : nullptr, pd.stride, apply ? &coef : nullptr};
  return tma::conv_full_wgrad(st, fs, in, id.stride, dy, o) ? 1 : 0;
}

int cudaF_affine_fprop_fused(cudaStream_t st, const float *in, MatrixDim id, const float *w, MatrixDim wd,
                             const float *bias, float *out, MatrixDim od, int relu, float *drop_out,
                             MatrixDim dd, float dp, float low, float high, const unsigned long long *seed_dev) {
  const int M = id.rows, N = wd.rows, K = wd.cols;
  if (M == 0 || N == 0) return 1;
  if (!tma::enabled()) return 0;
  tma::Epilogue epi;
  epi.rows.bias_n = bias; epi.rows.relu = relu ? 1 : 0;
  if (drop_out) {
    epi.rows.out2 = drop_out; epi.rows.ld_o2 = dd.stride;
    epi.rows.dp = dp; epi.rows.low = low; epi.rows.high = high; epi.rows.seed = seed_dev;
  }
  return tma::gemm<false, false>(st, tma::Matrix{in, M, K, id.stride}, tma::Matrix{w, N, K, wd.stride}, M, N, K, out,
                                 od.stride, epi, true) ? 1 : 0;
}

int cudaF_affine_dgrad_fused(cudaStream_t st, const float *out_deriv, MatrixDim odd, const float *w, MatrixDim wd,
                             float *in_deriv, MatrixDim idd, const float *relu_out, int relu_stride,
                             const float *drop_out, int drop_stride, int perm_r) {
  const int M = odd.rows, N = wd.cols, K = wd.rows;
  if (M == 0 || N == 0) return 1;
  if (!tma::enabled()) return 0;
  if (perm_r > 0 && (N % perm_r) != 0) return 0;
  tma::Epilogue epi;
  if (relu_out) { epi.rows.mask_x = relu_out; epi.rows.ld_mx = relu_stride; }
  if (relu_out && drop_out) { epi.rows.mask_y = drop_out; epi.rows.ld_my = drop_stride; }
  if (perm_r > 0) { epi.rows.perm_r = perm_r; epi.rows.perm_g = N / perm_r; epi.rows.div_r = FastDiv((uint32_t)perm_r); }
  return tma::gemm<false, true>(st, tma::Matrix{out_deriv, M, K, odd.stride}, tma::Matrix{w, K, N, wd.stride}, M, N, K,
                                in_deriv, idd.stride, epi, true) ? 1 : 0;
}

}  // extern "C"
