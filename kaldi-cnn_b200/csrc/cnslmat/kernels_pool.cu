// kaldi-cnn_b200/csrc/cnslmat/kernels_pool.cu
//
// 3-D (frequency x time x intermap-channel) max pooling for sm_100a.
//
// What they replace: _maxpool_prop / _maxpool_backprop and the overlap
// variants of the reference (src/cnslmat/cnsl-cu-kernels.cu:231-503), which
// launch one thread per output on a 16x16 block with the row index on
// threadIdx.y, so adjacent threads of a warp straddle two matrix rows.
//
// Design (HBM-bound, roofline = bytes / 6.5 TB/s):
//  * Flat 1-D grid over (row, output) with the output index fastest: a warp
//    reads and writes contiguous spans of ONE row.
//  * Time-axis layers (H == 1, the shape of every layer in
//    egs/exp/nnet/nnet.config) take the vector path: one thread produces four
//    consecutive outputs, reads pw 128-bit words per pooled channel and writes
//    one 128-bit word; the whole window lives in registers, no shuffles needed
//    because a window never straddles threads.
//  * Everything else takes the scalar path, same mapping, 32-bit accesses.
//  * Bit-exactness: the running value starts at -1e20f and is replaced only on
//    strict '<', scanning c -> w -> h, exactly as cnsl-cu-kernels.cu:251-262.
//    That fixes the result for ties between -0.0f / +0.0f (first seen wins),
//    for NaN (never wins) and for windows entirely below the sentinel.
//  * Backward, reference-exact mode: "dest = err" at EVERY element equal to the
//    pooled value (:302-303).  With zero_others the kernel also writes the
//    zeros, folding MaxpoolComponent::Backprop's kSetZero pass
//    (nnet0/nnet-component-nnet0.cc:889) into the same sweep.
//  * Index mode: forward records the window position of the first maximum in
//    one byte; backward routes from it without reading the activations.

#include "kcnn_common.cuh"

namespace kcnn {

struct PoolGeom {
  int H, W, ph, pw, pc;      // input plane and window
  int OH, OW;                // output plane
  int out_cols;
  FastDiv div_ohw, div_oh;   // j -> (oc, pos) ; pos -> (ow, oh)
};

static PoolGeom make_geom(int H, int W, int ph, int pw, int pc, int out_cols) {
  PoolGeom g;
  g.H = H; g.W = W; g.ph = ph; g.pw = pw; g.pc = pc;
  g.OH = H / ph; g.OW = W / pw;
  g.out_cols = out_cols;
  g.div_ohw = FastDiv((uint32_t)(g.OH * g.OW));
  g.div_oh = FastDiv((uint32_t)g.OH);
  return g;
}

// ---------------------------------------------------------------- forward --

// Scalar path, plain (non-overlapping) windows.  One thread per output.
template <bool kIndex>
__global__ void __launch_bounds__(256)
maxpool_prop_scalar(const float *__restrict__ src, int src_stride, float *__restrict__ pool,
                    int pool_stride, unsigned char *__restrict__ index, int index_stride,
                    int rows, PoolGeom g, FastDiv div_cols) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * g.out_cols) return;
  uint32_t i, j;
  div_cols.divmod((uint32_t)t, i, j);
  uint32_t oc, pos, ow, oh;
  g.div_ohw.divmod(j, oc, pos);
  g.div_oh.divmod(pos, ow, oh);
  const int HW = g.H * g.W;
  const float *p = src + (size_t)i * src_stride + (size_t)oc * g.pc * HW + ow * g.pw * g.H + oh * g.ph;
  float val = -1e20f;
  int best = 0, k = 0;
  for (int c = 0; c < g.pc; c++) {
    const float *pc_ = p + (size_t)c * HW;
    for (int w = 0; w < g.pw; w++) {
      const float *pw_ = pc_ + w * g.H;
      for (int h = 0; h < g.ph; h++, k++) {
        float s = __ldg(pw_ + h);
        if (val < s) { val = s; if (kIndex) best = k; }
      }
    }
  }
  pool[(size_t)i * pool_stride + j] = val;
  if (kIndex) index[(size_t)i * index_stride + j] = (unsigned char)best;
}

// Vector path for H == 1 (so ph == 1, OH == 1): the output index within a row
// is j = oc*OW + ow and the window of (oc, ow) is, for each of pc planes, the pw
// contiguous floats at (oc*pc + c)*W + ow*pw.  One thread makes outputs
// ow0..ow0+3: PW float4 loads per plane, one float4 store.
// Requires OW % 4 == 0 and 16-byte aligned rows (checked by the launcher).
template <int PW, bool kIndex>
__global__ void __launch_bounds__(256)
maxpool_prop_time_vec4(const float *__restrict__ src, int src_stride, float *__restrict__ pool,
                       int pool_stride, unsigned char *__restrict__ index, int index_stride,
                       int rows, int W, int pc, int quads_per_row, FastDiv div_quads,
                       FastDiv div_owq) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * quads_per_row) return;
  uint32_t i, q, oc, owq;
  div_quads.divmod((uint32_t)t, i, q);
  div_owq.divmod(q, oc, owq);                      // owq = ow0 / 4
  const float *p = src + (size_t)i * src_stride + (size_t)oc * pc * W + owq * (4 * PW);
  float val[4] = {-1e20f, -1e20f, -1e20f, -1e20f};
  int best[4] = {0, 0, 0, 0};
#pragma unroll 2
  for (int c = 0; c < pc; c++) {
    float s[4 * PW];
    const float4 *p4 = reinterpret_cast<const float4 *>(p + (size_t)c * W);
#pragma unroll
    for (int v = 0; v < PW; v++) {
      float4 x = __ldg(p4 + v);
      s[4 * v + 0] = x.x; s[4 * v + 1] = x.y; s[4 * v + 2] = x.z; s[4 * v + 3] = x.w;
    }
#pragma unroll
    for (int o = 0; o < 4; o++) {
#pragma unroll
      for (int w = 0; w < PW; w++) {
        float e = s[o * PW + w];
        if (val[o] < e) { val[o] = e; if (kIndex) best[o] = c * PW + w; }
      }
    }
  }
  *reinterpret_cast<float4 *>(pool + (size_t)i * pool_stride + 4 * q) =
      make_float4(val[0], val[1], val[2], val[3]);
  if (kIndex) {
    uchar4 b = make_uchar4((unsigned char)best[0], (unsigned char)best[1],
                           (unsigned char)best[2], (unsigned char)best[3]);
    *reinterpret_cast<uchar4 *>(index + (size_t)i * index_stride + 4 * q) = b;
  }
}

// Overlap (1-D sliding window over channels, stride 1; cnsl-cu-kernels.cu:310-356)
// and overlap2D (channels as a sqrt(C) x sqrt(C) map, pc x pc window; :405-452).
template <int MODE>
__global__ void __launch_bounds__(256)
maxpool_prop_overlap(const float *__restrict__ src, int src_stride, float *__restrict__ pool,
                     int pool_stride, int rows, int out_cols, int HW, int pc, int o2, int i2,
                     FastDiv div_cols, FastDiv div_hw, FastDiv div_o2) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * out_cols) return;
  uint32_t i, j, oc, pos;
  div_cols.divmod((uint32_t)t, i, j);
  div_hw.divmod(j, oc, pos);
  const float *p = src + (size_t)i * src_stride + pos;
  float val = -1e20f;
  if (MODE == KCNN_POOL_OVERLAP) {
    for (int c = 0; c < pc; c++) {
      float s = __ldg(p + (size_t)(oc + c) * HW);
      if (val < s) val = s;
    }
  } else {
    uint32_t cx0, cy0;
    div_o2.divmod(oc, cx0, cy0);
    for (int cx = 0; cx < pc; cx++)
      for (int cy = 0; cy < pc; cy++) {
        float s = __ldg(p + (size_t)((cx0 + cx) * i2 + (cy0 + cy)) * HW);
        if (val < s) val = s;
      }
  }
  pool[(size_t)i * pool_stride + j] = val;
}

// --------------------------------------------------------------- backward --

// Reference-exact routing, plain windows, scalar.  One thread per output.
template <bool kZeroOthers>
__global__ void __launch_bounds__(256)
maxpool_backprop_scalar(const float *__restrict__ in_val, int in_stride,
                        const float *__restrict__ out_val, int ov_stride,
                        const float *__restrict__ out_deriv, int od_stride,
                        float *__restrict__ dest, int dest_stride, int rows, PoolGeom g,
                        FastDiv div_cols) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * g.out_cols) return;
  uint32_t i, j, oc, pos, ow, oh;
  div_cols.divmod((uint32_t)t, i, j);
  g.div_ohw.divmod(j, oc, pos);
  g.div_oh.divmod(pos, ow, oh);
  const int HW = g.H * g.W;
  size_t off = (size_t)oc * g.pc * HW + ow * g.pw * g.H + oh * g.ph;
  const float *p = in_val + (size_t)i * in_stride + off;
  float *d = dest + (size_t)i * dest_stride + off;
  float ov = __ldg(out_val + (size_t)i * ov_stride + j);
  float err = __ldg(out_deriv + (size_t)i * od_stride + j);
  for (int c = 0; c < g.pc; c++)
    for (int w = 0; w < g.pw; w++)
      for (int h = 0; h < g.ph; h++) {
        size_t e = (size_t)c * HW + w * g.H + h;
        float s = __ldg(p + e);
        if (ov == s) d[e] = err;
        else if (kZeroOthers) d[e] = 0.0f;
      }
}

// Reference-exact routing fused with the zero fill, H == 1 vector path.
template <int PW>
__global__ void __launch_bounds__(256)
maxpool_backprop_time_vec4(const float *__restrict__ in_val, int in_stride,
                           const float *__restrict__ out_val, int ov_stride,
                           const float *__restrict__ out_deriv, int od_stride,
                           float *__restrict__ dest, int dest_stride, int rows, int W, int pc,
                           int quads_per_row, FastDiv div_quads, FastDiv div_owq) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * quads_per_row) return;
  uint32_t i, q, oc, owq;
  div_quads.divmod((uint32_t)t, i, q);
  div_owq.divmod(q, oc, owq);
  size_t off = (size_t)oc * pc * W + owq * (4 * PW);
  const float *p = in_val + (size_t)i * in_stride + off;
  float *d = dest + (size_t)i * dest_stride + off;
  float4 ov4 = __ldg(reinterpret_cast<const float4 *>(out_val + (size_t)i * ov_stride + 4 * q));
  float4 er4 = __ldg(reinterpret_cast<const float4 *>(out_deriv + (size_t)i * od_stride + 4 * q));
  float ov[4] = {ov4.x, ov4.y, ov4.z, ov4.w};
  float er[4] = {er4.x, er4.y, er4.z, er4.w};
#pragma unroll 2
  for (int c = 0; c < pc; c++) {
    float s[4 * PW];
    const float4 *p4 = reinterpret_cast<const float4 *>(p + (size_t)c * W);
#pragma unroll
    for (int v = 0; v < PW; v++) {
      float4 x = __ldg(p4 + v);
      s[4 * v + 0] = x.x; s[4 * v + 1] = x.y; s[4 * v + 2] = x.z; s[4 * v + 3] = x.w;
    }
#pragma unroll
    for (int e = 0; e < 4 * PW; e++) s[e] = (ov[e / PW] == s[e]) ? er[e / PW] : 0.0f;
    float4 *d4 = reinterpret_cast<float4 *>(d + (size_t)c * W);
#pragma unroll
    for (int v = 0; v < PW; v++)
      d4[v] = make_float4(s[4 * v + 0], s[4 * v + 1], s[4 * v + 2], s[4 * v + 3]);
  }
}

// Overlapping windows accumulate.  The reference kernels do a racy
// read-modify-write from the OUTPUT side (cnsl-cu-kernels.cu:396-397, 497-498);
// here one thread owns one INPUT element and adds the matching windows in
// ascending output order, which is the serial order and is deterministic.
template <int MODE>
__global__ void __launch_bounds__(256)
maxpool_backprop_overlap(const float *__restrict__ in_val, int in_stride,
                         const float *__restrict__ out_val, int ov_stride,
                         const float *__restrict__ out_deriv, int od_stride,
                         float *__restrict__ dest, int dest_stride, int rows, int in_cols,
                         int HW, int pc, int OC, int o2, int i2, FastDiv div_cols,
                         FastDiv div_hw, FastDiv div_i2) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * in_cols) return;
  uint32_t i, col, c, pos;
  div_cols.divmod((uint32_t)t, i, col);
  div_hw.divmod(col, c, pos);
  float s = __ldg(in_val + (size_t)i * in_stride + col);
  const float *ov = out_val + (size_t)i * ov_stride + pos;
  const float *od = out_deriv + (size_t)i * od_stride + pos;
  float acc = dest[(size_t)i * dest_stride + col];
  bool any = false;
  if (MODE == KCNN_POOL_OVERLAP) {
    int lo = (int)c - pc + 1; if (lo < 0) lo = 0;
    int hi = (int)c; if (hi > OC - 1) hi = OC - 1;
    for (int oc = lo; oc <= hi; oc++)
      if (__ldg(ov + (size_t)oc * HW) == s) { acc = acc + __ldg(od + (size_t)oc * HW); any = true; }
  } else {
    uint32_t x, y;
    div_i2.divmod(c, x, y);
    int xlo = (int)x - pc + 1; if (xlo < 0) xlo = 0;
    int xhi = (int)x; if (xhi > o2 - 1) xhi = o2 - 1;
    int ylo = (int)y - pc + 1; if (ylo < 0) ylo = 0;
    int yhi = (int)y; if (yhi > o2 - 1) yhi = o2 - 1;
    for (int cx = xlo; cx <= xhi; cx++)
      for (int cy = ylo; cy <= yhi; cy++) {
        size_t oc = (size_t)cx * o2 + cy;
        if (__ldg(ov + oc * HW) == s) { acc = acc + __ldg(od + oc * HW); any = true; }
      }
  }
  if (any) dest[(size_t)i * dest_stride + col] = acc;
}

// Index-routed backward: whole dest written, activations not read.
__global__ void __launch_bounds__(256)
maxpool_backprop_index_scalar(const unsigned char *__restrict__ index, int index_stride,
                              const float *__restrict__ out_deriv, int od_stride,
                              float *__restrict__ dest, int dest_stride, int rows, PoolGeom g,
                              FastDiv div_cols) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * g.out_cols) return;
  uint32_t i, j, oc, pos, ow, oh;
  div_cols.divmod((uint32_t)t, i, j);
  g.div_ohw.divmod(j, oc, pos);
  g.div_oh.divmod(pos, ow, oh);
  const int HW = g.H * g.W;
  float *d = dest + (size_t)i * dest_stride + (size_t)oc * g.pc * HW + ow * g.pw * g.H + oh * g.ph;
  int best = index[(size_t)i * index_stride + j];
  float err = __ldg(out_deriv + (size_t)i * od_stride + j);
  int k = 0;
  for (int c = 0; c < g.pc; c++)
    for (int w = 0; w < g.pw; w++)
      for (int h = 0; h < g.ph; h++, k++)
        d[(size_t)c * HW + w * g.H + h] = (k == best) ? err : 0.0f;
}

template <int PW>
__global__ void __launch_bounds__(256)
maxpool_backprop_index_time_vec4(const unsigned char *__restrict__ index, int index_stride,
                                 const float *__restrict__ out_deriv, int od_stride,
                                 float *__restrict__ dest, int dest_stride, int rows, int W,
                                 int pc, int quads_per_row, FastDiv div_quads, FastDiv div_owq) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * quads_per_row) return;
  uint32_t i, q, oc, owq;
  div_quads.divmod((uint32_t)t, i, q);
  div_owq.divmod(q, oc, owq);
  float *d = dest + (size_t)i * dest_stride + (size_t)oc * pc * W + owq * (4 * PW);
  uchar4 b4 = *reinterpret_cast<const uchar4 *>(index + (size_t)i * index_stride + 4 * q);
  float4 er4 = __ldg(reinterpret_cast<const float4 *>(out_deriv + (size_t)i * od_stride + 4 * q));
  int best[4] = {b4.x, b4.y, b4.z, b4.w};
  float er[4] = {er4.x, er4.y, er4.z, er4.w};
  for (int c = 0; c < pc; c++) {
    float s[4 * PW];
#pragma unroll
    for (int e = 0; e < 4 * PW; e++)
      s[e] = (best[e / PW] == c * PW + (e % PW)) ? er[e / PW] : 0.0f;
    float4 *d4 = reinterpret_cast<float4 *>(d + (size_t)c * W);
#pragma unroll
    for (int v = 0; v < PW; v++)
      d4[v] = make_float4(s[4 * v + 0], s[4 * v + 1], s[4 * v + 2], s[4 * v + 3]);
  }
}

// ---------------------------------------------------------------- host side --

static bool time_vec_ok(int H, int W, int ph, int pw, const void *a, int sa, const void *b,
                        int sb) {
  if (H != 1 || ph != 1) return false;
  if (!(pw == 1 || pw == 2 || pw == 3 || pw == 4)) return false;
  int OW = W / pw;
  if (OW % 4 != 0 || W % 4 != 0) return false;
  if (sa % 4 != 0 || sb % 4 != 0) return false;
  return host_aligned16(a) && host_aligned16(b);
}

template <bool kIndex>
static void launch_prop_plain(cudaStream_t st, const float *src, MatrixDim sd, float *pool,
                              MatrixDim pd, unsigned char *index, int index_stride, int H, int W,
                              int ph, int pw, int pc) {
  if (pd.rows == 0 || pd.cols == 0) return;
  bool vec = time_vec_ok(H, W, ph, pw, src, sd.stride, pool, pd.stride) &&
             (!kIndex || (index_stride % 4 == 0 && host_aligned16(index)));
  if (vec) {
    int OW = W / pw, quads = pd.cols / 4, owq = OW / 4;
    long long total = (long long)pd.rows * quads;
    unsigned int grid = ceil_div_u(total, 256);
    FastDiv dq((uint32_t)quads), dw((uint32_t)owq);
#define KCNN_POOL_VEC(PW_)                                                                  \
  KCNN_LAUNCH((maxpool_prop_time_vec4<PW_, kIndex>), grid, 256, 0, st, src, sd.stride, pool, \
              pd.stride, index, index_stride, pd.rows, W, pc, quads, dq, dw)
    switch (pw) {
      case 1: KCNN_POOL_VEC(1); break;
      case 2: KCNN_POOL_VEC(2); break;
      case 3: KCNN_POOL_VEC(3); break;
      default: KCNN_POOL_VEC(4); break;
    }
#undef KCNN_POOL_VEC
    return;
  }
  PoolGeom g = make_geom(H, W, ph, pw, pc, pd.cols);
  long long total = (long long)pd.rows * pd.cols;
  KCNN_LAUNCH((maxpool_prop_scalar<kIndex>), ceil_div_u(total, 256), 256, 0, st, src, sd.stride,
              pool, pd.stride, index, index_stride, pd.rows, g, FastDiv((uint32_t)pd.cols));
}

}  // namespace kcnn

using namespace kcnn;

extern "C" {

void cudaF_maxpool_prop_s(cudaStream_t st, const float *src, MatrixDim sd, float *pool,
                          MatrixDim pd, int H, int W, int ph, int pw, int pc, int mode) {
  if (pd.rows == 0 || pd.cols == 0) return;
  if (mode == KCNN_POOL_PLAIN) {
    launch_prop_plain<false>(st, src, sd, pool, pd, nullptr, 0, H, W, ph, pw, pc);
    return;
  }
  int HW = H * W, OC = pd.cols / HW, o2 = 0, i2 = 0;
  if (mode == KCNN_POOL_OVERLAP2D) { o2 = (int)sqrt((double)OC); i2 = o2 + pc - 1; }
  long long total = (long long)pd.rows * pd.cols;
  unsigned int grid = ceil_div_u(total, 256);
  FastDiv dc((uint32_t)pd.cols), dh((uint32_t)HW), d2((uint32_t)(o2 > 0 ? o2 : 1));
  if (mode == KCNN_POOL_OVERLAP)
    KCNN_LAUNCH((maxpool_prop_overlap<KCNN_POOL_OVERLAP>), grid, 256, 0, st, src, sd.stride, pool,
                pd.stride, pd.rows, pd.cols, HW, pc, o2, i2, dc, dh, d2);
  else
    KCNN_LAUNCH((maxpool_prop_overlap<KCNN_POOL_OVERLAP2D>), grid, 256, 0, st, src, sd.stride,
                pool, pd.stride, pd.rows, pd.cols, HW, pc, o2, i2, dc, dh, d2);
}

void cudaF_maxpool_prop_index(cudaStream_t st, const float *src, MatrixDim sd, float *pool,
                              MatrixDim pd, unsigned char *index, int index_stride, int H, int W,
                              int ph, int pw, int pc) {
  launch_prop_plain<true>(st, src, sd, pool, pd, index, index_stride, H, W, ph, pw, pc);
}

void cudaF_maxpool_backprop_s(cudaStream_t st, const float *in_val, MatrixDim id,
                              const float *out_val, MatrixDim ovd, const float *out_deriv,
                              MatrixDim odd, float *dest, MatrixDim dd, int H, int W, int ph,
                              int pw, int pc, int mode, int zero_others) {
  if (odd.rows == 0 || odd.cols == 0) return;
  if (mode == KCNN_POOL_PLAIN) {
    bool vec = zero_others && time_vec_ok(H, W, ph, pw, in_val, id.stride, dest, dd.stride) &&
               ovd.stride % 4 == 0 && odd.stride % 4 == 0 && host_aligned16(out_val) &&
               host_aligned16(out_deriv);
    if (vec) {
      int OW = W / pw, quads = odd.cols / 4, owq = OW / 4;
      long long total = (long long)odd.rows * quads;
      unsigned int grid = ceil_div_u(total, 256);
      FastDiv dq((uint32_t)quads), dw((uint32_t)owq);
#define KCNN_POOLB_VEC(PW_)                                                                   \
  KCNN_LAUNCH((maxpool_backprop_time_vec4<PW_>), grid, 256, 0, st, in_val, id.stride, out_val, \
              ovd.stride, out_deriv, odd.stride, dest, dd.stride, odd.rows, W, pc, quads, dq, dw)
      switch (pw) {
        case 1: KCNN_POOLB_VEC(1); break;
        case 2: KCNN_POOLB_VEC(2); break;
        case 3: KCNN_POOLB_VEC(3); break;
        default: KCNN_POOLB_VEC(4); break;
      }
#undef KCNN_POOLB_VEC
      return;
    }
    PoolGeom g = make_geom(H, W, ph, pw, pc, odd.cols);
    long long total = (long long)odd.rows * odd.cols;
    unsigned int grid = ceil_div_u(total, 256);
    FastDiv dc((uint32_t)odd.cols);
    if (zero_others)
      KCNN_LAUNCH((maxpool_backprop_scalar<true>), grid, 256, 0, st, in_val, id.stride, out_val,
                  ovd.stride, out_deriv, odd.stride, dest, dd.stride, odd.rows, g, dc);
    else
      KCNN_LAUNCH((maxpool_backprop_scalar<false>), grid, 256, 0, st, in_val, id.stride, out_val,
                  ovd.stride, out_deriv, odd.stride, dest, dd.stride, odd.rows, g, dc);
    return;
  }
  int HW = H * W, OC = odd.cols / HW, o2 = 0, i2 = 1;
  if (mode == KCNN_POOL_OVERLAP2D) { o2 = (int)sqrt((double)OC); i2 = o2 + pc - 1; }
  long long total = (long long)id.rows * id.cols;
  unsigned int grid = ceil_div_u(total, 256);
  FastDiv dc((uint32_t)id.cols), dh((uint32_t)HW), d2((uint32_t)i2);
  if (mode == KCNN_POOL_OVERLAP)
    KCNN_LAUNCH((maxpool_backprop_overlap<KCNN_POOL_OVERLAP>), grid, 256, 0, st, in_val, id.stride,
                out_val, ovd.stride, out_deriv, odd.stride, dest, dd.stride, id.rows, id.cols, HW,
                pc, OC, o2, i2, dc, dh, d2);
  else
    KCNN_LAUNCH((maxpool_backprop_overlap<KCNN_POOL_OVERLAP2D>), grid, 256, 0, st, in_val,
                id.stride, out_val, ovd.stride, out_deriv, odd.stride, dest, dd.stride, id.rows,
                id.cols, HW, pc, OC, o2, i2, dc, dh, d2);
}

void cudaF_maxpool_backprop_index(cudaStream_t st, const unsigned char *index, int index_stride,
                                  const float *out_deriv, MatrixDim odd, float *dest, MatrixDim dd,
                                  int H, int W, int ph, int pw, int pc) {
  if (odd.rows == 0 || odd.cols == 0) return;
  bool vec = time_vec_ok(H, W, ph, pw, out_deriv, odd.stride, dest, dd.stride) &&
             index_stride % 4 == 0 && host_aligned16(index);
  if (vec) {
    int OW = W / pw, quads = odd.cols / 4, owq = OW / 4;
    long long total = (long long)odd.rows * quads;
    unsigned int grid = ceil_div_u(total, 256);
    FastDiv dq((uint32_t)quads), dw((uint32_t)owq);
#define KCNN_POOLI_VEC(PW_)                                                                  \
  KCNN_LAUNCH((maxpool_backprop_index_time_vec4<PW_>), grid, 256, 0, st, index, index_stride, \
              out_deriv, odd.stride, dest, dd.stride, odd.rows, W, pc, quads, dq, dw)
    switch (pw) {
      case 1: KCNN_POOLI_VEC(1); break;
      case 2: KCNN_POOLI_VEC(2); break;
      case 3: KCNN_POOLI_VEC(3); break;
      default: KCNN_POOLI_VEC(4); break;
    }
#undef KCNN_POOLI_VEC
    return;
  }
  PoolGeom g = make_geom(H, W, ph, pw, pc, odd.cols);
  long long total = (long long)odd.rows * odd.cols;
  KCNN_LAUNCH(maxpool_backprop_index_scalar, ceil_div_u(total, 256), 256, 0, st, index,
              index_stride, out_deriv, odd.stride, dest, dd.stride, odd.rows, g,
              FastDiv((uint32_t)odd.cols));
}

// ---- legacy launchers (Gr / Bl ignored) -----------------------------------

void cudaF_maxpool_prop(dim3, dim3, const float *src, MatrixDim sd, float *pool, MatrixDim pd,
                        int H, int W, int ph, int pw, int pc) {
  cudaF_maxpool_prop_s(g_legacy_stream, src, sd, pool, pd, H, W, ph, pw, pc, KCNN_POOL_PLAIN);
}
void cudaF_maxpool_backprop(dim3, dim3, const float *in_val, MatrixDim id, const float *out_val,
                            MatrixDim ovd, const float *out_deriv, MatrixDim odd, float *dest,
                            MatrixDim dd, int H, int W, int ph, int pw, int pc) {
  cudaF_maxpool_backprop_s(g_legacy_stream, in_val, id, out_val, ovd, out_deriv, odd, dest, dd, H,
                           W, ph, pw, pc, KCNN_POOL_PLAIN, 0);
}
void cudaF_maxpoolchannel_overlap_prop(dim3, dim3, const float *src, MatrixDim sd, float *pool,
                                       MatrixDim pd, int H, int W, int ph, int pw, int pc) {
  cudaF_maxpool_prop_s(g_legacy_stream, src, sd, pool, pd, H, W, ph, pw, pc, KCNN_POOL_OVERLAP);
}
void cudaF_maxpoolchannel_overlap_backprop(dim3, dim3, const float *in_val, MatrixDim id,
                                           const float *out_val, MatrixDim ovd,
                                           const float *out_deriv, MatrixDim odd, float *dest,
                                           MatrixDim dd, int H, int W, int ph, int pw, int pc) {
  cudaF_maxpool_backprop_s(g_legacy_stream, in_val, id, out_val, ovd, out_deriv, odd, dest, dd, H,
                           W, ph, pw, pc, KCNN_POOL_OVERLAP, 0);
}
void cudaF_maxpoolchannel_overlap2D_prop(dim3, dim3, const float *src, MatrixDim sd, float *pool,
                                         MatrixDim pd, int H, int W, int ph, int pw, int pc) {
  cudaF_maxpool_prop_s(g_legacy_stream, src, sd, pool, pd, H, W, ph, pw, pc, KCNN_POOL_OVERLAP2D);
}
void cudaF_maxpoolchannel_overlap2D_backprop(dim3, dim3, const float *in_val, MatrixDim id,
                                             const float *out_val, MatrixDim ovd,
                                             const float *out_deriv, MatrixDim odd, float *dest,
                                             MatrixDim dd, int H, int W, int ph, int pw, int pc) {
  cudaF_maxpool_backprop_s(g_legacy_stream, in_val, id, out_val, ovd, out_deriv, odd, dest, dd, H,
                           W, ph, pw, pc, KCNN_POOL_OVERLAP2D, 0);
}

}  // extern "C"
