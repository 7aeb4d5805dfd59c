// kaldi-cnn_b200/csrc/cnslmat/gemm_operands.cuh
//
// Operand addressing of the implicit GEMMs.
//
// The reference materialises every layout change as a full copy before its
// SGEMM: PaddingZero, the im2col span (_span_row_to_convmat), FlipMat, TpBlock,
// TpInsideBlock, the transposed weights, and after it the col2im scatter
// (_convmat_to_out) and ModPermuteRow (cnslmat/conv2D.cc:100-185,
// nnet0/nnet-component-nnet0.cc:499-540, 745-765).  Here each of those is a
// piece of index algebra evaluated while a GEMM tile is loaded or stored:
//
//     element(A, m, k) = base[ M(m).off + K(k).off ]      if the tap is inside the image
//                      = 0                                 otherwise (zero padding)
//
// Every operand of every GEMM of the path is SEPARABLE like this: the GEMM-row
// index decodes to an offset plus a (w, h) coordinate, the GEMM-column index
// decodes to an offset plus a (dw, dh) displacement, and the element exists iff
// 0 <= w + dw < Wlim and 0 <= h + dh < Hlim.  A decoder is three integer
// divisions by launch-time constants (FastDiv) -- done once per tile row /
// column, not once per element.
//
// Tensor layout (cnsl-cu-kernels.cu:28-32): one matrix row per sample,
// [C][W][H] with H fastest; kernel rows [c][kw][kh] with kh fastest, columns g.

#ifndef KCNN_GEMM_OPERANDS_CUH_
#define KCNN_GEMM_OPERANDS_CUH_

#include "kcnn_common.cuh"

namespace kcnn {

// Decoded GEMM index: element offset contribution + window coordinate.
// Offsets are 32-bit element counts: the reference indexes with int32 everywhere
// (cnsl-cu-kernels.cu:32 "index = Ir*in_dim.stride + ..."), so rows*stride < 2^31 is
// already the contract; the launchers check it.
struct Ctx {
  int off;
  int dw, dh;
};

constexpr int kInvalidCoord = -(1 << 28);   // makes the window test fail

// idx -> (a, b, c) with idx = (a * nb + b) * nc + c ; off = a*sa + b*sb + c*sc + off0 ;
// dw = b*wb + w0 ; dh = c*hc + h0.  Covers:
//   sample/position rows   a = n, b = w-like, c = h-like
//   channel/tap columns    a = c, b = kw,     c = kh       (sign = -1 for the flipped taps of dgrad)
//   plain matrix rows      nb = nc = 1
struct Dec3 {
  FastDiv div_bc, div_c;       // by nb*nc and by nc
  int sa, sb, sc, off0;
  int wb, hc, w0, h0;
  int count;                   // number of valid indices
  __device__ __forceinline__ Ctx operator()(int idx) const {
    Ctx r;
    if (idx >= count) { r.off = 0; r.dw = kInvalidCoord; r.dh = 0; return r; }
    uint32_t a, bc, b, c;
    div_bc.divmod((uint32_t)idx, a, bc);
    div_c.divmod(bc, b, c);
    r.off = (int)a * sa + (int)b * sb + (int)c * sc + off0;
    r.dw = (int)b * wb + w0;
    r.dh = (int)c * hc + h0;
    return r;
  }
};

// idx -> (t, a) with idx = t * na + a (a fastest), t = b * nc + c.  Used for the
// K index of dgrad, ordered (kw, kh, g) with g fastest so the weight rows are read
// along their contiguous axis.
struct Dec3Inner {
  FastDiv div_a, div_c;
  int sa, sb, sc, off0;
  int wb, hc, w0, h0;
  int count;
  __device__ __forceinline__ Ctx operator()(int idx) const {
    Ctx r;
    if (idx >= count) { r.off = 0; r.dw = kInvalidCoord; r.dh = 0; return r; }
    uint32_t t, a, b, c;
    div_a.divmod((uint32_t)idx, t, a);
    div_c.divmod(t, b, c);
    r.off = (int)a * sa + (int)b * sb + (int)c * sc + off0;
    r.dw = (int)b * wb + w0;
    r.dh = (int)c * hc + h0;
    return r;
  }
};

inline Dec3 make_dec3(int na, int nb, int nc, long long sa, long long sb, long long sc,
                      long long off0, int wb, int hc, int w0, int h0) {
  Dec3 d;
  d.div_bc = FastDiv((uint32_t)(nb * nc));
  d.div_c = FastDiv((uint32_t)nc);
  d.sa = (int)sa; d.sb = (int)sb; d.sc = (int)sc; d.off0 = (int)off0;
  d.wb = wb; d.hc = hc; d.w0 = w0; d.h0 = h0;
  d.count = na * nb * nc;
  return d;
}

inline Dec3 make_linear(int count, long long stride) {
  return make_dec3(count, 1, 1, stride, 0, 0, 0, 0, 0, 0, 0);
}

inline Dec3Inner make_dec3_inner(int nb, int nc, int na, long long sa, long long sb, long long sc,
                                 long long off0, int wb, int hc, int w0, int h0) {
  Dec3Inner d;
  d.div_a = FastDiv((uint32_t)na);
  d.div_c = FastDiv((uint32_t)nc);
  d.sa = (int)sa; d.sb = (int)sb; d.sc = (int)sc; d.off0 = (int)off0;
  d.wb = wb; d.hc = hc; d.w0 = w0; d.h0 = h0;
  d.count = na * nb * nc;
  return d;
}

// One GEMM operand: base pointer, a decoder per axis, the window limits.
template <class DecMN, class DecK>
struct Operand {
  const float *base;
  DecMN mn;
  DecK k;
  int wlim, hlim;
  __device__ __forceinline__ float load(const Ctx &r, const Ctx &c) const {
    bool ok = (unsigned)(r.dw + c.dw) < (unsigned)wlim && (unsigned)(r.dh + c.dh) < (unsigned)hlim;
    return ok ? __ldg(base + r.off + c.off) : 0.0f;
  }
};

// Output side: C(m, n) lands at out[M(m).off + N(n).off]; bias (if any) is indexed
// by the GEMM column.
template <class DecM, class DecN>
struct OutputMap {
  float *base;
  DecM m;
  DecN n;
  const float *bias_n;    // added per GEMM column, or nullptr
  const float *bias_m;    // added per GEMM row, or nullptr
};

}  // namespace kcnn

#endif
