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
//  * Intermap pooling (pool_channel_dim > 1) and 2-D planes take the STAGED path: the
//    channel stride H*W makes a warp of the direct kernels touch 32 different 128-byte
//    lines per load, so they are bound by L1 wavefronts (measured 38-75 % of the HBM
//    peak), not by HBM.  A pooled channel's window slab (pc*H*W floats) is contiguous in
//    the row, so the staged kernels move whole slabs with 1-D TMA bulk copies
//    (cp.async.bulk global -> shared, mbarrier complete_tx; backward also shared ->
//    global), 4 stages x 24 KB per CTA, 2 persistent CTAs per SM, and do the strided
//    window walk in shared memory.  Scan order, strict '<' and the sentinel are the same.

#include "kcnn_common.cuh"

#include <stdlib.h>

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

// ------------------------------------------------------------- staged path --

namespace staged {

constexpr int kStageFloats = 6144;          // 24 KB
constexpr int kStages = 4;
constexpr int kThreads = 1024;         // x 2 CTAs per SM: full occupancy for the shared-memory window walk
constexpr int kSmemBytes = kStages * kStageFloats * 4 + kStages * 8 + 128;

struct Geom {
  int H, W, ph, pw, pc, OH, OW;
  int HW, L, OHW;            // plane, slab = pc*HW floats, outputs per slab
  int OC;                    // slabs (pooled channels) per row
  int S;                     // slabs per chunk
  int chunks_per_row;
  FastDiv div_cpr, div_ohw, div_oh;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "POOL_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra POOL_WAIT_DONE;\n\t"
      "bra POOL_WAIT_LOOP;\n\t"
      "POOL_WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// kBackward = false: pool[j] = max over the window (optional arg-max byte).
// kBackward = true : the staged slab is rewritten in place -- err where the element equals the
//                    pooled value, 0 elsewhere (cnsl-cu-kernels.cu:302-303 + the caller's
//                    kSetZero) -- and goes back to dest with a bulk store.
// VW >= 1: H == 1 and pw == VW, the window row is one 32 / 64 / 128-bit shared-memory access;
// VW == 0: any plane, scalar walk c -> w -> h; PH > 0 fixes pool_height_dim at compile time so
// the innermost loop unrolls (the walk is instruction-bound, not bandwidth-bound).
template <bool kBackward, bool kIndex, int VW, int PH>
__global__ void __launch_bounds__(kThreads)
maxpool_staged_kernel(const float *__restrict__ src, int src_stride, float *__restrict__ pool, int pool_stride,
                      const float *__restrict__ out_deriv, int od_stride, float *__restrict__ dest,
                      int dest_stride, unsigned char *__restrict__ index, int index_stride, int rows, Geom g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *stage_f = reinterpret_cast<float *>(smem_raw);
  const uint32_t stage_u = smem_u32(smem_raw);
  const uint32_t bar_u = stage_u + kStages * kStageFloats * 4;
  const int tid = threadIdx.x;
  const int nchunks = rows * g.chunks_per_row;

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(bar_u + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  kcnn::pdl_prologue();

  auto issue = [&](int k) {                       // thread 0: fetch the k-th chunk of this CTA
    const long long chunk = (long long)blockIdx.x + (long long)k * gridDim.x;
    if (chunk >= nchunks) return;
    uint32_t i, cr;
    g.div_cpr.divmod((uint32_t)chunk, i, cr);
    const int s0 = cr * g.S, ns = min(g.S, g.OC - s0);
    const uint32_t bytes = (uint32_t)ns * g.L * 4u;
    const int st = k % kStages;
    mbar_expect_tx(bar_u + 8 * st, bytes);
    bulk_load(stage_u + st * kStageFloats * 4, src + (size_t)i * src_stride + (size_t)s0 * g.L, bytes,
              bar_u + 8 * st);
  };
  constexpr int kPrefetch = kBackward ? kStages - 1 : kStages;
  if (tid == 0)
    for (int k = 0; k < kPrefetch; k++) issue(k);

  for (int k = 0;; k++) {
    const long long chunk = (long long)blockIdx.x + (long long)k * gridDim.x;
    if (chunk >= nchunks) break;
    uint32_t i, cr;
    g.div_cpr.divmod((uint32_t)chunk, i, cr);
    const int s0 = cr * g.S, ns = min(g.S, g.OC - s0);
    const int st = k % kStages;
    float *sm = stage_f + st * kStageFloats;
    mbar_wait(bar_u + 8 * st, (uint32_t)(k / kStages) & 1u);

    const int nout = ns * g.OHW;
    const size_t j0 = (size_t)s0 * g.OHW;
    for (int lo = tid; lo < nout; lo += kThreads) {
      uint32_t ocl, pos, ow, oh;
      g.div_ohw.divmod((uint32_t)lo, ocl, pos);
      g.div_oh.divmod(pos, ow, oh);
      float *p = sm + ocl * g.L + ow * g.pw * g.H + oh * g.ph;
      if (!kBackward) {
        float val = -1e20f;
        int best = 0;
        if (VW == 0) {
          const int ph = PH > 0 ? PH : g.ph;
          int kk = 0;
          for (int c = 0; c < g.pc; c++)
            for (int w = 0; w < g.pw; w++) {
              const float *r = p + c * g.HW + w * g.H;
#pragma unroll
              for (int h = 0; h < ph; h++, kk++) {
                const float e = r[h];
                if (val < e) { val = e; if (kIndex) best = kk; }
              }
            }
        } else {
#pragma unroll 4
          for (int c = 0; c < g.pc; c++) {
            float e[VW > 0 ? VW : 1];
            if (VW == 1) {
              e[0] = p[c * g.HW];
            } else if (VW == 2) {
              const float2 v = *reinterpret_cast<const float2 *>(p + c * g.HW);
              e[0] = v.x; e[1 % (VW > 0 ? VW : 1)] = v.y;
            } else {
              const float4 v = *reinterpret_cast<const float4 *>(p + c * g.HW);
              e[0] = v.x; e[1 % (VW > 0 ? VW : 1)] = v.y; e[2 % (VW > 0 ? VW : 1)] = v.z; e[3 % (VW > 0 ? VW : 1)] = v.w;
            }
#pragma unroll
            for (int w = 0; w < (VW > 0 ? VW : 1); w++)
              if (val < e[w]) { val = e[w]; if (kIndex) best = c * VW + w; }
          }
        }
        pool[(size_t)i * pool_stride + j0 + lo] = val;
        if (kIndex) index[(size_t)i * index_stride + j0 + lo] = (unsigned char)best;
      } else {
        const float ov = __ldg(pool + (size_t)i * pool_stride + j0 + lo);
        const float err = __ldg(out_deriv + (size_t)i * od_stride + j0 + lo);
        if (VW == 0) {
          const int ph = PH > 0 ? PH : g.ph;
          for (int c = 0; c < g.pc; c++)
            for (int w = 0; w < g.pw; w++) {
              float *q = p + c * g.HW + w * g.H;
#pragma unroll
              for (int h = 0; h < ph; h++) q[h] = (ov == q[h]) ? err : 0.0f;
            }
        } else if (VW == 1) {
#pragma unroll 4
          for (int c = 0; c < g.pc; c++) {
            float *q = p + c * g.HW;
            *q = (ov == *q) ? err : 0.0f;
          }
        } else if (VW == 2) {
#pragma unroll 4
          for (int c = 0; c < g.pc; c++) {
            float2 *q = reinterpret_cast<float2 *>(p + c * g.HW);
            float2 v = *q;
            v.x = (ov == v.x) ? err : 0.0f; v.y = (ov == v.y) ? err : 0.0f;
            *q = v;
          }
        } else {
#pragma unroll 4
          for (int c = 0; c < g.pc; c++) {
            float4 *q = reinterpret_cast<float4 *>(p + c * g.HW);
            float4 v = *q;
            v.x = (ov == v.x) ? err : 0.0f; v.y = (ov == v.y) ? err : 0.0f;
            v.z = (ov == v.z) ? err : 0.0f; v.w = (ov == v.w) ? err : 0.0f;
            *q = v;
          }
        }
      }
    }
    if (kBackward) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic stores -> bulk-store reads
      __syncthreads();
      if (tid == 0) {
        bulk_store(dest + (size_t)i * dest_stride + (size_t)s0 * g.L, stage_u + st * kStageFloats * 4,
                   (uint32_t)ns * g.L * 4u);
        // the stage of chunk k-1 is free once ITS store has read it: all but the newest group
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        issue(k + kStages - 1);
      }
    } else {
      __syncthreads();                                                 // everyone is done reading the stage
      if (tid == 0) issue(k + kStages);
    }
  }
  if (kBackward && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Chunking of a row of OC slabs of L floats: every chunk starts and ends on a 16-byte boundary.
static bool make_geom(Geom *g, int H, int W, int ph, int pw, int pc, int in_cols, int out_cols) {
  if (H <= 0 || W <= 0 || ph <= 0 || pw <= 0 || pc <= 0) return false;
  g->H = H; g->W = W; g->ph = ph; g->pw = pw; g->pc = pc;
  g->OH = H / ph; g->OW = W / pw;
  g->HW = H * W; g->OHW = g->OH * g->OW;
  if (g->OHW <= 0) return false;
  const long long L = (long long)pc * g->HW;
  if (L > kStageFloats) return false;
  g->L = (int)L;
  if (out_cols % g->OHW != 0) return false;
  g->OC = out_cols / g->OHW;
  if (g->OC <= 0 || (long long)g->OC * L > in_cols) return false;
  if (((long long)g->OC * L) % 4 != 0) return false;              // the last chunk ends on 16 bytes
  const int m = (L % 4 == 0) ? 1 : (L % 2 == 0 ? 2 : 4);          // S*L % 4 == 0
  int S = (int)(kStageFloats / L) / m * m;
  if (S <= 0) return false;
  if (S >= g->OC) S = g->OC;
  g->S = S;
  g->chunks_per_row = (g->OC + S - 1) / S;
  g->div_cpr = FastDiv((uint32_t)g->chunks_per_row);
  g->div_ohw = FastDiv((uint32_t)g->OHW);
  g->div_oh = FastDiv((uint32_t)g->OH);
  return true;
}

static bool enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_POOL_STAGED");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <bool kBackward, bool kIndex, int VW, int PH>
static void launch_one(cudaStream_t st, const float *src, int src_stride, float *pool, int pool_stride,
                       const float *out_deriv, int od_stride, float *dest, int dest_stride, unsigned char *index,
                       int index_stride, int rows, const Geom &g) {
  auto kern = maxpool_staged_kernel<kBackward, kIndex, VW, PH>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    attr_set = true;
  }
  const long long nchunks = (long long)rows * g.chunks_per_row;
  const unsigned grid = (unsigned)(nchunks < 2 * kNumSMs ? nchunks : 2 * kNumSMs);
  KCNN_LAUNCH(kern, grid, kThreads, kSmemBytes, st, src, src_stride, pool, pool_stride, out_deriv, od_stride, dest,
              dest_stride, index, index_stride, rows, g);
}

template <bool kBackward, bool kIndex>
static void launch(cudaStream_t st, const float *src, int src_stride, float *pool, int pool_stride,
                   const float *out_deriv, int od_stride, float *dest, int dest_stride, unsigned char *index,
                   int index_stride, int rows, const Geom &g) {
  const int vw = (g.H == 1 && g.ph == 1 && (g.pw == 1 || g.pw == 2 || g.pw == 4) && g.W % g.pw == 0) ? g.pw : 0;
#define KCNN_POOL_STAGED(VW_, PH_)                                                                            \
  launch_one<kBackward, kIndex, VW_, PH_>(st, src, src_stride, pool, pool_stride, out_deriv, od_stride, dest, \
                                          dest_stride, index, index_stride, rows, g)
  if (vw == 4) KCNN_POOL_STAGED(4, 0);
  else if (vw == 2) KCNN_POOL_STAGED(2, 0);
  else if (vw == 1) KCNN_POOL_STAGED(1, 0);
  else if (g.ph == 1) KCNN_POOL_STAGED(0, 1);
  else if (g.ph == 2) KCNN_POOL_STAGED(0, 2);
  else if (g.ph == 3) KCNN_POOL_STAGED(0, 3);
  else if (g.ph == 4) KCNN_POOL_STAGED(0, 4);
  else KCNN_POOL_STAGED(0, 0);
#undef KCNN_POOL_STAGED
}

}  // namespace staged

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
  bool vec_first = vec && (pc == 1 || !staged::enabled());       // contiguous windows: already at the HBM peak
  if (!vec_first) {
    staged::Geom sg;
    if (staged::enabled() && sd.stride % 4 == 0 && host_aligned16(src) &&
        (long long)pd.rows * sd.stride < (1ll << 31) &&
        staged::make_geom(&sg, H, W, ph, pw, pc, sd.cols, pd.cols)) {
      staged::launch<false, kIndex>(st, src, sd.stride, pool, pd.stride, nullptr, 0, nullptr, 0, index,
                                    index_stride, pd.rows, sg);
      return;
    }
  }
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
    if (zero_others && !(vec && pc == 1) && staged::enabled()) {
      // every element of dest must lie in exactly one window: the slab is stored back whole
      staged::Geom sg;
      if (W % pw == 0 && H % ph == 0 && id.stride % 4 == 0 && dd.stride % 4 == 0 && host_aligned16(in_val) &&
          host_aligned16(dest) && (long long)id.rows * id.stride < (1ll << 31) &&
          (long long)dd.rows * dd.stride < (1ll << 31) &&
          staged::make_geom(&sg, H, W, ph, pw, pc, id.cols, odd.cols) &&
          (long long)sg.OC * sg.L == dd.cols && dd.cols == id.cols) {
        staged::launch<true, false>(st, in_val, id.stride, const_cast<float *>(out_val), ovd.stride, out_deriv,
                                    odd.stride, dest, dd.stride, nullptr, 0, odd.rows, sg);
        return;
      }
    }
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
