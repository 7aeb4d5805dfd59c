// kaldi-cnn_b200/csrc/cnslmat/kernels_fused.cu
//
// The bandwidth-bound kernels of NnetMinibatchUpdater's fused step (csrc/nnet2/nnet-fused.cc).
// In that step the time-axis layers keep activations and derivatives channels-last --
// [N][W][C], C fastest, the layout the convolution tensor maps read -- so none of the
// reference's copies (PaddingZero / TpBlock / TpInsideBlock / FlipMat, nnet0/nnet-component-
// nnet0.cc:423-446, 461-544, 738-777) and none of round 1's staging packs exist any more:
//
//   maxpool_prop_cl / maxpool_backprop_cl   MaxpoolComponent (cnsl-cu-kernels.cu:231-308) on
//       channels-last data, same comparison order (c -> w) and the same -1e20 sentinel, strict '<'
//       and '==' routing, so values, tie behaviour and signed zeros are those of the reference;
//       the ReLU that follows / precedes the pool in nnet.config rides along.
//   colsum_batch        every per-column sum of the step in ONE launch each way: bias gradients
//       (+ their SGD step) of all convolution / affine layers, and the NonlinearComponent
//       statistics (upstream nnet2/nnet-component.cc:337-363) of all ReLU / softmax layers;
//       two-stage and deterministic (the last block of a column group adds the partials in order).
//   softmax_xent        SoftmaxComponent::Propagate (:930-950) + the cross-entropy objective and
//       derivative + SoftmaxComponent::Backprop (:952-1000) on one row held in registers, with the
//       arithmetic (and therefore the bits) of the three separate kernels of kernels_elementwise.cu.

#include <math.h>

#include "kcnn_common.cuh"

namespace kcnn {

// ------------------------------------------------------------------ max-pool, channels-last --

// in [N][W][C] -> out [N][W/pw][C/pc]; window order c (outer) -> w, as _maxpool_prop with H = 1.
// ref_ld > 0: the outputs are written in the reference layout instead, rows [co][wo] with pitch ref_ld
// (what a following affine layer reads).
template <bool kRelu>
__global__ void __launch_bounds__(256)
maxpool_prop_cl_kernel(const float *__restrict__ in, long long total, int W, int C, int pw, int pc,
                       float *__restrict__ out, float *__restrict__ out_relu, FastDiv div_co, FastDiv div_wo,
                       int ref_ld) {
  kcnn::pdl_prologue();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  uint32_t rest, co, n, wo;
  div_co.divmod((uint32_t)t, rest, co);
  div_wo.divmod(rest, n, wo);
  const float *base = in + ((size_t)n * W + (size_t)wo * pw) * C + (size_t)co * pc;
  float val = -1e20f;
  for (int c = 0; c < pc; c++)
    for (int w = 0; w < pw; w++) {
      const float s = __ldg(base + (size_t)w * C + c);
      if (val < s) val = s;
    }
  const size_t o = ref_ld > 0 ? (size_t)n * ref_ld + (size_t)co * div_wo.d + wo : (size_t)t;
  out[o] = val;
  if (kRelu) out_relu[o] = val > 0.0f ? val : 0.0f;
}

// Gather form of _maxpool_backprop + the caller's zero fill (nnet0/nnet-component-nnet0.cc:889):
// in_deriv[i] = (in[i] == out[window(i)]) ? out_deriv[window(i)] : 0   -- every element equal to the
// maximum receives the derivative, as in the reference.  kRelu: then gated by [in[i] > 0], the
// backward pass of the ReLU that produced `in`.
template <bool kRelu>
__global__ void __launch_bounds__(256)
maxpool_backprop_cl_kernel(const float *__restrict__ in, const float *__restrict__ out,
                           const float *__restrict__ out_deriv, long long total, int W, int C, int pw, int pc,
                           float *__restrict__ in_deriv, FastDiv div_c, FastDiv div_w, FastDiv div_pw,
                           FastDiv div_pc, int ref_ld) {
  kcnn::pdl_prologue();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  uint32_t rest, c, n, w;
  div_c.divmod((uint32_t)t, rest, c);
  div_w.divmod(rest, n, w);
  const int WO = W / pw, CO = C / pc;
  const uint32_t wo = div_pw.div(w), co = div_pc.div(c);
  float d = 0.0f;
  const float x = __ldg(in + t);
  if ((int)wo < WO && (int)co < CO) {
    const size_t o = ((size_t)n * WO + wo) * CO + co;                   // out_deriv is channels-last
    const size_t ov = ref_ld > 0 ? (size_t)n * ref_ld + (size_t)co * WO + wo : o;
    if (__ldg(out + ov) == x) d = __ldg(out_deriv + o);
  }
  if (kRelu) d = x > 0.0f ? d : 0.0f;
  in_deriv[t] = d;
}

// out[n][c*W + w] = in[n][w][c]: a channels-last activation back in the reference layout (only for
// callers that look at an intermediate activation; not part of the step).
__global__ void __launch_bounds__(256)
cl_to_ref_kernel(const float *__restrict__ in, long long total, int W, int C, float *__restrict__ out, int ldo,
                 FastDiv div_w, FastDiv div_wc) {
  kcnn::pdl_prologue();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over the OUTPUT: (n, c, w)
  if (t >= total) return;
  uint32_t n, rest, c, w;
  div_wc.divmod((uint32_t)t, n, rest);
  div_w.divmod(rest, c, w);
  out[(size_t)n * ldo + rest] = __ldg(in + ((size_t)n * W + w) * C + c);
}

// -------------------------------------------------------------------------- batched colsum --

constexpr int kColsumMaxJobs = 12;
constexpr int kColsumRowsPerBlock = 64;     // 8 warps x 8 rows: one fully unrolled pass of colsum_rows

struct ColsumJobDev {
  const float *src;
  int rows, cols, ld, op, perm_w, perm_c, vec;
  void *dst0, *dst1;
  float alpha;
  int first_block, col_blocks, row_splits, rows_per;
  long long scratch_off;          // floats
  int counter_off;
};
struct ColsumBatchDev {
  ColsumJobDev job[kColsumMaxJobs];
  int njobs;
};

constexpr int kColsumCols = 128;          // columns per block: 32 lanes x 4

__device__ __forceinline__ void add4(float4 &a, const float4 &b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void cnt4(float4 &c, const float4 &v) {
  c.x += v.x > 0.0f ? 1.0f : 0.0f; c.y += v.y > 0.0f ? 1.0f : 0.0f;
  c.z += v.z > 0.0f ? 1.0f : 0.0f; c.w += v.w > 0.0f ? 1.0f : 0.0f;
}
// four adjacent columns of one row: one 128-bit load when the job allows it (16-byte aligned base and pitch;
// lanes past `cols` then read the row's padding, which is never stored), else guarded scalar loads
template <bool kVec>
__device__ __forceinline__ float4 colsum_ld(const float *p, int col, int cols) {
  if (kVec) return __ldg(reinterpret_cast<const float4 *>(p));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  v.x = __ldg(p);
  if (col + 1 < cols) v.y = __ldg(p + 1);
  if (col + 2 < cols) v.z = __ldg(p + 2);
  if (col + 3 < cols) v.w = __ldg(p + 3);
  return v;
}

// rows [r0, r1) of 4 columns, rows r0 + ty, r0 + ty + 8, ...: eight 128-bit loads in flight per thread, four
// independent accumulators combined in a fixed order (the result does not depend on timing)
template <bool kVec>
__device__ __forceinline__ void colsum_rows(const float *p, size_t ld, int col, int cols, int r0, int r1, int ty,
                                            bool stats2, float4 &sum, float4 &cnt) {
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0, c0 = s0, c1 = s0;
  int r = r0 + ty;
  for (; r + 56 < r1; r += 64) {
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = colsum_ld<kVec>(p + (size_t)(r + 8 * i) * ld, col, cols);
    add4(s0, v[0]); add4(s1, v[1]); add4(s2, v[2]); add4(s3, v[3]);
    add4(s0, v[4]); add4(s1, v[5]); add4(s2, v[6]); add4(s3, v[7]);
    if (stats2) {
#pragma unroll
      for (int i = 0; i < 8; i++) cnt4((i & 1) ? c1 : c0, v[i]);
    }
  }
  for (; r < r1; r += 8) {
    const float4 v = colsum_ld<kVec>(p + (size_t)r * ld, col, cols);
    add4(s0, v);
    if (stats2) cnt4(c0, v);
  }
  add4(s0, s1); add4(s2, s3); add4(s0, s2);
  add4(c0, c1);
  sum = s0; cnt = c0;
}

__global__ void __launch_bounds__(256)
colsum_batch_kernel(const ColsumBatchDev batch, float *__restrict__ scratch, unsigned int *__restrict__ counters) {
  kcnn::pdl_prologue();
  __shared__ float4 red0[8][33], red1[8][33];
  __shared__ int is_last;
  int ji = 0;
  while (ji + 1 < batch.njobs && (int)blockIdx.x >= batch.job[ji + 1].first_block) ji++;
  const ColsumJobDev &J = batch.job[ji];
  const int local = (int)blockIdx.x - J.first_block;
  const int cb = local % J.col_blocks, rs = local / J.col_blocks;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = cb * kColsumCols + tx * 4;
  const bool stats2 = J.op == KCNN_COLSUM_STATS_RELU;
  const int r0 = rs * J.rows_per, r1 = min(J.rows, r0 + J.rows_per);
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f), cnt = sum;
  if (col < J.cols) {
    if (J.vec) colsum_rows<true>(J.src + col, (size_t)J.ld, col, J.cols, r0, r1, ty, stats2, sum, cnt);
    else       colsum_rows<false>(J.src + col, (size_t)J.ld, col, J.cols, r0, r1, ty, stats2, sum, cnt);
  }
  red0[ty][tx] = sum;
  red1[ty][tx] = cnt;
  __syncthreads();
  float *part = scratch + J.scratch_off;
  const size_t second = (size_t)J.row_splits * J.cols;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (ty == 0 && col < J.cols) {
    float4 sa = red0[0][tx], sb = red1[0][tx];
#pragma unroll
    for (int i = 1; i < 8; i++) { add4(sa, red0[i][tx]); add4(sb, red1[i][tx]); }
    a[0] = sa.x; a[1] = sa.y; a[2] = sa.z; a[3] = sa.w;
    b[0] = sb.x; b[1] = sb.y; b[2] = sb.z; b[3] = sb.w;
    if (J.row_splits > 1) {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (col + k < J.cols) {
          part[(size_t)rs * J.cols + col + k] = a[k];
          if (stats2) part[second + (size_t)rs * J.cols + col + k] = b[k];
        }
    }
  }
  if (J.row_splits > 1) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(counters + J.counter_off + cb, 1u);
      is_last = prev == (unsigned int)(J.row_splits - 1);
      if (is_last) counters[J.counter_off + cb] = 0u;         // ready for the next launch / graph replay
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
  }
  if (J.row_splits > 1) {
    // second stage, by the block that arrived last: warp ty sums splits ty, ty + 8, ... of its 4 columns
    // (loads of different splits are independent: four in flight per column), then the 8 warps' sums
    // are combined in warp order -- a fixed order, like the first stage
    float4 fa = make_float4(0.f, 0.f, 0.f, 0.f), fb = fa;
    if (col < J.cols) {
      float pa[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f};
      for (int z0 = ty; z0 < J.row_splits; z0 += 32) {
        float va[4][4], vb[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int z = z0 + 8 * u;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const bool ok = z < J.row_splits && col + k < J.cols;
            va[u][k] = ok ? __ldcg(part + (size_t)z * J.cols + col + k) : 0.f;
            vb[u][k] = (ok && stats2) ? __ldcg(part + second + (size_t)z * J.cols + col + k) : 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int k = 0; k < 4; k++) { pa[k] += va[u][k]; pb[k] += vb[u][k]; }
      }
      fa = make_float4(pa[0], pa[1], pa[2], pa[3]);
      fb = make_float4(pb[0], pb[1], pb[2], pb[3]);
    }
    __syncthreads();                       // red0 / red1 of the first stage have been consumed
    red0[ty][tx] = fa;
    red1[ty][tx] = fb;
    __syncthreads();
    if (ty != 0 || col >= J.cols) return;
    float4 sa = red0[0][tx], sb = red1[0][tx];
#pragma unroll
    for (int i = 1; i < 8; i++) { add4(sa, red0[i][tx]); add4(sb, red1[i][tx]); }
    a[0] = sa.x; a[1] = sa.y; a[2] = sa.z; a[3] = sa.w;
    b[0] = sb.x; b[1] = sb.y; b[2] = sb.z; b[3] = sb.w;
  }
  if (ty != 0 || col >= J.cols) return;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int c = col + k;
    if (c >= J.cols) break;
    int oc = c;
    if (J.perm_w > 0) oc = (c % J.perm_c) * J.perm_w + c / J.perm_c;     // channels-last column -> [c][w] index
    switch (J.op) {
      case KCNN_COLSUM_STORE: static_cast<float *>(J.dst0)[oc] = a[k]; break;
      case KCNN_COLSUM_AXPY: {
        float *d = static_cast<float *>(J.dst0);
        d[oc] = fmaf(J.alpha, a[k], d[oc]);
        break;
      }
      case KCNN_COLSUM_STATS_RELU:
        static_cast<double *>(J.dst0)[oc] += (double)a[k];
        static_cast<double *>(J.dst1)[oc] += (double)b[k];
        break;
      default: static_cast<double *>(J.dst0)[oc] += (double)a[k]; break;      // KCNN_COLSUM_STATS_VALUE
    }
  }
}

static void colsum_layout(const KcnnColsumJob &j, int &col_blocks, int &row_splits, int &rows_per) {
  col_blocks = (j.cols + kColsumCols - 1) / kColsumCols;
  row_splits = (j.rows + kColsumRowsPerBlock - 1) / kColsumRowsPerBlock;
  if (row_splits > 64) row_splits = 64;
  if (row_splits < 1) row_splits = 1;
  rows_per = (j.rows + row_splits - 1) / row_splits;
  row_splits = rows_per > 0 ? (j.rows + rows_per - 1) / rows_per : 1;
  if (row_splits < 1) row_splits = 1;
}

// ------------------------------------------------------------------- softmax + cross-entropy --

__device__ __forceinline__ float block_reduce256(float v, float *scratch, bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    float x = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, x) : v + x;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  v = lane < nw ? scratch[lane] : (is_max ? -INFINITY : 0.0f);
  for (int o = 16; o > 0; o >>= 1) {
    float x = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, x) : v + x;
  }
  return v;
}

struct SeedList {
  unsigned long long *p[4];
  int n;
};

// One block per row, the row (<= 256 * 16 columns) in registers.
//   kFromLogits: post = softmax(logits), floored at 1e-20          (softmax_fprop_kernel<true>)
//   deriv wrt the posteriors: 1 / p at the label, 0 elsewhere; objf += log p   (xent_deriv_kernel)
//   d_logits = post * (deriv - dot(post, deriv))                               (softmax_bprop_kernel<true>)
// Same operations in the same order as those three kernels: dot(post, deriv) has a single non-zero
// term fl(p * fl(1/p)), all other addends are exact zeros.
// Block 0 also advances the dropout seeds of the step (the forward pass that used them is done).
constexpr int kSxRegs = 16;
template <bool kFromLogits>
__global__ void __launch_bounds__(256)
softmax_xent_kernel(const float *__restrict__ logits, int ld_l, float *__restrict__ post, int ld_p,
                    const int *__restrict__ labels, float *__restrict__ d_logits, int ld_d, int cols,
                    double *objf_accum, SeedList seeds) {
  kcnn::pdl_prologue();
  __shared__ float scratch[32];
  const int row = blockIdx.x;
  if (row == 0 && threadIdx.x == 0)
    for (int i = 0; i < seeds.n; i++) *seeds.p[i] += 1ull;
  float *y = post + (size_t)row * ld_p;
  float v[kSxRegs];
  if (kFromLogits) {
    const float *x = logits + (size_t)row * ld_l;
#pragma unroll
    for (int k = 0; k < kSxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      v[k] = j < cols ? __ldg(x + j) : -INFINITY;
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < kSxRegs; k++) m = fmaxf(m, v[k]);
    m = block_reduce256(m, scratch, true);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kSxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      if (j < cols) {
        v[k] = expf(v[k] - m);
        s += v[k];
      }
    }
    s = block_reduce256(s, scratch, false);
    const float inv = 1.0f / s;
#pragma unroll
    for (int k = 0; k < kSxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      if (j < cols) {
        const float r = v[k] * inv;
        v[k] = r < 1e-20f ? 1e-20f : r;
        y[j] = v[k];
      } else {
        v[k] = 0.0f;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < kSxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      v[k] = j < cols ? y[j] : 0.0f;
    }
  }
  const int label = __ldg(labels + row);
  float pinv = 0.0f, dot = 0.0f;
#pragma unroll
  for (int k = 0; k < kSxRegs; k++) {
    const int j = (int)threadIdx.x + k * 256;
    if (j < cols && j == label) {
      pinv = 1.0f / v[k];
      dot = fmaf(v[k], pinv, dot);
      if (objf_accum) atomicAdd(objf_accum, (double)logf(v[k]));
    }
  }
  dot = block_reduce256(dot, scratch, false);
  float *o = d_logits + (size_t)row * ld_d;
#pragma unroll
  for (int k = 0; k < kSxRegs; k++) {
    const int j = (int)threadIdx.x + k * 256;
    if (j < cols) o[j] = v[k] * ((j == label ? pinv : 0.0f) - dot);
  }
}

__global__ void bump_seeds_kernel(SeedList seeds) {
  kcnn::pdl_prologue();
  for (int i = 0; i < seeds.n; i++) *seeds.p[i] += 1ull;
}

}  // namespace kcnn

using namespace kcnn;

extern "C" {

void cudaF_maxpool_prop_cl(cudaStream_t st, const float *in, int N, int W, int C, int pw, int pc, float *out,
                           float *out_relu, int ref_ld) {
  const int WO = W / pw, CO = C / pc;
  const long long total = (long long)N * WO * CO;
  if (total == 0) return;
  const unsigned grid = ceil_div_u(total, 256);
  if (out_relu)
    KCNN_LAUNCH(maxpool_prop_cl_kernel<true>, grid, 256, 0, st, in, total, W, C, pw, pc, out, out_relu,
                FastDiv((uint32_t)CO), FastDiv((uint32_t)WO), ref_ld);
  else
    KCNN_LAUNCH(maxpool_prop_cl_kernel<false>, grid, 256, 0, st, in, total, W, C, pw, pc, out, out_relu,
                FastDiv((uint32_t)CO), FastDiv((uint32_t)WO), ref_ld);
}

void cudaF_maxpool_backprop_cl(cudaStream_t st, const float *in, const float *out, int ref_ld,
                               const float *out_deriv, int N, int W, int C, int pw, int pc, float *in_deriv,
                               int relu_gate) {
  const long long total = (long long)N * W * C;
  if (total == 0) return;
  const unsigned grid = ceil_div_u(total, 256);
  if (relu_gate)
    KCNN_LAUNCH(maxpool_backprop_cl_kernel<true>, grid, 256, 0, st, in, out, out_deriv, total, W, C, pw, pc, in_deriv,
                FastDiv((uint32_t)C), FastDiv((uint32_t)W), FastDiv((uint32_t)pw), FastDiv((uint32_t)pc), ref_ld);
  else
    KCNN_LAUNCH(maxpool_backprop_cl_kernel<false>, grid, 256, 0, st, in, out, out_deriv, total, W, C, pw, pc, in_deriv,
                FastDiv((uint32_t)C), FastDiv((uint32_t)W), FastDiv((uint32_t)pw), FastDiv((uint32_t)pc), ref_ld);
}

void cudaF_cl_to_ref(cudaStream_t st, const float *in, int N, int W, int C, float *out, int ldo) {
  const long long total = (long long)N * W * C;
  if (total == 0) return;
  KCNN_LAUNCH(cl_to_ref_kernel, ceil_div_u(total, 256), 256, 0, st, in, total, W, C, out, ldo, FastDiv((uint32_t)W),
              FastDiv((uint32_t)(W * C)));
}

size_t kcnn_colsum_batch_scratch_bytes(const KcnnColsumJob *jobs, int njobs) {
  size_t floats = 0, counters = 0;
  for (int i = 0; i < njobs; i++) {
    int cb, rs, rp;
    colsum_layout(jobs[i], cb, rs, rp);
    floats += (size_t)rs * jobs[i].cols * (jobs[i].op == KCNN_COLSUM_STATS_RELU ? 2 : 1);
    counters += (size_t)cb;
  }
  return (floats + counters) * 4 + 16;
}

void cudaF_colsum_batch(cudaStream_t st, const KcnnColsumJob *jobs, int njobs, void *scratch) {
  // layout of `scratch` (zero-filled once by the caller): all counters first, then the partial sums
  size_t counters = 0;
  for (int i = 0; i < njobs; i++) counters += (size_t)((jobs[i].cols + kColsumCols - 1) / kColsumCols);
  unsigned int *cnt = static_cast<unsigned int *>(scratch);
  float *part = reinterpret_cast<float *>(cnt + ((counters + 3) & ~(size_t)3));
  size_t coff = 0;
  long long soff = 0;
  for (int base = 0; base < njobs; base += kColsumMaxJobs) {
    ColsumBatchDev b;
    b.njobs = 0;
    int blocks = 0;
    for (int i = base; i < njobs && i < base + kColsumMaxJobs; i++) {
      const KcnnColsumJob &j = jobs[i];
      int cb, rs, rp;
      colsum_layout(j, cb, rs, rp);
      if (j.rows > 0 && j.cols > 0) {
        ColsumJobDev &d = b.job[b.njobs++];
        d.src = j.src; d.rows = j.rows; d.cols = j.cols; d.ld = j.ld; d.op = j.op; d.perm_w = j.perm_w;
        d.perm_c = j.perm_c; d.dst0 = j.dst0; d.dst1 = j.dst1; d.alpha = j.alpha;
        d.vec = ((reinterpret_cast<uintptr_t>(j.src) & 15u) == 0 && j.ld % 4 == 0) ? 1 : 0;
        d.first_block = blocks; d.col_blocks = cb; d.row_splits = rs; d.rows_per = rp;
        d.scratch_off = soff; d.counter_off = (int)coff;
        blocks += cb * rs;
      }
      soff += (long long)rs * j.cols * (j.op == KCNN_COLSUM_STATS_RELU ? 2 : 1);
      coff += (size_t)cb;
    }
    if (b.njobs > 0) KCNN_LAUNCH(colsum_batch_kernel, blocks, 256, 0, st, b, part, cnt);
  }
}

int cudaF_softmax_xent(cudaStream_t st, const float *logits, MatrixDim ld, float *post, MatrixDim pd,
                       const int *labels, float *d_logits, MatrixDim dd, double *objf_accum,
                       unsigned long long *const *seeds, int num_seeds) {
  if (pd.cols > 256 * kSxRegs || num_seeds > 4) return 0;
  if (pd.rows == 0 || pd.cols == 0) return 1;
  SeedList sl;
  sl.n = num_seeds;
  for (int i = 0; i < 4; i++) sl.p[i] = i < num_seeds ? seeds[i] : nullptr;
  if (logits)
    KCNN_LAUNCH(softmax_xent_kernel<true>, pd.rows, 256, 0, st, logits, ld.stride, post, pd.stride, labels, d_logits,
                dd.stride, pd.cols, objf_accum, sl);
  else
    KCNN_LAUNCH(softmax_xent_kernel<false>, pd.rows, 256, 0, st, logits, 0, post, pd.stride, labels, d_logits,
                dd.stride, pd.cols, objf_accum, sl);
  return 1;
}

void cudaF_bump_seeds(cudaStream_t st, unsigned long long *const *seeds, int num_seeds) {
  for (int base = 0; base < num_seeds; base += 4) {
    SeedList sl;
    sl.n = num_seeds - base < 4 ? num_seeds - base : 4;
    for (int i = 0; i < 4; i++) sl.p[i] = i < sl.n ? seeds[base + i] : nullptr;
    KCNN_LAUNCH(bump_seeds_kernel, 1, 1, 0, st, sl);
  }
}

}  // extern "C"
