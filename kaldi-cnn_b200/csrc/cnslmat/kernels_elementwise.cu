// kaldi-cnn_b200/csrc/cnslmat/kernels_elementwise.cu
//
// Bandwidth-bound elementwise pieces of the training step:
//  * the momentum / weight-decay SGD of ConvolutionComponent::Update and
//    FullyConnectedComponent::UpdateSimple, four stock Kaldi passes
//    (Scale, AddMat, AddMat, AddMat; nnet0/nnet-component-nnet0.cc:769-772,
//    1139-1142) folded into one sweep: 3 reads + 2 writes = 20 B / parameter;
//  * the glue between hot-path layers (ReLU, softmax, cross-entropy).

#include <math.h>

#include "kcnn_common.cuh"

namespace kcnn {

__device__ __forceinline__ void sgd_one(float &w, float &p, float g, float momentum, float a_decay,
                                        float a_grad) {
  p = p * momentum;            // prev_grad_.Scale(momentum_)
  p = fmaf(a_decay, w, p);     // prev_grad_.AddMat(-lr*wd, linear_params_)
  p = fmaf(a_grad, g, p);      // prev_grad_.AddMat(lr, grad)
  w = w + p;                   // linear_params_.AddMat(1.0, prev_grad_)
}

template <bool kVec4>
__global__ void __launch_bounds__(256)
sgd_momentum_kernel(float *__restrict__ w, int w_stride, float *__restrict__ p, int p_stride,
                    const float *__restrict__ g, int g_stride, int rows, int cols, float momentum,
                    float a_decay, float a_grad, FastDiv div_units) {
  kcnn::pdl_prologue();
  const int units = kVec4 ? cols / 4 : cols;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * units) return;
  uint32_t i, u;
  div_units.divmod((uint32_t)t, i, u);
  if (kVec4) {
    float4 *wp = reinterpret_cast<float4 *>(w + (size_t)i * w_stride) + u;
    float4 *pp = reinterpret_cast<float4 *>(p + (size_t)i * p_stride) + u;
    float4 gv = __ldg(reinterpret_cast<const float4 *>(g + (size_t)i * g_stride) + u);
    float4 wv = *wp, pv = *pp;
    sgd_one(wv.x, pv.x, gv.x, momentum, a_decay, a_grad);
    sgd_one(wv.y, pv.y, gv.y, momentum, a_decay, a_grad);
    sgd_one(wv.z, pv.z, gv.z, momentum, a_decay, a_grad);
    sgd_one(wv.w, pv.w, gv.w, momentum, a_decay, a_grad);
    *wp = wv;
    *pp = pv;
  } else {
    float wv = w[(size_t)i * w_stride + u], pv = p[(size_t)i * p_stride + u];
    sgd_one(wv, pv, __ldg(g + (size_t)i * g_stride + u), momentum, a_decay, a_grad);
    w[(size_t)i * w_stride + u] = wv;
    p[(size_t)i * p_stride + u] = pv;
  }
}

__global__ void __launch_bounds__(256)
vec_axpy_kernel(float *__restrict__ v, const float *__restrict__ g, int dim, float alpha) {
  kcnn::pdl_prologue();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < dim) v[t] = fmaf(alpha, __ldg(g + t), v[t]);
}

template <bool kVec4>
__global__ void __launch_bounds__(256)
relu_fprop_kernel(const float *__restrict__ in, int in_stride, float *__restrict__ out,
                  int out_stride, int rows, int cols, FastDiv div_units) {
  kcnn::pdl_prologue();
  const int units = kVec4 ? cols / 4 : cols;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * units) return;
  uint32_t i, u;
  div_units.divmod((uint32_t)t, i, u);
  if (kVec4) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(in + (size_t)i * in_stride) + u);
    v.x = v.x > 0.0f ? v.x : 0.0f; v.y = v.y > 0.0f ? v.y : 0.0f;
    v.z = v.z > 0.0f ? v.z : 0.0f; v.w = v.w > 0.0f ? v.w : 0.0f;
    reinterpret_cast<float4 *>(out + (size_t)i * out_stride)[u] = v;
  } else {
    float v = __ldg(in + (size_t)i * in_stride + u);
    out[(size_t)i * out_stride + u] = v > 0.0f ? v : 0.0f;
  }
}

template <bool kVec4>
__global__ void __launch_bounds__(256)
relu_bprop_kernel(const float *__restrict__ ov, int ov_stride, const float *__restrict__ od,
                  int od_stride, float *__restrict__ id, int id_stride, int rows, int cols,
                  FastDiv div_units) {
  kcnn::pdl_prologue();
  const int units = kVec4 ? cols / 4 : cols;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * units) return;
  uint32_t i, u;
  div_units.divmod((uint32_t)t, i, u);
  if (kVec4) {
    float4 y = __ldg(reinterpret_cast<const float4 *>(ov + (size_t)i * ov_stride) + u);
    float4 d = __ldg(reinterpret_cast<const float4 *>(od + (size_t)i * od_stride) + u);
    d.x = y.x > 0.0f ? d.x : 0.0f; d.y = y.y > 0.0f ? d.y : 0.0f;
    d.z = y.z > 0.0f ? d.z : 0.0f; d.w = y.w > 0.0f ? d.w : 0.0f;
    reinterpret_cast<float4 *>(id + (size_t)i * id_stride)[u] = d;
  } else {
    float y = __ldg(ov + (size_t)i * ov_stride + u);
    id[(size_t)i * id_stride + u] = y > 0.0f ? __ldg(od + (size_t)i * od_stride + u) : 0.0f;
  }
}

__device__ __forceinline__ float block_reduce(float v, float *scratch, bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    float x = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, x) : v + x;
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
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

// One block per row: max, exp-sum, normalise, floor at 1e-20
// (SoftmaxComponent::Propagate, nnet2/nnet-component.cc:930-950).
// kInRegs: rows of up to 256 x kSoftmaxRegs columns are read ONCE and stay in registers between
// the three passes (all loads of a thread in flight together instead of three dependent sweeps
// over global memory); the arithmetic and its order are those of the generic form, so the two
// give identical bits.  Launched with 256 threads.
constexpr int kSoftmaxRegs = 16;
template <bool kInRegs>
__global__ void __launch_bounds__(256)
softmax_fprop_kernel(const float *__restrict__ in, int in_stride, float *__restrict__ out,
                     int out_stride, int cols) {
  kcnn::pdl_prologue();
  __shared__ float scratch[32];
  const float *x = in + (size_t)blockIdx.x * in_stride;
  float *y = out + (size_t)blockIdx.x * out_stride;
  if (kInRegs) {
    float v[kSoftmaxRegs];
#pragma unroll
    for (int k = 0; k < kSoftmaxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      v[k] = j < cols ? __ldg(x + j) : -INFINITY;
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < kSoftmaxRegs; k++) m = fmaxf(m, v[k]);
    m = block_reduce(m, scratch, true);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kSoftmaxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      if (j < cols) {
        v[k] = expf(v[k] - m);
        s += v[k];
      }
    }
    s = block_reduce(s, scratch, false);
    const float inv = 1.0f / s;
#pragma unroll
    for (int k = 0; k < kSoftmaxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      if (j < cols) {
        const float r = v[k] * inv;
        y[j] = r < 1e-20f ? 1e-20f : r;
      }
    }
    return;
  }
  float m = -INFINITY;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) m = fmaxf(m, __ldg(x + j));
  m = block_reduce(m, scratch, true);
  float s = 0.0f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    float e = expf(__ldg(x + j) - m);
    y[j] = e;
    s += e;
  }
  s = block_reduce(s, scratch, false);
  float inv = 1.0f / s;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    float v = y[j] * inv;
    y[j] = v < 1e-20f ? 1e-20f : v;
  }
}

// in_deriv = y * (d - dot(y, d))   (SoftmaxComponent::Backprop, :952-1000); kInRegs as above.
template <bool kInRegs>
__global__ void __launch_bounds__(256)
softmax_bprop_kernel(const float *__restrict__ ov, int ov_stride, const float *__restrict__ od,
                     int od_stride, float *__restrict__ id, int id_stride, int cols) {
  kcnn::pdl_prologue();
  __shared__ float scratch[32];
  const float *y = ov + (size_t)blockIdx.x * ov_stride;
  const float *d = od + (size_t)blockIdx.x * od_stride;
  float *o = id + (size_t)blockIdx.x * id_stride;
  if (kInRegs) {
    float yv[kSoftmaxRegs], dv[kSoftmaxRegs];
#pragma unroll
    for (int k = 0; k < kSoftmaxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      yv[k] = j < cols ? __ldg(y + j) : 0.0f;
      dv[k] = j < cols ? __ldg(d + j) : 0.0f;
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kSoftmaxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      if (j < cols) s = fmaf(yv[k], dv[k], s);
    }
    s = block_reduce(s, scratch, false);
#pragma unroll
    for (int k = 0; k < kSoftmaxRegs; k++) {
      const int j = (int)threadIdx.x + k * 256;
      if (j < cols) o[j] = yv[k] * (dv[k] - s);
    }
    return;
  }
  float s = 0.0f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) s = fmaf(__ldg(y + j), __ldg(d + j), s);
  s = block_reduce(s, scratch, false);
  for (int j = threadIdx.x; j < cols; j += blockDim.x) o[j] = __ldg(y + j) * (__ldg(d + j) - s);
}

__global__ void __launch_bounds__(256)
xent_deriv_kernel(const float *__restrict__ post, int post_stride, const int *__restrict__ labels,
                  float *__restrict__ deriv, int deriv_stride, int rows, int cols,
                  double *objf_accum, FastDiv div_cols) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * cols) return;
  uint32_t i, j;
  div_cols.divmod((uint32_t)t, i, j);
  float v = 0.0f;
  if ((int)j == __ldg(labels + i)) {
    float p = __ldg(post + (size_t)i * post_stride + j);
    v = 1.0f / p;
    if (objf_accum) atomicAdd(objf_accum, (double)logf(p));
  }
  deriv[(size_t)i * deriv_stride + j] = v;
}

// NormalizeComponent (upstream nnet2/nnet-component.cc:576-639): one block per row.
//   f = max(2^-66, |x|^2 / D)^-0.5 ; fprop: y = f x ; bprop: dx = f dy - (f == 2^33 ? 0 : f^3) / D (dy . x) x
template <bool kBackward>
__global__ void __launch_bounds__(256)
normalize_kernel(const float *__restrict__ in, int in_stride, const float *__restrict__ od, int od_stride,
                 float *__restrict__ out, int out_stride, int cols) {
  kcnn::pdl_prologue();
  __shared__ float scratch[32];
  const float *x = in + (size_t)blockIdx.x * in_stride;
  const float *d = kBackward ? od + (size_t)blockIdx.x * od_stride : nullptr;
  float *o = out + (size_t)blockIdx.x * out_stride;
  float ss = 0.0f, dot = 0.0f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    const float v = __ldg(x + j);
    ss = fmaf(v, v, ss);
    if (kBackward) dot = fmaf(__ldg(d + j), v, dot);
  }
  ss = block_reduce(ss, scratch, false);
  if (kBackward) dot = block_reduce(dot, scratch, false);
  const float floor_p = 1.3552527156068805e-20f;                     // 2^-66, kNormFloor
  float p = ss * (1.0f / (float)cols);
  if (p < floor_p) p = floor_p;
  const float f = 1.0f / sqrtf(p);
  if (!kBackward) {
    for (int j = threadIdx.x; j < cols; j += blockDim.x) o[j] = __ldg(x + j) * f;
    return;
  }
  const float g = (p == floor_p) ? 0.0f : f * f * f;                 // ReplaceValue(1 / sqrt(floor), 0), ApplyPow(3)
  const float c = (-1.0f / (float)cols) * (dot * g);
  for (int j = threadIdx.x; j < cols; j += blockDim.x) o[j] = fmaf(c, __ldg(x + j), f * __ldg(d + j));
}

static bool vec4_ok(int cols, std::initializer_list<int> strides,
                    std::initializer_list<const void *> ptrs) {
  if (cols % 4) return false;
  for (int s : strides) if (s % 4) return false;
  for (const void *p : ptrs) if (!host_aligned16(p)) return false;
  return true;
}

}  // namespace kcnn

using namespace kcnn;

extern "C" {

void cudaF_sgd_momentum_update(cudaStream_t st, float *w, MatrixDim wd, float *p, MatrixDim pd,
                               const float *g, MatrixDim gd, float momentum, float a_decay,
                               float a_grad) {
  if (wd.rows == 0 || wd.cols == 0) return;
  bool v4 = vec4_ok(wd.cols, {wd.stride, pd.stride, gd.stride}, {w, p, g});
  int units = v4 ? wd.cols / 4 : wd.cols;
  unsigned int grid = ceil_div_u((long long)wd.rows * units, 256);
  FastDiv du((uint32_t)units);
  if (v4)
    KCNN_LAUNCH(sgd_momentum_kernel<true>, grid, 256, 0, st, w, wd.stride, p, pd.stride, g,
                gd.stride, wd.rows, wd.cols, momentum, a_decay, a_grad, du);
  else
    KCNN_LAUNCH(sgd_momentum_kernel<false>, grid, 256, 0, st, w, wd.stride, p, pd.stride, g,
                gd.stride, wd.rows, wd.cols, momentum, a_decay, a_grad, du);
}

void cudaF_vec_axpy(cudaStream_t st, float *vec, const float *grad, int dim, float alpha) {
  if (dim == 0) return;
  KCNN_LAUNCH(vec_axpy_kernel, ceil_div_u(dim, 256), 256, 0, st, vec, grad, dim, alpha);
}

void cudaF_relu_fprop(cudaStream_t st, const float *in, MatrixDim id, float *out, MatrixDim od) {
  if (od.rows == 0 || od.cols == 0) return;
  bool v4 = vec4_ok(od.cols, {id.stride, od.stride}, {in, out});
  int units = v4 ? od.cols / 4 : od.cols;
  unsigned int grid = ceil_div_u((long long)od.rows * units, 256);
  FastDiv du((uint32_t)units);
  if (v4)
    KCNN_LAUNCH(relu_fprop_kernel<true>, grid, 256, 0, st, in, id.stride, out, od.stride, od.rows,
                od.cols, du);
  else
    KCNN_LAUNCH(relu_fprop_kernel<false>, grid, 256, 0, st, in, id.stride, out, od.stride, od.rows,
                od.cols, du);
}

void cudaF_relu_bprop(cudaStream_t st, const float *ov, MatrixDim ovd, const float *od,
                      MatrixDim odd, float *id, MatrixDim idd) {
  if (idd.rows == 0 || idd.cols == 0) return;
  bool v4 = vec4_ok(idd.cols, {ovd.stride, odd.stride, idd.stride}, {ov, od, id});
  int units = v4 ? idd.cols / 4 : idd.cols;
  unsigned int grid = ceil_div_u((long long)idd.rows * units, 256);
  FastDiv du((uint32_t)units);
  if (v4)
    KCNN_LAUNCH(relu_bprop_kernel<true>, grid, 256, 0, st, ov, ovd.stride, od, odd.stride, id,
                idd.stride, idd.rows, idd.cols, du);
  else
    KCNN_LAUNCH(relu_bprop_kernel<false>, grid, 256, 0, st, ov, ovd.stride, od, odd.stride, id,
                idd.stride, idd.rows, idd.cols, du);
}

void cudaF_softmax_fprop(cudaStream_t st, const float *in, MatrixDim id, float *out,
                         MatrixDim od) {
  if (od.rows == 0 || od.cols == 0) return;
  if (od.cols <= 256 * kSoftmaxRegs)
    KCNN_LAUNCH(softmax_fprop_kernel<true>, od.rows, 256, 0, st, in, id.stride, out, od.stride, od.cols);
  else
    KCNN_LAUNCH(softmax_fprop_kernel<false>, od.rows, 256, 0, st, in, id.stride, out, od.stride, od.cols);
}

void cudaF_softmax_bprop(cudaStream_t st, const float *ov, MatrixDim ovd, const float *od,
                         MatrixDim odd, float *id, MatrixDim idd) {
  if (idd.rows == 0 || idd.cols == 0) return;
  if (idd.cols <= 256 * kSoftmaxRegs)
    KCNN_LAUNCH(softmax_bprop_kernel<true>, idd.rows, 256, 0, st, ov, ovd.stride, od, odd.stride, id,
                idd.stride, idd.cols);
  else
    KCNN_LAUNCH(softmax_bprop_kernel<false>, idd.rows, 256, 0, st, ov, ovd.stride, od, odd.stride, id,
                idd.stride, idd.cols);
}

void cudaF_normalize_fprop(cudaStream_t st, const float *in, MatrixDim id, float *out, MatrixDim od) {
  if (od.rows == 0 || od.cols == 0) return;
  KCNN_LAUNCH(normalize_kernel<false>, od.rows, 256, 0, st, in, id.stride, nullptr, 0, out, od.stride, od.cols);
}

void cudaF_normalize_bprop(cudaStream_t st, const float *in_value, MatrixDim ivd, const float *out_deriv,
                           MatrixDim odd, float *in_deriv, MatrixDim idd) {
  if (idd.rows == 0 || idd.cols == 0) return;
  KCNN_LAUNCH(normalize_kernel<true>, idd.rows, 256, 0, st, in_value, ivd.stride, out_deriv, odd.stride, in_deriv,
              idd.stride, idd.cols);
}

void cudaF_xent_deriv(cudaStream_t st, const float *post, MatrixDim pd, const int *labels,
                      float *deriv, MatrixDim dd, double *objf_accum) {
  if (pd.rows == 0 || pd.cols == 0) return;
  unsigned int grid = ceil_div_u((long long)pd.rows * pd.cols, 256);
  KCNN_LAUNCH(xent_deriv_kernel, grid, 256, 0, st, post, pd.stride, labels, deriv, dd.stride,
              pd.rows, pd.cols, objf_accum, FastDiv((uint32_t)pd.cols));
}

}  // extern "C"
