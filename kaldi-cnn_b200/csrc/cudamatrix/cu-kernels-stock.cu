// cudamatrix/cu-kernels-stock.cu -- shim (see cu-kernels-stock.h).
#include "cudamatrix/cu-kernels-stock.h"
#include "kcnn_common.cuh"

namespace kaldi {
namespace cu_stock {
using kcnn::FastDiv;
using kcnn::ceil_div_u;

template <class F>
__global__ void __launch_bounds__(256) map_mat_kernel(MatrixDim d, FastDiv div_cols, F f) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)d.rows * d.cols) return;
  uint32_t i, j;
  div_cols.divmod((uint32_t)t, i, j);
  f(i, j);
}

template <class F>
static void map_mat(cudaStream_t st, MatrixDim d, F f) {
  if (d.rows == 0 || d.cols == 0) return;
  KCNN_LAUNCH(map_mat_kernel<F>, ceil_div_u((long long)d.rows * d.cols, 256), 256, 0, st, d,
              FastDiv((uint32_t)d.cols), f);
}

template <class F>
__global__ void __launch_bounds__(256) map_vec_kernel(int dim, F f) {
  kcnn::pdl_prologue();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < dim) f(t);
}
template <class F>
static void map_vec(cudaStream_t st, int dim, F f) {
  if (dim == 0) return;
  KCNN_LAUNCH(map_vec_kernel<F>, ceil_div_u(dim, 256), 256, 0, st, dim, f);
}

void set_mat(cudaStream_t st, float *m, MatrixDim d, float value) {
  map_mat(st, d, [=] __device__(uint32_t i, uint32_t j) { m[(size_t)i * d.stride + j] = value; });
}
void scale_mat(cudaStream_t st, float *m, MatrixDim d, float alpha) {
  map_mat(st, d, [=] __device__(uint32_t i, uint32_t j) { m[(size_t)i * d.stride + j] *= alpha; });
}
void add_mat(cudaStream_t st, float *dst, MatrixDim d, float alpha, const float *src, MatrixDim sd, bool trans) {
  if (trans)
    map_mat(st, d, [=] __device__(uint32_t i, uint32_t j) {
      dst[(size_t)i * d.stride + j] = fmaf(alpha, src[(size_t)j * sd.stride + i], dst[(size_t)i * d.stride + j]);
    });
  else
    map_mat(st, d, [=] __device__(uint32_t i, uint32_t j) {
      dst[(size_t)i * d.stride + j] = fmaf(alpha, src[(size_t)i * sd.stride + j], dst[(size_t)i * d.stride + j]);
    });
}
void copy_mat(cudaStream_t st, float *dst, MatrixDim d, const float *src, MatrixDim sd, bool trans) {
  if (trans)
    map_mat(st, d, [=] __device__(uint32_t i, uint32_t j) { dst[(size_t)i * d.stride + j] = src[(size_t)j * sd.stride + i]; });
  else
    map_mat(st, d, [=] __device__(uint32_t i, uint32_t j) { dst[(size_t)i * d.stride + j] = src[(size_t)i * sd.stride + j]; });
}
void copy_rows_from_vec(cudaStream_t st, float *dst, MatrixDim d, const float *vec) {
  map_mat(st, d, [=] __device__(uint32_t i, uint32_t j) { dst[(size_t)i * d.stride + j] = vec[j]; });
}
void mul_elements(cudaStream_t st, float *dst, MatrixDim d, const float *src, MatrixDim sd) {
  map_mat(st, d, [=] __device__(uint32_t i, uint32_t j) { dst[(size_t)i * d.stride + j] *= src[(size_t)i * sd.stride + j]; });
}
void max_elements(cudaStream_t st, float *dst, MatrixDim d, const float *src, MatrixDim sd) {
  map_mat(st, d, [=] __device__(uint32_t i, uint32_t j) {
    float a = dst[(size_t)i * d.stride + j], b = src[(size_t)i * sd.stride + j];
    dst[(size_t)i * d.stride + j] = a < b ? b : a;
  });
}
void equal_mask(cudaStream_t st, const float *a, MatrixDim ad, const float *b, MatrixDim bd, float *mask, MatrixDim md) {
  map_mat(st, md, [=] __device__(uint32_t i, uint32_t j) {
    mask[(size_t)i * md.stride + j] = a[(size_t)i * ad.stride + j] == b[(size_t)i * bd.stride + j] ? 1.0f : 0.0f;
  });
}
void set_vec(cudaStream_t st, float *v, int dim, float value) {
  map_vec(st, dim, [=] __device__(int i) { v[i] = value; });
}
void add_const_vec(cudaStream_t st, float *v, int dim, float value) {
  map_vec(st, dim, [=] __device__(int i) { v[i] += value; });
}
void scale_vec(cudaStream_t st, float *v, int dim, float alpha) {
  map_vec(st, dim, [=] __device__(int i) { v[i] *= alpha; });
}
void axpby_vec(cudaStream_t st, float *v, int dim, float alpha, const float *src, float beta) {
  map_vec(st, dim, [=] __device__(int i) { v[i] = fmaf(alpha, src[i], beta * v[i]); });
}
void copy_col(cudaStream_t st, float *v, const float *m, MatrixDim d, int col) {
  map_vec(st, d.rows, [=] __device__(int i) { v[i] = m[(size_t)i * d.stride + col]; });
}

}  // namespace cu_stock
}  // namespace kaldi
