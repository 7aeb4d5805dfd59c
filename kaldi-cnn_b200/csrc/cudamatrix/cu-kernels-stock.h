// cudamatrix/cu-kernels-stock.h -- shim: the handful of STOCK CuMatrix / CuVector
// element-wise operations the hot-path host code calls (Set, Scale, AddMat,
// CopyFromMat, CopyRowsFromVec, MulElements, AddRowSumMat, ...).  In a real Kaldi
// tree these are cudamatrix/cu-kernels.cu; here they are small grid-stride kernels.
#ifndef KALDI_CUDAMATRIX_CU_KERNELS_STOCK_H_
#define KALDI_CUDAMATRIX_CU_KERNELS_STOCK_H_
#include <cuda_runtime_api.h>
#include "cudamatrix/cu-matrixdim.h"
namespace kaldi {
namespace cu_stock {
void set_mat(cudaStream_t st, float *m, MatrixDim d, float value);
void scale_mat(cudaStream_t st, float *m, MatrixDim d, float alpha);
// dst += alpha * (trans ? src^T : src); src_dim describes src as stored.
void add_mat(cudaStream_t st, float *dst, MatrixDim d, float alpha, const float *src, MatrixDim sd, bool trans);
// dst = (trans ? src^T : src)
void copy_mat(cudaStream_t st, float *dst, MatrixDim d, const float *src, MatrixDim sd, bool trans);
void copy_rows_from_vec(cudaStream_t st, float *dst, MatrixDim d, const float *vec);
void mul_elements(cudaStream_t st, float *dst, MatrixDim d, const float *src, MatrixDim sd);
void max_elements(cudaStream_t st, float *dst, MatrixDim d, const float *src, MatrixDim sd);
void equal_mask(cudaStream_t st, const float *a, MatrixDim ad, const float *b, MatrixDim bd, float *mask, MatrixDim md);
void set_vec(cudaStream_t st, float *v, int dim, float value);
void add_const_vec(cudaStream_t st, float *v, int dim, float value);
void scale_vec(cudaStream_t st, float *v, int dim, float alpha);
// v = alpha * src + beta * v
void axpby_vec(cudaStream_t st, float *v, int dim, float alpha, const float *src, float beta);
// v[i] = m[i, col]
void copy_col(cudaStream_t st, float *v, const float *m, MatrixDim d, int col);
}  // namespace cu_stock
}  // namespace kaldi
#endif
