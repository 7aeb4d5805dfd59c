/*
 * include/cu-matrixdim.h
 *
 * The matrix descriptor every kernel launcher of the hot path takes BY VALUE.
 * It is Kaldi's upstream cudamatrix/cu-matrixdim.h type (the reference includes
 * it at cnslmat/cnsl-cu-kernels.h:13 but does not ship it): element (r, c) of a
 * matrix is data[r * stride + c], stride >= cols (pitched rows).
 */
#ifndef KCNN_CU_MATRIXDIM_H_
#define KCNN_CU_MATRIXDIM_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef int int32_cuda;

typedef struct MatrixDim_ {
  int32_cuda rows;
  int32_cuda cols;
  int32_cuda stride;
} MatrixDim;

#ifdef __cplusplus
}
#endif

/* Thread-block edge the reference's host code uses for every launch
 * (upstream cu-matrixdim.h; cnslmat/conv2D.cc:102).  The B200 launchers pick
 * their own geometry and ignore the Gr/Bl a legacy caller passes. */
#ifndef CU2DBLOCK
#define CU2DBLOCK 16
#endif
#ifndef CU1DBLOCK
#define CU1DBLOCK 256
#endif

#endif
