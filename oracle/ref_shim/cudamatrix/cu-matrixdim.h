/* oracle/ref_shim/cudamatrix/cu-matrixdim.h -- TEST INFRASTRUCTURE ONLY.
 * The reference (a patch on Kaldi r4510) includes Kaldi's cudamatrix/cu-matrixdim.h, which is
 * not part of /root/reference.  This restates the one type its kernels need, with the layout
 * SURVEY 8(b) records: MatrixDim = {int rows, cols, stride}, passed by value, element (r, c) at
 * data[r * stride + c]. */
#ifndef KCNN_ORACLE_REF_SHIM_CU_MATRIXDIM_H_
#define KCNN_ORACLE_REF_SHIM_CU_MATRIXDIM_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int32_t int32_cuda;
typedef struct MatrixDim_ {
  int32_cuda rows;
  int32_cuda cols;
  int32_cuda stride;
} MatrixDim;
#ifdef __cplusplus
}
#endif
#define CU1DBLOCK 256
#define CU2DBLOCK 16
#endif
