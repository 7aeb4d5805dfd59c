// cudamatrix/cu-matrix-lib.h -- shim umbrella.
#ifndef KALDI_CUDAMATRIX_CU_MATRIX_LIB_H_
#define KALDI_CUDAMATRIX_CU_MATRIX_LIB_H_
#include "cudamatrix/cu-vector.h"
#include "cudamatrix/cu-matrix.h"
#include "cudamatrix/cu-device.h"
#endif
