// cudamatrix/cu-matrixdim.h -- shim: forwards to the public header.
#ifndef KALDI_CUDAMATRIX_CU_MATRIXDIM_H_
#define KALDI_CUDAMATRIX_CU_MATRIXDIM_H_
#include <cu-matrixdim.h>   // include/cu-matrixdim.h (MatrixDim, CU2DBLOCK); <> skips this directory
#endif
