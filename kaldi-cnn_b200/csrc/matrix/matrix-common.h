// matrix/matrix-common.h -- shim: enums and index type of Kaldi's matrix library.
#ifndef KALDI_MATRIX_MATRIX_COMMON_H_
#define KALDI_MATRIX_MATRIX_COMMON_H_
#include "base/kaldi-common.h"
namespace kaldi {
typedef enum { kTrans = 112, kNoTrans = 111 } MatrixTransposeType;    // == CblasTrans / CblasNoTrans
typedef enum { kSetZero, kUndefined, kCopyData } MatrixResizeType;
typedef enum { kDefaultStride, kStrideEqualNumCols } MatrixStrideType;
typedef int32 MatrixIndexT;
typedef int32 SignedMatrixIndexT;
typedef uint32 UnsignedMatrixIndexT;
template <typename Real> class VectorBase;
template <typename Real> class Vector;
template <typename Real> class MatrixBase;
template <typename Real> class Matrix;
template <typename Real> class SubMatrix;
}  // namespace kaldi
#endif
