// cudamatrix/cu-matrix.h -- shim: CuMatrixBase / CuMatrix / CuSubMatrix on device
// memory, with the TEN CNN-layer member functions of the reference declared with
// the reference's own signatures (src/cudamatrix/cu-matrix.h:446-482, the
// "// added by hwaran" block).  Their definitions are in cnslmat/conv2D.cc.
#ifndef KALDI_CUDAMATRIX_CU_MATRIX_H_
#define KALDI_CUDAMATRIX_CU_MATRIX_H_

#include "matrix/matrix-lib.h"
#include "cudamatrix/cu-common.h"
#include "cudamatrix/cu-device.h"
#include "cudamatrix/cu-vector.h"

namespace kaldi {

template <typename Real> class CuMatrix;
template <typename Real> class CuSubMatrix;

template <typename Real>
class CuMatrixBase {
 public:
  friend class CuVectorBase<Real>;
  friend class CuSubMatrix<Real>;

  MatrixIndexT NumRows() const { return num_rows_; }
  MatrixIndexT NumCols() const { return num_cols_; }
  MatrixIndexT Stride() const { return stride_; }
  ::MatrixDim Dim() const {
    ::MatrixDim d = {num_rows_, num_cols_, stride_};
    return d;
  }
  inline const Real *Data() const { return data_; }
  inline Real *Data() { return data_; }
  inline const Real *RowData(MatrixIndexT r) const { return data_ + (size_t)r * stride_; }

  // ---- stock operations used by the hot-path host code ----------------------
  void SetZero();
  void Set(Real value);
  void Scale(Real value);
  void SetRandn();
  /// *this += alpha * A  (A transposed if transA == kTrans)
  void AddMat(Real alpha, const CuMatrixBase<Real> &A, MatrixTransposeType transA = kNoTrans);
  /// *this = alpha * op(A) * op(B) + beta * *this
  void AddMatMat(Real alpha, const CuMatrixBase<Real> &A, MatrixTransposeType transA,
                 const CuMatrixBase<Real> &B, MatrixTransposeType transB, Real beta);
  void CopyFromMat(const CuMatrixBase<Real> &src, MatrixTransposeType trans = kNoTrans);
  void CopyFromMat(const MatrixBase<Real> &src, MatrixTransposeType trans = kNoTrans);
  void CopyToMat(MatrixBase<Real> *dst) const;
  void CopyRowsFromVec(const CuVectorBase<Real> &v);
  void MulElements(const CuMatrixBase<Real> &A);
  void Max(const CuMatrixBase<Real> &A);
  void EqualElementMask(const CuMatrixBase<Real> &mat, CuMatrix<Real> *mask) const;
  Real Sum() const;
  /// Single element read (device -> host copy; debugging only).
  Real operator()(MatrixIndexT r, MatrixIndexT c) const;

  inline CuSubMatrix<Real> Range(const MatrixIndexT row_offset, const MatrixIndexT num_rows,
                                 const MatrixIndexT col_offset, const MatrixIndexT num_cols) const {
    return CuSubMatrix<Real>(*this, row_offset, num_rows, col_offset, num_cols);
  }
  inline CuSubMatrix<Real> RowRange(const MatrixIndexT row_offset, const MatrixIndexT num_rows) const {
    return CuSubMatrix<Real>(*this, row_offset, num_rows, 0, num_cols_);
  }
  inline CuSubMatrix<Real> ColRange(const MatrixIndexT col_offset, const MatrixIndexT num_cols) const {
    return CuSubMatrix<Real>(*this, 0, num_rows_, col_offset, num_cols);
  }

  // ---- added by hwaran (reference cudamatrix/cu-matrix.h:446-482) ------------

  // Convolution 'this' with kernel => out
  // this matrix : row = num_chunks, col=in_height * in_width * in_channel
  void Conv2D(const CuMatrixBase<Real> &kernel,
              int32 in_height,
              int32 in_width,
              int32 in_channel,
              int32 kernel_height,
              int32 kernel_width,
              int32 group,
              CuMatrixBase<Real> *out,
              bool concat) const;

  // if vec = [1 2 3] and rep = 2 => vec2 = [ 1 1 2 2 3 3];
  // this = this + repmat(vec2, NumRows(), 1);
  void AddMatRepVec(const CuVectorBase<Real> &vec, int32 rep) const;

  // this [(kernel_height*kernel_width*in_channel) x group]
  // flip [(kernel_height*kernel_width*group) x in_channel]
  void FlipMat(int32 kernel_height, int32 kernel_width, int32 in_channel, int32 group,
               CuMatrix<Real> *flip) const;

  // zero border of (kernel_height-1) rows / (kernel_width-1) columns on every side
  void PaddingZero(int32 orig_height, int32 orig_width, int32 orig_channel, int32 kernel_height,
                   int32 kernel_width, CuMatrix<Real> *padmat) const;

  void TpBlock(int32 in_channel, int32 block_size, CuMatrix<Real> *out) const;

  void TpInsideBlock(int32 group, int32 block_size, CuMatrix<Real> *out) const;

  void ModPermuteRow(int32 in_channel, int32 block_size, CuMatrix<Real> *out) const;

  void Maxpool_prop(int32 in_height, int32 in_width, int32 pool_height_dim, int32 pool_width_dim,
                    int32 pool_channel_dim, bool overlap, bool overlap2D,
                    CuMatrixBase<Real> *out) const;
  void Maxpool_backprop(const CuMatrixBase<Real> &out_value, const CuMatrixBase<Real> &out_deriv,
                        CuMatrix<Real> *in_deriv, int32 in_height, int32 in_width,
                        int32 pool_height_dim, int32 pool_width_dim, int32 pool_channel_dim,
                        bool overlap, bool overlap2D) const;
  void ModPermuteChannel(int32 comp_idx, int32 num_component, int32 in_height, int32 in_width,
                         CuMatrixBase<Real> *container, bool fromCompToContainer);

 protected:
  CuMatrixBase() : data_(NULL), num_cols_(0), num_rows_(0), stride_(0) {}
  CuMatrixBase(Real *data, MatrixIndexT num_rows, MatrixIndexT num_cols, MatrixIndexT stride)
      : data_(data), num_cols_(num_cols), num_rows_(num_rows), stride_(stride) {}

  Real *data_;
  MatrixIndexT num_cols_;
  MatrixIndexT num_rows_;
  MatrixIndexT stride_;

 private:
  KALDI_DISALLOW_COPY_AND_ASSIGN(CuMatrixBase);
};

template <typename Real>
class CuMatrix : public CuMatrixBase<Real> {
 public:
  CuMatrix() : owns_(true) {}
  CuMatrix(MatrixIndexT rows, MatrixIndexT cols, MatrixResizeType resize_type = kSetZero)
      : owns_(true) {
    Resize(rows, cols, resize_type);
  }
  CuMatrix(const CuMatrix<Real> &other, MatrixTransposeType trans = kNoTrans);
  explicit CuMatrix(const CuMatrixBase<Real> &other, MatrixTransposeType trans = kNoTrans);
  explicit CuMatrix(const MatrixBase<Real> &other, MatrixTransposeType trans = kNoTrans);
  ~CuMatrix() { Destroy(); }

  CuMatrix<Real> &operator=(const CuMatrixBase<Real> &other) {
    this->Resize(other.NumRows(), other.NumCols(), kUndefined);
    this->CopyFromMat(other);
    return *this;
  }
  CuMatrix<Real> &operator=(const CuMatrix<Real> &other) {
    this->Resize(other.NumRows(), other.NumCols(), kUndefined);
    this->CopyFromMat(other);
    return *this;
  }
  CuMatrix<Real> &operator=(const MatrixBase<Real> &other) {
    this->Resize(other.NumRows(), other.NumCols(), kUndefined);
    this->CopyFromMat(other);
    return *this;
  }

  /// Allocates (if the size changes) with rows pitched to 16 bytes.  As in Kaldi,
  /// a same-size Resize only zeroes (kSetZero) or does nothing (kUndefined).
  void Resize(MatrixIndexT rows, MatrixIndexT cols, MatrixResizeType resize_type = kSetZero);
  void Swap(CuMatrix<Real> *mat);
  /// Shim extension: make this object a [rows x cols] view of caller-owned device memory
  /// (never freed here).  Lets the C API hand a foreign buffer to interfaces that take a
  /// CuMatrix<Real>* (Backprop's in_deriv); a Resize to the same shape keeps the buffer.
  void Borrow(Real *data, MatrixIndexT rows, MatrixIndexT cols, MatrixIndexT stride);
  void Read(std::istream &is, bool binary);
  void Write(std::ostream &os, bool binary) const;

 private:
  void Destroy();
  bool owns_;
};

template <typename Real>
class CuSubMatrix : public CuMatrixBase<Real> {
 public:
  inline CuSubMatrix(const CuMatrixBase<Real> &mat, const MatrixIndexT row_offset,
                     const MatrixIndexT num_rows, const MatrixIndexT col_offset,
                     const MatrixIndexT num_cols) {
    KALDI_ASSERT(row_offset >= 0 && col_offset >= 0 && num_rows >= 0 && num_cols >= 0 &&
                 row_offset + num_rows <= mat.num_rows_ && col_offset + num_cols <= mat.num_cols_);
    this->data_ = mat.data_ + (size_t)row_offset * mat.stride_ + col_offset;
    this->num_cols_ = num_cols;
    this->num_rows_ = num_rows;
    this->stride_ = mat.stride_;
  }
  /// View of caller-owned device memory (what the C API builds from a raw pointer).
  inline CuSubMatrix(const Real *data, MatrixIndexT num_rows, MatrixIndexT num_cols, MatrixIndexT stride)
      : CuMatrixBase<Real>(const_cast<Real *>(data), num_rows, num_cols, stride) {}
  inline CuSubMatrix(const CuSubMatrix &other)
      : CuMatrixBase<Real>(other.data_, other.num_rows_, other.num_cols_, other.stride_) {}
 private:
  CuSubMatrix<Real> &operator=(const CuSubMatrix<Real> &other);
};

template <typename Real>
Real TraceMatMat(const CuMatrixBase<Real> &A, const CuMatrixBase<Real> &B,
                 MatrixTransposeType trans = kNoTrans);

}  // namespace kaldi
#endif
