// matrix/matrix-lib.h -- shim: minimal HOST Vector / Matrix, used only for parameter
// I/O (Read / Write, SetParams, Vectorize) -- never on the timed path.
#ifndef KALDI_MATRIX_MATRIX_LIB_H_
#define KALDI_MATRIX_MATRIX_LIB_H_

#include <vector>
#include "matrix/matrix-common.h"

namespace kaldi {

template <typename Real>
class VectorBase {
 public:
  MatrixIndexT Dim() const { return dim_; }
  Real *Data() { return data_; }
  const Real *Data() const { return data_; }
  Real &operator()(MatrixIndexT i) { return data_[i]; }
  Real operator()(MatrixIndexT i) const { return data_[i]; }
  void SetZero() { for (MatrixIndexT i = 0; i < dim_; i++) data_[i] = 0; }
  void Write(std::ostream &os, bool binary) const;
 protected:
  VectorBase() : data_(NULL), dim_(0) {}
  Real *data_;
  MatrixIndexT dim_;
};

template <typename Real>
class Vector : public VectorBase<Real> {
 public:
  Vector() {}
  explicit Vector(MatrixIndexT dim, MatrixResizeType t = kSetZero) { Resize(dim, t); }
  Vector(const Vector<Real> &o) : VectorBase<Real>() { *this = o; }
  Vector<Real> &operator=(const Vector<Real> &o) {
    Resize(o.Dim(), kUndefined);
    for (MatrixIndexT i = 0; i < this->dim_; i++) this->data_[i] = o(i);
    return *this;
  }
  void Resize(MatrixIndexT dim, MatrixResizeType t = kSetZero) {
    store_.assign(dim, Real(0));
    (void)t;
    this->data_ = dim ? store_.data() : NULL;
    this->dim_ = dim;
  }
  void Read(std::istream &is, bool binary);
 private:
  std::vector<Real> store_;
};

template <typename Real>
class MatrixBase {
 public:
  MatrixIndexT NumRows() const { return rows_; }
  MatrixIndexT NumCols() const { return cols_; }
  MatrixIndexT Stride() const { return stride_; }
  Real *Data() { return data_; }
  const Real *Data() const { return data_; }
  Real *RowData(MatrixIndexT r) { return data_ + (size_t)r * stride_; }
  const Real *RowData(MatrixIndexT r) const { return data_ + (size_t)r * stride_; }
  Real &operator()(MatrixIndexT r, MatrixIndexT c) { return data_[(size_t)r * stride_ + c]; }
  Real operator()(MatrixIndexT r, MatrixIndexT c) const { return data_[(size_t)r * stride_ + c]; }
  void Write(std::ostream &os, bool binary) const;
 protected:
  MatrixBase() : data_(NULL), rows_(0), cols_(0), stride_(0) {}
  Real *data_;
  MatrixIndexT rows_, cols_, stride_;
};

template <typename Real>
class Matrix : public MatrixBase<Real> {
 public:
  Matrix() {}
  Matrix(MatrixIndexT r, MatrixIndexT c, MatrixResizeType t = kSetZero) { Resize(r, c, t); }
  Matrix(const Matrix<Real> &o) : MatrixBase<Real>() { *this = o; }
  Matrix<Real> &operator=(const Matrix<Real> &o) {
    Resize(o.NumRows(), o.NumCols(), kUndefined);
    for (MatrixIndexT r = 0; r < this->rows_; r++)
      for (MatrixIndexT c = 0; c < this->cols_; c++) (*this)(r, c) = o(r, c);
    return *this;
  }
  void Resize(MatrixIndexT r, MatrixIndexT c, MatrixResizeType t = kSetZero) {
    (void)t;
    store_.assign((size_t)r * c, Real(0));
    this->data_ = store_.empty() ? NULL : store_.data();
    this->rows_ = r; this->cols_ = c; this->stride_ = c;
  }
  void Read(std::istream &is, bool binary);
 private:
  std::vector<Real> store_;
};

}  // namespace kaldi
#endif
