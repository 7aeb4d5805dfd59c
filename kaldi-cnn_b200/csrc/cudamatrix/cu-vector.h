// cudamatrix/cu-vector.h -- shim: CuVectorBase / CuVector / CuSubVector (device memory).
#ifndef KALDI_CUDAMATRIX_CU_VECTOR_H_
#define KALDI_CUDAMATRIX_CU_VECTOR_H_

#include "matrix/matrix-lib.h"
#include "cudamatrix/cu-common.h"
#include "cudamatrix/cu-device.h"

namespace kaldi {

template <typename Real> class CuMatrixBase;
template <typename Real> class CuSubVector;

template <typename Real>
class CuVectorBase {
 public:
  friend class CuMatrixBase<Real>;
  MatrixIndexT Dim() const { return dim_; }
  inline const Real *Data() const { return data_; }
  inline Real *Data() { return data_; }

  void SetZero();
  void Set(Real value);
  void Add(Real value);
  void Scale(Real value);
  void SetRandn();
  /// *this += alpha * vec
  void AddVec(Real alpha, const CuVectorBase<Real> &vec, Real beta = 1.0);
  /// *this = alpha * (sum of the rows of mat) + beta * *this
  void AddRowSumMat(Real alpha, const CuMatrixBase<Real> &mat, Real beta = 1.0);
  void CopyFromVec(const CuVectorBase<Real> &src);
  void CopyFromVec(const VectorBase<Real> &src);
  void CopyToVec(VectorBase<Real> *dst) const;
  void CopyColFromMat(const CuMatrixBase<Real> &mat, MatrixIndexT col);
  CuSubVector<Real> Range(const MatrixIndexT o, const MatrixIndexT l) {
    return CuSubVector<Real>(*this, o, l);
  }
  /// Single element read (device -> host copy; debugging / Info() only).
  Real operator()(MatrixIndexT i) const;

 protected:
  CuVectorBase() : data_(NULL), dim_(0) {}
  Real *data_;
  MatrixIndexT dim_;
 private:
  KALDI_DISALLOW_COPY_AND_ASSIGN(CuVectorBase);
};

template <typename Real>
class CuVector : public CuVectorBase<Real> {
 public:
  CuVector() : owns_(true) {}
  CuVector(MatrixIndexT dim, MatrixResizeType t = kSetZero) : owns_(true) { Resize(dim, t); }
  CuVector(const CuVectorBase<Real> &v) : owns_(true) { Resize(v.Dim(), kUndefined); this->CopyFromVec(v); }
  CuVector(const CuVector<Real> &v) : CuVectorBase<Real>(), owns_(true) { Resize(v.Dim(), kUndefined); this->CopyFromVec(v); }
  CuVector(const VectorBase<Real> &v) : owns_(true) { Resize(v.Dim(), kUndefined); this->CopyFromVec(v); }
  ~CuVector() { Destroy(); }
  CuVector<Real> &operator=(const CuVectorBase<Real> &o) { Resize(o.Dim(), kUndefined); this->CopyFromVec(o); return *this; }
  CuVector<Real> &operator=(const CuVector<Real> &o) { Resize(o.Dim(), kUndefined); this->CopyFromVec(o); return *this; }
  CuVector<Real> &operator=(const VectorBase<Real> &o) { Resize(o.Dim(), kUndefined); this->CopyFromVec(o); return *this; }
  void Resize(MatrixIndexT dim, MatrixResizeType t = kSetZero);
  /// B200 extension (as CuMatrix::Borrow): become a non-owning view of caller-provided device memory --
  /// the data-parallel trainer keeps every parameter in one NVLink-visible arena.  A Resize to the same
  /// dimension keeps the view.
  void Borrow(Real *data, MatrixIndexT dim) { Destroy(); this->data_ = data; this->dim_ = dim; owns_ = false; }
  void Read(std::istream &is, bool binary);
  void Write(std::ostream &os, bool binary) const;
 private:
  void Destroy();
  bool owns_;
};

template <typename Real>
class CuSubVector : public CuVectorBase<Real> {
 public:
  CuSubVector(const CuVectorBase<Real> &t, const MatrixIndexT origin, const MatrixIndexT length) {
    KALDI_ASSERT(origin >= 0 && length >= 0 && origin + length <= t.Dim());
    this->data_ = const_cast<Real *>(t.Data()) + origin;
    this->dim_ = length;
  }
  CuSubVector(const CuSubVector &o) : CuVectorBase<Real>() { this->data_ = o.data_; this->dim_ = o.dim_; }
  CuSubVector(const Real *data, MatrixIndexT length) { this->data_ = const_cast<Real *>(data); this->dim_ = length; }
};

template <typename Real>
Real VecVec(const CuVectorBase<Real> &a, const CuVectorBase<Real> &b);

}  // namespace kaldi
#endif
