// cudamatrix/cu-matrix.cc -- shim: stock CuMatrix / CuVector operations (GPU only).
#include <random>
#include <vector>

#include "cudamatrix/cu-matrix.h"
#include "cudamatrix/cu-kernels-stock.h"
#include "cnsl-cu-kernels.h"

namespace kaldi {

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }
static inline void Need(const char *what) { CuDevice::Instantiate().RequireEnabled(what); }

// ------------------------------------------------------------------ CuVector --

template <typename Real> void CuVector<Real>::Destroy() {
  if (this->data_ && owns_) CuDevice::Instantiate().Free(this->data_);
  this->data_ = NULL;
  this->dim_ = 0;
  owns_ = true;
}

template <typename Real> void CuVector<Real>::Resize(MatrixIndexT dim, MatrixResizeType t) {
  KALDI_ASSERT(dim >= 0);
  if (this->dim_ == dim) { if (t == kSetZero) this->SetZero(); return; }
  Destroy();
  if (dim == 0) return;
  Need("CuVector::Resize");
  this->data_ = static_cast<Real *>(CuDevice::Instantiate().Malloc(sizeof(Real) * dim));
  this->dim_ = dim;
  if (t == kSetZero) this->SetZero();
}

template <typename Real> void CuVectorBase<Real>::SetZero() {
  if (dim_ == 0) return;
  Need("CuVector::SetZero");
  CU_SAFE_CALL(cudaMemsetAsync(data_, 0, sizeof(Real) * dim_, Str()));
}
template <typename Real> void CuVectorBase<Real>::Set(Real v) { Need("CuVector::Set"); cu_stock::set_vec(Str(), data_, dim_, v); }
template <typename Real> void CuVectorBase<Real>::Add(Real v) { Need("CuVector::Add"); cu_stock::add_const_vec(Str(), data_, dim_, v); }
template <typename Real> void CuVectorBase<Real>::Scale(Real v) { Need("CuVector::Scale"); cu_stock::scale_vec(Str(), data_, dim_, v); }

template <typename Real> void CuVectorBase<Real>::AddVec(Real alpha, const CuVectorBase<Real> &vec, Real beta) {
  KALDI_ASSERT(vec.Dim() == dim_);
  Need("CuVector::AddVec");
  cu_stock::axpby_vec(Str(), data_, dim_, alpha, vec.Data(), beta);
}

template <typename Real>
void CuVectorBase<Real>::AddRowSumMat(Real alpha, const CuMatrixBase<Real> &mat, Real beta) {
  KALDI_ASSERT(mat.NumCols() == dim_);
  Need("CuVector::AddRowSumMat");
  CuVector<Real> sums(dim_, kUndefined);
  cudaF_sum_rows_per_map(Str(), mat.Data(), mat.Dim(), 1, sums.Data());
  cu_stock::axpby_vec(Str(), data_, dim_, alpha, sums.Data(), beta);
}

template <typename Real> void CuVectorBase<Real>::CopyFromVec(const CuVectorBase<Real> &src) {
  KALDI_ASSERT(src.Dim() == dim_);
  if (dim_ == 0) return;
  CU_SAFE_CALL(cudaMemcpyAsync(data_, src.Data(), sizeof(Real) * dim_, cudaMemcpyDeviceToDevice, Str()));
}
template <typename Real> void CuVectorBase<Real>::CopyFromVec(const VectorBase<Real> &src) {
  KALDI_ASSERT(src.Dim() == dim_);
  if (dim_ == 0) return;
  CU_SAFE_CALL(cudaMemcpyAsync(data_, src.Data(), sizeof(Real) * dim_, cudaMemcpyHostToDevice, Str()));
  CU_SAFE_CALL(cudaStreamSynchronize(Str()));
}
template <typename Real> void CuVectorBase<Real>::CopyToVec(VectorBase<Real> *dst) const {
  KALDI_ASSERT(dst->Dim() == dim_);
  if (dim_ == 0) return;
  CU_SAFE_CALL(cudaMemcpyAsync(dst->Data(), data_, sizeof(Real) * dim_, cudaMemcpyDeviceToHost, Str()));
  CU_SAFE_CALL(cudaStreamSynchronize(Str()));
}
template <typename Real> void CuVectorBase<Real>::CopyColFromMat(const CuMatrixBase<Real> &mat, MatrixIndexT col) {
  KALDI_ASSERT(col < mat.NumCols() && dim_ == mat.NumRows());
  cu_stock::copy_col(Str(), data_, mat.Data(), mat.Dim(), col);
}
template <typename Real> Real CuVectorBase<Real>::operator()(MatrixIndexT i) const {
  KALDI_ASSERT(i >= 0 && i < dim_);
  Real v;
  CU_SAFE_CALL(cudaMemcpyAsync(&v, data_ + i, sizeof(Real), cudaMemcpyDeviceToHost, Str()));
  CU_SAFE_CALL(cudaStreamSynchronize(Str()));
  return v;
}
template <typename Real> void CuVectorBase<Real>::SetRandn() {
  if (dim_ == 0) return;
  Vector<Real> tmp(dim_);
  std::mt19937_64 gen(CuDevice::Instantiate().NextRandSeed());
  std::normal_distribution<double> nd(0.0, 1.0);
  for (MatrixIndexT i = 0; i < dim_; i++) tmp(i) = static_cast<Real>(nd(gen));
  CopyFromVec(tmp);
}
template <typename Real> void CuVector<Real>::Read(std::istream &is, bool binary) {
  Vector<Real> tmp;
  tmp.Read(is, binary);
  Resize(tmp.Dim(), kUndefined);
  this->CopyFromVec(tmp);
}
template <typename Real> void CuVector<Real>::Write(std::ostream &os, bool binary) const {
  Vector<Real> tmp(this->dim_, kUndefined);
  this->CopyToVec(&tmp);
  tmp.Write(os, binary);
}
template <typename Real> Real VecVec(const CuVectorBase<Real> &a, const CuVectorBase<Real> &b) {
  KALDI_ASSERT(a.Dim() == b.Dim());
  Vector<Real> ha(a.Dim()), hb(b.Dim());
  a.CopyToVec(&ha); b.CopyToVec(&hb);
  double s = 0;
  for (MatrixIndexT i = 0; i < ha.Dim(); i++) s += (double)ha(i) * hb(i);
  return static_cast<Real>(s);
}

// ------------------------------------------------------------------ CuMatrix --

template <typename Real> void CuMatrix<Real>::Destroy() {
  if (this->data_ && owns_) CuDevice::Instantiate().Free(this->data_);
  this->data_ = NULL;
  this->num_rows_ = this->num_cols_ = this->stride_ = 0;
  owns_ = true;
}

template <typename Real>
void CuMatrix<Real>::Borrow(Real *data, MatrixIndexT rows, MatrixIndexT cols, MatrixIndexT stride) {
  Destroy();
  this->data_ = data;
  this->num_rows_ = rows;
  this->num_cols_ = cols;
  this->stride_ = stride;
  owns_ = false;
}

template <typename Real>
void CuMatrix<Real>::Resize(MatrixIndexT rows, MatrixIndexT cols, MatrixResizeType resize_type) {
  KALDI_ASSERT(rows >= 0 && cols >= 0);
  KALDI_ASSERT(resize_type == kSetZero || resize_type == kUndefined);   // kCopyData not needed here
  if (rows * cols == 0) KALDI_ASSERT(rows == 0 && cols == 0);
  if (this->num_rows_ == rows && this->num_cols_ == cols) {
    if (resize_type == kSetZero) this->SetZero();
    return;
  }
  Destroy();
  if (rows == 0) return;
  Need("CuMatrix::Resize");
  MatrixIndexT stride = CuDevice::PitchInElements(cols, sizeof(Real));
  this->data_ = static_cast<Real *>(CuDevice::Instantiate().Malloc(sizeof(Real) * (size_t)rows * stride));
  this->num_rows_ = rows;
  this->num_cols_ = cols;
  this->stride_ = stride;
  if (resize_type == kSetZero) this->SetZero();
}

template <typename Real> void CuMatrix<Real>::Swap(CuMatrix<Real> *mat) {
  std::swap(mat->owns_, this->owns_);
  std::swap(mat->data_, this->data_);
  std::swap(mat->num_cols_, this->num_cols_);
  std::swap(mat->num_rows_, this->num_rows_);
  std::swap(mat->stride_, this->stride_);
}

template <typename Real>
CuMatrix<Real>::CuMatrix(const CuMatrix<Real> &other, MatrixTransposeType trans) : owns_(true) {
  if (trans == kNoTrans) this->Resize(other.NumRows(), other.NumCols(), kUndefined);
  else this->Resize(other.NumCols(), other.NumRows(), kUndefined);
  this->CopyFromMat(other, trans);
}
template <typename Real>
CuMatrix<Real>::CuMatrix(const CuMatrixBase<Real> &other, MatrixTransposeType trans) : owns_(true) {
  if (trans == kNoTrans) this->Resize(other.NumRows(), other.NumCols(), kUndefined);
  else this->Resize(other.NumCols(), other.NumRows(), kUndefined);
  this->CopyFromMat(other, trans);
}
template <typename Real>
CuMatrix<Real>::CuMatrix(const MatrixBase<Real> &other, MatrixTransposeType trans) : owns_(true) {
  if (trans == kNoTrans) this->Resize(other.NumRows(), other.NumCols(), kUndefined);
  else this->Resize(other.NumCols(), other.NumRows(), kUndefined);
  this->CopyFromMat(other, trans);
}

template <typename Real> void CuMatrixBase<Real>::SetZero() {
  if (num_rows_ == 0) return;
  Need("CuMatrix::SetZero");
  CU_SAFE_CALL(cudaMemset2DAsync(data_, sizeof(Real) * stride_, 0, sizeof(Real) * num_cols_, num_rows_, Str()));
}
template <typename Real> void CuMatrixBase<Real>::Set(Real v) { Need("CuMatrix::Set"); cu_stock::set_mat(Str(), data_, Dim(), v); }
template <typename Real> void CuMatrixBase<Real>::Scale(Real v) { Need("CuMatrix::Scale"); cu_stock::scale_mat(Str(), data_, Dim(), v); }

template <typename Real>
void CuMatrixBase<Real>::AddMat(Real alpha, const CuMatrixBase<Real> &A, MatrixTransposeType transA) {
  if (transA == kNoTrans) KALDI_ASSERT(A.NumRows() == num_rows_ && A.NumCols() == num_cols_);
  else KALDI_ASSERT(A.NumCols() == num_rows_ && A.NumRows() == num_cols_);
  Need("CuMatrix::AddMat");
  cu_stock::add_mat(Str(), data_, Dim(), alpha, A.Data(), A.Dim(), transA == kTrans);
}

template <typename Real>
void CuMatrixBase<Real>::CopyFromMat(const CuMatrixBase<Real> &src, MatrixTransposeType trans) {
  if (trans == kNoTrans) KALDI_ASSERT(src.NumRows() == num_rows_ && src.NumCols() == num_cols_);
  else KALDI_ASSERT(src.NumCols() == num_rows_ && src.NumRows() == num_cols_);
  if (num_rows_ == 0) return;
  Need("CuMatrix::CopyFromMat");
  if (trans == kNoTrans) {
    CU_SAFE_CALL(cudaMemcpy2DAsync(data_, sizeof(Real) * stride_, src.Data(), sizeof(Real) * src.Stride(),
                                   sizeof(Real) * num_cols_, num_rows_, cudaMemcpyDeviceToDevice, Str()));
  } else {
    cu_stock::copy_mat(Str(), data_, Dim(), src.Data(), src.Dim(), true);
  }
}

template <typename Real>
void CuMatrixBase<Real>::CopyFromMat(const MatrixBase<Real> &src, MatrixTransposeType trans) {
  if (num_rows_ == 0) return;
  Need("CuMatrix::CopyFromMat(host)");
  if (trans == kNoTrans) {
    KALDI_ASSERT(src.NumRows() == num_rows_ && src.NumCols() == num_cols_);
    CU_SAFE_CALL(cudaMemcpy2DAsync(data_, sizeof(Real) * stride_, src.Data(), sizeof(Real) * src.Stride(),
                                   sizeof(Real) * num_cols_, num_rows_, cudaMemcpyHostToDevice, Str()));
    CU_SAFE_CALL(cudaStreamSynchronize(Str()));
  } else {
    CuMatrix<Real> tmp(src);
    this->CopyFromMat(tmp, kTrans);
  }
}

template <typename Real> void CuMatrixBase<Real>::CopyToMat(MatrixBase<Real> *dst) const {
  KALDI_ASSERT(dst->NumRows() == num_rows_ && dst->NumCols() == num_cols_);
  if (num_rows_ == 0) return;
  CU_SAFE_CALL(cudaMemcpy2DAsync(dst->Data(), sizeof(Real) * dst->Stride(), data_, sizeof(Real) * stride_,
                                 sizeof(Real) * num_cols_, num_rows_, cudaMemcpyDeviceToHost, Str()));
  CU_SAFE_CALL(cudaStreamSynchronize(Str()));
}

template <typename Real> void CuMatrixBase<Real>::CopyRowsFromVec(const CuVectorBase<Real> &v) {
  KALDI_ASSERT(v.Dim() == num_cols_);
  Need("CuMatrix::CopyRowsFromVec");
  cu_stock::copy_rows_from_vec(Str(), data_, Dim(), v.Data());
}
template <typename Real> void CuMatrixBase<Real>::MulElements(const CuMatrixBase<Real> &A) {
  KALDI_ASSERT(A.NumRows() == num_rows_ && A.NumCols() == num_cols_);
  cu_stock::mul_elements(Str(), data_, Dim(), A.Data(), A.Dim());
}
template <typename Real> void CuMatrixBase<Real>::Max(const CuMatrixBase<Real> &A) {
  KALDI_ASSERT(A.NumRows() == num_rows_ && A.NumCols() == num_cols_);
  cu_stock::max_elements(Str(), data_, Dim(), A.Data(), A.Dim());
}
template <typename Real>
void CuMatrixBase<Real>::EqualElementMask(const CuMatrixBase<Real> &mat, CuMatrix<Real> *mask) const {
  KALDI_ASSERT(mat.NumRows() == num_rows_ && mat.NumCols() == num_cols_ && mask != NULL);
  mask->Resize(num_rows_, num_cols_, kSetZero);
  cu_stock::equal_mask(Str(), data_, Dim(), mat.Data(), mat.Dim(), mask->Data(), mask->Dim());
}

template <typename Real>
void CuMatrixBase<Real>::AddMatMat(Real alpha, const CuMatrixBase<Real> &A, MatrixTransposeType transA,
                                   const CuMatrixBase<Real> &B, MatrixTransposeType transB, Real beta) {
  MatrixIndexT m = (transA == kTrans ? A.NumCols() : A.NumRows()),
               k = (transA == kTrans ? A.NumRows() : A.NumCols()),
               k1 = (transB == kTrans ? B.NumCols() : B.NumRows()),
               n = (transB == kTrans ? B.NumRows() : B.NumCols());
  KALDI_ASSERT(k == k1 && m == num_rows_ && n == num_cols_);
  if (m == 0) return;
  Need("CuMatrix::AddMatMat");
  const int math = CuDevice::Instantiate().MathMode();
  const bool direct = (alpha == Real(1) && beta == Real(0));
  CuMatrix<Real> prod_store;
  if (!direct) prod_store.Resize(m, n, kUndefined);
  Real *prod = direct ? data_ : prod_store.Data();
  ::MatrixDim pd = direct ? Dim() : prod_store.Dim();
  cudaStream_t st = Str();
  if (transA == kNoTrans && transB == kTrans) {
    cudaF_affine_fprop(st, math, A.Data(), A.Dim(), B.Data(), B.Dim(), NULL, prod, pd);
  } else if (transA == kNoTrans && transB == kNoTrans) {
    cudaF_affine_dgrad(st, math, A.Data(), A.Dim(), B.Data(), B.Dim(), prod, pd);
  } else if (transA == kTrans && transB == kNoTrans) {
    cudaF_affine_wgrad(st, math, B.Data(), B.Dim(), A.Data(), A.Dim(), prod, pd, NULL);
  } else {
    CuMatrix<Real> At(A, kTrans);
    cudaF_affine_fprop(st, math, At.Data(), At.Dim(), B.Data(), B.Dim(), NULL, prod, pd);
  }
  if (!direct) {
    if (beta == Real(0)) this->SetZero();
    else if (beta != Real(1)) this->Scale(beta);
    this->AddMat(alpha, prod_store, kNoTrans);
  }
}

template <typename Real> Real CuMatrixBase<Real>::operator()(MatrixIndexT r, MatrixIndexT c) const {
  KALDI_ASSERT(r >= 0 && r < num_rows_ && c >= 0 && c < num_cols_);
  Real v;
  CU_SAFE_CALL(cudaMemcpyAsync(&v, data_ + (size_t)r * stride_ + c, sizeof(Real), cudaMemcpyDeviceToHost, Str()));
  CU_SAFE_CALL(cudaStreamSynchronize(Str()));
  return v;
}

template <typename Real> Real CuMatrixBase<Real>::Sum() const {
  Matrix<Real> h(num_rows_, num_cols_);
  CopyToMat(&h);
  double s = 0;
  for (MatrixIndexT i = 0; i < num_rows_; i++)
    for (MatrixIndexT j = 0; j < num_cols_; j++) s += h(i, j);
  return static_cast<Real>(s);
}

template <typename Real> void CuMatrixBase<Real>::SetRandn() {
  if (num_rows_ == 0) return;
  Matrix<Real> tmp(num_rows_, num_cols_);
  std::mt19937_64 gen(CuDevice::Instantiate().NextRandSeed());
  std::normal_distribution<double> nd(0.0, 1.0);
  for (MatrixIndexT i = 0; i < num_rows_; i++)
    for (MatrixIndexT j = 0; j < num_cols_; j++) tmp(i, j) = static_cast<Real>(nd(gen));
  CopyFromMat(tmp);
}

template <typename Real> void CuMatrix<Real>::Read(std::istream &is, bool binary) {
  Matrix<Real> tmp;
  tmp.Read(is, binary);
  Resize(tmp.NumRows(), tmp.NumCols(), kUndefined);
  this->CopyFromMat(tmp);
}
template <typename Real> void CuMatrix<Real>::Write(std::ostream &os, bool binary) const {
  Matrix<Real> tmp(this->num_rows_, this->num_cols_, kUndefined);
  this->CopyToMat(&tmp);
  tmp.Write(os, binary);
}

template <typename Real>
Real TraceMatMat(const CuMatrixBase<Real> &A, const CuMatrixBase<Real> &B, MatrixTransposeType trans) {
  Matrix<Real> ha(A.NumRows(), A.NumCols()), hb(B.NumRows(), B.NumCols());
  A.CopyToMat(&ha); B.CopyToMat(&hb);
  double s = 0;
  if (trans == kNoTrans) {
    KALDI_ASSERT(A.NumRows() == B.NumCols() && A.NumCols() == B.NumRows());
    for (MatrixIndexT i = 0; i < ha.NumRows(); i++)
      for (MatrixIndexT j = 0; j < ha.NumCols(); j++) s += (double)ha(i, j) * hb(j, i);
  } else {
    KALDI_ASSERT(A.NumRows() == B.NumRows() && A.NumCols() == B.NumCols());
    for (MatrixIndexT i = 0; i < ha.NumRows(); i++)
      for (MatrixIndexT j = 0; j < ha.NumCols(); j++) s += (double)ha(i, j) * hb(i, j);
  }
  return static_cast<Real>(s);
}

template class CuVectorBase<float>;
template class CuVector<float>;
template class CuMatrix<float>;
template float VecVec(const CuVectorBase<float> &, const CuVectorBase<float> &);
template float TraceMatMat(const CuMatrixBase<float> &, const CuMatrixBase<float> &, MatrixTransposeType);

// Stock members of CuMatrixBase<float> (the ten CNN members are instantiated in
// cnslmat/conv2D.cc, so the class is not explicitly instantiated as a whole here).
template void CuMatrixBase<float>::SetZero();
template void CuMatrixBase<float>::Set(float);
template void CuMatrixBase<float>::Scale(float);
template void CuMatrixBase<float>::SetRandn();
template void CuMatrixBase<float>::AddMat(float, const CuMatrixBase<float> &, MatrixTransposeType);
template void CuMatrixBase<float>::AddMatMat(float, const CuMatrixBase<float> &, MatrixTransposeType,
                                             const CuMatrixBase<float> &, MatrixTransposeType, float);
template void CuMatrixBase<float>::CopyFromMat(const CuMatrixBase<float> &, MatrixTransposeType);
template void CuMatrixBase<float>::CopyFromMat(const MatrixBase<float> &, MatrixTransposeType);
template void CuMatrixBase<float>::CopyToMat(MatrixBase<float> *) const;
template void CuMatrixBase<float>::CopyRowsFromVec(const CuVectorBase<float> &);
template void CuMatrixBase<float>::MulElements(const CuMatrixBase<float> &);
template void CuMatrixBase<float>::Max(const CuMatrixBase<float> &);
template void CuMatrixBase<float>::EqualElementMask(const CuMatrixBase<float> &, CuMatrix<float> *) const;
template float CuMatrixBase<float>::Sum() const;
template float CuMatrixBase<float>::operator()(MatrixIndexT, MatrixIndexT) const;

}  // namespace kaldi
