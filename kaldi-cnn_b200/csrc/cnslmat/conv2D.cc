// cnslmat/conv2D.cc -- L1: definitions of the CNN-layer members of CuMatrixBase<Real>
// for the B200 build.
//
// Same member names, parameter lists, shape assertions, resize rules and error
// behaviour as the reference's src/cnslmat/conv2D.cc (cited per function); the
// bodies differ:
//   * one stream-ordered launch through the extern "C" layer of
//     include/cnsl-cu-kernels.h on CuDevice::Stream(), instead of a 16x16-block
//     launch on the default stream;
//   * Conv2D is ONE implicit-GEMM launch -- no cudaMemGetInfo, no im2col / GEMM /
//     copy_rows_at split loop, no convMat temporary (conv2D.cc:69-185);
//   * there is no CPU "else" branch: without a device the call fails loudly.
// Unlike the reference (which #includes this file into nnet-component-nnet0.cc,
// cnslmat/Makefile:15-18) it is compiled on its own and instantiated for float.

#include "cudamatrix/cu-matrix.h"
#include "cnsl-cu-kernels.h"
#include "kcnn_common.cuh"

namespace kaldi {

namespace {
struct Launch {       // Timer + error check + AccuProfile bracket (conv2D.cc:101-110)
  const char *func;
  Timer tim;
  explicit Launch(const char *f) : func(f) { CuDevice::Instantiate().RequireEnabled(f); }
  ~Launch() noexcept(false) {
    CU_SAFE_CALL(cudaGetLastError());
    CuDevice::Instantiate().AccuProfile(func, tim.Elapsed());
  }
};
inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }
inline int PoolMode(bool overlap, bool overlap2D) {
  return overlap ? KCNN_POOL_OVERLAP : (overlap2D ? KCNN_POOL_OVERLAP2D : KCNN_POOL_PLAIN);
}
}  // namespace

// reference: conv2D.cc:44-201.  'this' [num_chunks x in_height*in_width*in_channel],
// kernel [kernel_height*kernel_width*in_channel x group]; out is NEVER resized (:62-63).
template <typename Real>
void CuMatrixBase<Real>::Conv2D(const CuMatrixBase<Real> &kernel, int32 in_height, int32 in_width,
                                int32 in_channel, int32 kernel_height, int32 kernel_width,
                                int32 group, CuMatrixBase<Real> *out, bool concat) const {
  KALDI_ASSERT(NumCols() == in_height * in_width * in_channel);
  KALDI_ASSERT(kernel.NumCols() == group);
  KALDI_ASSERT(kernel.NumRows() == kernel_height * kernel_width * in_channel);
  int32 out_height = in_height - kernel_height + 1, out_width = in_width - kernel_width + 1;
  KALDI_ASSERT(out != NULL);
  if (concat) {
    KALDI_ASSERT(out->NumRows() == NumRows() && out->NumCols() == out_height * out_width * group);
  } else {
    KALDI_ASSERT(out->NumRows() == out_height * out_width * NumRows() && out->NumCols() == group);
  }
  Launch l(__func__);
  cudaF_conv2d_fprop(Str(), CuDevice::Instantiate().MathMode(), data_, Dim(), kernel.Data(),
                     kernel.Dim(), NULL, out->Data(), out->Dim(), in_height, in_width, in_channel, 0,
                     0, kernel_height, kernel_width, group, concat ? 1 : 0);
}

// reference: conv2D.cc:213-242 (const, yet mutates *this -- kept).
template <typename Real>
void CuMatrixBase<Real>::AddMatRepVec(const CuVectorBase<Real> &vec, int32 rep) const {
  KALDI_ASSERT(vec.Dim() * rep == this->NumCols());
  Launch l(__func__);
  cudaF_add_mat_rep_vec_s(Str(), vec.Data(), rep, this->data_, this->Dim());
}

// reference: conv2D.cc:244-287.
template <typename Real>
void CuMatrixBase<Real>::FlipMat(int32 kernel_height, int32 kernel_width, int32 in_channel,
                                 int32 group, CuMatrix<Real> *flip) const {
  KALDI_ASSERT(NumRows() == (kernel_height * kernel_width * in_channel));
  KALDI_ASSERT(flip != NULL);
  if ((flip->NumRows() != (kernel_height * kernel_width * group)) || (flip->NumCols() != in_channel))
    flip->Resize((kernel_height * kernel_width * group), in_channel, kSetZero);
  Launch l(__func__);
  cudaF_flip_mat_s(Str(), this->data_, this->Dim(), kernel_height, kernel_width, group, flip->Data(),
                   flip->Dim());
}

// reference: conv2D.cc:289-344.
template <typename Real>
void CuMatrixBase<Real>::PaddingZero(int32 orig_height, int32 orig_width, int32 orig_channel,
                                     int32 kernel_height, int32 kernel_width,
                                     CuMatrix<Real> *padmat) const {
  KALDI_ASSERT(NumCols() == (orig_height * orig_width * orig_channel));
  KALDI_ASSERT(padmat != NULL);
  int32 padmat_height = orig_height + 2 * (kernel_height - 1),
        padmat_width = orig_width + 2 * (kernel_width - 1);
  if (padmat->NumRows() != NumRows() || padmat->NumCols() != padmat_height * padmat_width * orig_channel)
    padmat->Resize(NumRows(), padmat_height * padmat_width * orig_channel, kSetZero);
  Launch l(__func__);
  cudaF_pad_zero_s(Str(), this->data_, this->Dim(), orig_height, orig_width, kernel_height,
                   kernel_width, padmat->Data(), padmat->Dim());
}

// reference: conv2D.cc:348-386.
template <typename Real>
void CuMatrixBase<Real>::TpBlock(int32 in_channel, int32 block_size, CuMatrix<Real> *out) const {
  KALDI_ASSERT(this->NumCols() == block_size * in_channel);
  KALDI_ASSERT(out != NULL);
  if ((out->NumRows() != in_channel) || (out->NumCols() != NumRows() * block_size))
    out->Resize(in_channel, NumRows() * block_size, kSetZero);
  Launch l(__func__);
  cudaF_tp_block_s(Str(), this->data_, this->Dim(), out->Data(), out->Dim(), block_size);
}

// reference: conv2D.cc:388-426.
template <typename Real>
void CuMatrixBase<Real>::TpInsideBlock(int32 group, int32 block_size, CuMatrix<Real> *out) const {
  KALDI_ASSERT(this->NumCols() == block_size * group);
  KALDI_ASSERT(out != NULL);
  if ((out->NumRows() != block_size * NumRows()) || (out->NumCols() != group))
    out->Resize(block_size * NumRows(), group, kSetZero);
  Launch l(__func__);
  cudaF_tp_inside_block_s(Str(), this->data_, this->Dim(), out->Data(), out->Dim(), block_size);
}

// reference: conv2D.cc:429-463.
template <typename Real>
void CuMatrixBase<Real>::ModPermuteRow(int32 in_channel, int32 block_size, CuMatrix<Real> *out) const {
  KALDI_ASSERT(out != NULL);
  if ((out->NumRows() != NumRows()) || (out->NumCols() != NumCols()))
    out->Resize(NumRows(), NumCols(), kSetZero);
  Launch l(__func__);
  cudaF_mod_permute_row_s(Str(), this->data_, this->Dim(), out->Data(), out->Dim(), block_size,
                          in_channel);
}

// reference: conv2D.cc:465-559.  out is not resized; out->NumCols() is trusted (:469).
template <typename Real>
void CuMatrixBase<Real>::Maxpool_prop(int32 in_height, int32 in_width, int32 pool_height_dim,
                                      int32 pool_width_dim, int32 pool_channel_dim, bool overlap,
                                      bool overlap2D, CuMatrixBase<Real> *out) const {
  KALDI_ASSERT(out != NULL);
  KALDI_ASSERT(out->NumRows() == NumRows());
  Launch l(__func__);
  cudaF_maxpool_prop_s(Str(), this->data_, this->Dim(), out->Data(), out->Dim(), in_height, in_width,
                       pool_height_dim, pool_width_dim, pool_channel_dim, PoolMode(overlap, overlap2D));
}

// reference: conv2D.cc:565-684.  in_deriv is resized (kSetZero) only when its shape
// differs (:571-573); matching elements get "dest = err", the rest is left untouched.
// The reference's GPU branch lacks an "else" so overlap=true ALSO launches the plain
// kernel with out-of-range indices (:581-589); that defect is not copied.
template <typename Real>
void CuMatrixBase<Real>::Maxpool_backprop(const CuMatrixBase<Real> &out_value,
                                          const CuMatrixBase<Real> &out_deriv,
                                          CuMatrix<Real> *in_deriv, int32 in_height, int32 in_width,
                                          int32 pool_height_dim, int32 pool_width_dim,
                                          int32 pool_channel_dim, bool overlap, bool overlap2D) const {
  KALDI_ASSERT(in_deriv != NULL);
  if ((in_deriv->NumRows() != NumRows()) || (in_deriv->NumCols() != NumCols()))
    in_deriv->Resize(NumRows(), NumCols(), kSetZero);
  if (overlap || overlap2D) KALDI_ASSERT(pool_height_dim == 1 && pool_width_dim == 1);
  KALDI_ASSERT(out_value.NumRows() == NumRows() && out_deriv.NumRows() == NumRows() &&
               out_value.NumCols() == out_deriv.NumCols());
  Launch l(__func__);
  cudaF_maxpool_backprop_s(Str(), this->data_, this->Dim(), out_value.Data(), out_value.Dim(),
                           out_deriv.Data(), out_deriv.Dim(), in_deriv->Data(), in_deriv->Dim(),
                           in_height, in_width, pool_height_dim, pool_width_dim, pool_channel_dim,
                           PoolMode(overlap, overlap2D), 0);
}

// reference: conv2D.cc:687-727.  Only the unregistered ConvolutionComponentContainer
// calls it (out of scope); kept so the class declaration matches.
namespace {
__global__ void mod_permute_channels_kernel(float *comp, ::MatrixDim cd, float *container,
                                            ::MatrixDim kd, int comp_idx, int num_component, int hw,
                                            bool to_container) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)cd.rows * cd.cols) return;
  int i = (int)(t / cd.cols), j = (int)(t % cd.cols);
  int chan = j / hw, pos = j % hw;
  size_t ci = (size_t)i * cd.stride + j;
  size_t ki = (size_t)i * kd.stride + (size_t)(chan * num_component + comp_idx) * hw + pos;
  if (to_container) container[ki] = comp[ci];
  else comp[ci] = container[ki];
}
}  // namespace

template <typename Real>
void CuMatrixBase<Real>::ModPermuteChannel(int32 comp_idx, int32 num_component, int32 in_height,
                                           int32 in_width, CuMatrixBase<Real> *container,
                                           bool fromCompToContainer) {
  KALDI_ASSERT(container != NULL);
  if (NumRows() == 0 || NumCols() == 0) return;
  Launch l(__func__);
  long long total = (long long)NumRows() * NumCols();
  KCNN_LAUNCH(mod_permute_channels_kernel, kcnn::ceil_div_u(total, 256), 256, 0, Str(), this->data_,
              this->Dim(), container->Data(), container->Dim(), comp_idx, num_component,
              in_height * in_width, fromCompToContainer);
}

// BaseFloat only: the reference hard-codes CuMatrix<BaseFloat> temporaries inside the
// Real template (conv2D.cc:81, 96, 138), so only float was ever usable.
template void CuMatrixBase<float>::Conv2D(const CuMatrixBase<float> &, int32, int32, int32, int32,
                                          int32, int32, CuMatrixBase<float> *, bool) const;
template void CuMatrixBase<float>::AddMatRepVec(const CuVectorBase<float> &, int32) const;
template void CuMatrixBase<float>::FlipMat(int32, int32, int32, int32, CuMatrix<float> *) const;
template void CuMatrixBase<float>::PaddingZero(int32, int32, int32, int32, int32, CuMatrix<float> *) const;
template void CuMatrixBase<float>::TpBlock(int32, int32, CuMatrix<float> *) const;
template void CuMatrixBase<float>::TpInsideBlock(int32, int32, CuMatrix<float> *) const;
template void CuMatrixBase<float>::ModPermuteRow(int32, int32, CuMatrix<float> *) const;
template void CuMatrixBase<float>::Maxpool_prop(int32, int32, int32, int32, int32, bool, bool,
                                                CuMatrixBase<float> *) const;
template void CuMatrixBase<float>::Maxpool_backprop(const CuMatrixBase<float> &,
                                                    const CuMatrixBase<float> &, CuMatrix<float> *,
                                                    int32, int32, int32, int32, int32, bool, bool) const;
template void CuMatrixBase<float>::ModPermuteChannel(int32, int32, int32, int32,
                                                     CuMatrixBase<float> *, bool);

}  // namespace kaldi
