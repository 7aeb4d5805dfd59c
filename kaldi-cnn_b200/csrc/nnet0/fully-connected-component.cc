// nnet0/fully-connected-component.cc -- FullyConnectedComponent for the B200 build.
// Follows reference src/nnet0/nnet-component-nnet0.cc:980-1150.  Propagate / Backprop
// are inherited from AffineComponent (nnet2/nnet-component.cc:1216-1258).

#include <sstream>

#include "nnet0/nnet-component-nnet0.h"
#include "util/common-utils.h"
#include "cnsl-cu-kernels.h"

namespace cnsl {
namespace nnet0 {

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }
static inline int Math() { return CuDevice::Instantiate().MathMode(); }

// reference :980-999 (no <IsGradient> in this component's stream)
void FullyConnectedComponent::Read(std::istream &is, bool binary) {
  const std::string beg = "<" + Type() + ">", end = "</" + Type() + ">";
  ExpectOneOrTwoTokens(is, binary, beg, "<LearningRate>");
  ReadBasicType(is, binary, &learning_rate_);
  ExpectToken(is, binary, "<LinearParams>");
  linear_params_.Read(is, binary);
  ExpectToken(is, binary, "<BiasParams>");
  bias_params_.Read(is, binary);
  ExpectToken(is, binary, "<WeightDecay>");
  ReadBasicType(is, binary, &weight_decay_);
  ExpectToken(is, binary, "<Momentum>");
  ReadBasicType(is, binary, &momentum_);
  ExpectToken(is, binary, "<PrevGrad>");
  prev_grad_.Read(is, binary);
  ExpectToken(is, binary, end);
}

// reference :1001-1020
void FullyConnectedComponent::Write(std::ostream &os, bool binary) const {
  const std::string beg = "<" + Type() + ">", end = "</" + Type() + ">";
  WriteToken(os, binary, beg);
  WriteToken(os, binary, "<LearningRate>");
  WriteBasicType(os, binary, learning_rate_);
  WriteToken(os, binary, "<LinearParams>");
  linear_params_.Write(os, binary);
  WriteToken(os, binary, "<BiasParams>");
  bias_params_.Write(os, binary);
  WriteToken(os, binary, "<WeightDecay>");
  WriteBasicType(os, binary, weight_decay_);
  WriteToken(os, binary, "<Momentum>");
  WriteBasicType(os, binary, momentum_);
  WriteToken(os, binary, "<PrevGrad>");
  prev_grad_.Write(os, binary);
  WriteToken(os, binary, end);
}

// reference :1022-1045: the bias is CONSTANT bias_stddev (not random), wd and momentum
// must be positive.
void FullyConnectedComponent::Init(BaseFloat learning_rate, int32 input_dim, int32 output_dim,
                                   BaseFloat param_stddev, BaseFloat bias_stddev,
                                   BaseFloat weight_decay, BaseFloat momentum) {
  UpdatableComponent::Init(learning_rate);
  KALDI_ASSERT(input_dim > 0 && output_dim > 0);
  linear_params_.Resize(output_dim, input_dim);
  bias_params_.Resize(output_dim);
  KALDI_ASSERT(output_dim > 0 && input_dim > 0 && param_stddev >= 0.0);
  linear_params_.SetRandn();
  linear_params_.Scale(param_stddev);
  bias_params_.SetZero();
  bias_params_.Add(bias_stddev);
  weight_decay_ = weight_decay;
  KALDI_ASSERT(weight_decay_ > 0.0);
  momentum_ = momentum;
  KALDI_ASSERT(momentum_ > 0.0);
  prev_grad_.Resize(output_dim, input_dim);
  prev_grad_.SetZero();
}

// reference :1048-1064 (prev_grad_ is seeded with the weights there, App. C.3; kept)
void FullyConnectedComponent::Init(BaseFloat learning_rate, BaseFloat weight_decay,
                                   BaseFloat momentum, std::string matrix_filename) {
  UpdatableComponent::Init(learning_rate);
  weight_decay_ = weight_decay;
  momentum_ = momentum;
  CuMatrix<BaseFloat> mat;
  ReadKaldiObject(matrix_filename, &mat);
  KALDI_ASSERT(mat.NumCols() >= 2);
  int32 input_dim = mat.NumCols() - 1, output_dim = mat.NumRows();
  linear_params_.Resize(output_dim, input_dim);
  bias_params_.Resize(output_dim);
  linear_params_.CopyFromMat(mat.Range(0, output_dim, 0, input_dim));
  bias_params_.CopyColFromMat(mat, input_dim);
  prev_grad_.Resize(output_dim, input_dim);
  prev_grad_.CopyFromMat(mat.Range(0, output_dim, 0, input_dim));
}

// reference :1066-1100 (unlike the convolution, weight-decay / momentum ARE applied)
void FullyConnectedComponent::InitFromString(std::string args) {
  std::string orig_args(args);
  std::string matrix_filename;
  BaseFloat learning_rate = learning_rate_;
  BaseFloat weight_decay = weight_decay_, momentum = momentum_;
  int32 input_dim = -1, output_dim = -1;
  ParseFromString("learning-rate", &args, &learning_rate);   // optional.
  ParseFromString("weight-decay", &args, &weight_decay);
  ParseFromString("momentum", &args, &momentum);
  if (ParseFromString("matrix", &args, &matrix_filename)) {
    Init(learning_rate, weight_decay, momentum, matrix_filename);
    if (ParseFromString("input-dim", &args, &input_dim))
      KALDI_ASSERT(input_dim == InputDim() && "input-dim mismatch vs. matrix.");
    if (ParseFromString("output-dim", &args, &output_dim))
      KALDI_ASSERT(output_dim == OutputDim() && "output-dim mismatch vs. matrix.");
  } else {
    bool ok = true;
    ok = ok && ParseFromString("input-dim", &args, &input_dim);
    ok = ok && ParseFromString("output-dim", &args, &output_dim);
    BaseFloat param_stddev = 1.0 / std::sqrt(input_dim), bias_stddev = 1.0;
    ParseFromString("param-stddev", &args, &param_stddev);
    ParseFromString("bias-stddev", &args, &bias_stddev);
    if (!ok) KALDI_ERR << "Bad initializer " << orig_args;
    Init(learning_rate, input_dim, output_dim, param_stddev, bias_stddev, weight_decay, momentum);
  }
  if (!args.empty()) KALDI_ERR << "Could not process these elements in initializer: " << args;
}

// reference :1102-1119
std::string FullyConnectedComponent::Info() const {
  std::stringstream stream;
  BaseFloat linear_params_size = static_cast<BaseFloat>(linear_params_.NumRows()) *
                                 static_cast<BaseFloat>(linear_params_.NumCols());
  BaseFloat linear_stddev = std::sqrt(TraceMatMat(linear_params_, linear_params_, kTrans) / linear_params_size),
            bias_stddev = std::sqrt(VecVec(bias_params_, bias_params_) / bias_params_.Dim());
  stream << Type() << ", input-dim=" << InputDim() << ", output-dim=" << OutputDim()
         << ", linear-params-stddev=" << linear_stddev << ", bias-params-stddev=" << bias_stddev
         << ", learning-rate=" << LearningRate() << ", weight-decay=" << weight_decay_
         << ", momentum=" << momentum_;
  return stream.str();
}

// reference :1121-1131
Component *FullyConnectedComponent::Copy() const {
  FullyConnectedComponent *ans = new FullyConnectedComponent();
  ans->learning_rate_ = learning_rate_;
  ans->linear_params_ = linear_params_;
  ans->bias_params_ = bias_params_;
  ans->weight_decay_ = weight_decay_;
  ans->momentum_ = momentum_;
  ans->prev_grad_ = prev_grad_;
  ans->is_gradient_ = is_gradient_;
  return ans;
}

// The SGD step of reference :1133-1143 on an already computed gradient:
//   lr = learning_rate_/N ; bias += lr * colsum(out_deriv)
//   prev = m*prev ; prev += (-lr*wd) * W ; prev += lr * out_deriv^T in ; W += prev
void FullyConnectedComponent::ApplyGradient(int32 total_num_samples) {
  EnsureGradBuffers();
  KALDI_ASSERT(total_num_samples > 0);
  double learning_rate = learning_rate_ / total_num_samples;
  BaseFloat a_decay = -1 * learning_rate * weight_decay_, a_grad = learning_rate;
  if (prev_grad_.NumRows() != linear_params_.NumRows() || prev_grad_.NumCols() != linear_params_.NumCols())
    prev_grad_.Resize(linear_params_.NumRows(), linear_params_.NumCols(), kSetZero);
  cudaF_vec_axpy(Str(), bias_params_.Data(), b_grad_.data, bias_params_.Dim(), a_grad);
  ::MatrixDim gd = {w_grad_.rows, w_grad_.cols, w_grad_.stride};
  cudaF_sgd_momentum_update(Str(), linear_params_.Data(), linear_params_.Dim(), prev_grad_.Data(),
                            prev_grad_.Dim(), w_grad_.data, gd, momentum_, a_decay, a_grad);
  CU_SAFE_CALL(cudaGetLastError());
}

bool FullyConnectedComponent::GetStepTarget(int32 num_rows, StepTarget *t) {
  if (is_gradient_ || num_rows <= 0) return false;
  EnsureGradBuffers();
  if (prev_grad_.NumRows() != linear_params_.NumRows() || prev_grad_.NumCols() != linear_params_.NumCols())
    prev_grad_.Resize(linear_params_.NumRows(), linear_params_.NumCols(), kSetZero);
  double learning_rate = learning_rate_ / num_rows;                 // reference :1136
  t->w = linear_params_.Data(); t->wd = linear_params_.Dim();
  t->prev = prev_grad_.Data(); t->pd = prev_grad_.Dim();
  t->bias = bias_params_.Data(); t->bias_dim = bias_params_.Dim();
  t->w_grad = w_grad_.data; t->gd.rows = w_grad_.rows; t->gd.cols = w_grad_.cols; t->gd.stride = w_grad_.stride;
  t->b_grad = b_grad_.data;
  t->deferred = deferred_;
  t->momentum = momentum_;
  t->a_decay = -1 * learning_rate * weight_decay_;
  t->a_grad = learning_rate;
  return true;
}

// reference :1133-1150
void FullyConnectedComponent::UpdateSimple(const CuMatrixBase<BaseFloat> &in_value,
                                           const CuMatrixBase<BaseFloat> &out_deriv) {
  if (!deferred_ && in_value.NumRows() > 0) {
    // Single-GPU step: the update runs in the epilogue of the weight-gradient GEMM, the
    // gradient matrix is never written (16 instead of 24 bytes of HBM traffic per weight).
    double learning_rate = learning_rate_ / in_value.NumRows();
    BaseFloat a_decay = -1 * learning_rate * weight_decay_, a_grad = learning_rate;
    if (prev_grad_.NumRows() != linear_params_.NumRows() || prev_grad_.NumCols() != linear_params_.NumCols())
      prev_grad_.Resize(linear_params_.NumRows(), linear_params_.NumCols(), kSetZero);
    if (cudaF_affine_wgrad_sgd(Str(), Math(), in_value.Data(), in_value.Dim(), out_deriv.Data(),
                               out_deriv.Dim(), linear_params_.Data(), linear_params_.Dim(),
                               prev_grad_.Data(), prev_grad_.Dim(), bias_params_.Data(), momentum_,
                               a_decay, a_grad)) {
      CU_SAFE_CALL(cudaGetLastError());
      return;
    }
  }
  ComputeGradient(in_value, out_deriv);
  if (!deferred_) ApplyGradient(in_value.NumRows());
}

}  // namespace nnet0
}  // namespace cnsl
