// nnet0/fully-connected-component.cc -- FullyConnectedComponent for the B200 build.
// Follows reference src/nnet0/nnet-component-nnet0.cc:980-1150.  Propagate / Backprop
// are inherited from AffineComponent (nnet2/nnet-component.cc:1216-1258).

#include <sstream>

#include "nnet0/nnet-component-nnet0.h"
#include "nnet0/component-fields.h"
#include "util/common-utils.h"
#include "cnsl-cu-kernels.h"

namespace cnsl {
namespace nnet0 {

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }
static inline int Math() { return CuDevice::Instantiate().MathMode(); }

// ---- construction, configuration, (de)serialisation ---------------------------------------------
// What the reference does here (nnet0/nnet-component-nnet0.cc:980-1131) is fixed by its files: the model
// stream is <LearningRate> <LinearParams> <BiasParams> <WeightDecay> <Momentum> <PrevGrad> (no
// <IsGradient>), the config keys are learning-rate / weight-decay / momentum / matrix / input-dim /
// output-dim / param-stddev / bias-stddev.  Both are declared once, as FieldLists, and walked.

FieldList FullyConnectedComponent::StreamFields() {
  FieldList f;
  f.Float("learning-rate", "<LearningRate>", &learning_rate_)
      .Matrix("<LinearParams>", &linear_params_)
      .Vector("<BiasParams>", &bias_params_)
      .Float("weight-decay", "<WeightDecay>", &weight_decay_)
      .Float("momentum", "<Momentum>", &momentum_)
      .Matrix("<PrevGrad>", &prev_grad_);
  return f;
}

void FullyConnectedComponent::Read(std::istream &is, bool binary) {
  const FieldList f = StreamFields();
  ExpectOneOrTwoTokens(is, binary, "<" + Type() + ">", f.FirstToken());
  f.Read(is, binary, /*skip_first_token=*/true);
  ExpectToken(is, binary, "</" + Type() + ">");
}

void FullyConnectedComponent::Write(std::ostream &os, bool binary) const {
  WriteToken(os, binary, "<" + Type() + ">");
  const_cast<FullyConnectedComponent *>(this)->StreamFields().Write(os, binary);
  WriteToken(os, binary, "</" + Type() + ">");
}

void FullyConnectedComponent::SetHyper(BaseFloat learning_rate, BaseFloat weight_decay, BaseFloat momentum) {
  UpdatableComponent::Init(learning_rate);
  weight_decay_ = weight_decay;
  momentum_ = momentum;
}

// Random start (reference :1022-1045): W ~ N(0, param_stddev^2); the bias is the CONSTANT bias_stddev;
// weight decay and momentum must be positive; the momentum matrix starts at zero.
void FullyConnectedComponent::Init(BaseFloat learning_rate, int32 input_dim, int32 output_dim,
                                   BaseFloat param_stddev, BaseFloat bias_stddev,
                                   BaseFloat weight_decay, BaseFloat momentum) {
  KALDI_ASSERT(input_dim > 0 && output_dim > 0 && param_stddev >= 0.0);
  KALDI_ASSERT(weight_decay > 0.0 && momentum > 0.0);
  SetHyper(learning_rate, weight_decay, momentum);
  linear_params_.Resize(output_dim, input_dim, kUndefined);
  linear_params_.SetRandn();
  linear_params_.Scale(param_stddev);
  bias_params_.Resize(output_dim, kUndefined);
  bias_params_.Set(bias_stddev);
  prev_grad_.Resize(output_dim, input_dim, kSetZero);
}

// Start from a [W | b] matrix file (reference :1048-1064).  The reference seeds the momentum matrix
// with the weights themselves; kept, since the first update depends on it.
void FullyConnectedComponent::Init(BaseFloat learning_rate, BaseFloat weight_decay,
                                   BaseFloat momentum, std::string matrix_filename) {
  SetHyper(learning_rate, weight_decay, momentum);
  CuMatrix<BaseFloat> w_and_b;
  ReadKaldiObject(matrix_filename, &w_and_b);
  KALDI_ASSERT(w_and_b.NumCols() >= 2);
  const int32 rows = w_and_b.NumRows(), cols = w_and_b.NumCols() - 1;
  linear_params_ = w_and_b.Range(0, rows, 0, cols);
  prev_grad_ = linear_params_;
  bias_params_.Resize(rows, kUndefined);
  bias_params_.CopyColFromMat(w_and_b, cols);
}

// reference :1066-1100.  Unlike the convolution, weight-decay / momentum of the config line ARE applied.
void FullyConnectedComponent::InitFromString(std::string args) {
  const std::string line(args);
  struct {
    BaseFloat learning_rate, weight_decay, momentum, param_stddev, bias_stddev;
    int32 input_dim, output_dim;
  } o = {learning_rate_, weight_decay_, momentum_, -1.0f, 1.0f, -1, -1};
  std::string matrix;
  FieldList keys;
  keys.Float("learning-rate", NULL, &o.learning_rate, FieldList::kOptional)
      .Float("weight-decay", NULL, &o.weight_decay, FieldList::kOptional)
      .Float("momentum", NULL, &o.momentum, FieldList::kOptional)
      .Int("input-dim", NULL, &o.input_dim, FieldList::kOptional)
      .Int("output-dim", NULL, &o.output_dim, FieldList::kOptional)
      .Float("param-stddev", NULL, &o.param_stddev, FieldList::kOptional)
      .Float("bias-stddev", NULL, &o.bias_stddev, FieldList::kOptional);
  keys.ParseConfig(&args);
  const bool from_matrix = ParseFromString("matrix", &args, &matrix);
  if (from_matrix) {
    Init(o.learning_rate, o.weight_decay, o.momentum, matrix);
    KALDI_ASSERT((o.input_dim < 0 || o.input_dim == InputDim()) && "input-dim mismatch vs. matrix.");
    KALDI_ASSERT((o.output_dim < 0 || o.output_dim == OutputDim()) && "output-dim mismatch vs. matrix.");
  } else {
    if (o.input_dim < 0 || o.output_dim < 0) KALDI_ERR << "Bad initializer " << line;
    if (o.param_stddev < 0) o.param_stddev = 1.0 / std::sqrt(static_cast<BaseFloat>(o.input_dim));
    Init(o.learning_rate, o.input_dim, o.output_dim, o.param_stddev, o.bias_stddev, o.weight_decay,
         o.momentum);
  }
  if (!args.empty()) KALDI_ERR << "Could not process these elements in initializer: " << args;
}

// Root mean square of the entries (what nnet2's Info() strings call "stddev").
static BaseFloat Rms(const CuMatrixBase<BaseFloat> &m) {
  return std::sqrt(TraceMatMat(m, m, kTrans) / (static_cast<BaseFloat>(m.NumRows()) * m.NumCols()));
}

// Same fields as the reference prints (:1102-1119): dims, parameter RMS, then the hyper-parameters.
std::string FullyConnectedComponent::Info() const {
  std::ostringstream os;
  os << Type() << ", input-dim=" << InputDim() << ", output-dim=" << OutputDim()
     << ", linear-params-stddev=" << Rms(linear_params_)
     << ", bias-params-stddev=" << std::sqrt(VecVec(bias_params_, bias_params_) / bias_params_.Dim()) << ", "
     << const_cast<FullyConnectedComponent *>(this)->StreamFields().Describe();
  return os.str();
}

Component *FullyConnectedComponent::Copy() const {
  FullyConnectedComponent *c = new FullyConnectedComponent();
  // the state is exactly what the stream carries, plus the gradient flag
  c->SetHyper(learning_rate_, weight_decay_, momentum_);
  c->linear_params_ = linear_params_;
  c->bias_params_ = bias_params_;
  c->prev_grad_ = prev_grad_;
  c->is_gradient_ = is_gradient_;
  return c;
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
