// nnet0/convolution-component.cc -- ConvolutionComponent for the B200 build.
// Follows reference src/nnet0/nnet-component-nnet0.cc:178-777 (cited per function).

#include <sstream>

#include "nnet0/nnet-component-nnet0.h"
#include "nnet0/component-fields.h"
#include "util/common-utils.h"
#include "cnsl-cu-kernels.h"

namespace cnsl {
namespace nnet0 {

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }
static inline int Math() { return CuDevice::Instantiate().MathMode(); }

ConvolutionComponent::ConvolutionComponent()     // defaults of reference .h:27
    : is_gradient_(false), in_height_(0), in_width_(0), in_channel_(0), in_pad_height_(0),
      in_pad_width_(0), kernel_height_(0), kernel_width_(0), stride_(1), group_(0), out_height_(0),
      out_width_(0), weight_decay_(0.0002), momentum_(0.9), deferred_(false),
      grad_external_(false), workspace_rows_(-1), staged_src_(NULL), staged_rows_(0), staged_stride_(0),
      input_persists_(false) {}

// reference :178-195 (the copy constructor leaves prev_grad_ empty, App. C.3; here it
// is sized and zeroed so that a copied component can be updated).
ConvolutionComponent::ConvolutionComponent(const ConvolutionComponent &c)
    : UpdatableComponent(c), linear_params_(c.linear_params_), bias_params_(c.bias_params_),
      is_gradient_(c.is_gradient_), in_height_(c.in_height_), in_width_(c.in_width_),
      in_channel_(c.in_channel_), in_pad_height_(c.in_pad_height_), in_pad_width_(c.in_pad_width_),
      kernel_height_(c.kernel_height_), kernel_width_(c.kernel_width_), stride_(c.stride_),
      group_(c.group_), out_height_(c.out_height_), out_width_(c.out_width_),
      weight_decay_(c.weight_decay_), momentum_(c.momentum_), deferred_(false),
      grad_external_(false), workspace_rows_(-1), staged_src_(NULL), staged_rows_(0), staged_stride_(0),
      input_persists_(false) {
  prev_grad_.Resize(linear_params_.NumRows(), linear_params_.NumCols(), kSetZero);
}

// reference :199-229
ConvolutionComponent::ConvolutionComponent(const CuMatrix<BaseFloat> &linear_params,
                                           const CuVector<BaseFloat> &bias_params,
                                           BaseFloat learning_rate, int32 in_height, int32 in_width,
                                           int32 in_channels, int32 in_pad_height,
                                           int32 in_pad_width, int32 kernel_height,
                                           int32 kernel_width, int32 stride, int32 group,
                                           int32 out_height, int32 out_width, BaseFloat weight_decay,
                                           BaseFloat momentum)
    : UpdatableComponent(learning_rate), linear_params_(linear_params), bias_params_(bias_params),
      is_gradient_(false), in_height_(in_height), in_width_(in_width), in_channel_(in_channels),
      in_pad_height_(in_pad_height), in_pad_width_(in_pad_width), kernel_height_(kernel_height),
      kernel_width_(kernel_width), stride_(stride), group_(group), out_height_(out_height),
      out_width_(out_width), weight_decay_(weight_decay), momentum_(momentum), deferred_(false),
      grad_external_(false), workspace_rows_(-1), staged_src_(NULL), staged_rows_(0), staged_stride_(0),
      input_persists_(false) {
  KALDI_ASSERT(linear_params.NumCols() == bias_params.Dim() && bias_params.Dim() != 0);
  prev_grad_.Resize(linear_params_.NumRows(), linear_params_.NumCols(), kSetZero);
}

// ---- geometry, configuration ---------------------------------------------------------------------
// One table for the scalar members: config key (nnet.config line, reference :323-385), model-file token
// (reference :556-666) and the member.  The ORDER is the model stream's.  "in-pad-*" are the only optional
// geometry keys.  Weight decay and momentum are in the stream but -- see InitFromString -- not settable
// from a config line.
FieldList ConvolutionComponent::StreamFields() {
  const FieldList::Need opt = FieldList::kOptional;
  FieldList f;
  f.Int("in-height", "<in_height>", &in_height_)
      .Int("in-width", "<in_width>", &in_width_)
      .Int("in-channel", "<in_channel>", &in_channel_)
      .Int("kernel-height", "<kernel_height>", &kernel_height_)
      .Int("kernel-width", "<kernel_width>", &kernel_width_)
      .Int("stride", "<stride>", &stride_)
      .Int("in-pad-height", "<padding_height>", &in_pad_height_, opt)
      .Int("in-pad-width", "<padding_width>", &in_pad_width_, opt)
      .Int("group", "<group>", &group_)
      .Int("out-height", "<out_height>", &out_height_)
      .Int("out-width", "<out_width>", &out_width_)
      .Float("learning-rate", "<LearningRate>", &learning_rate_)
      .Float(NULL, "<WeightDecay>", &weight_decay_)
      .Float(NULL, "<Momentum>", &momentum_)
      .Matrix("<LinearParams>", &linear_params_)
      .Vector("<BiasParams>", &bias_params_)
      .Matrix("<PrevGrad>", &prev_grad_)
      .Bool(NULL, "<IsGradient>", &is_gradient_, opt);       // [17]: optional on input
  return f;
}
static const size_t kConvOptionalFrom = 17;

// The checks of reference :232-275 / :349-360 on the shape members.  The reference parses "stride" but
// its Conv2D only implements stride 1 (cnslmat/conv2D.cc:59-60) and silently mis-sizes the output for
// anything else; that is an error here.
void ConvolutionComponent::CheckGeometry() const {
  KALDI_ASSERT(in_pad_height_ >= 0 && "in-pad-height should be positive");
  KALDI_ASSERT(in_pad_width_ >= 0 && "in-pad-width should be positive");
  KALDI_ASSERT(stride_ != 0);
  KALDI_ASSERT(out_height_ == 1 + (in_height_ + 2 * in_pad_height_ - kernel_height_) / stride_ &&
               "out-height == 1 + (in-height + 2*in-pad-height - kernel-height) / stride");
  KALDI_ASSERT(out_width_ == 1 + (in_width_ + 2 * in_pad_width_ - kernel_width_) / stride_ &&
               "out-width == 1 + (in-width + 2*in-pad-width - kernel-width) / stride");
  if (stride_ != 1)
    KALDI_ERR << "ConvolutionComponent: stride=" << stride_ << " is not implemented (only stride 1)";
}

void ConvolutionComponent::SetShape(BaseFloat learning_rate, int32 in_height, int32 in_width,
                                    int32 in_channels, int32 in_pad_height, int32 in_pad_width,
                                    int32 kernel_height, int32 kernel_width, int32 stride, int32 group,
                                    int32 out_height, int32 out_width, BaseFloat weight_decay,
                                    BaseFloat momentum) {
  in_height_ = in_height; in_width_ = in_width; in_channel_ = in_channels;
  in_pad_height_ = in_pad_height; in_pad_width_ = in_pad_width;
  kernel_height_ = kernel_height; kernel_width_ = kernel_width;
  stride_ = stride; group_ = group;
  out_height_ = out_height; out_width_ = out_width;
  weight_decay_ = weight_decay; momentum_ = momentum;
  CheckGeometry();
  UpdatableComponent::Init(learning_rate);
}

// Random start (reference :232-275): W ~ N(0, param_stddev^2), bias ~ N(0, bias_stddev^2) (random, unlike
// the fully connected layer), momentum matrix zero.  Weights are drawn before the bias.
void ConvolutionComponent::Init(BaseFloat learning_rate, int32 in_height, int32 in_width,
                                int32 in_channels, int32 in_pad_height, int32 in_pad_width,
                                int32 kernel_height, int32 kernel_width, int32 stride, int32 group,
                                int32 out_height, int32 out_width, BaseFloat param_stddev,
                                BaseFloat bias_stddev, BaseFloat weight_decay, BaseFloat momentum) {
  SetShape(learning_rate, in_height, in_width, in_channels, in_pad_height, in_pad_width, kernel_height,
           kernel_width, stride, group, out_height, out_width, weight_decay, momentum);
  KALDI_ASSERT(param_stddev >= 0.0);
  linear_params_.Resize(KernelDim(), group, kUndefined);
  linear_params_.SetRandn();
  linear_params_.Scale(param_stddev);
  bias_params_.Resize(group, kUndefined);
  bias_params_.SetRandn();
  bias_params_.Scale(bias_stddev);
  prev_grad_.Resize(KernelDim(), group, kSetZero);
}

// Start from a matrix file (reference :277-321) holding [KernelDim()+1 x group]: the weights, then one
// row of biases.  (The reference sizes the bias as kernel_dim and reads a column, App. C.3 -- a defect of
// a path nothing uses; here the bias is the last ROW, one value per output map.)
void ConvolutionComponent::Init(BaseFloat learning_rate, int32 in_height, int32 in_width,
                                int32 in_channels, int32 in_pad_height, int32 in_pad_width,
                                int32 kernel_height, int32 kernel_width, int32 stride, int32 group,
                                int32 out_height, int32 out_width, BaseFloat weight_decay,
                                BaseFloat momentum, std::string matrix_filename) {
  SetShape(learning_rate, in_height, in_width, in_channels, in_pad_height, in_pad_width, kernel_height,
           kernel_width, stride, group, out_height, out_width, weight_decay, momentum);
  Matrix<BaseFloat> w_then_b;
  ReadKaldiObject(matrix_filename, &w_then_b);
  KALDI_ASSERT(w_then_b.NumCols() == Group() && w_then_b.NumRows() == KernelDim() + 1);
  Matrix<BaseFloat> w(KernelDim(), Group(), kUndefined);
  Vector<BaseFloat> b(Group(), kUndefined);
  for (int32 c = 0; c < Group(); c++) {
    for (int32 r = 0; r < KernelDim(); r++) w(r, c) = w_then_b(r, c);
    b(c) = w_then_b(KernelDim(), c);
  }
  linear_params_ = w;
  bias_params_ = b;
  prev_grad_.Resize(KernelDim(), group, kSetZero);
}

// Config line (reference :323-385).  A quirk of the reference is kept: "weight-decay"
// and "momentum" are accepted on the line but NOT applied -- a freshly initialised convolution runs with
// the constructor defaults (0.0002 / 0.9) until SetWeightDecay / SetMomentum / Read change them; models
// trained with the reference depend on it.
void ConvolutionComponent::InitFromString(std::string args) {
  const std::string line(args);
  ConvolutionComponent parsed;                         // the table's members, filled from the line
  parsed.learning_rate_ = learning_rate_;
  if (!parsed.StreamFields().ParseConfig(&args)) KALDI_ERR << "Bad initializer " << line;
  parsed.CheckGeometry();
  std::string matrix;
  BaseFloat param_stddev = 1.0 / std::sqrt(static_cast<BaseFloat>(parsed.kernel_height_ * parsed.kernel_width_)),
            bias_stddev = 1.0, ignored = 0.0;
  const bool from_matrix = ParseFromString("matrix", &args, &matrix);
  if (!from_matrix) {
    ParseFromString("param-stddev", &args, &param_stddev);
    ParseFromString("bias-stddev", &args, &bias_stddev);
  }
  ParseFromString("weight-decay", &args, &ignored);     // accepted, not applied
  ParseFromString("momentum", &args, &ignored);
  if (!args.empty()) KALDI_ERR << "Could not process these elements in initializer: " << args;
  const ConvolutionComponent &g = parsed;
  if (from_matrix)
    Init(g.learning_rate_, g.in_height_, g.in_width_, g.in_channel_, g.in_pad_height_, g.in_pad_width_,
         g.kernel_height_, g.kernel_width_, g.stride_, g.group_, g.out_height_, g.out_width_, weight_decay_,
         momentum_, matrix);
  else
    Init(g.learning_rate_, g.in_height_, g.in_width_, g.in_channel_, g.in_pad_height_, g.in_pad_width_,
         g.kernel_height_, g.kernel_width_, g.stride_, g.group_, g.out_height_, g.out_width_, param_stddev,
         bias_stddev, weight_decay_, momentum_);
}

static BaseFloat Rms(const CuMatrixBase<BaseFloat> &m) {
  return std::sqrt(TraceMatMat(m, m, kTrans) / (static_cast<BaseFloat>(m.NumRows()) * m.NumCols()));
}

// The fields the reference prints (:387-421), grouped the same way.
std::string ConvolutionComponent::Info() const {
  std::ostringstream os;
  os << Type() << ", input-dim=" << InputDim() << " ( in-height=" << in_height_ << ", in-width=" << in_width_
     << ", in-channels=" << in_channel_ << "), output-dim=" << OutputDim() << " ( out-height=" << out_height_
     << ", out-width=" << out_width_ << ", group-num=" << group_ << "), kernel-dim=" << KernelDim()
     << " ( kernel-height=" << kernel_height_ << ", kernel-width=" << kernel_width_
     << "), ( padding-height=" << in_pad_height_ << ", padding-width=" << in_pad_width_
     << "), linear-params-stddev=" << Rms(linear_params_)
     << ", bias-params-stddev=" << std::sqrt(VecVec(bias_params_, bias_params_) / bias_params_.Dim())
     << ", learning-rate=" << learning_rate_ << ", weight-decay=" << weight_decay_
     << ", momentum=" << momentum_;
  return os.str();
}

// reference :423-446: [PaddingZero] -> Conv2D -> AddMatRepVec, here ONE implicit GEMM
// whose addressing supplies the zero border and whose epilogue adds the bias.
void ConvolutionComponent::Propagate(const ChunkInfo &in_info, const ChunkInfo &,
                                     const CuMatrixBase<BaseFloat> &in,
                                     CuMatrixBase<BaseFloat> *out) const {
  PropagateAct(in_info, in, out, KCNN_ACT_NONE);
}

// Fusion hook: the ReLU that follows every convolution of egs/exp/nnet/nnet.config applied in
// the GEMM epilogue (one launch and one pass over the activation less per layer).
bool ConvolutionComponent::PropagateRelu(const ChunkInfo &in_info, const ChunkInfo &,
                                         const CuMatrixBase<BaseFloat> &in,
                                         CuMatrixBase<BaseFloat> *out) const {
  PropagateAct(in_info, in, out, KCNN_ACT_RELU);
  return true;
}

void ConvolutionComponent::PropagateAct(const ChunkInfo &in_info, const CuMatrixBase<BaseFloat> &in,
                                        CuMatrixBase<BaseFloat> *out, int act) const {
  KALDI_ASSERT(in.NumCols() == InputDim() && out != NULL);
  KALDI_ASSERT(in.NumRows() == in_info.NumChunks() || in_pad_height_ + in_pad_width_ == 0);
  KALDI_ASSERT(out->NumRows() == in.NumRows() && out->NumCols() == OutputDim());
  CuDevice::Instantiate().RequireEnabled("ConvolutionComponent::Propagate");
  Timer tim;
  // When the caller has promised that in_value persists until Backprop (SetInputPersists), the
  // channels-last staging copy of `in` that the TMA path makes anyway is kept in a
  // per-component buffer, and Backprop does not pack in_value a second time.
  BaseFloat *staging = NULL;
  staged_src_ = NULL;
  if (input_persists_ && Math() == KCNN_MATH_TF32_TC) {
    size_t floats = kcnn_conv2d_staging_floats(in.NumRows(), in_height_, in_width_, in_channel_,
                                               in_pad_height_, in_pad_width_, kernel_height_,
                                               kernel_width_, group_);
    if (floats > 0) {
      if (static_cast<size_t>(staged_in_.Dim()) != floats) staged_in_.Resize(floats, kUndefined);
      staging = staged_in_.Data();
    }
  }
  int staged = cudaF_conv2d_fprop_act(Str(), Math(), in.Data(), in.Dim(), linear_params_.Data(),
                                      linear_params_.Dim(), bias_params_.Data(), out->Data(),
                                      out->Dim(), in_height_, in_width_, in_channel_, in_pad_height_,
                                      in_pad_width_, kernel_height_, kernel_width_, group_, 1, staging, act);
  if (staged) {
    staged_src_ = in.Data(); staged_rows_ = in.NumRows(); staged_stride_ = in.Stride();
  }
  CU_SAFE_CALL(cudaGetLastError());
  CuDevice::Instantiate().AccuProfile(__func__, tim.Elapsed());
}

void ConvolutionComponent::Scale(BaseFloat scale) {      // reference :448-451
  linear_params_.Scale(scale);
  bias_params_.Scale(scale);
}

void ConvolutionComponent::Add(BaseFloat alpha, const UpdatableComponent &other_in) {   // :453-459
  const ConvolutionComponent *other = dynamic_cast<const ConvolutionComponent *>(&other_in);
  KALDI_ASSERT(other != NULL);
  linear_params_.AddMat(alpha, other->linear_params_);
  bias_params_.AddVec(alpha, other->bias_params_);
}

// reference :461-544.  The reference picks between two data movements by comparing the
// padded-kernel and padded-out_deriv sizes (:489-497); both evaluate the same sum
//   dX[n,c,w,h] = sum_{g,kw,kh} dY[n,g,w+pw-kw,h+ph-kh] K[c,kw,kh,g]
// which is one implicit GEMM here, so the branch disappears.  Then Update, as there.
void ConvolutionComponent::Backprop(const ChunkInfo &, const ChunkInfo &,
                                    const CuMatrixBase<BaseFloat> &in_value,
                                    const CuMatrixBase<BaseFloat> &,   // out_value
                                    const CuMatrixBase<BaseFloat> &out_deriv,
                                    Component *to_update_in, CuMatrix<BaseFloat> *in_deriv) const {
  ConvolutionComponent *to_update = dynamic_cast<ConvolutionComponent *>(to_update_in);
  KALDI_ASSERT(out_deriv.NumCols() == OutputDim());
  CuDevice::Instantiate().RequireEnabled("ConvolutionComponent::Backprop");
  if (in_deriv != NULL &&
      (in_deriv->NumRows() != out_deriv.NumRows() || in_deriv->NumCols() != InputDim()))
    in_deriv->Resize(out_deriv.NumRows(), InputDim(), kUndefined);
  if (to_update == this && out_deriv.NumRows() > 0) {
    // Ordinary SGD (to_update is the component itself): input gradient, weight gradient,
    // bias gradient and -- unless the update is deferred for the data-parallel all-reduce --
    // the momentum / weight-decay step, from one staging copy of out_deriv and in_value.
    ConvolutionComponent *self = to_update;
    self->EnsureGradBuffers();
    const bool apply = !deferred_;
    double learning_rate = learning_rate_ / out_deriv.NumRows();
    BaseFloat a_decay = -1 * learning_rate * weight_decay_, a_grad = learning_rate;
    ::MatrixDim gd = {self->w_grad_.rows, self->w_grad_.cols, self->w_grad_.stride};
    ::MatrixDim idd = {0, 0, 0};
    if (in_deriv != NULL) idd = in_deriv->Dim();
    const BaseFloat *staged = NULL;
    if (input_persists_ && staged_src_ != NULL && staged_src_ == in_value.Data() && staged_rows_ == in_value.NumRows() &&
        staged_stride_ == in_value.Stride() && Math() == KCNN_MATH_TF32_TC)
      staged = staged_in_.Data();
    Timer tim;
    int done = cudaF_conv2d_backward(
        Str(), Math(), in_value.Data(), in_value.Dim(), out_deriv.Data(), out_deriv.Dim(),
        self->linear_params_.Data(), self->linear_params_.Dim(),
        in_deriv != NULL ? in_deriv->Data() : NULL, idd, self->w_grad_.data, gd, self->b_grad_.data,
        self->prev_grad_.Data(), self->prev_grad_.Dim(), self->bias_params_.Data(), apply ? 1 : 0,
        momentum_, a_decay, a_grad, staged, in_height_, in_width_, in_channel_, in_pad_height_,
        in_pad_width_, kernel_height_, kernel_width_, group_);
    if (done) {
      CU_SAFE_CALL(cudaGetLastError());
      CuDevice::Instantiate().AccuProfile(__func__, tim.Elapsed());
      return;
    }
  }
  if (in_deriv != NULL) {
    Timer tim;
    cudaF_conv2d_dgrad(Str(), Math(), out_deriv.Data(), out_deriv.Dim(), linear_params_.Data(),
                       linear_params_.Dim(), in_deriv->Data(), in_deriv->Dim(), in_height_, in_width_,
                       in_channel_, in_pad_height_, in_pad_width_, kernel_height_, kernel_width_,
                       group_);
    CU_SAFE_CALL(cudaGetLastError());
    CuDevice::Instantiate().AccuProfile(__func__, tim.Elapsed());
  }
  if (to_update != NULL) to_update->Update(in_value, out_deriv);
}

void ConvolutionComponent::SetZero(bool treat_as_gradient) {      // reference :546-554
  if (treat_as_gradient) SetLearningRate(1.0);
  linear_params_.SetZero();
  bias_params_.SetZero();
  if (treat_as_gradient) is_gradient_ = true;
}

// Model stream (reference :556-666): the table above, then -- on input only -- the <AvgInput> /
// <AvgInputCount> pair some old files carry (read and discarded, :603-610) and an optional <IsGradient>.
void ConvolutionComponent::Read(std::istream &is, bool binary) {
  const FieldList f = StreamFields();
  const std::string end = "</" + Type() + ">";
  ExpectOneOrTwoTokens(is, binary, "<" + Type() + ">", f.FirstToken());
  f.Read(is, binary, /*skip_first_token=*/true, 0, kConvOptionalFrom);
  is_gradient_ = false;
  std::string tok;
  for (ReadToken(is, binary, &tok); tok != end; ReadToken(is, binary, &tok)) {
    if (tok == "<AvgInput>") {
      CuVector<BaseFloat> unused;
      unused.Read(is, binary);
    } else if (tok == "<AvgInputCount>") {
      BaseFloat unused;
      ReadBasicType(is, binary, &unused);
    } else if (tok == "<IsGradient>") {
      ReadBasicType(is, binary, &is_gradient_);
    } else {
      KALDI_ERR << "Unexpected token " << tok << " in " << Type();
    }
  }
  workspace_rows_ = -1;
}

void ConvolutionComponent::Write(std::ostream &os, bool binary) const {
  WriteToken(os, binary, "<" + Type() + ">");
  const_cast<ConvolutionComponent *>(this)->StreamFields().Write(os, binary);
  WriteToken(os, binary, "</" + Type() + ">");
}

BaseFloat ConvolutionComponent::DotProduct(const UpdatableComponent &other_in) const {   // :670-675
  const ConvolutionComponent *other = dynamic_cast<const ConvolutionComponent *>(&other_in);
  KALDI_ASSERT(other != NULL);
  return TraceMatMat(linear_params_, other->linear_params_, kTrans) +
         VecVec(bias_params_, other->bias_params_);
}

// reference :679-705; the momentum matrix is part of the copy (the copy CONSTRUCTOR zeroes it).
Component *ConvolutionComponent::Copy() const {
  ConvolutionComponent *c = new ConvolutionComponent(*this);
  c->prev_grad_ = prev_grad_;
  return c;
}

void ConvolutionComponent::PerturbParams(BaseFloat stddev) {     // reference :706-714
  CuMatrix<BaseFloat> temp_linear_params(linear_params_);
  temp_linear_params.SetRandn();
  linear_params_.AddMat(stddev, temp_linear_params);
  CuVector<BaseFloat> temp_bias_params(bias_params_);
  temp_bias_params.SetRandn();
  bias_params_.AddVec(stddev, temp_bias_params);
}

// reference :716-736 use AffineComponent's shapes and are wrong for a convolution
// (App. C.3, off the hot path); these versions use the convolution's own shapes.
void ConvolutionComponent::SetParams(const VectorBase<BaseFloat> &bias,
                                     const MatrixBase<BaseFloat> &linear) {
  bias_params_ = bias;
  linear_params_ = linear;
  KALDI_ASSERT(bias_params_.Dim() == linear_params_.NumCols());
}

int32 ConvolutionComponent::GetParameterDim() const { return (KernelDim() + 1) * Group(); }

void ConvolutionComponent::Vectorize(VectorBase<BaseFloat> *params) const {
  KALDI_ASSERT(params->Dim() == GetParameterDim());
  Matrix<BaseFloat> w(linear_params_.NumRows(), linear_params_.NumCols());
  linear_params_.CopyToMat(&w);
  Vector<BaseFloat> b(bias_params_.Dim());
  bias_params_.CopyToVec(&b);
  int32 k = 0;
  for (int32 r = 0; r < w.NumRows(); r++)
    for (int32 c = 0; c < w.NumCols(); c++) (*params)(k++) = w(r, c);
  for (int32 i = 0; i < b.Dim(); i++) (*params)(k++) = b(i);
}

void ConvolutionComponent::UnVectorize(const VectorBase<BaseFloat> &params) {
  KALDI_ASSERT(params.Dim() == GetParameterDim());
  Matrix<BaseFloat> w(KernelDim(), Group());
  Vector<BaseFloat> b(Group());
  int32 k = 0;
  for (int32 r = 0; r < w.NumRows(); r++)
    for (int32 c = 0; c < w.NumCols(); c++) w(r, c) = params(k++);
  for (int32 i = 0; i < b.Dim(); i++) b(i) = params(k++);
  linear_params_ = w;
  bias_params_ = b;
}

// ---- gradient buffers -------------------------------------------------------

void ConvolutionComponent::EnsureGradBuffers() {
  if (grad_external_) return;
  if (w_grad_store_.NumRows() != KernelDim() || w_grad_store_.NumCols() != group_) {
    w_grad_store_.Resize(KernelDim(), group_, kUndefined);
    b_grad_store_.Resize(group_, kUndefined);
  }
  w_grad_.data = w_grad_store_.Data(); w_grad_.rows = KernelDim(); w_grad_.cols = group_;
  w_grad_.stride = w_grad_store_.Stride();
  b_grad_.data = b_grad_store_.Data(); b_grad_.rows = 1; b_grad_.cols = group_; b_grad_.stride = group_;
}

size_t ConvolutionComponent::GradientFloats() const {
  size_t stride = CuDevice::PitchInElements(group_, sizeof(BaseFloat));
  return stride * KernelDim() + stride;
}

void ConvolutionComponent::SetGradientStorage(float *base) {
  if (base == NULL) { grad_external_ = false; return; }
  int32 stride = CuDevice::PitchInElements(group_, sizeof(BaseFloat));
  w_grad_.data = base; w_grad_.rows = KernelDim(); w_grad_.cols = group_; w_grad_.stride = stride;
  b_grad_.data = base + (size_t)stride * KernelDim(); b_grad_.rows = 1; b_grad_.cols = group_;
  b_grad_.stride = group_;
  grad_external_ = true;
}

void ConvolutionComponent::SetParameterStorage(float *base) {
  const int32 rows = linear_params_.NumRows(), cols = linear_params_.NumCols(), dim = bias_params_.Dim();
  const int32 stride = CuDevice::PitchInElements(cols, sizeof(BaseFloat));
  CuMatrix<BaseFloat> w(linear_params_);
  CuVector<BaseFloat> b(bias_params_);
  if (base != NULL) {
    linear_params_.Borrow(base, rows, cols, stride);
    bias_params_.Borrow(base + (size_t)stride * rows, dim);
    linear_params_.CopyFromMat(w);
    bias_params_.CopyFromVec(b);
  } else {
    linear_params_.Swap(&w);
    bias_params_.Borrow(NULL, 0);
    bias_params_ = b;
  }
  staged_src_ = NULL;
}

std::vector<UpdatableComponent::GradBuffer> ConvolutionComponent::GradientBuffers() {
  EnsureGradBuffers();
  std::vector<GradBuffer> v;
  v.push_back(w_grad_);
  v.push_back(b_grad_);
  return v;
}

// Weight and bias gradient, reference :745-765 and the row sums of :775: one
// reduction kernel (bias) + one implicit GEMM whose output rows are already in
// linear_params_ order (no ModPermuteRow), reading in_value / out_deriv in place.
void ConvolutionComponent::ComputeGradient(const CuMatrixBase<BaseFloat> &in_value,
                                           const CuMatrixBase<BaseFloat> &out_deriv) {
  KALDI_ASSERT(in_value.NumCols() == InputDim() && out_deriv.NumCols() == OutputDim() &&
               in_value.NumRows() == out_deriv.NumRows());
  EnsureGradBuffers();
  const int32 rows = in_value.NumRows();
  if (workspace_rows_ != rows) {
    size_t bytes = kcnn_conv2d_wgrad_workspace(rows, in_height_, in_width_, in_channel_,
                                               in_pad_height_, in_pad_width_, kernel_height_,
                                               kernel_width_, group_);
    workspace_.Resize((bytes + sizeof(BaseFloat) - 1) / sizeof(BaseFloat), kUndefined);
    workspace_rows_ = rows;
  }
  ::MatrixDim gd = {w_grad_.rows, w_grad_.cols, w_grad_.stride};
  Timer tim;
  cudaF_conv2d_wgrad(Str(), Math(), in_value.Data(), in_value.Dim(), out_deriv.Data(),
                     out_deriv.Dim(), w_grad_.data, gd, b_grad_.data,
                     workspace_.Dim() ? workspace_.Data() : NULL, in_height_, in_width_, in_channel_,
                     in_pad_height_, in_pad_width_, kernel_height_, kernel_width_, group_);
  CU_SAFE_CALL(cudaGetLastError());
  CuDevice::Instantiate().AccuProfile(__func__, tim.Elapsed());
}

// The SGD step of reference :767-775:
//   lr = learning_rate_ / num_sample (BaseFloat division, widened to double)
//   prev = momentum*prev ; prev += (-lr*wd) * W ; prev += lr * grad ; W += prev   (one fused pass)
//   bias += lr * db                                                             (no momentum / decay)
void ConvolutionComponent::ApplyGradient(int32 total_num_samples) {
  EnsureGradBuffers();
  KALDI_ASSERT(total_num_samples > 0);
  double learning_rate = learning_rate_ / total_num_samples;
  BaseFloat a_decay = -1 * learning_rate * weight_decay_, a_grad = learning_rate;
  ::MatrixDim gd = {w_grad_.rows, w_grad_.cols, w_grad_.stride};
  cudaF_sgd_momentum_update(Str(), linear_params_.Data(), linear_params_.Dim(), prev_grad_.Data(),
                            prev_grad_.Dim(), w_grad_.data, gd, momentum_, a_decay, a_grad);
  cudaF_vec_axpy(Str(), bias_params_.Data(), b_grad_.data, group_, a_grad);
  CU_SAFE_CALL(cudaGetLastError());
}

bool ConvolutionComponent::GetStepTarget(int32 num_rows, StepTarget *t) {
  if (is_gradient_ || num_rows <= 0) return false;
  EnsureGradBuffers();
  double learning_rate = learning_rate_ / num_rows;                 // reference :767
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

// reference :738-777
void ConvolutionComponent::Update(const CuMatrixBase<BaseFloat> &in_value,
                                  const CuMatrixBase<BaseFloat> &out_deriv) {
  ComputeGradient(in_value, out_deriv);
  if (!deferred_) ApplyGradient(in_value.NumRows());
}

}  // namespace nnet0
}  // namespace cnsl
