// nnet0/nnet-component-nnet0.h -- L2: the three CNN components of the nnet0 add-on,
// B200 build.  Class names, namespaces, virtual signatures, config keys, Type()
// strings and the serialised token streams are those of the reference
// (src/nnet0/nnet-component-nnet0.h:23-232), so nnet2's NewComponentOfType
// (nnet2/nnet-component.cc:112-117), existing nnet.config lines and existing model
// files work unchanged.  What differs is what Propagate / Backprop / Update launch:
//
//   ConvolutionComponent::Propagate  1 implicit-GEMM (pad + conv + bias)        vs 5 kernels + SGEMM + 3 allocs
//   ConvolutionComponent::Backprop   1 implicit-GEMM dgrad (either branch)      vs 5-7 copies + SGEMM
//   ConvolutionComponent::Update     bias-grad + implicit-GEMM wgrad + 1 SGD    vs 4 copies + SGEMM + 5 passes
//   MaxpoolComponent                 1 kernel each way (zero fill fused)        vs memset + kernel
//   FullyConnectedComponent          GEMM with bias epilogue; GEMM; GEMM + SGD  vs 2 + 1 + 5 passes
//
// ConvolutionComponentContainer (reference :234-328) is not built: it is not
// registered in the factory and cannot compile against the shipped ChunkInfo
// (SURVEY section 2, #8).
#ifndef CNSL_NNET0_NNET_CONV_H_
#define CNSL_NNET0_NNET_CONV_H_

#include <iostream>

#include "base/kaldi-common.h"
#include "itf/options-itf.h"
#include "matrix/matrix-lib.h"
#include "cudamatrix/cu-matrix-lib.h"
#include "thread/kaldi-mutex.h"
#include "nnet2/nnet-component.h"

using namespace kaldi;          // as the reference header does (:17-18)
using namespace kaldi::nnet2;

namespace cnsl {
namespace nnet0 {

class FieldList;   // nnet0/component-fields.h: config keys + model-file tokens of a component, declared once

/// 2-D convolution over [C][W][H] activations (H fastest), stride 1, optional zero
/// padding.  "group" is the number of OUTPUT maps (filters), not channel grouping.
///   in  [num_chunks x in_height*in_width*in_channel]
///   out [num_chunks x out_height*out_width*group]
///   linear_params_ [(kernel_height*kernel_width*in_channel) x group], bias_params_ [group]
class ConvolutionComponent : public nnet2::UpdatableComponent {
 public:
  explicit ConvolutionComponent(const ConvolutionComponent &other);
  ConvolutionComponent();   // use Init to really initialize.
  ConvolutionComponent(const CuMatrix<BaseFloat> &linear_params,
                       const CuVector<BaseFloat> &bias_params, BaseFloat learning_rate,
                       int32 in_height, int32 in_width, int32 in_channels, int32 in_pad_height,
                       int32 in_pad_width, int32 kernel_height, int32 kernel_width, int32 stride,
                       int32 group, int32 out_height, int32 out_width, BaseFloat weight_decay,
                       BaseFloat momentum);
  virtual ~ConvolutionComponent() {}

  virtual int32 InputDim() const { return in_height_ * in_width_ * in_channel_; }
  virtual int32 OutputDim() const { return out_height_ * out_width_ * group_; }
  inline int32 In_height() const { return in_height_; }
  inline int32 In_width() const { return in_width_; }
  inline int32 In_channels() const { return in_channel_; }
  inline int32 Out_height() const { return out_height_; }
  inline int32 Out_width() const { return out_width_; }
  inline int32 Group() const { return group_; }
  inline int32 KernelDim() const { return kernel_height_ * kernel_width_ * in_channel_; }
  inline int32 Kernel_height() const { return kernel_height_; }
  inline int32 Kernel_width() const { return kernel_width_; }
  inline int32 In_pad_height() const { return in_pad_height_; }
  inline int32 In_pad_width() const { return in_pad_width_; }

  void Init(BaseFloat learning_rate, int32 in_height, int32 in_width, int32 in_channels,
            int32 in_pad_height, int32 in_pad_width, int32 kernel_height, int32 kernel_width,
            int32 stride, int32 group, int32 out_height, int32 out_width, BaseFloat param_stddev,
            BaseFloat bias_stddev, BaseFloat weight_decay, BaseFloat momentum);
  void Init(BaseFloat learning_rate, int32 in_height, int32 in_width, int32 in_channels,
            int32 in_pad_height, int32 in_pad_width, int32 kernel_height, int32 kernel_width,
            int32 stride, int32 group, int32 out_height, int32 out_width, BaseFloat weight_decay,
            BaseFloat momentum, std::string matrix_filename);

  virtual void InitFromString(std::string args);
  virtual std::string Info() const;
  virtual std::string Type() const { return "ConvolutionComponent"; }
  virtual bool BackpropNeedsInput() const { return true; }
  virtual bool BackpropNeedsOutput() const { return false; }
  using Component::Propagate;   // to avoid name hiding
  virtual void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                         const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual bool PropagateRelu(const ChunkInfo &in_info, const ChunkInfo &out_info,
                             const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual void Scale(BaseFloat scale);
  virtual void Add(BaseFloat alpha, const UpdatableComponent &other);
  virtual void Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info,
                        const CuMatrixBase<BaseFloat> &in_value,
                        const CuMatrixBase<BaseFloat> &out_value,   // dummy
                        const CuMatrixBase<BaseFloat> &out_deriv,
                        Component *to_update,   // may be identical to "this".
                        CuMatrix<BaseFloat> *in_deriv) const;
  virtual void SetZero(bool treat_as_gradient);
  virtual void Read(std::istream &is, bool binary);
  virtual void Write(std::ostream &os, bool binary) const;
  virtual BaseFloat DotProduct(const UpdatableComponent &other) const;
  virtual Component *Copy() const;
  virtual void PerturbParams(BaseFloat stddev);
  virtual void SetParams(const VectorBase<BaseFloat> &bias, const MatrixBase<BaseFloat> &linear);
  const CuVector<BaseFloat> &BiasParams() { return bias_params_; }
  const CuMatrix<BaseFloat> &LinearParams() { return linear_params_; }
  const CuMatrix<BaseFloat> &PrevGrad() { return prev_grad_; }
  virtual int32 GetParameterDim() const;
  virtual void Vectorize(VectorBase<BaseFloat> *params) const;
  virtual void UnVectorize(const VectorBase<BaseFloat> &params);
  void SetWeightDecay(BaseFloat weight_decay) { weight_decay_ = weight_decay; }
  void SetMomentum(BaseFloat momentum) { momentum_ = momentum; }
  BaseFloat WeightDecay() const { return weight_decay_; }
  BaseFloat Momentum() const { return momentum_; }

  // data-parallel extension (nnet2::UpdatableComponent)
  virtual void SetDeferredUpdate(bool on) { deferred_ = on; }
  virtual bool DeferredUpdate() const { return deferred_; }
  virtual std::vector<GradBuffer> GradientBuffers();
  virtual void ApplyGradient(int32 total_num_samples);
  virtual size_t GradientFloats() const;
  virtual void SetGradientStorage(float *base);
  virtual void SetParameterStorage(float *base);
  virtual bool GetStepTarget(int32 num_rows, StepTarget *t);
  virtual uint64 StepSignature() const {
    uint64 h = HashValue(learning_rate_, 13);
    h = HashValue(weight_decay_, h); h = HashValue(momentum_, h);
    h = HashValue(linear_params_.Data(), h); h = HashValue(bias_params_.Data(), h);
    h = HashValue(prev_grad_.Data(), h); h = HashValue(deferred_, h);
    h = HashValue(w_grad_.data, h); h = HashValue(b_grad_.data, h);
    h = HashValue(input_persists_, h);
    return HashValue(is_gradient_, h);
  }

 protected:
  /// (nnet2::UpdatableComponent, reachable by NnetMinibatchUpdater only.)  With the promise that the
  /// matrix given to Backprop as in_value is, unmodified, the one last propagated, Backprop reuses
  /// Propagate's channels-last staging copy instead of packing in_value again.  Off by default: a bare
  /// Component makes no such assumption.
  virtual void SetInputPersists(bool on) { input_persists_ = on; staged_src_ = NULL; }
  virtual void Update(const CuMatrixBase<BaseFloat> &in_value,
                      const CuMatrixBase<BaseFloat> &out_deriv);
  void ComputeGradient(const CuMatrixBase<BaseFloat> &in_value,
                       const CuMatrixBase<BaseFloat> &out_deriv);
  void EnsureGradBuffers();
  FieldList StreamFields();
  void CheckGeometry() const;
  void SetShape(BaseFloat learning_rate, int32 in_height, int32 in_width, int32 in_channels,
                int32 in_pad_height, int32 in_pad_width, int32 kernel_height, int32 kernel_width,
                int32 stride, int32 group, int32 out_height, int32 out_width, BaseFloat weight_decay,
                BaseFloat momentum);
  void PropagateAct(const ChunkInfo &in_info, const CuMatrixBase<BaseFloat> &in,
                    CuMatrixBase<BaseFloat> *out, int act) const;

  const ConvolutionComponent &operator=(const ConvolutionComponent &other);   // Disallow.

  CuMatrix<BaseFloat> linear_params_;
  CuVector<BaseFloat> bias_params_;   // each output map shares one bias value
  bool is_gradient_;                  // if true, treat this as just a gradient.

  int32 in_height_, in_width_, in_channel_;
  int32 in_pad_height_, in_pad_width_;
  int32 kernel_height_, kernel_width_;
  int32 stride_;   // parsed, stored, serialised -- and, as in the reference, never applied
  int32 group_;
  int32 out_height_, out_width_;

  BaseFloat weight_decay_;
  BaseFloat momentum_;
  CuMatrix<BaseFloat> prev_grad_;   // momentum state, checkpointed as <PrevGrad>

  // scratch owned by the component: allocated once, reused every minibatch
  bool deferred_, grad_external_;
  CuMatrix<BaseFloat> w_grad_store_;
  CuVector<BaseFloat> b_grad_store_;
  GradBuffer w_grad_, b_grad_;
  // channels-last staging copy of the last propagated input (see Propagate)
  mutable CuVector<BaseFloat> staged_in_;
  mutable const BaseFloat *staged_src_;
  mutable int32 staged_rows_, staged_stride_;
  bool input_persists_;
  CuVector<BaseFloat> workspace_;     // split-K partials of the weight-gradient GEMM
  int32 workspace_rows_;
};

/// 3-D (height x width x intermap-channel) max pooling, non-overlapping windows; optional
/// "overlap" (1-D sliding window over channels) and "overlap2D" modes.
class MaxpoolComponent : public nnet2::Component {
 public:
  void Init(int32 input_dim, int32 output_dim, int32 in_height, int32 in_width, int32 in_channel,
            int32 pool_height_dim, int32 pool_width_dim, int32 pool_channel_dim, bool overlap,
            bool overlap2D);
  explicit MaxpoolComponent(int32 input_dim, int32 output_dim, int32 in_height, int32 in_width,
                            int32 in_channel, int32 pool_height_dim, int32 pool_width_dim,
                            int32 pool_channel_dim, bool overlap, bool overlap2D)
      : index_routing_(false) {
    Init(input_dim, output_dim, in_height, in_width, in_channel, pool_height_dim, pool_width_dim,
         pool_channel_dim, overlap, overlap2D);
  }
  MaxpoolComponent()
      : input_dim_(0), output_dim_(0), in_height_(0), in_width_(0), in_channel_(0),
        pool_height_dim_(0), pool_width_dim_(0), pool_channel_dim_(0), overlap_(false),
        overlap2D_(false), index_routing_(false) {}
  virtual std::string Type() const { return "MaxpoolComponent"; }
  virtual void InitFromString(std::string args);
  virtual int32 InputDim() const { return input_dim_; }
  virtual int32 OutputDim() const { return output_dim_; }
  using Component::Propagate;   // to avoid name hiding
  virtual void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                         const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual void Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info,
                        const CuMatrixBase<BaseFloat> &in_value,
                        const CuMatrixBase<BaseFloat> &out_value,
                        const CuMatrixBase<BaseFloat> &out_deriv,
                        Component *to_update,   // may be identical to "this".
                        CuMatrix<BaseFloat> *in_deriv) const;
  virtual bool BackpropNeedsInput() const { return true; }
  virtual bool BackpropNeedsOutput() const { return true; }
  virtual Component *Copy() const {
    MaxpoolComponent *c = new MaxpoolComponent(input_dim_, output_dim_, in_height_, in_width_,
                                               in_channel_, pool_height_dim_, pool_width_dim_,
                                               pool_channel_dim_, overlap_, overlap2D_);
    c->index_routing_ = index_routing_;
    return c;
  }
  virtual void Read(std::istream &is, bool binary);
  virtual void Write(std::ostream &os, bool binary) const;
  virtual std::string Info() const;

  /// B200 extension: route the backward pass from a one-byte arg-max index recorded in
  /// Propagate (first maximum in c -> w -> h order) instead of comparing values.
  /// Identical to the reference whenever each window has a unique maximum; under ties
  /// the reference routes to EVERY maximal element, this mode to the first.  Default
  /// off (reference-exact routing); plain mode only.
  void SetIndexRouting(bool on) { index_routing_ = on; }
  bool IndexRouting() const { return index_routing_; }
  int32 In_height() const { return in_height_; }
  int32 In_width() const { return in_width_; }
  int32 In_channel() const { return in_channel_; }
  int32 Pool_height_dim() const { return pool_height_dim_; }
  int32 Pool_width_dim() const { return pool_width_dim_; }
  int32 Pool_channel_dim() const { return pool_channel_dim_; }
  bool Overlap() const { return overlap_; }
  bool Overlap2D() const { return overlap2D_; }

 protected:
  FieldList StreamFields();
  void Check() const;
  int32 input_dim_;
  int32 output_dim_;
  int32 in_height_;
  int32 in_width_;
  int32 in_channel_;
  int32 pool_height_dim_;
  int32 pool_width_dim_;
  int32 pool_channel_dim_;
  bool overlap_;
  bool overlap2D_;
  bool index_routing_;
  mutable unsigned char *index_ = NULL;     // device, [rows x index_stride_]
  mutable int32 index_rows_ = 0, index_stride_ = 0;
 public:
  virtual ~MaxpoolComponent();
};

/// AffineComponent with momentum + weight-decay SGD (reference :193-232).
class FullyConnectedComponent : public nnet2::AffineComponent {
 public:
  virtual std::string Type() const { return "FullyConnectedComponent"; }
  virtual void Read(std::istream &is, bool binary);
  virtual void Write(std::ostream &os, bool binary) const;
  void Init(BaseFloat learning_rate, int32 input_dim, int32 output_dim, BaseFloat param_stddev,
            BaseFloat bias_stddev, BaseFloat weight_decay, BaseFloat momentum);
  void Init(BaseFloat learning_rate, BaseFloat weight_decay, BaseFloat momentum,
            std::string matrix_filename);
  virtual void InitFromString(std::string args);
  virtual std::string Info() const;
  virtual Component *Copy() const;
  FullyConnectedComponent() : weight_decay_(0.0002), momentum_(0.9) {}
  void SetWeightDecay(BaseFloat weight_decay) { weight_decay_ = weight_decay; }
  void SetMomentum(BaseFloat momentum) { momentum_ = momentum; }
  BaseFloat WeightDecay() const { return weight_decay_; }
  BaseFloat Momentum() const { return momentum_; }
  const CuMatrix<BaseFloat> &PrevGrad() { return prev_grad_; }

  virtual void Update(const CuMatrixBase<BaseFloat> &in_value,
                      const CuMatrixBase<BaseFloat> &out_deriv) {
    UpdateSimple(in_value, out_deriv);
  }
  virtual void UpdateSimple(const CuMatrixBase<BaseFloat> &in_value,
                            const CuMatrixBase<BaseFloat> &out_deriv);
  virtual void ApplyGradient(int32 total_num_samples);
  virtual bool GetStepTarget(int32 num_rows, StepTarget *t);
  virtual uint64 StepSignature() const {
    uint64 h = nnet2::AffineComponent::StepSignature();
    h = HashValue(weight_decay_, h); h = HashValue(momentum_, h);
    return HashValue(prev_grad_.Data(), h);
  }

 protected:
  KALDI_DISALLOW_COPY_AND_ASSIGN(FullyConnectedComponent);
  FieldList StreamFields();
  void SetHyper(BaseFloat learning_rate, BaseFloat weight_decay, BaseFloat momentum);
  BaseFloat weight_decay_;
  BaseFloat momentum_;
  CuMatrix<BaseFloat> prev_grad_;   // for momentum
};

}  // namespace nnet0
}  // namespace cnsl

#endif
