// nnet2/nnet-component.h -- shim of the nnet2 Component API the nnet0 add-on plugs into.
//
// The reference ships stock Kaldi r4510's nnet2/nnet-component.{h,cc} (6.2 kLoC, ~38
// component classes) with four added lines: the include of nnet0 and the three factory
// branches (nnet2/nnet-component.cc:32, 112-117).  Everything the CNN hot path touches
// is declared here with the SAME names and signatures:
//   ChunkInfo                     nnet2/nnet-component.h:72-146
//   Component                     :157-269   (Propagate / Backprop virtuals :197-233)
//   UpdatableComponent            :279-348
//   NonlinearComponent            :352-409
//   AffineComponent               :843-941   (base class of FullyConnectedComponent)
//   RectifiedLinear / Softmax / Dropout components: the glue between CNN layers in
//   egs/exp/nnet/nnet.config.
// The other stock components are unrelated to the path and are not re-created
// (out of scope).
#ifndef KALDI_NNET2_NNET_COMPONENT_H_
#define KALDI_NNET2_NNET_COMPONENT_H_

#include <string>
#include <vector>

#include "base/kaldi-common.h"
#include "itf/options-itf.h"
#include "matrix/matrix-lib.h"
#include "cudamatrix/cu-matrix-lib.h"
#include "thread/kaldi-mutex.h"

namespace kaldi {
namespace nnet2 {

/// Shape descriptor of a minibatch (reference nnet2/nnet-component.h:72-146): the
/// matrix has NumChunks() * (frames per chunk) rows of feat_dim columns.  The CNN hot
/// path reads only NumChunks() (nnet0/nnet-component-nnet0.cc:431) and CheckSize().
class ChunkInfo {
 public:
  ChunkInfo() : feat_dim_(0), num_chunks_(0), first_offset_(0), last_offset_(0), offsets_() {}
  ChunkInfo(int32 feat_dim, int32 num_chunks, int32 first_offset, int32 last_offset)
      : feat_dim_(feat_dim), num_chunks_(num_chunks), first_offset_(first_offset),
        last_offset_(last_offset), offsets_() { Check(); }
  ChunkInfo(int32 feat_dim, int32 num_chunks, const std::vector<int32> offsets)
      : feat_dim_(feat_dim), num_chunks_(num_chunks), first_offset_(offsets.front()),
        last_offset_(offsets.back()), offsets_(offsets) {
    if (last_offset_ - first_offset_ + 1 == static_cast<int32>(offsets_.size())) offsets_.clear();
    Check();
  }
  int32 GetIndex(int32 offset) const;
  int32 GetOffset(int32 index) const;
  void MakeOffsetsContiguous() { offsets_.clear(); Check(); }
  inline int32 ChunkSize() const { return NumRows() / num_chunks_; }
  inline int32 NumChunks() const { return num_chunks_; }
  int32 NumRows() const {
    return num_chunks_ * (!offsets_.empty() ? static_cast<int32>(offsets_.size())
                                            : last_offset_ - first_offset_ + 1);
  }
  int32 NumCols() const { return feat_dim_; }
  void CheckSize(const CuMatrixBase<BaseFloat> &mat) const;
  void Check() const;

 private:
  int32 feat_dim_;
  int32 num_chunks_;
  int32 first_offset_;
  int32 last_offset_;
  std::vector<int32> offsets_;
};

class NnetMinibatchUpdater;

class Component {
 public:
  Component() : index_(-1) {}
  virtual std::string Type() const = 0;
  virtual int32 Index() const { return index_; }
  virtual void SetIndex(int32 index) { index_ = index; }
  virtual void InitFromString(std::string args) = 0;
  virtual int32 InputDim() const = 0;
  virtual int32 OutputDim() const = 0;
  virtual std::vector<int32> Context() const { return std::vector<int32>(1, 0); }

  virtual void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                         const CuMatrixBase<BaseFloat> &in,
                         CuMatrixBase<BaseFloat> *out) const = 0;

  /// Non-virtual overload that first resizes the output (reference :203-215).
  void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                 const CuMatrixBase<BaseFloat> &in, CuMatrix<BaseFloat> *out) const {
    if (out->NumRows() != out_info.NumRows() || out->NumCols() != out_info.NumCols())
      out->Resize(out_info.NumRows(), out_info.NumCols());
    Propagate(in_info, out_info, in, static_cast<CuMatrixBase<BaseFloat> *>(out));
  }

  virtual void Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info,
                        const CuMatrixBase<BaseFloat> &in_value,
                        const CuMatrixBase<BaseFloat> &out_value,
                        const CuMatrixBase<BaseFloat> &out_deriv,
                        Component *to_update,  // may be identical to "this".
                        CuMatrix<BaseFloat> *in_deriv) const = 0;

  virtual bool BackpropNeedsInput() const { return true; }
  virtual bool BackpropNeedsOutput() const { return true; }

  /// B200 extension (fusion hook): this component's Propagate followed by
  /// RectifiedLinearComponent::Propagate as ONE launch -- `out` receives max(Propagate(in), 0),
  /// the pre-activation is never stored.  Returns false when the component has no such path;
  /// the caller then runs the two components one after the other.  Only valid for a caller
  /// that will not look at the pre-activation again (this component's BackpropNeedsOutput()
  /// is false and the ReLU's BackpropNeedsInput() is false: NnetMinibatchUpdater checks both).
  virtual bool PropagateRelu(const ChunkInfo & /*in_info*/, const ChunkInfo & /*out_info*/,
                             const CuMatrixBase<BaseFloat> & /*in*/, CuMatrixBase<BaseFloat> * /*out*/) const {
    return false;
  }

  /// B200 extension: a hash of everything a captured CUDA graph of this component's
  /// Propagate / Backprop bakes in besides the activation buffers -- parameter addresses,
  /// learning rate, momentum, ... (NnetMinibatchUpdater::TrainStep re-captures its graph
  /// when the sum over the components changes).
  virtual uint64 StepSignature() const { return 0; }
  static uint64 HashBytes(const void *p, size_t n, uint64 h = 1469598103934665603ull) {
    const unsigned char *b = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
  }
  template <class T> static uint64 HashValue(const T &v, uint64 h) { return HashBytes(&v, sizeof(T), h); }

  static Component *ReadNew(std::istream &is, bool binary);
  virtual Component *Copy() const = 0;
  static Component *NewFromString(const std::string &initializer_line);
  static Component *NewComponentOfType(const std::string &type);
  virtual void Read(std::istream &is, bool binary) = 0;
  virtual void Write(std::ostream &os, bool binary) const = 0;
  virtual std::string Info() const;
  virtual ~Component() {}

 private:
  int32 index_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(Component);
};

class UpdatableComponent : public Component {
 public:
  UpdatableComponent(const UpdatableComponent &other) : Component(), learning_rate_(other.learning_rate_) {}
  void Init(BaseFloat learning_rate) { learning_rate_ = learning_rate; }
  UpdatableComponent(BaseFloat learning_rate) { Init(learning_rate); }
  virtual void SetZero(bool treat_as_gradient) = 0;
  UpdatableComponent() : learning_rate_(0.001) {}
  virtual ~UpdatableComponent() {}
  virtual BaseFloat DotProduct(const UpdatableComponent &other) const = 0;
  virtual void PerturbParams(BaseFloat stddev) = 0;
  virtual void Scale(BaseFloat scale) = 0;
  virtual void Add(BaseFloat alpha, const UpdatableComponent &other) = 0;
  void SetLearningRate(BaseFloat lrate) { learning_rate_ = lrate; }
  BaseFloat LearningRate() const { return learning_rate_; }
  virtual std::string Info() const;
  virtual int32 GetParameterDim() const { KALDI_ASSERT(0); return 0; }
  virtual void Vectorize(VectorBase<BaseFloat> *) const { KALDI_ASSERT(0); }
  virtual void UnVectorize(const VectorBase<BaseFloat> &) { KALDI_ASSERT(0); }

  // ---- B200 data-parallel extension (not in the reference) ----------------------
  // The reference updates parameters INSIDE Backprop (to_update->Update).  To shard a
  // minibatch over GPUs the update is split: with deferred updates on, Update() only
  // leaves the un-normalised gradient in GradientBuffers(); the caller all-reduces
  // those buffers and then calls ApplyGradient(total_rows), which performs exactly the
  // reference's update with lr / total_rows.  Default: off (reference behaviour).
  struct GradBuffer { float *data; int32 rows, cols, stride; };
  virtual void SetDeferredUpdate(bool) {}
  virtual bool DeferredUpdate() const { return false; }
  virtual std::vector<GradBuffer> GradientBuffers() { return std::vector<GradBuffer>(); }
  virtual void ApplyGradient(int32 /*total_num_samples*/) {}
  /// Place the gradient buffers in caller-provided storage (one flat all-reduce bucket).
  virtual size_t GradientFloats() const { return 0; }
  virtual void SetGradientStorage(float * /*base*/) {}
  /// Raw views of the parameters, the momentum state and the gradient buffers, with the SGD
  /// coefficients of one minibatch of num_rows rows (lr = learning_rate_ / num_rows, reference
  /// nnet0/nnet-component-nnet0.cc:767, 1136): what NnetMinibatchUpdater's fused step hands to the
  /// kernels directly.  Returns false for components without a momentum / weight-decay update.
  struct StepTarget {
    float *w; ::MatrixDim wd;          // linear_params_
    float *prev; ::MatrixDim pd;       // prev_grad_
    float *bias; int32 bias_dim;       // bias_params_
    float *w_grad; ::MatrixDim gd;     // gradient buffers (deferred / data-parallel mode)
    float *b_grad;
    bool deferred;
    float momentum, a_decay, a_grad;
  };
  virtual bool GetStepTarget(int32 /*num_rows*/, StepTarget * /*t*/) { return false; }
  /// Move linear_params_ and bias_params_ into caller-provided device memory laid out like the
  /// gradient bucket (GradientFloats(): W rows x pitch, then the bias), keeping their values; NULL moves
  /// them back into memory of their own.  The data-parallel trainer keeps all parameters in one
  /// NVLink-visible arena so that the owner of a slice can write the updated weights to every replica.
  virtual void SetParameterStorage(float * /*base*/) {}

 protected:
  /// "Backprop's in_value is, unmodified, the matrix last given to Propagate" -- lets a component keep
  /// per-input scratch between the two calls (ConvolutionComponent: the channels-last staging copy).
  /// Only NnetMinibatchUpdater, which OWNS the activation buffers, can make that promise; it is not part
  /// of the public interface because nothing checks the contents.
  friend class NnetMinibatchUpdater;
  virtual void SetInputPersists(bool) {}
  BaseFloat learning_rate_;
 private:
  const UpdatableComponent &operator=(const UpdatableComponent &other);  // Disallow.
};

/// Element-wise nonlinearities (reference :352-409).  The value / derivative sums are
/// diagnostics written to the model file (<ValueSum> <DerivSum> <Count>); they live on
/// the device in double, as in Kaldi.
class NonlinearComponent : public Component {
 public:
  void Init(int32 dim) { dim_ = dim; count_ = 0.0; }
  explicit NonlinearComponent(int32 dim) : stats_(NULL), stats_dim_(0) { Init(dim); }
  NonlinearComponent() : dim_(0), count_(0.0), stats_(NULL), stats_dim_(0) {}
  explicit NonlinearComponent(const NonlinearComponent &other);
  virtual ~NonlinearComponent();
  virtual int32 InputDim() const { return dim_; }
  virtual int32 OutputDim() const { return dim_; }
  virtual void InitFromString(std::string args);
  virtual void Read(std::istream &is, bool binary);
  virtual void Write(std::ostream &os, bool binary) const;
  double Count() const { return count_; }
  /// Frames a replayed CUDA graph of Backprop has accumulated into the statistics.
  void AddToCount(double frames) { count_ += frames; }
  /// Host copies of the diagnostics.
  void GetStats(Vector<double> *value_sum, Vector<double> *deriv_sum) const;
  /// Device accumulators [2 x dim]: value sums, then derivative sums (allocated on first use).
  double *StatsDevice() { EnsureStats(); return stats_; }

 protected:
  friend class RectifiedLinearComponent;
  friend class SoftmaxComponent;
  /// Adds column sums of out_value (and, for ReLU, of the 0/1 derivative) to the stats.
  void UpdateStats(const CuMatrixBase<BaseFloat> &out_value, bool relu_deriv);
  /// in_deriv = out_deriv * [out_value > 0] and the UpdateStats sums, one kernel.
  void BackpropReluWithStats(const CuMatrixBase<BaseFloat> &out_value, const CuMatrixBase<BaseFloat> &out_deriv,
                             CuMatrix<BaseFloat> *in_deriv);
  void EnsureStats();
  const NonlinearComponent &operator=(const NonlinearComponent &other);  // Disallow.
  int32 dim_;
  double count_;
  double *stats_;        // device: [2 x stats_dim_] value sums, derivative sums
  int32 stats_dim_;
  Vector<double> value_sum_host_, deriv_sum_host_;   // as read from a model file
};

class RectifiedLinearComponent : public NonlinearComponent {
 public:
  explicit RectifiedLinearComponent(int32 dim) : NonlinearComponent(dim) {}
  explicit RectifiedLinearComponent(const RectifiedLinearComponent &other) : NonlinearComponent(other) {}
  RectifiedLinearComponent() {}
  virtual std::string Type() const { return "RectifiedLinearComponent"; }
  virtual Component *Copy() const { return new RectifiedLinearComponent(*this); }
  virtual bool BackpropNeedsInput() const { return false; }
  virtual bool BackpropNeedsOutput() const { return true; }
  using Component::Propagate;
  virtual void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                         const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual void Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info,
                        const CuMatrixBase<BaseFloat> &in_value,
                        const CuMatrixBase<BaseFloat> &out_value,
                        const CuMatrixBase<BaseFloat> &out_deriv, Component *to_update,
                        CuMatrix<BaseFloat> *in_deriv) const;
 private:
  RectifiedLinearComponent &operator=(const RectifiedLinearComponent &other);  // Disallow.
};

class SoftmaxComponent : public NonlinearComponent {
 public:
  explicit SoftmaxComponent(int32 dim) : NonlinearComponent(dim) {}
  explicit SoftmaxComponent(const SoftmaxComponent &other) : NonlinearComponent(other) {}
  SoftmaxComponent() {}
  virtual std::string Type() const { return "SoftmaxComponent"; }
  virtual bool BackpropNeedsInput() const { return false; }
  virtual bool BackpropNeedsOutput() const { return true; }
  using Component::Propagate;
  virtual void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                         const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual void Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info,
                        const CuMatrixBase<BaseFloat> &in_value,
                        const CuMatrixBase<BaseFloat> &out_value,
                        const CuMatrixBase<BaseFloat> &out_deriv, Component *to_update,
                        CuMatrix<BaseFloat> *in_deriv) const;
  virtual Component *Copy() const { return new SoftmaxComponent(*this); }
 private:
  SoftmaxComponent &operator=(const SoftmaxComponent &other);  // Disallow.
};

/// reference nnet2/nnet-component.cc:576-639.  Scales every row to unit root-mean-square (floor 2^-66
/// on the mean square); the statistics of NonlinearComponent are carried but not updated, as upstream.
class NormalizeComponent : public NonlinearComponent {
 public:
  explicit NormalizeComponent(int32 dim) : NonlinearComponent(dim) {}
  explicit NormalizeComponent(const NormalizeComponent &other) : NonlinearComponent(other) {}
  NormalizeComponent() {}
  virtual std::string Type() const { return "NormalizeComponent"; }
  virtual Component *Copy() const { return new NormalizeComponent(*this); }
  virtual bool BackpropNeedsInput() const { return true; }
  virtual bool BackpropNeedsOutput() const { return true; }
  using Component::Propagate;
  virtual void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                         const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual void Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info,
                        const CuMatrixBase<BaseFloat> &in_value,
                        const CuMatrixBase<BaseFloat> &out_value,
                        const CuMatrixBase<BaseFloat> &out_deriv, Component *to_update,
                        CuMatrix<BaseFloat> *in_deriv) const;
 private:
  NormalizeComponent &operator=(const NormalizeComponent &other);  // Disallow.
};

/// reference nnet2/nnet-component.h:1092-1129, .cc:2524-2866.  Splices the frames at the offsets in
/// `context` around each output frame: the front end that turns [chunks x frames, input-dim] filterbank
/// rows into the [C][W][H] window the first convolution reads (nnet.config line 1).  The last
/// const-component-dim columns are copied once instead of being spliced.
class SpliceComponent : public Component {
 public:
  SpliceComponent() : input_dim_(0), const_component_dim_(0), map_dev_(NULL), map_len_(0), map_key_(0) {}
  virtual ~SpliceComponent();
  void Init(int32 input_dim, std::vector<int32> context, int32 const_component_dim = 0);
  virtual std::string Type() const { return "SpliceComponent"; }
  virtual std::string Info() const;
  virtual void InitFromString(std::string args);
  virtual int32 InputDim() const { return input_dim_; }
  virtual int32 OutputDim() const;
  virtual std::vector<int32> Context() const { return context_; }
  int32 ConstComponentDim() const { return const_component_dim_; }
  /// True when the context is a run of consecutive offsets and nothing is held constant: a chunk of
  /// context.size() input frames with ONE output frame is then a plain reshape of the input rows.
  bool IsContiguousWindow() const;
  using Component::Propagate;
  virtual void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                         const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual void Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info,
                        const CuMatrixBase<BaseFloat> &in_value,
                        const CuMatrixBase<BaseFloat> &out_value,
                        const CuMatrixBase<BaseFloat> &out_deriv, Component *to_update,
                        CuMatrix<BaseFloat> *in_deriv) const;
  virtual bool BackpropNeedsInput() const { return false; }
  virtual bool BackpropNeedsOutput() const { return false; }
  virtual Component *Copy() const;
  virtual void Read(std::istream &is, bool binary);
  virtual void Write(std::ostream &os, bool binary) const;
 private:
  KALDI_DISALLOW_COPY_AND_ASSIGN(SpliceComponent);
  /// Device table [context.size() x out_chunk_size] of the input frame (within a chunk) each spliced
  /// block of each output frame of a chunk comes from; cached per (in_info, out_info) shape.
  const int32 *FrameMap(const ChunkInfo &in_info, const ChunkInfo &out_info) const;
  int32 input_dim_;
  std::vector<int32> context_;
  int32 const_component_dim_;
  mutable int32 *map_dev_;
  mutable size_t map_len_;
  mutable uint64 map_key_;
};

/// reference nnet2/nnet-component.cc:3560-3640.  out = in .* mask, where a proportion
/// dp of the mask is dropout_scale and the rest (1 - dp*scale)/(1 - dp).
class DropoutComponent : public Component {
 public:
  void Init(int32 dim, BaseFloat dropout_proportion = 0.5, BaseFloat dropout_scale = 0.0);
  DropoutComponent(int32 dim, BaseFloat dp = 0.5, BaseFloat sc = 0.0) : seed_dev_(NULL) { Init(dim, dp, sc); }
  DropoutComponent() : dim_(0), dropout_proportion_(0.5), dropout_scale_(0.0), seed_dev_(NULL) {}
  virtual ~DropoutComponent();
  virtual int32 InputDim() const { return dim_; }
  virtual int32 OutputDim() const { return dim_; }
  virtual void InitFromString(std::string args);
  virtual void Read(std::istream &is, bool binary);
  virtual void Write(std::ostream &os, bool binary) const;
  virtual std::string Type() const { return "DropoutComponent"; }
  void SetDropoutScale(BaseFloat scale) { dropout_scale_ = scale; }
  BaseFloat DropoutProportion() const { return dropout_proportion_; }
  BaseFloat DropoutScale() const { return dropout_scale_; }
  /// Device-resident seed counter of the mask (created, from CuDevice's seed, on first use).
  unsigned long long *SeedDevice() const;
  virtual uint64 StepSignature() const {
    return HashValue(seed_dev_, HashValue(dropout_scale_, HashValue(dropout_proportion_, 7)));
  }
  virtual bool BackpropNeedsInput() const { return true; }
  virtual bool BackpropNeedsOutput() const { return true; }
  virtual Component *Copy() const;
  using Component::Propagate;
  virtual void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                         const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual void Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info,
                        const CuMatrixBase<BaseFloat> &in_value,
                        const CuMatrixBase<BaseFloat> &out_value,
                        const CuMatrixBase<BaseFloat> &out_deriv, Component *to_update,
                        CuMatrix<BaseFloat> *in_deriv) const;
  virtual std::string Info() const;
 private:
  int32 dim_;
  BaseFloat dropout_proportion_;
  BaseFloat dropout_scale_;
  mutable unsigned long long *seed_dev_;   // device counter: (seed, element) -> uniform
};

/// reference :843-941.  linear_params_ is [output_dim x input_dim].
class AffineComponent : public UpdatableComponent {
 public:
  explicit AffineComponent(const AffineComponent &other);
  AffineComponent(const CuMatrixBase<BaseFloat> &linear_params,
                  const CuVectorBase<BaseFloat> &bias_params, BaseFloat learning_rate);
  virtual int32 InputDim() const { return linear_params_.NumCols(); }
  virtual int32 OutputDim() const { return linear_params_.NumRows(); }
  void Init(BaseFloat learning_rate, int32 input_dim, int32 output_dim, BaseFloat param_stddev,
            BaseFloat bias_stddev);
  void Init(BaseFloat learning_rate, std::string matrix_filename);
  virtual std::string Info() const;
  virtual void InitFromString(std::string args);
  AffineComponent() : is_gradient_(false), deferred_(false), grad_external_(false) {}
  virtual std::string Type() const { return "AffineComponent"; }
  virtual bool BackpropNeedsInput() const { return true; }
  virtual bool BackpropNeedsOutput() const { return false; }
  using Component::Propagate;
  virtual void Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                         const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual bool PropagateRelu(const ChunkInfo &in_info, const ChunkInfo &out_info,
                             const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const;
  virtual void Scale(BaseFloat scale);
  virtual void Add(BaseFloat alpha, const UpdatableComponent &other);
  virtual void Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info,
                        const CuMatrixBase<BaseFloat> &in_value,
                        const CuMatrixBase<BaseFloat> &out_value,
                        const CuMatrixBase<BaseFloat> &out_deriv, Component *to_update,
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
  virtual int32 GetParameterDim() const;
  virtual void Vectorize(VectorBase<BaseFloat> *params) const;
  virtual void UnVectorize(const VectorBase<BaseFloat> &params);

  // data-parallel extension (see UpdatableComponent)
  virtual void SetDeferredUpdate(bool on) { deferred_ = on; }
  virtual bool DeferredUpdate() const { return deferred_; }
  virtual std::vector<GradBuffer> GradientBuffers();
  virtual size_t GradientFloats() const;
  virtual void SetGradientStorage(float *base);
  virtual void SetParameterStorage(float *base);
  virtual uint64 StepSignature() const {
    uint64 h = HashValue(learning_rate_, 11);
    h = HashValue(linear_params_.Data(), h); h = HashValue(bias_params_.Data(), h);
    h = HashValue(linear_params_.Stride(), h); h = HashValue(deferred_, h);
    h = HashValue(w_grad_.data, h); h = HashValue(b_grad_.data, h);
    return HashValue(is_gradient_, h);
  }

 protected:
  virtual void Update(const CuMatrixBase<BaseFloat> &in_value,
                      const CuMatrixBase<BaseFloat> &out_deriv) {
    UpdateSimple(in_value, out_deriv);
  }
  virtual void UpdateSimple(const CuMatrixBase<BaseFloat> &in_value,
                            const CuMatrixBase<BaseFloat> &out_deriv);
  /// w_grad_ = out_deriv^T in_value, b_grad_ = column sums of out_deriv (un-normalised).
  void ComputeGradient(const CuMatrixBase<BaseFloat> &in_value,
                       const CuMatrixBase<BaseFloat> &out_deriv);
  void EnsureGradBuffers();
  void PropagateAct(const ChunkInfo &in_info, const ChunkInfo &out_info, const CuMatrixBase<BaseFloat> &in,
                    CuMatrixBase<BaseFloat> *out, int act) const;

  const AffineComponent &operator=(const AffineComponent &other);  // Disallow.
  CuMatrix<BaseFloat> linear_params_;
  CuVector<BaseFloat> bias_params_;
  bool is_gradient_;

  bool deferred_, grad_external_;
  CuMatrix<BaseFloat> w_grad_store_;
  CuVector<BaseFloat> b_grad_store_;
  GradBuffer w_grad_, b_grad_;
};

// Config-line parsing helpers ("key=value" tokens; reference nnet2/nnet-component.cc:161-300
// and the duplicate at nnet0/nnet-component-nnet0.cc:42-176).
bool ParseFromString(const std::string &name, std::string *string, int32 *param);
bool ParseFromString(const std::string &name, std::string *string, bool *param);
bool ParseFromString(const std::string &name, std::string *string, BaseFloat *param);
bool ParseFromString(const std::string &name, std::string *string, std::string *param);
bool ParseFromString(const std::string &name, std::string *string, std::vector<int32> *param);
void ExpectOneOrTwoTokens(std::istream &is, bool binary, const std::string &token1,
                          const std::string &token2);

}  // namespace nnet2
}  // namespace kaldi

#endif
