// nnet2/nnet-nnet.h -- shim: a minimal nnet2::Nnet (sequence of Components) and the
// minibatch step that nnet2's NnetUpdater performs (upstream nnet-update.cc, absent from
// the patch set): Propagate through every component, cross-entropy
// objective + derivative at the output, Backprop in reverse with to_update == the
// component itself, so the parameter update happens inside Backprop.
// These are the CALLERS of the hot path, kept just big enough to drive it.
#ifndef KALDI_NNET2_NNET_NNET_H_
#define KALDI_NNET2_NNET_NNET_H_

#include <string>
#include <vector>

#include "nnet2/nnet-component.h"

namespace kaldi {
namespace nnet2 {

class Nnet {
 public:
  Nnet() {}
  ~Nnet() { Destroy(); }
  /// One component per config line ("Type key=value ..."), as nnet-init does.
  /// skip_splice drops SpliceComponent lines (the caller then feeds already-spliced windows,
  /// [N x 40*21] for nnet.config); by default the Splice front end is part of the network.
  void Init(std::istream &is, bool skip_splice = false);
  void Read(std::istream &is, bool binary);
  void Write(std::ostream &os, bool binary) const;
  int32 NumComponents() const { return static_cast<int32>(components_.size()); }
  const Component &GetComponent(int32 c) const { return *components_[c]; }
  Component &GetComponent(int32 c) { return *components_[c]; }
  int32 InputDim() const { return components_.empty() ? 0 : components_.front()->InputDim(); }
  int32 OutputDim() const { return components_.empty() ? 0 : components_.back()->OutputDim(); }
  void Append(Component *c) { components_.push_back(c); }
  int32 NumUpdatableComponents() const;
  std::string Info() const;
  void Check() const;
 private:
  void Destroy();
  std::vector<Component *> components_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(Nnet);
};

/// Forward / backward over one minibatch with persistent activation buffers (sized on
/// the first call for a given number of rows; no allocation afterwards).
class NnetMinibatchUpdater {
 public:
  explicit NnetMinibatchUpdater(Nnet *nnet);
  ~NnetMinibatchUpdater();
  /// feats: device [num_rows x InputDim()].  Runs Propagate through all components.
  /// With a SpliceComponent in front (nnet.config line 1) feats holds the FRAMES of the training
  /// examples, FramesPerExample() consecutive rows per example (nnet2's NnetExample: left-context + 1 +
  /// right-context input frames for one output frame), and the network has
  /// feats.NumRows() / FramesPerExample() rows from the Splice output on.  When the context is a run of
  /// consecutive offsets and the rows of feats are dense, that output is the SAME memory viewed as
  /// [examples x frames * dim] -- the [C][W][H] window of the first convolution -- so the front end
  /// costs no copy and no launch; otherwise SpliceComponent::Propagate gathers it.
  void Forward(const CuMatrixBase<BaseFloat> &feats);
  /// Input rows per training example: the span of the SpliceComponent's context, 1 without one.
  int32 FramesPerExample() const;
  /// Propagate through components [first, last] only.  first == 0 binds `feats` as the input;
  /// otherwise the activations below `first` must come from an earlier call on the same batch.
  /// (Lets the data-parallel step run the layers whose weights are up to date while the
  /// gradients of the others are still being all-reduced.)
  /// labels_dev (optional, device int32 [num_rows]): when the range ends with the softmax, the fused plan
  /// then computes objective and derivative in the same kernel (ComputeObjfAndDeriv becomes a no-op).
  void ForwardRange(const CuMatrixBase<BaseFloat> &feats, int32 first, int32 last, const int32 *labels_dev = NULL);
  /// labels: device int32 [num_rows].  Writes the cross-entropy derivative at the output
  /// and accumulates sum_i log p[i, label_i] into the device objective.
  void ComputeObjfAndDeriv(const int32 *labels_dev);
  /// Backprop through components [first, last] in reverse order (last defaults to the
  /// top).  Splitting the range lets a data-parallel caller start all-reducing the
  /// gradients of the upper layers while the lower layers are still running.
  void Backward(int32 last = -1, int32 first = 0);
  /// One whole training step on device buffers: Forward(feats), ComputeObjfAndDeriv(labels),
  /// Backward() -- what NnetUpdater does per minibatch.  The first call for a given (feats,
  /// labels, configuration) runs the launches eagerly (buffers and scratch get sized); the
  /// second captures them into a CUDA graph; from then on a step is ONE cudaGraphLaunch.  The
  /// graph is dropped and re-captured when the buffers or any component's StepSignature()
  /// (parameter addresses, learning rate, momentum, ...) change.  Needs a non-default compute
  /// stream (the legacy default stream cannot be captured): otherwise, or with
  /// KCNN_NNET_GRAPH=0, every step runs eagerly.
  void TrainStep(const CuMatrixBase<BaseFloat> &feats, const int32 *labels_dev);
  /// Fusion (default on; KCNN_NNET_FUSE=0 turns it off).  With the tensor-core math mode and a model
  /// made of the hot path's layers (Convolution / Maxpool / FullyConnected with ReLU, dropout and a
  /// final softmax) the whole step runs as the FUSED PLAN of nnet-fused.cc: channels-last
  /// activations between the time-axis layers, element-wise components inside the GEMM epilogues,
  /// weight gradients beside the next input gradient, batched column sums (KCNN_NNET_PLAN=0 keeps
  /// the component-by-component path).  Otherwise only [Convolution | FullyConnected | Affine] +
  /// RectifiedLinear are merged, through Component::PropagateRelu.  Either way the pre-activation
  /// Activation(c + 1) of a fused pair is NOT filled.
  void SetFusion(bool on) { fuse_ = on; }
  bool Fusion() const { return fuse_; }
  /// True when the last TrainStep was a graph replay.
  bool LastStepReplayed() const { return last_replayed_; }
  /// TrainStep records / replays CUDA graphs (default, KCNN_NNET_GRAPH=0 turns it off for the process);
  /// off: every TrainStep runs its launches eagerly (profiling, debugging).
  void SetGraphs(bool on) { if (!on) DropGraph(); graphs_on_ = on; }
  /// ApplyGradient(total_rows) on every updatable component (deferred-update mode).
  void ApplyGradients(int32 total_rows);
  /// Total floats of gradient storage; SetGradientArena places every component's
  /// gradient buffers in one contiguous arena (one bucket per component, top first).
  size_t GradientFloats() const;
  void SetGradientArena(float *base);
  /// Offset (floats) of component c's bucket inside the arena and its length.
  void GradientBucket(int32 c, size_t *offset, size_t *length) const;
  void SetDeferredUpdate(bool on);
  /// Data-parallel trainer: Backward(last, first) is called once per updatable layer.  With the deferred
  /// join the weight-gradient branch of a range (side stream) is NOT joined into the compute stream when
  /// the call returns -- the next range's input-gradient GEMMs run beside it, as they do inside one
  /// whole-network call; GradientStream() is the stream on which the range's gradients complete (record
  /// the "gradients ready" event there) and JoinSide() joins the branch once the pass is over.
  void SetDeferredJoin(bool on) { deferred_join_ = on; }
  cudaStream_t GradientStream() const { return grad_stream_; }
  void JoinSide();
  const CuMatrix<BaseFloat> &Output() const { return forward_.back(); }
  /// Output of component i - 1 in the reference layout (a channels-last buffer of the fused plan is
  /// converted into a side buffer on demand).
  const CuMatrix<BaseFloat> &Activation(int32 i) { return FusedActive() ? FusedActivation(i) : forward_[i]; }
  /// Derivative with respect to the network input.  The fused plan does not compute it (nothing
  /// trains below the first layer): the matrix is then empty.
  const CuMatrix<BaseFloat> &InputDeriv() const { return derivs_.empty() ? empty_ : derivs_[0]; }
  /// True when the current configuration runs as the fused plan.
  bool FusedActive() const;
  double *ObjfDevice() { return objf_dev_; }
  double GetObjfAndReset();      // synchronises
  int32 NumRows() const { return num_rows_; }
 private:
  void SetInputPersists(bool on);
  void EagerStep(const CuMatrixBase<BaseFloat> &feats, const int32 *labels_dev);
  void DropGraph();
  uint64 StepKey(const CuMatrixBase<BaseFloat> &feats, const int32 *labels_dev) const;
  uint64 ConfigKey() const;
  struct GraphState;                 // recorded steps + the host-side effects of one step
  GraphState *graph_;
  bool last_replayed_;
  bool graphs_on_;
  bool fuse_;
  Nnet *nnet_;
  int32 num_rows_;
  // ---- fused plan (nnet-fused.cc)
  struct FusedOp;
  struct FusedState;
  FusedState *fused_;
  void FusedInit();
  void FusedDestroy();
  bool PlanFused();
  const CuMatrix<BaseFloat> &FusedActivation(int32 i);
  void FusedForward(int32 first, int32 last, const int32 *labels_dev);
  bool FusedObjf(const int32 *labels_dev);
  void FusedBackward(int32 last, int32 first);
  bool deferred_join_;
  cudaStream_t grad_stream_;
  const int32 *step_labels_;                    // TrainStep: labels known while the forward pass runs
  int32 base_;                                  // 1 when component 0 is the Splice front end, else 0

  std::vector<CuMatrix<BaseFloat> > forward_;   // [0] = copy-free view of the input
  std::vector<ChunkInfo> info_;
  // derivs_[i] = d objf / d forward_[i]: one buffer per activation, sized once -- nothing is
  // allocated or returned to the device cache while a step (or its recorded graph) runs
  std::vector<CuMatrix<BaseFloat> > derivs_;
  CuMatrix<BaseFloat> empty_;
  CuMatrix<BaseFloat> splice_out_;              // Splice output when it cannot be a view of the input
  const int32 *labels_;
  double *objf_dev_;
  std::vector<size_t> bucket_off_, bucket_len_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(NnetMinibatchUpdater);
};

}  // namespace nnet2
}  // namespace kaldi
#endif
