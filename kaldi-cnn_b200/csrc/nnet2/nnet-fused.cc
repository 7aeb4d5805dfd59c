// nnet2/nnet-fused.cc -- the fused training step of NnetMinibatchUpdater.
//
// nnet2's NnetUpdater drives one Propagate / Backprop call per component; for
// egs/exp/nnet/nnet.config that was 76 launches in round 1, a third of them staging copies and
// split-K reductions around GEMMs of a few microseconds.  The updater owns every activation and
// derivative buffer of the step, so it may do what a bare Component cannot assume:
//
//  * plan the LAYOUT of each buffer: between time-axis layers (in_height = 1) activations and
//    derivatives stay channels-last [N][W][C], the layout the convolution tensor maps read, so a
//    convolution's epilogue writes the next convolution's operand directly (no PaddingZero /
//    TpBlock / TpInsideBlock / FlipMat as in nnet0/nnet-component-nnet0.cc:423-446, 461-544,
//    738-777, and no staging pack either).  The reference layout appears only where an affine
//    layer reads the activation, and its input gradient is written back channels-last;
//  * fuse the element-wise components into the neighbouring GEMM: ReLU and dropout forward in the
//    producer's epilogue (upstream nnet2/nnet-component.cc:799-806, 3592-3620), their backward
//    as a gate in the CONSUMER's input-gradient epilogue (:813-827, 3634-3636); the ReLU around a
//    max-pool rides in the pool kernels; softmax + cross-entropy + softmax backward are one kernel;
//  * run each layer's weight gradient (+ momentum SGD in its epilogue, K-splits reduced inside the
//    kernel) on a side stream next to the NEXT layer's input gradient: both are small grids;
//  * take every column sum of the step -- bias gradients with their update, NonlinearComponent
//    statistics (:337-363) -- in two batched launches instead of one or two per layer;
//  * skip the input gradient of the first layer (nothing consumes it; nnet2's updater stops at the
//    first updatable component too).
//
// The component-by-component path (NnetMinibatchUpdater::ForwardRange / Backward in nnet-nnet.cc)
// stays as it was: it runs whenever the model, the math mode (FP32) or an alignment is outside
// this plan, and it is what the parity tests compare the fused step with.

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "nnet2/nnet-nnet.h"
#include "nnet0/nnet-component-nnet0.h"
#include "cnsl-cu-kernels.h"

namespace kaldi {
namespace nnet2 {

using cnsl::nnet0::ConvolutionComponent;
using cnsl::nnet0::FullyConnectedComponent;
using cnsl::nnet0::MaxpoolComponent;

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }

// Label + algorithmic work of the next launch, for the benchmark's per-launch roofline table
// (kcnn_profile_*; a no-op unless a recording is running).  FLOPs / bytes as SURVEY 8d counts them.
static void Tag(int32 comp, const char *what, double flops, double bytes) {
  char buf[96];
  snprintf(buf, sizeof(buf), "comp%d %s", comp, what);
  kcnn_profile_label(buf, flops, bytes);
}

struct NnetMinibatchUpdater::FusedOp {
  enum Kind { kConvFull, kConvTime, kPool, kAffine, kSoftmax };
  Kind kind;
  int32 comp;              // the convolution / pool / affine / softmax component
  int32 relu;              // RectifiedLinearComponent fused behind it (-1: none)
  int32 dropout;           // DropoutComponent fused behind the ReLU (-1: none)
  int32 last;              // last component index this op covers
  int32 in, out;           // forward_[in] is read; forward_[out] is the op's final output
  int32 producer;          // op that wrote forward_[in] (-1: the network input)
  bool need_dgrad;         // an updatable layer sits below: the input gradient is wanted
  // time-axis geometry of the op's input / output (kAffine / kSoftmax: unused)
  int32 W, C, OW, G;       // conv: input [W][C] -> output [OW][G]; pool: input [W][C] -> [OW][G]
};

struct NnetMinibatchUpdater::FusedState {
  bool valid;
  uint64 key;                                   // what the plan was built for
  std::vector<FusedOp> ops;
  std::vector<int32> op_of_comp;                // component index -> op index
  std::vector<char> act_cl, der_cl;             // layout of forward_[i] / derivs_[i]
  std::vector<CuMatrix<BaseFloat> > act_ref;    // Activation(i) of a channels-last buffer, on demand
  cudaStream_t side;
  std::vector<cudaEvent_t> fork_ev;
  cudaEvent_t join_ev;
  void *colsum_scratch[2];                      // [0] statistics launch, [1] bias launch
  size_t colsum_bytes[2];
  std::vector<KcnnColsumJob> pending_stats;     // statistics of a backward pass issued in several ranges
  bool objf_done;                               // the forward pass already ran softmax + objective
  bool fork_fc;
  bool stats_issued;                            // this backward pass sent its statistics out at the top
  FusedState() : valid(false), key(0), side(NULL), join_ev(NULL), objf_done(false), fork_fc(false), stats_issued(false) {
    colsum_scratch[0] = colsum_scratch[1] = NULL;
    colsum_bytes[0] = colsum_bytes[1] = 0;
  }
};

void NnetMinibatchUpdater::FusedInit() { fused_ = new FusedState; }

void NnetMinibatchUpdater::FusedDestroy() {
  if (!fused_) return;
  for (size_t i = 0; i < fused_->fork_ev.size(); i++) cudaEventDestroy(fused_->fork_ev[i]);
  if (fused_->join_ev) cudaEventDestroy(fused_->join_ev);
  if (fused_->side) cudaStreamDestroy(fused_->side);
  for (int i = 0; i < 2; i++)
    if (fused_->colsum_scratch[i]) CuDevice::Instantiate().Free(fused_->colsum_scratch[i]);
  delete fused_;
  fused_ = NULL;
}

static bool Dense(const CuMatrixBase<BaseFloat> &m) {
  return m.Stride() == m.NumCols() && (reinterpret_cast<uintptr_t>(m.Data()) & 15u) == 0;
}

// Builds the plan for the current model / row count, or marks it invalid.  Cheap when nothing
// changed (one hash over the component types and shapes).
bool NnetMinibatchUpdater::PlanFused() {
  FusedState &F = *fused_;
  const int32 L = nnet_->NumComponents();
  uint64 key = Component::HashValue(num_rows_, 29);
  key = Component::HashValue(CuDevice::Instantiate().MathMode(), key);
  key = Component::HashValue(fuse_, key);
  key = Component::HashValue(L, key);
  // the first op reads the caller's buffer through a tensor map: its alignment is part of the plan
  key = Component::HashValue(reinterpret_cast<uintptr_t>(forward_[base_].Data()) & 15u, key);
  key = Component::HashValue(forward_[base_].Stride() & 3, key);
  for (int32 c = 0; c < L; c++) {
    const Component &comp = nnet_->GetComponent(c);
    key = Component::HashValue(&comp, key);
    key = Component::HashValue(comp.InputDim(), key);
    key = Component::HashValue(comp.OutputDim(), key);
  }
  key |= 1;
  if (key == F.key) return F.valid;
  F.key = key;
  F.valid = false;
  F.ops.clear();
  static int enabled = -1, fork_fc = 0;
  if (enabled < 0) {
    const char *e = getenv("KCNN_NNET_PLAN");
    enabled = (e && e[0] == '0') ? 0 : 1;
    const char *f = getenv("KCNN_FUSED_FORK_FC");
    fork_fc = (f && f[0] == '1') ? 1 : 0;
  }
  F.fork_fc = fork_fc != 0;
  if (!enabled || !fuse_ || CuDevice::Instantiate().MathMode() != KCNN_MATH_TF32_TC || num_rows_ <= 0 || L < 2)
    return false;
  if ((reinterpret_cast<uintptr_t>(forward_[base_].Data()) & 15u) != 0 || (forward_[base_].Stride() & 3) != 0)
    return false;

  // ---- group the components into ops
  std::vector<FusedOp> ops;
  std::vector<int32> op_of_comp(L, -1);
  int32 producer = -1;
  bool seen_updatable = false;
  for (int32 c = base_; c < L;) {          // (a Splice front end is a view of the input, not an op)
    Component &comp = nnet_->GetComponent(c);
    FusedOp op;
    memset(&op, 0, sizeof(op));
    op.comp = c; op.relu = -1; op.dropout = -1; op.in = c; op.producer = producer;
    op.need_dgrad = seen_updatable;
    int32 next = c + 1;
    auto take_relu = [&]() {
      if (next < L && dynamic_cast<RectifiedLinearComponent *>(&nnet_->GetComponent(next)) != NULL) op.relu = next++;
    };
    if (ConvolutionComponent *cv = dynamic_cast<ConvolutionComponent *>(&comp)) {
      UpdatableComponent::StepTarget t;
      if (!cv->GetStepTarget(num_rows_, &t)) return false;
      op.W = cv->In_width(); op.C = cv->In_channels(); op.OW = cv->Out_width(); op.G = cv->Group();
      if (cv->In_height() == 1 && cv->Kernel_height() == 1 && cv->In_pad_height() == 0 && cv->Out_height() == 1) {
        op.kind = FusedOp::kConvTime;
        if (c == base_) return false;                  // a time-axis layer on the raw input: component path
        if (!kcnn_conv_time_shape_ok(num_rows_, op.W, op.C, cv->In_pad_width(), cv->Kernel_width(), op.G))
          return false;
      } else if (cv->Kernel_height() == cv->In_height() && cv->In_pad_height() == 0 && cv->In_pad_width() == 0 &&
                 cv->Out_height() == 1 && c == base_) {
        op.kind = FusedOp::kConvFull;
        if (!kcnn_conv_full_shape_ok(num_rows_, cv->In_height(), op.W, op.C, cv->Kernel_width(), op.G)) return false;
      } else {
        return false;
      }
      take_relu();
      seen_updatable = true;
    } else if (MaxpoolComponent *mp = dynamic_cast<MaxpoolComponent *>(&comp)) {
      if (mp->In_height() != 1 || mp->Pool_height_dim() != 1 || mp->Overlap() || mp->Overlap2D() ||
          mp->IndexRouting())
        return false;
      if (mp->In_width() % mp->Pool_width_dim() != 0 || mp->In_channel() % mp->Pool_channel_dim() != 0) return false;
      op.kind = FusedOp::kPool;
      op.W = mp->In_width(); op.C = mp->In_channel();
      op.OW = mp->In_width() / mp->Pool_width_dim(); op.G = mp->In_channel() / mp->Pool_channel_dim();
      if (op.OW * op.G != mp->OutputDim() || op.W * op.C != mp->InputDim()) return false;
      take_relu();
    } else if (FullyConnectedComponent *fc = dynamic_cast<FullyConnectedComponent *>(&comp)) {
      UpdatableComponent::StepTarget t;
      if (!fc->GetStepTarget(num_rows_, &t)) return false;
      op.kind = FusedOp::kAffine;
      take_relu();
      if (op.relu >= 0 && next < L && dynamic_cast<DropoutComponent *>(&nnet_->GetComponent(next)) != NULL)
        op.dropout = next++;
      seen_updatable = true;
    } else if (dynamic_cast<SoftmaxComponent *>(&comp) != NULL) {
      if (c != L - 1 || comp.OutputDim() > 4096) return false;
      op.kind = FusedOp::kSoftmax;
    } else {
      return false;                                    // anything else: component path
    }
    op.last = next - 1;
    op.out = next;
    for (int32 k = c; k < next; k++) op_of_comp[k] = static_cast<int32>(ops.size());
    producer = static_cast<int32>(ops.size());
    ops.push_back(op);
    c = next;
  }
  if (ops.empty() || ops.back().kind != FusedOp::kSoftmax) return false;

  // ---- layouts.  Activations: channels-last between time-axis ops; the reference layout at the
  // input of an affine layer.  Derivatives: channels-last wherever a time-axis op consumes them.
  std::vector<char> act_cl(L + 1, 0), der_cl(L + 1, 0);
  for (size_t i = 0; i < ops.size(); i++) {
    const FusedOp &op = ops[i];
    const bool time_in = op.kind == FusedOp::kConvTime || op.kind == FusedOp::kPool;
    const bool time_out = time_in || op.kind == FusedOp::kConvFull;
    if (time_in) {
      if (op.producer < 0) return false;
      const FusedOp &p = ops[op.producer];
      if (p.kind != FusedOp::kConvFull && p.kind != FusedOp::kConvTime && p.kind != FusedOp::kPool) return false;
      if (p.OW != op.W || p.G != op.C) return false;
      if (p.dropout >= 0) return false;
    }
    if (time_out) {
      if (i + 1 >= ops.size()) return false;
      const FusedOp &nx = ops[i + 1];
      const bool next_time = nx.kind == FusedOp::kConvTime || nx.kind == FusedOp::kPool;
      if (!next_time && nx.kind != FusedOp::kAffine) return false;
      if ((op.G & 3) != 0) return false;
      const bool cl = next_time;
      for (int32 k = op.comp + 1; k <= op.out; k++) act_cl[k] = cl;
      der_cl[op.out] = 1;                              // an affine consumer writes its input gradient channels-last
      if (op.kind == FusedOp::kConvFull && !cl) return false;   // the full-height epilogue here only writes channels-last
    }
  }

  // ---- buffers: activations are sized by ForwardRange; channels-last ones must be dense
  derivs_.resize(L + 1);
  for (size_t i = 0; i < ops.size(); i++) {
    const FusedOp &op = ops[i];
    for (int32 k = op.comp + 1; k <= op.out; k++)
      if (act_cl[k] && !Dense(forward_[k])) return false;
    if (op.kind == FusedOp::kSoftmax) continue;
    // derivative with respect to the op's (pre-ReLU) output
    derivs_[op.out].Resize(num_rows_, forward_[op.out].NumCols(), kUndefined);
    if (der_cl[op.out] && !Dense(derivs_[op.out])) return false;
  }
  F.act_ref.clear();
  F.act_ref.resize(L + 1);

  // ---- side stream + events (never created inside a stream capture: planning runs in the first,
  // eager, step of a configuration)
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(Str(), &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return false;
  }
  static int side_enabled = -1;
  if (side_enabled < 0) {
    const char *e = getenv("KCNN_SIDE_STREAM");
    side_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (side_enabled && F.side == NULL) {
    if (cudaStreamCreateWithFlags(&F.side, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); F.side = NULL; }
    if (F.side && cudaEventCreateWithFlags(&F.join_ev, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      cudaStreamDestroy(F.side);
      F.side = NULL;
    }
  }
  while (F.side && F.fork_ev.size() < ops.size() + 2) {
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
    F.fork_ev.push_back(ev);
  }
  F.ops.swap(ops);
  F.op_of_comp.swap(op_of_comp);
  F.act_cl.swap(act_cl);
  F.der_cl.swap(der_cl);
  F.valid = true;
  return true;
}

bool NnetMinibatchUpdater::FusedActive() const { return fused_ != NULL && fused_->valid; }

// Reference-layout copy of a channels-last activation, for Activation(i).
const CuMatrix<BaseFloat> &NnetMinibatchUpdater::FusedActivation(int32 i) {
  FusedState &F = *fused_;
  if (!F.valid || i < 0 || i >= static_cast<int32>(F.act_cl.size()) || !F.act_cl[i]) return forward_[i];
  const FusedOp &op = F.ops[F.op_of_comp[i - 1]];
  CuMatrix<BaseFloat> &dst = F.act_ref[i];
  dst.Resize(forward_[i].NumRows(), forward_[i].NumCols(), kUndefined);
  cudaF_cl_to_ref(Str(), forward_[i].Data(), forward_[i].NumRows(), op.OW, op.G, dst.Data(), dst.Stride());
  CU_SAFE_CALL(cudaGetLastError());
  return dst;
}

// ------------------------------------------------------------------------------- forward --

void NnetMinibatchUpdater::FusedForward(int32 first, int32 last, const int32 *labels_dev) {
  FusedState &F = *fused_;
  cudaStream_t st = Str();
  F.objf_done = false;
  const int32 o_first = F.op_of_comp[first], o_last = F.op_of_comp[last];
  KALDI_ASSERT(F.ops[o_first].comp == first && F.ops[o_last].last == last &&
               "ForwardRange: the range must not cut through a fused group of components");
  std::vector<unsigned long long *> seeds;
  for (int32 i = o_first; i <= o_last; i++) {
    const FusedOp &op = F.ops[i];
    Component &comp = nnet_->GetComponent(op.comp);
    const CuMatrix<BaseFloat> &in = forward_[op.in];
    int ok = 1;
    switch (op.kind) {
      case FusedOp::kConvFull: {
        ConvolutionComponent &cv = static_cast<ConvolutionComponent &>(comp);
        Tag(op.comp, "conv fprop (+bias+ReLU)", 2.0 * num_rows_ * op.OW * op.G * cv.KernelDim(),
            4.0 * ((double)num_rows_ * (cv.InputDim() + op.OW * op.G) + (double)cv.KernelDim() * op.G));
        ok = cudaF_conv_full_fprop_cl(st, in.Data(), in.Dim(), cv.In_height(), cv.In_width(), cv.In_channels(),
                                      cv.Kernel_width(), cv.Group(), cv.LinearParams().Data(), cv.LinearParams().Dim(),
                                      cv.BiasParams().Data(), forward_[op.out].Data(), op.relu >= 0);
        break;
      }
      case FusedOp::kConvTime: {
        ConvolutionComponent &cv = static_cast<ConvolutionComponent &>(comp);
        CuMatrix<BaseFloat> &out = forward_[op.out];
        Tag(op.comp, "conv fprop (+bias+ReLU)", 2.0 * num_rows_ * op.OW * op.G * cv.KernelDim(),
            4.0 * ((double)num_rows_ * (op.W * op.C + op.OW * op.G) + (double)cv.KernelDim() * op.G));
        ok = cudaF_conv_time_fprop_cl(st, in.Data(), num_rows_, op.W, op.C, cv.In_pad_width(), cv.Kernel_width(), op.G,
                                      cv.LinearParams().Data(), cv.LinearParams().Dim(), cv.BiasParams().Data(),
                                      out.Data(), F.act_cl[op.out], out.Stride(), op.relu >= 0);
        break;
      }
      case FusedOp::kPool: {
        MaxpoolComponent &mp = static_cast<MaxpoolComponent &>(comp);
        CuMatrix<BaseFloat> &pool = forward_[op.comp + 1];
        Tag(op.comp, "maxpool fwd (+ReLU)", 0.0, 4.0 * num_rows_ * (op.W * op.C + op.OW * op.G));
        cudaF_maxpool_prop_cl(st, in.Data(), num_rows_, op.W, op.C, mp.Pool_width_dim(), mp.Pool_channel_dim(),
                              pool.Data(), op.relu >= 0 ? forward_[op.out].Data() : NULL,
                              F.act_cl[op.out] ? 0 : pool.Stride());
        if (op.relu >= 0 && !F.act_cl[op.out]) KALDI_ASSERT(pool.Stride() == forward_[op.out].Stride());
        break;
      }
      case FusedOp::kAffine: {
        FullyConnectedComponent &fc = static_cast<FullyConnectedComponent &>(comp);
        CuMatrix<BaseFloat> &out = forward_[op.relu >= 0 ? op.relu + 1 : op.comp + 1];
        float *drop = NULL;
        ::MatrixDim dd = {0, 0, 0};
        float dp = 0.f, low = 0.f, high = 0.f;
        unsigned long long *seed = NULL;
        if (op.dropout >= 0) {
          DropoutComponent &dc = static_cast<DropoutComponent &>(nnet_->GetComponent(op.dropout));
          dp = dc.DropoutProportion();
          KALDI_ASSERT(dp < 1.0 && dp >= 0.0 && dc.DropoutScale() <= 1.0 && dc.DropoutScale() >= 0.0);
          low = dc.DropoutScale();
          high = (1.0 - (dp * low)) / (1.0 - dp);
          seed = dc.SeedDevice();
          seeds.push_back(seed);
          drop = forward_[op.out].Data();
          dd = forward_[op.out].Dim();
        }
        Tag(op.comp, "affine fprop (+bias+ReLU+dropout)", 2.0 * num_rows_ * fc.InputDim() * fc.OutputDim(),
            4.0 * ((double)num_rows_ * (fc.InputDim() + fc.OutputDim()) + (double)fc.InputDim() * fc.OutputDim()));
        ok = cudaF_affine_fprop_fused(st, in.Data(), in.Dim(), fc.LinearParams().Data(), fc.LinearParams().Dim(),
                                      fc.BiasParams().Data(), out.Data(), out.Dim(), op.relu >= 0, drop, dd, dp, low,
                                      high, seed);
        break;
      }
      case FusedOp::kSoftmax: {
        CuMatrix<BaseFloat> &post = forward_[op.out];
        bool fused = false;
        Tag(op.comp, "softmax + x-ent + softmax bwd", 0.0, 12.0 * num_rows_ * in.NumCols());
        if (labels_dev != NULL) {
          // the whole step is known: softmax, objective, derivative and softmax backward at once
          derivs_[op.in].Resize(num_rows_, in.NumCols(), kUndefined);
          fused = cudaF_softmax_xent(st, in.Data(), in.Dim(), post.Data(), post.Dim(), labels_dev,
                                     derivs_[op.in].Data(), derivs_[op.in].Dim(), objf_dev_,
                                     seeds.empty() ? NULL : &seeds[0], static_cast<int>(seeds.size())) != 0;
          if (fused) { F.objf_done = true; seeds.clear(); }
        }
        if (!fused) cudaF_softmax_fprop(st, in.Data(), in.Dim(), post.Data(), post.Dim());
        break;
      }
    }
    if (!ok) KALDI_ERR << "fused step: component " << op.comp << " (" << comp.Type() << ") was planned but its "
                       << "kernel rejected the shape";
  }
  if (!seeds.empty()) cudaF_bump_seeds(st, &seeds[0], static_cast<int>(seeds.size()));
  CU_SAFE_CALL(cudaGetLastError());
}

// Objective + derivative at the output.  Returns false when the plan is not active.
bool NnetMinibatchUpdater::FusedObjf(const int32 *labels_dev) {
  FusedState &F = *fused_;
  if (!F.valid) return false;
  if (F.objf_done) { F.objf_done = false; return true; }        // done inside the forward pass
  const FusedOp &op = F.ops.back();
  CuMatrix<BaseFloat> &post = forward_[op.out];
  derivs_[op.in].Resize(num_rows_, post.NumCols(), kUndefined);
  ::MatrixDim none = {0, 0, 0};
  Tag(op.comp, "x-ent + softmax bwd", 0.0, 8.0 * num_rows_ * post.NumCols());
  if (!cudaF_softmax_xent(Str(), NULL, none, post.Data(), post.Dim(), labels_dev, derivs_[op.in].Data(),
                          derivs_[op.in].Dim(), objf_dev_, NULL, 0))
    KALDI_ERR << "fused step: softmax row too long";
  CU_SAFE_CALL(cudaGetLastError());
  return true;
}

// ------------------------------------------------------------------------------ backward --

void NnetMinibatchUpdater::FusedBackward(int32 last, int32 first) {
  FusedState &F = *fused_;
  cudaStream_t st = Str();
  const int32 o_first = F.op_of_comp[first], o_last = F.op_of_comp[last];
  KALDI_ASSERT(F.ops[o_first].comp == first && F.ops[o_last].last == last &&
               "Backward: the range must not cut through a fused group of components");
  std::vector<KcnnColsumJob> &stat_jobs = F.pending_stats;
  std::vector<KcnnColsumJob> bias_jobs;
  if (last == nnet_->NumComponents() - 1) stat_jobs.clear();     // a new backward pass starts at the top
  bool forked = false;
  bool dirty = false;                          // st has produced a derivative the side branch is not ordered behind yet
  size_t ev = 0;
  auto fork = [&]() -> cudaStream_t {          // work issued on the returned stream runs beside what follows on st
    if (F.side == NULL || ev >= F.fork_ev.size()) return st;
    if (kcnn_profile_active()) return st;      // per-launch timing: one stream, nothing shares the SMs
    if (cudaEventRecord(F.fork_ev[ev], st) != cudaSuccess || cudaStreamWaitEvent(F.side, F.fork_ev[ev], 0) != cudaSuccess) {
      cudaGetLastError();
      return st;
    }
    ev++;
    forked = true;
    dirty = false;
    return F.side;
  };
  // The bias column sums of the range run on the side branch at the end and read EVERY out_deriv of the range.
  // A weight gradient that stays on st (an affine layer's; the first layer's, which has no input gradient to
  // run beside) does not fork, so derivatives produced on st since the last fork -- e.g. by the max-pool backward
  // between conv2 and conv1 -- would be read by the side branch without an ordering.  Called before such a
  // weight gradient is issued for the LAST op of the range (every producer of the range has been issued by
  // then): one more edge st -> side, placed where it costs the side branch nothing.
  auto order_side = [&]() { if (forked && dirty) fork(); };
  // gate of the op that produced forward_[op.in]: its ReLU (and dropout) backward, applied by the consumer
  auto gate = [&](const FusedOp &op, const float **x, int *ldx, const float **y, int *ldy) {
    *x = NULL; *y = NULL; *ldx = 0; *ldy = 0;
    if (op.producer < 0) return;
    const FusedOp &p = F.ops[op.producer];
    if (p.relu < 0) return;
    const CuMatrix<BaseFloat> &r = forward_[p.relu + 1];
    *x = r.Data(); *ldx = r.Stride();
    if (p.dropout >= 0) { *y = forward_[p.out].Data(); *ldy = forward_[p.out].Stride(); }
  };
  // statistics of the fused nonlinearities (NonlinearComponent::UpdateStats, done in Backprop upstream):
  // they read forward activations only
  auto stats_of = [&](const FusedOp &op) {
    if (op.relu >= 0) {
      NonlinearComponent &nl = static_cast<NonlinearComponent &>(nnet_->GetComponent(op.relu));
      const CuMatrix<BaseFloat> &y = forward_[op.relu + 1];
      double *s = nl.StatsDevice();
      KcnnColsumJob j = {y.Data(), y.NumRows(), y.NumCols(), y.Stride(), KCNN_COLSUM_STATS_RELU, 0, 0, s,
                         s + nl.InputDim(), 0.f};
      if (F.act_cl[op.relu + 1]) { j.perm_w = op.OW; j.perm_c = op.G; }
      stat_jobs.push_back(j);
      nl.AddToCount(num_rows_);
    }
    if (op.kind == FusedOp::kSoftmax) {
      NonlinearComponent &nl = static_cast<NonlinearComponent &>(nnet_->GetComponent(op.comp));
      const CuMatrix<BaseFloat> &y = forward_[op.out];
      KcnnColsumJob j = {y.Data(), y.NumRows(), y.NumCols(), y.Stride(), KCNN_COLSUM_STATS_VALUE, 0, 0,
                         nl.StatsDevice(), NULL, 0.f};
      stat_jobs.push_back(j);
      nl.AddToCount(num_rows_);
    }
  };
  auto launch_colsums = [&](std::vector<KcnnColsumJob> &jobs, int k, cudaStream_t cst) {
    if (jobs.empty()) return;
    const size_t need = kcnn_colsum_batch_scratch_bytes(&jobs[0], static_cast<int>(jobs.size()));
    if (need > F.colsum_bytes[k]) {
      // (first, eager, step only: a captured step finds the buffer in place)
      if (F.colsum_scratch[k]) CuDevice::Instantiate().Free(F.colsum_scratch[k]);
      F.colsum_scratch[k] = CuDevice::Instantiate().Malloc(need);
      F.colsum_bytes[k] = need;
      CU_SAFE_CALL(cudaMemsetAsync(F.colsum_scratch[k], 0, need, cst));
    }
    double bytes = 0.0;
    for (size_t j = 0; j < jobs.size(); j++) bytes += 4.0 * jobs[j].rows * jobs[j].cols;
    Tag(-1, k == 0 ? "nonlinearity statistics (all layers, batched column sums)"
                   : "bias gradients + update (batched column sums)", 0.0, bytes);
    cudaF_colsum_batch(cst, &jobs[0], static_cast<int>(jobs.size()), F.colsum_scratch[k]);
  };
  // A pass that starts at the top and will reach the bottom (one call, or the data-parallel trainer's
  // layer-by-layer calls): the statistics of ALL layers go out first, on the side branch, instead of at the
  // end of the pass on the compute stream (32 us of the critical path for the benchmarked model).
  const bool whole_pass = last == nnet_->NumComponents() - 1 && (first == base_ || deferred_join_);
  if (whole_pass && F.side != NULL) {
    for (int32 i = o_last; i >= 0; i--) stats_of(F.ops[i]);
    cudaStream_t ss = fork();
    launch_colsums(stat_jobs, 0, ss);
    stat_jobs.clear();
    F.stats_issued = true;
  } else if (last == nnet_->NumComponents() - 1) {
    F.stats_issued = false;
  }
  for (int32 i = o_last; i >= o_first; i--) {
    const FusedOp &op = F.ops[i];
    Component &comp = nnet_->GetComponent(op.comp);
    if (!F.stats_issued) stats_of(op);
    if (op.kind == FusedOp::kSoftmax) continue;  // derivs_[op.in] was written with the objective
    const CuMatrix<BaseFloat> &dy = derivs_[op.out];
    const CuMatrix<BaseFloat> &x = forward_[op.in];
    const float *gx, *gy;
    int ldx, ldy;
    gate(op, &gx, &ldx, &gy, &ldy);
    int ok = 1;
    if (op.kind == FusedOp::kPool) {
      MaxpoolComponent &mp = static_cast<MaxpoolComponent &>(comp);
      const CuMatrix<BaseFloat> &pool = forward_[op.comp + 1];
      KALDI_ASSERT(gy == NULL);
      Tag(op.comp, "maxpool bwd (+ReLU gate)", 0.0, 4.0 * num_rows_ * (2.0 * op.W * op.C + 2.0 * op.OW * op.G));
      cudaF_maxpool_backprop_cl(st, x.Data(), pool.Data(), F.act_cl[op.comp + 1] ? 0 : pool.Stride(), dy.Data(),
                                num_rows_, op.W, op.C, mp.Pool_width_dim(), mp.Pool_channel_dim(),
                                derivs_[op.in].Data(), gx != NULL);
      dirty = true;
      continue;
    }
    UpdatableComponent &uc = static_cast<UpdatableComponent &>(comp);
    UpdatableComponent::StepTarget t;
    if (!uc.GetStepTarget(num_rows_, &t)) KALDI_ERR << "fused step: component " << op.comp << " lost its update target";
    const int apply = t.deferred ? 0 : 1;
    float *wout = apply ? t.w : t.w_grad;
    ::MatrixDim wod = apply ? t.wd : t.gd;
    if (op.kind == FusedOp::kAffine) {
      if (op.need_dgrad) {
        int perm_r = 0;
        if (F.der_cl[op.in]) perm_r = F.ops[op.producer].OW;
        CuMatrix<BaseFloat> &dx = derivs_[op.in];
        Tag(op.comp, "affine dgrad (+ReLU/dropout gate)", 2.0 * num_rows_ * t.wd.rows * t.wd.cols,
            4.0 * ((double)num_rows_ * (t.wd.rows + 2.0 * t.wd.cols) + (double)t.wd.rows * t.wd.cols));
        ok = cudaF_affine_dgrad_fused(st, dy.Data(), dy.Dim(), t.w, t.wd, dx.Data(), dx.Dim(), gx, ldx, gy, ldy, perm_r);
        dirty = true;
      }
      cudaStream_t ws = (F.fork_fc && op.need_dgrad) ? fork() : st;
      if (ws == st && i == o_first) order_side();
      // with the SGD step in the epilogue W and prev_grad are read and written: 16 B per weight
      Tag(op.comp, apply ? "affine wgrad + SGD" : "affine wgrad", 2.0 * num_rows_ * t.wd.rows * t.wd.cols,
          (apply ? 16.0 : 4.0) * t.wd.rows * t.wd.cols + 4.0 * num_rows_ * (t.wd.rows + t.wd.cols));
      if (apply)
        ok = ok && cudaF_affine_wgrad_sgd(ws, KCNN_MATH_TF32_TC, x.Data(), x.Dim(), dy.Data(), dy.Dim(), t.w, t.wd, t.prev,
                                          t.pd, NULL, t.momentum, t.a_decay, t.a_grad);
      else
        cudaF_affine_wgrad(ws, KCNN_MATH_TF32_TC, x.Data(), x.Dim(), dy.Data(), dy.Dim(), t.w_grad, t.gd, NULL);
      KcnnColsumJob j = {dy.Data(), dy.NumRows(), dy.NumCols(), dy.Stride(), apply ? KCNN_COLSUM_AXPY : KCNN_COLSUM_STORE,
                         0, 0, apply ? t.bias : t.b_grad, NULL, t.a_grad};
      bias_jobs.push_back(j);
    } else {
      ConvolutionComponent &cv = static_cast<ConvolutionComponent &>(comp);
      KALDI_ASSERT(gy == NULL);
      if (op.need_dgrad) {
        KALDI_ASSERT(op.kind == FusedOp::kConvTime);
        Tag(op.comp, "conv dgrad (+ReLU gate)", 2.0 * num_rows_ * op.OW * op.G * cv.KernelDim(),
            4.0 * ((double)num_rows_ * (2.0 * op.W * op.C + op.OW * op.G) + (double)cv.KernelDim() * op.G));
        ok = cudaF_conv_time_dgrad_cl(st, dy.Data(), num_rows_, op.W, op.C, cv.In_pad_width(), cv.Kernel_width(), op.G,
                                      t.w, t.wd, derivs_[op.in].Data(), gx);
        dirty = true;
      }
      cudaStream_t ws = op.need_dgrad ? fork() : st;      // the update writes the kernel dgrad reads: behind it
      if (ws == st && i == o_first) order_side();
      Tag(op.comp, apply ? "conv wgrad + SGD" : "conv wgrad", 2.0 * num_rows_ * op.OW * op.G * cv.KernelDim(),
          4.0 * ((double)num_rows_ * (x.NumCols() + op.OW * op.G)) + (apply ? 16.0 : 4.0) * cv.KernelDim() * op.G);
      if (op.kind == FusedOp::kConvTime)
        ok = ok && cudaF_conv_time_wgrad_cl(ws, x.Data(), dy.Data(), num_rows_, op.W, op.C, cv.In_pad_width(),
                                            cv.Kernel_width(), op.G, wout, wod, t.prev, t.pd, apply, t.momentum,
                                            t.a_decay, t.a_grad);
      else
        ok = ok && cudaF_conv_full_wgrad_cl(ws, x.Data(), x.Dim(), dy.Data(), cv.In_height(), cv.In_width(),
                                            cv.In_channels(), cv.Kernel_width(), cv.Group(), wout, wod, t.prev, t.pd,
                                            apply, t.momentum, t.a_decay, t.a_grad);
      KcnnColsumJob j = {dy.Data(), num_rows_ * op.OW, op.G, op.G, apply ? KCNN_COLSUM_AXPY : KCNN_COLSUM_STORE,
                         0, 0, apply ? t.bias : t.b_grad, NULL, t.a_grad};
      bias_jobs.push_back(j);
    }
    if (!ok) KALDI_ERR << "fused step: component " << op.comp << " (" << comp.Type() << ") was planned but a "
                       << "backward kernel rejected the shape";
  }
  // the bias gradients of the range (and, for a pass issued in partial ranges, the statistics once it has
  // reached the bottom): one launch each, beside the tail of the GEMM chain
  order_side();                                // (nothing to do unless the range ended on a pooling layer)
  cudaStream_t cst = (F.side != NULL && forked) ? F.side : st;
  if (!F.stats_issued && first == base_) launch_colsums(stat_jobs, 0, cst);
  launch_colsums(bias_jobs, 1, cst);
  if (first == base_) stat_jobs.clear();
  grad_stream_ = forked ? F.side : st;
  if (forked && !deferred_join_) {
    cudaEventRecord(F.join_ev, F.side);
    cudaStreamWaitEvent(st, F.join_ev, 0);
  }
  CU_SAFE_CALL(cudaGetLastError());
}

void NnetMinibatchUpdater::JoinSide() {
  if (fused_ == NULL || fused_->side == NULL || fused_->join_ev == NULL) return;
  CU_SAFE_CALL(cudaEventRecord(fused_->join_ev, fused_->side));
  CU_SAFE_CALL(cudaStreamWaitEvent(Str(), fused_->join_ev, 0));
}

}  // namespace nnet2
}  // namespace kaldi
