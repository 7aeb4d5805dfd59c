// nnet2/nnet-dp.cc -- see nnet-dp.h.

#include <stdlib.h>

#include <algorithm>

#include <cuda_runtime_api.h>

#include "nnet2/nnet-dp.h"
#include "kcnn_capi.h"
#include "cnsl-cu-kernels.h"

namespace kaldi {
namespace nnet2 {

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }

// Programmatic dependent launch only in the FORWARD half of what the trainer launches or records: in the
// backward half dependents parked on the SMs get in the way of the communication kernels and of the
// weight-gradient branch.
namespace {
// KCNN_DP_PDL: fwd (default) on in the forward half of the rotation only, 0 off everywhere, 1 on everywhere,
// bwd backward half only.  Measured at 2 GPUs: fwd 0.810 ms, off 0.821 ms, bwd 0.982 ms, on 0.969 ms.
static int DpPdlMode() {
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("KCNN_DP_PDL");
    mode = 1;
    if (e && e[0] == '0') mode = 0;
    else if (e && e[0] == '1') mode = 3;
    else if (e && e[0] == 'f') mode = 1;
    else if (e && e[0] == 'b') mode = 2;
  }
  return mode;
}
struct NoPdl {
  int before;
  bool active;
  explicit NoPdl(int part = 3) : before(1), active(true) {     // part: 1 forward, 2 backward, 3 both
    active = (DpPdlMode() & part) != part;
    if (active) before = kcnn_set_pdl(0);
  }
  ~NoPdl() { if (active) kcnn_set_pdl(before); }
};
}  // namespace

size_t NnetDataParallel::ArenaFloats(NnetMinibatchUpdater *updater) {
  return 2 * updater->GradientFloats() + kcnn_p2p_flag_floats();
}

NnetDataParallel::NnetDataParallel(Nnet *nnet, NnetMinibatchUpdater *updater, int32 rank, int32 world,
                                   float *local_base, const unsigned long long *peer_bases,
                                   unsigned long long multicast_base)
    : nnet_(nnet), updater_(updater), rank_(rank), world_(world), base_(local_base), multicast_(multicast_base),
      error_pinned_(NULL), primed_(false), last_replayed_(false) {
  KALDI_ASSERT(world >= 1 && world <= 8 && rank >= 0 && rank < world && local_base != NULL);
  peers_.assign(peer_bases, peer_bases + world);
  KALDI_ASSERT(peers_[rank] == reinterpret_cast<unsigned long long>(local_base));
  grad_floats_ = updater_->GradientFloats();
  flag_off_ = 2 * grad_floats_;
  updater_->SetDeferredUpdate(true);
  updater_->SetDeferredJoin(true);
  updater_->SetGradientArena(base_);                         // [0, grad_floats_): one bucket per layer, top first
  comm_[0] = comm_[1] = NULL;
  int lo = 0, hi = 0;
  CU_SAFE_CALL(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // reductions first: short CTAs, on the critical path
  for (int i = 0; i < 2; i++) CU_SAFE_CALL(cudaStreamCreateWithPriority(&comm_[i], cudaStreamNonBlocking, hi));
  for (int32 c = 0; c < nnet_->NumComponents(); c++) {
    UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(&nnet_->GetComponent(c));
    if (u == NULL) continue;
    UpdatableComponent::StepTarget t;
    if (!u->GetStepTarget(1, &t))
      KALDI_ERR << "NnetDataParallel: component " << c << " (" << u->Type() << ") has no momentum update target";
    Layer l;
    l.comp = c;
    updater_->GradientBucket(c, &l.off, &l.len);
    u->SetParameterStorage(base_ + grad_floats_ + l.off);    // the parameter arena mirrors the gradient arena
    if (!u->GetStepTarget(1, &t)) KALDI_ERR << "NnetDataParallel: lost the update target";
    l.weight_floats = (size_t)t.wd.rows * t.wd.stride;
    KALDI_ASSERT(t.w == base_ + grad_floats_ + l.off && t.pd.stride == t.wd.stride && l.weight_floats <= l.len);
    layers_.push_back(l);
  }
  static int group_small = -1;
  if (group_small < 0) {
    const char *e = getenv("KCNN_DP_GROUP");             // KCNN_DP_GROUP=0: one launch per layer
    group_small = (e && e[0] == '0') ? 0 : 1;
  }
  for (size_t i = 0; i < layers_.size(); i++) {
    const bool small = layers_[i].len < (1u << 20);
    if (group_small && small && !groups_.empty() && groups_.back().channel == 1 &&
        groups_.back().last == (int32)i - 1 && groups_.back().last - groups_.back().first + 1 < 8) {
      groups_.back().last = (int32)i;
      continue;
    }
    Group g;
    g.first = g.last = (int32)i;
    g.channel = small ? 1 : 0;                           // convolution buckets must not queue behind the FC stack
    CU_SAFE_CALL(cudaEventCreateWithFlags(&g.ready, cudaEventDisableTiming));
    CU_SAFE_CALL(cudaEventCreateWithFlags(&g.ready_compute, cudaEventDisableTiming));
    CU_SAFE_CALL(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
    groups_.push_back(g);
  }
  CU_SAFE_CALL(cudaMallocHost(reinterpret_cast<void **>(&error_pinned_), 2 * sizeof(unsigned int)));
  error_pinned_[0] = error_pinned_[1] = 0u;
  CU_SAFE_CALL(cudaStreamSynchronize(Str()));
}

NnetDataParallel::~NnetDataParallel() {
  cudaStreamSynchronize(Str());
  DropGraphs();
  for (int i = 0; i < 2; i++)
    if (comm_[i]) { cudaStreamSynchronize(comm_[i]); cudaStreamDestroy(comm_[i]); }
  for (size_t i = 0; i < layers_.size(); i++) {
    // parameters leave the arena (which belongs to the caller) with their current values
    UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(&nnet_->GetComponent(layers_[i].comp));
    if (u) { u->SetParameterStorage(NULL); u->SetGradientStorage(NULL); }
  }
  for (size_t i = 0; i < groups_.size(); i++) {
    cudaEventDestroy(groups_[i].ready);
    cudaEventDestroy(groups_[i].ready_compute);
    cudaEventDestroy(groups_[i].done);
  }
  updater_->SetDeferredUpdate(false);
  updater_->SetDeferredJoin(false);
  if (error_pinned_) cudaFreeHost(error_pinned_);
}

void NnetDataParallel::DropGraphs() {
  for (size_t i = 0; i < graphs_.size(); i++)
    if (graphs_[i].exec) { cudaGraphExecDestroy(graphs_[i].exec); CuDevice::Instantiate().GraphDestroyed(); }
  graphs_.clear();
  seen_.clear();
}

// One kernel per group on the group's communication stream, behind everything the backward pass has issued
// for its layers (weight gradients, bias gradients, and the input-gradient GEMMs that read the weights).
void NnetDataParallel::ReduceAndUpdate(const Group &g, int32 rows_global) {
  KcnnSgdBucket b[8];
  int n = 0;
  for (int32 i = g.first; i <= g.last; i++) {
    const Layer &l = layers_[i];
    UpdatableComponent *u = static_cast<UpdatableComponent *>(&nnet_->GetComponent(l.comp));
    UpdatableComponent::StepTarget t;
    if (!u->GetStepTarget(rows_global, &t)) KALDI_ERR << "NnetDataParallel: lost the update target";
    b[n].offset_floats = l.off; b[n].count_floats = l.len; b[n].weight_floats = l.weight_floats;
    b[n].prev_grad = t.prev; b[n].momentum = t.momentum; b[n].decay_alpha = t.a_decay; b[n].grad_alpha = t.a_grad;
    n++;
  }
  // the gradients complete on the updater's weight-gradient branch (or on the compute stream when the
  // range had none); that branch starts behind the layers' input-gradient GEMMs, the last readers of W
  CU_SAFE_CALL(cudaEventRecord(g.ready, updater_->GradientStream()));
  CU_SAFE_CALL(cudaStreamWaitEvent(comm_[g.channel], g.ready, 0));
  // ... and behind the compute stream too: the branch exists as soon as the pass has sent its statistics out,
  // but an affine layer's weight gradient and the first layer's (no input gradient to run beside) are issued
  // on the compute stream, as are the input-gradient GEMMs that still read the weights this kernel replaces
  if (updater_->GradientStream() != Str()) {
    CU_SAFE_CALL(cudaEventRecord(g.ready_compute, Str()));
    CU_SAFE_CALL(cudaStreamWaitEvent(comm_[g.channel], g.ready_compute, 0));
  }
  int rc;
  if (n == 1)
    rc = kcnn_p2p_reduce_sgd_f32(comm_[g.channel], &peers_[0], multicast_, rank_, world_, b[0].offset_floats,
                                 b[0].count_floats, b[0].weight_floats, grad_floats_, b[0].prev_grad, b[0].momentum,
                                 b[0].decay_alpha, b[0].grad_alpha, flag_off_, g.channel);
  else
    rc = kcnn_p2p_reduce_sgd_multi_f32(comm_[g.channel], &peers_[0], multicast_, rank_, world_, n, b, grad_floats_,
                                       flag_off_, g.channel);
  if (rc != 0) KALDI_ERR << "kcnn_p2p_reduce_sgd rejected its arguments";
  CU_SAFE_CALL(cudaEventRecord(g.done, comm_[g.channel]));
}

void NnetDataParallel::BackwardWithUpdates(int32 rows_global) {
  NoPdl no_pdl(2);
  int32 hi = nnet_->NumComponents() - 1;
  for (size_t i = groups_.size(); i-- > 0;) {                // top group first: the order backward produces them
    const int32 lowest = layers_[groups_[i].first].comp;
    updater_->Backward(hi, lowest);
    ReduceAndUpdate(groups_[i], rows_global);
    hi = lowest - 1;
  }
  if (hi >= 0) updater_->Backward(hi, 0);
  updater_->JoinSide();
}

void NnetDataParallel::ForwardBehindUpdates(const CuMatrixBase<BaseFloat> &feats, const int32 *labels) {
  NoPdl no_pdl(1);
  const int32 L = nnet_->NumComponents();
  int32 first = 0;
  for (size_t i = 0; i < groups_.size(); i++) {
    const int32 c = layers_[groups_[i].first].comp;
    if (c > first) { updater_->ForwardRange(feats, first, c - 1, NULL); first = c; }
    CU_SAFE_CALL(cudaStreamWaitEvent(Str(), groups_[i].done, 0));     // this group's new weights are in place
    const int32 last = i + 1 < groups_.size() ? layers_[groups_[i + 1].first].comp - 1 : L - 1;
    updater_->ForwardRange(feats, first, last, last == L - 1 ? labels : NULL);
    first = last + 1;
  }
  if (first < L) updater_->ForwardRange(feats, first, L - 1, labels);
  updater_->ComputeObjfAndDeriv(labels);
}

void NnetDataParallel::Prime(const CuMatrixBase<BaseFloat> &feats, const int32 *labels_dev) {
  NoPdl no_pdl(1);
  updater_->ForwardRange(feats, 0, nnet_->NumComponents() - 1, labels_dev);
  updater_->ComputeObjfAndDeriv(labels_dev);
  primed_ = true;
}

void NnetDataParallel::RotateEager(const CuMatrixBase<BaseFloat> &feats_next, const int32 *labels_next,
                                   int32 rows_global) {
  BackwardWithUpdates(rows_global);
  ForwardBehindUpdates(feats_next, labels_next);
  // the barrier error words come back with every step (checked by Failed() without a synchronisation)
  for (int ch = 0; ch < 2; ch++)
    CU_SAFE_CALL(cudaMemcpyAsync(error_pinned_ + ch, kcnn_p2p_error_word(base_, flag_off_, ch), sizeof(unsigned int),
                                 cudaMemcpyDeviceToHost, Str()));
}

void NnetDataParallel::Rotate(const CuMatrixBase<BaseFloat> &feats_next, const int32 *labels_next,
                              int32 rows_global) {
  if (!primed_) KALDI_ERR << "NnetDataParallel::Rotate: Prime() the pipeline with the first batch";
  last_replayed_ = false;
  cudaStream_t st = Str();
  static int graphs_on = -1;
  if (graphs_on < 0) {
    const char *e = getenv("KCNN_NNET_GRAPH");
    graphs_on = (e && e[0] == '0') ? 0 : 1;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (!graphs_on || st == 0 || st == cudaStreamLegacy || st == cudaStreamPerThread || cs != cudaStreamCaptureStatusNone) {
    RotateEager(feats_next, labels_next, rows_global);
    return;
  }
  uint64 key = Component::HashValue(feats_next.Data(), 23);
  key = Component::HashValue(feats_next.NumRows(), key);
  key = Component::HashValue(feats_next.Stride(), key);
  key = Component::HashValue(labels_next, key);
  key = Component::HashValue(rows_global, key);
  for (int32 c = 0; c < nnet_->NumComponents(); c++)
    key = Component::HashValue(nnet_->GetComponent(c).StepSignature(), key);
  key |= 1;
  const int32 L = nnet_->NumComponents();
  for (size_t i = 0; i < graphs_.size(); i++) {
    if (graphs_[i].key != key) continue;
    if (graphs_[i].exec == NULL) { RotateEager(feats_next, labels_next, rows_global); return; }
    CU_SAFE_CALL(cudaGraphLaunch(graphs_[i].exec, st));
    for (int32 c = 0; c < L; c++) {
      NonlinearComponent *nl = dynamic_cast<NonlinearComponent *>(&nnet_->GetComponent(c));
      if (nl && graphs_[i].count_delta[c] != 0.0) nl->AddToCount(graphs_[i].count_delta[c]);
    }
    last_replayed_ = true;
    return;
  }
  bool seen = false;
  for (size_t i = 0; i < seen_.size(); i++) seen = seen || seen_[i] == key;
  if (!seen) {                               // first rotation with these buffers: eager (everything gets sized)
    RotateEager(feats_next, labels_next, rows_global);
    if (seen_.size() >= 8) seen_.erase(seen_.begin());
    seen_.push_back(key);
    return;
  }
  std::vector<double> before(L, 0.0);
  for (int32 c = 0; c < L; c++) {
    const NonlinearComponent *nl = dynamic_cast<const NonlinearComponent *>(&nnet_->GetComponent(c));
    if (nl) before[c] = nl->Count();
  }
  cudaGraph_t g = NULL;
  bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess;
  if (ok) {
    try {
      RotateEager(feats_next, labels_next, rows_global);
    } catch (...) {
      ok = false;
    }
    if (cudaStreamEndCapture(st, &g) != cudaSuccess || g == NULL) ok = false;
  }
  cudaGraphExec_t exec = NULL;
  if (ok && cudaGraphInstantiate(&exec, g, 0) != cudaSuccess) { ok = false; exec = NULL; }
  if (ok) CuDevice::Instantiate().GraphRecorded();
  else CuDevice::Instantiate().CaptureAbandoned();
  if (g) cudaGraphDestroy(g);
  Recorded r;
  r.key = key;
  r.exec = exec;
  r.count_delta.assign(L, 0.0);
  if (graphs_.size() >= 4) {
    if (graphs_[0].exec) { cudaGraphExecDestroy(graphs_[0].exec); CuDevice::Instantiate().GraphDestroyed(); }
    graphs_.erase(graphs_.begin());
  }
  if (!ok) {
    cudaGetLastError();
    for (int32 c = 0; c < L; c++) {
      NonlinearComponent *nl = dynamic_cast<NonlinearComponent *>(&nnet_->GetComponent(c));
      if (nl) nl->AddToCount(before[c] - nl->Count());
    }
    graphs_.push_back(r);                    // exec == NULL: eager from now on
    RotateEager(feats_next, labels_next, rows_global);
    return;
  }
  for (int32 c = 0; c < L; c++) {
    const NonlinearComponent *nl = dynamic_cast<const NonlinearComponent *>(&nnet_->GetComponent(c));
    if (nl) r.count_delta[c] = nl->Count() - before[c];
  }
  graphs_.push_back(r);
  CU_SAFE_CALL(cudaGraphLaunch(exec, st));
  last_replayed_ = true;
}

void NnetDataParallel::Finish(int32 rows_global) {
  if (!primed_) return;
  BackwardWithUpdates(rows_global);
  for (size_t i = 0; i < groups_.size(); i++) CU_SAFE_CALL(cudaStreamWaitEvent(Str(), groups_[i].done, 0));
  primed_ = false;
}

bool NnetDataParallel::Failed(bool synchronise) {
  if (synchronise) {
    for (int ch = 0; ch < 2; ch++)
      CU_SAFE_CALL(cudaMemcpyAsync(error_pinned_ + ch, kcnn_p2p_error_word(base_, flag_off_, ch), sizeof(unsigned int),
                                   cudaMemcpyDeviceToHost, Str()));
    CU_SAFE_CALL(cudaStreamSynchronize(Str()));
  }
  return (error_pinned_[0] | error_pinned_[1]) != 0u;
}

void NnetDataParallel::GatherMomentum() {
  cudaStream_t st = Str();
  for (size_t i = 0; i < groups_.size(); i++) CU_SAFE_CALL(cudaStreamWaitEvent(st, groups_[i].done, 0));
  for (size_t i = 0; i < layers_.size(); i++) {
    const Layer &l = layers_[i];
    UpdatableComponent *u = static_cast<UpdatableComponent *>(&nnet_->GetComponent(l.comp));
    UpdatableComponent::StepTarget t;
    if (!u->GetStepTarget(1, &t)) continue;
    // this rank's slice of the bucket, as the update kernel cuts it (float4 units)
    const size_t n4 = l.len >> 2, w4 = l.weight_floats >> 2, per = (n4 + world_ - 1) / world_;
    const size_t lo = std::min((size_t)rank_ * per, w4), hi = std::min((size_t)(rank_ + 1) * per, w4);
    CU_SAFE_CALL(cudaMemsetAsync(base_ + l.off, 0, sizeof(float) * l.len, st));
    if (hi > lo)
      CU_SAFE_CALL(cudaMemcpyAsync(base_ + l.off + 4 * lo, t.prev + 4 * lo, sizeof(float) * 4 * (hi - lo),
                                   cudaMemcpyDeviceToDevice, st));
    if (kcnn_p2p_allreduce_f32(st, &peers_[0], rank_, world_, l.off, l.len, flag_off_, 0) != 0)
      KALDI_ERR << "kcnn_p2p_allreduce_f32 rejected its arguments";
    CU_SAFE_CALL(cudaMemcpyAsync(t.prev, base_ + l.off, sizeof(float) * l.weight_floats, cudaMemcpyDeviceToDevice, st));
  }
  CU_SAFE_CALL(cudaStreamSynchronize(st));
}

}  // namespace nnet2
}  // namespace kaldi
