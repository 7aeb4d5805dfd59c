// nnet2/nnet-nnet.cc -- shim (see nnet-nnet.h).
#include <stdlib.h>
#include <sstream>

#include <cuda_runtime_api.h>

#include "nnet2/nnet-nnet.h"
#include "cnsl-cu-kernels.h"

namespace kaldi {
namespace nnet2 {

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }

void Nnet::Destroy() {
  for (size_t i = 0; i < components_.size(); i++) delete components_[i];
  components_.clear();
}

void Nnet::Init(std::istream &is, bool skip_splice) {
  Destroy();
  std::string line;
  while (std::getline(is, line)) {
    size_t b = line.find_first_not_of(" \t\r\n");
    if (b == std::string::npos || line[b] == '#') continue;
    size_t e = line.find_last_not_of(" \t\r\n");
    line = line.substr(b, e - b + 1);
    if (skip_splice && line.compare(0, 15, "SpliceComponent") == 0) continue;
    Component *c = Component::NewFromString(line);
    c->SetIndex(components_.size());
    components_.push_back(c);
  }
  Check();
}

void Nnet::Check() const {
  for (size_t i = 0; i + 1 < components_.size(); i++) {
    int32 output_dim = components_[i]->OutputDim(), next_input_dim = components_[i + 1]->InputDim();
    if (output_dim != next_input_dim)
      KALDI_ERR << "Dimension mismatch between output of component " << i << " ("
                << components_[i]->Type() << ", " << output_dim << ") and input of the next ("
                << components_[i + 1]->Type() << ", " << next_input_dim << ")";
  }
}

void Nnet::Write(std::ostream &os, bool binary) const {
  WriteToken(os, binary, "<Nnet>");
  int32 num_components = components_.size();
  WriteToken(os, binary, "<NumComponents>");
  WriteBasicType(os, binary, num_components);
  WriteToken(os, binary, "<Components>");
  for (int32 c = 0; c < num_components; c++) {
    components_[c]->Write(os, binary);
    if (!binary) os << std::endl;
  }
  WriteToken(os, binary, "</Components>");
  WriteToken(os, binary, "</Nnet>");
}

void Nnet::Read(std::istream &is, bool binary) {
  Destroy();
  ExpectToken(is, binary, "<Nnet>");
  int32 num_components;
  ExpectToken(is, binary, "<NumComponents>");
  ReadBasicType(is, binary, &num_components);
  ExpectToken(is, binary, "<Components>");
  components_.resize(num_components, NULL);
  for (int32 c = 0; c < num_components; c++) {
    components_[c] = Component::ReadNew(is, binary);
    components_[c]->SetIndex(c);
  }
  ExpectToken(is, binary, "</Components>");
  ExpectToken(is, binary, "</Nnet>");
  Check();
}

int32 Nnet::NumUpdatableComponents() const {
  int32 ans = 0;
  for (size_t i = 0; i < components_.size(); i++)
    if (dynamic_cast<const UpdatableComponent *>(components_[i]) != NULL) ans++;
  return ans;
}

std::string Nnet::Info() const {
  std::ostringstream ostr;
  ostr << "num-components " << NumComponents() << std::endl;
  ostr << "num-updatable-components " << NumUpdatableComponents() << std::endl;
  ostr << "input-dim " << InputDim() << std::endl;
  ostr << "output-dim " << OutputDim() << std::endl;
  for (int32 i = 0; i < NumComponents(); i++)
    ostr << "component " << i << " : " << components_[i]->Info() << std::endl;
  return ostr.str();
}

// ------------------------------------------------------- NnetMinibatchUpdater --

struct NnetMinibatchUpdater::GraphState {
  struct Entry {
    cudaGraphExec_t exec;                 // NULL: recording was refused, stay eager for this key
    uint64 key;
    std::vector<double> count_delta;      // NonlinearComponent::count_ added by one step (host state)
  };
  std::vector<Entry> entries;             // a few (input buffer, labels) pairs: double-buffered callers
  std::vector<uint64> seen;               // keys that have run eagerly once (recording needs a warm step)
  uint64 config;                          // signature of everything but the buffers
  GraphState() : config(0) {}
};

NnetMinibatchUpdater::NnetMinibatchUpdater(Nnet *nnet)
    : graph_(new GraphState), last_replayed_(false), graphs_on_(true), fuse_(true),
      nnet_(nnet), num_rows_(0), fused_(NULL), deferred_join_(false), grad_stream_(NULL), step_labels_(NULL), base_(0), labels_(NULL), objf_dev_(NULL) {
  FusedInit();
  if (nnet_->NumComponents() > 1 && dynamic_cast<SpliceComponent *>(&nnet_->GetComponent(0)) != NULL) base_ = 1;
  objf_dev_ = static_cast<double *>(CuDevice::Instantiate().Malloc(sizeof(double)));
  CU_SAFE_CALL(cudaMemsetAsync(objf_dev_, 0, sizeof(double), Str()));
  const char *fe = getenv("KCNN_NNET_FUSE");
  if (fe && fe[0] == '0') fuse_ = false;
  SetInputPersists(true);    // forward_[c] is ours and untouched between Forward and Backward
}

NnetMinibatchUpdater::~NnetMinibatchUpdater() {
  DropGraph();
  delete graph_;
  FusedDestroy();
  SetInputPersists(false);
  CuDevice::Instantiate().Free(objf_dev_);
}

void NnetMinibatchUpdater::SetInputPersists(bool on) {
  for (int32 c = 0; c < nnet_->NumComponents(); c++) {
    UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(&nnet_->GetComponent(c));
    if (u) u->SetInputPersists(on);
  }
}

void NnetMinibatchUpdater::Forward(const CuMatrixBase<BaseFloat> &feats) {
  ForwardRange(feats, 0, nnet_->NumComponents() - 1);
}

int32 NnetMinibatchUpdater::FramesPerExample() const {
  if (base_ == 0) return 1;
  const std::vector<int32> ctx = nnet_->GetComponent(0).Context();
  return ctx.back() - ctx.front() + 1;
}

void NnetMinibatchUpdater::ForwardRange(const CuMatrixBase<BaseFloat> &feats, int32 first, int32 last,
                                        const int32 *labels_dev) {
  const int32 L = nnet_->NumComponents();
  KALDI_ASSERT(L > 0 && feats.NumCols() == nnet_->InputDim());
  KALDI_ASSERT(first >= 0 && last < L);
  const int32 span = FramesPerExample();
  if (feats.NumRows() % span != 0)
    KALDI_ERR << "the input has " << feats.NumRows() << " rows, not a multiple of the " << span
              << " frames per example the SpliceComponent needs";
  const int32 rows = feats.NumRows() / span;
  if (num_rows_ != rows || static_cast<int32>(forward_.size()) != L + 1) {
    num_rows_ = rows;
    forward_.clear();
    forward_.resize(L + 1);
    info_.clear();
    info_.push_back(ChunkInfo(nnet_->InputDim(), num_rows_, 0, span - 1));
    for (int32 c = 0; c < L; c++) {
      int32 dim = nnet_->GetComponent(c).OutputDim();
      if (c < base_) {          // the Splice output: one frame per example, at offset left-context
        const int32 left = -nnet_->GetComponent(0).Context().front();
        info_.push_back(ChunkInfo(dim, num_rows_, left, left));
      } else {
        info_.push_back(ChunkInfo(dim, num_rows_, 0, 0));
        forward_[c + 1].Resize(num_rows_, dim, kUndefined);
      }
    }
    derivs_.clear();
    derivs_.resize(L + 1);
  }
  // the input is used in place (a borrowed view), not copied
  if (first == 0) {
    forward_[0].Borrow(const_cast<BaseFloat *>(feats.Data()), feats.NumRows(), feats.NumCols(),
                       feats.Stride());
    if (base_ == 1) {
      const SpliceComponent &sp = static_cast<const SpliceComponent &>(nnet_->GetComponent(0));
      if (sp.IsContiguousWindow() && feats.Stride() == feats.NumCols()) {
        // [examples * span x dim] dense IS [examples x span * dim]: the Splice is a change of view
        forward_[1].Borrow(const_cast<BaseFloat *>(feats.Data()), num_rows_, span * feats.NumCols(),
                           span * feats.NumCols());
      } else {
        // gapped context / constant columns / pitched input: gather into a buffer of our own
        if (splice_out_.NumRows() != num_rows_ || splice_out_.NumCols() != sp.OutputDim())
          splice_out_.Resize(num_rows_, sp.OutputDim(), kUndefined);
        forward_[1].Borrow(splice_out_.Data(), num_rows_, sp.OutputDim(), splice_out_.Stride());
        sp.Propagate(info_[0], info_[1], forward_[0], static_cast<CuMatrixBase<BaseFloat> *>(&forward_[1]));
      }
    }
  }
  if (first < base_) first = base_;
  if (last < first) return;
  if (PlanFused()) {
    FusedForward(first, last, labels_dev != NULL ? labels_dev : step_labels_);
    return;
  }
  for (int32 c = first; c <= last; c++) {
    const Component &comp = nnet_->GetComponent(c);
    if (fuse_ && c + 1 <= last && !comp.BackpropNeedsOutput()) {
      const RectifiedLinearComponent *relu =
          dynamic_cast<const RectifiedLinearComponent *>(&nnet_->GetComponent(c + 1));
      if (relu != NULL && !relu->BackpropNeedsInput() &&
          comp.PropagateRelu(info_[c], info_[c + 2], forward_[c],
                             static_cast<CuMatrixBase<BaseFloat> *>(&forward_[c + 2]))) {
        c++;                       // forward_[c + 1] (the pre-activation) is not needed by anyone
        continue;
      }
    }
    comp.Propagate(info_[c], info_[c + 1], forward_[c],
                   static_cast<CuMatrixBase<BaseFloat> *>(&forward_[c + 1]));
  }
}

void NnetMinibatchUpdater::ComputeObjfAndDeriv(const int32 *labels_dev) {
  if (FusedObjf(labels_dev)) return;
  const CuMatrix<BaseFloat> &post = forward_.back();
  CuMatrix<BaseFloat> &d = derivs_.back();
  d.Resize(post.NumRows(), post.NumCols(), kUndefined);
  cudaF_xent_deriv(Str(), post.Data(), post.Dim(), labels_dev, d.Data(), d.Dim(), objf_dev_);
  CU_SAFE_CALL(cudaGetLastError());
}

void NnetMinibatchUpdater::Backward(int32 last, int32 first) {
  const int32 L = nnet_->NumComponents();
  if (last < 0) last = L - 1;
  KALDI_ASSERT(first >= 0 && last < L && !forward_.empty());
  if (first < base_) first = base_;       // nothing trains below the Splice front end (nnet2 stops there too)
  grad_stream_ = CuDevice::Instantiate().Stream();
  if (last < first) return;
  // KCNN_NNET_PDL_BWD=0: programmatic dependent launch in the forward pass only (see nnet-dp.cc: in the
  // data-parallel rotation the backward half is better off without it)
  static int pdl_bwd = -1;
  if (pdl_bwd < 0) {
    const char *e = getenv("KCNN_NNET_PDL_BWD");
    pdl_bwd = (e && e[0] == '0') ? 0 : 1;
  }
  struct Scope {
    int before; bool on;
    explicit Scope(bool off) : before(1), on(off) { if (on) before = kcnn_set_pdl(0); }
    ~Scope() { if (on) kcnn_set_pdl(before); }
  } scope(pdl_bwd == 0);
  if (FusedActive()) {
    FusedBackward(last, first);
    return;
  }
  // derivs_[c + 1] holds d objf / d output-of-component[c]; every buffer keeps its storage from
  // step to step (a same-size Resize inside Backprop is a no-op), so a recorded graph of the step
  // never points at memory the device cache could hand to someone else.
  for (int32 c = last; c >= first; c--) {
    Component &comp = nnet_->GetComponent(c);
    if (derivs_[c].NumRows() != num_rows_ || derivs_[c].NumCols() != comp.InputDim())
      derivs_[c].Resize(num_rows_, comp.InputDim(), kUndefined);
    comp.Backprop(info_[c], info_[c + 1], forward_[c], forward_[c + 1], derivs_[c + 1], &comp, &derivs_[c]);
  }
}

// ------------------------------------------------------------- graph step --

static bool GraphsEnabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_NNET_GRAPH");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

void NnetMinibatchUpdater::DropGraph() {
  for (size_t i = 0; i < graph_->entries.size(); i++)
    if (graph_->entries[i].exec) { cudaGraphExecDestroy(graph_->entries[i].exec); CuDevice::Instantiate().GraphDestroyed(); }
  graph_->entries.clear();
  graph_->seen.clear();
}

// Everything a recorded step bakes in besides the input / label buffers.
uint64 NnetMinibatchUpdater::ConfigKey() const {
  uint64 h = Component::HashValue(Str(), 17);
  h = Component::HashValue(CuDevice::Instantiate().MathMode(), h);
  h = Component::HashValue(fuse_, h);
  for (int32 c = 0; c < nnet_->NumComponents(); c++)
    h = Component::HashValue(nnet_->GetComponent(c).StepSignature(), h);
  return h | 1;
}

uint64 NnetMinibatchUpdater::StepKey(const CuMatrixBase<BaseFloat> &feats, const int32 *labels_dev) const {
  uint64 h = Component::HashValue(feats.Data(), 19);
  h = Component::HashValue(feats.NumRows(), h);
  h = Component::HashValue(feats.NumCols(), h);
  h = Component::HashValue(feats.Stride(), h);
  h = Component::HashValue(labels_dev, h);
  return h | 1;      // never 0
}

void NnetMinibatchUpdater::EagerStep(const CuMatrixBase<BaseFloat> &feats, const int32 *labels_dev) {
  step_labels_ = labels_dev;          // lets the fused plan finish softmax + objective inside the forward pass
  try {
    Forward(feats);
  } catch (...) {
    step_labels_ = NULL;
    throw;
  }
  step_labels_ = NULL;
  ComputeObjfAndDeriv(labels_dev);
  Backward();
}

void NnetMinibatchUpdater::TrainStep(const CuMatrixBase<BaseFloat> &feats, const int32 *labels_dev) {
  last_replayed_ = false;
  cudaStream_t st = Str();
  const bool capturable = graphs_on_ && GraphsEnabled() && st != 0 && st != cudaStreamLegacy && st != cudaStreamPerThread;
  if (!capturable) { EagerStep(feats, labels_dev); return; }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs != cudaStreamCaptureStatusNone) { EagerStep(feats, labels_dev); return; }   // caller's own capture
  const uint64 config = ConfigKey();
  if (config != graph_->config) {       // learning rate, parameter storage, stream, ... changed
    DropGraph();
    graph_->config = config;
  }
  const uint64 key = StepKey(feats, labels_dev);
  const int32 L = nnet_->NumComponents();
  for (size_t i = 0; i < graph_->entries.size(); i++) {
    GraphState::Entry &e = graph_->entries[i];
    if (e.key != key) continue;
    if (e.exec == NULL) { EagerStep(feats, labels_dev); return; }
    CU_SAFE_CALL(cudaGraphLaunch(e.exec, st));
    for (int32 c = 0; c < L; c++) {
      NonlinearComponent *nl = dynamic_cast<NonlinearComponent *>(&nnet_->GetComponent(c));
      if (nl && e.count_delta[c] != 0.0) nl->AddToCount(e.count_delta[c]);
    }
    last_replayed_ = true;
    return;
  }
  bool seen = false;
  for (size_t i = 0; i < graph_->seen.size(); i++) seen = seen || graph_->seen[i] == key;
  if (!seen) {                           // first step with these buffers: eager (sizes every scratch buffer)
    EagerStep(feats, labels_dev);
    if (ConfigKey() != config) {         // the step itself allocated state (e.g. a dropout seed)
      DropGraph();
      graph_->config = ConfigKey();
    }
    if (graph_->seen.size() >= 8) graph_->seen.erase(graph_->seen.begin());
    graph_->seen.push_back(key);
    return;
  }
  // Second step: record it.  Host-side effects of a step (the frame counts of the
  // nonlinearity statistics) happen once here and are re-applied after every replay.
  std::vector<double> before(L, 0.0);
  for (int32 c = 0; c < L; c++) {
    const NonlinearComponent *nl = dynamic_cast<const NonlinearComponent *>(&nnet_->GetComponent(c));
    if (nl) before[c] = nl->Count();
  }
  cudaGraph_t g = NULL;
  bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess;
  if (ok) {
    try {
      EagerStep(feats, labels_dev);
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
  if (graph_->entries.size() >= 4) {                  // oldest out
    if (graph_->entries[0].exec) { cudaGraphExecDestroy(graph_->entries[0].exec); CuDevice::Instantiate().GraphDestroyed(); }
    graph_->entries.erase(graph_->entries.begin());
  }
  GraphState::Entry e;
  e.exec = exec;
  e.key = key;
  e.count_delta.assign(L, 0.0);
  if (!ok) {
    cudaGetLastError();                               // clear the sticky capture error
    for (int32 c = 0; c < L; c++) {                   // the aborted recording did not run
      NonlinearComponent *nl = dynamic_cast<NonlinearComponent *>(&nnet_->GetComponent(c));
      if (nl) nl->AddToCount(before[c] - nl->Count());
    }
    graph_->entries.push_back(e);                     // exec == NULL: eager from now on
    EagerStep(feats, labels_dev);
    return;
  }
  for (int32 c = 0; c < L; c++) {
    const NonlinearComponent *nl = dynamic_cast<const NonlinearComponent *>(&nnet_->GetComponent(c));
    if (nl) e.count_delta[c] = nl->Count() - before[c];
  }
  graph_->entries.push_back(e);
  CU_SAFE_CALL(cudaGraphLaunch(exec, st));            // the recording itself executed nothing
  last_replayed_ = true;
}

void NnetMinibatchUpdater::SetDeferredUpdate(bool on) {
  for (int32 c = 0; c < nnet_->NumComponents(); c++) {
    UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(&nnet_->GetComponent(c));
    if (u) u->SetDeferredUpdate(on);
  }
}

void NnetMinibatchUpdater::ApplyGradients(int32 total_rows) {
  for (int32 c = 0; c < nnet_->NumComponents(); c++) {
    UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(&nnet_->GetComponent(c));
    if (u && u->DeferredUpdate()) u->ApplyGradient(total_rows);
  }
}

size_t NnetMinibatchUpdater::GradientFloats() const {
  size_t n = 0;
  for (int32 c = 0; c < nnet_->NumComponents(); c++) {
    const UpdatableComponent *u = dynamic_cast<const UpdatableComponent *>(&nnet_->GetComponent(c));
    if (u) n += (u->GradientFloats() + 63) / 64 * 64;      // 256-byte aligned buckets
  }
  return n;
}

void NnetMinibatchUpdater::SetGradientArena(float *base) {
  const int32 L = nnet_->NumComponents();
  bucket_off_.assign(L, 0);
  bucket_len_.assign(L, 0);
  size_t off = 0;
  for (int32 c = L - 1; c >= 0; c--) {     // top layer first: the order backward produces them
    UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(&nnet_->GetComponent(c));
    if (!u) continue;
    size_t len = (u->GradientFloats() + 63) / 64 * 64;
    bucket_off_[c] = off;
    bucket_len_[c] = len;
    u->SetGradientStorage(base ? base + off : NULL);
    off += len;
  }
}

void NnetMinibatchUpdater::GradientBucket(int32 c, size_t *offset, size_t *length) const {
  KALDI_ASSERT(c >= 0 && c < static_cast<int32>(bucket_off_.size()));
  *offset = bucket_off_[c];
  *length = bucket_len_[c];
}

double NnetMinibatchUpdater::GetObjfAndReset() {
  double v = 0;
  CU_SAFE_CALL(cudaMemcpyAsync(&v, objf_dev_, sizeof(double), cudaMemcpyDeviceToHost, Str()));
  CU_SAFE_CALL(cudaMemsetAsync(objf_dev_, 0, sizeof(double), Str()));
  CU_SAFE_CALL(cudaStreamSynchronize(Str()));
  return v;
}

}  // namespace nnet2
}  // namespace kaldi
