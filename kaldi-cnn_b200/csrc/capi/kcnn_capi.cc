// capi/kcnn_capi.cc -- the C ABI of include/kcnn_capi.h: thin forwarders onto the Kaldi-
// style C++ interface (CuMatrixBase members, nnet2::Component virtuals).  Kaldi
// assertions / KALDI_ERR raised underneath are caught and reported through
// kcnn_last_error() so a foreign host is not aborted.

#include <string.h>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <sstream>
#include <string>

#include "kcnn_capi.h"
#include "cnsl-cu-kernels.h"
#include "nnet0/nnet-component-nnet0.h"
#include "nnet2/nnet-nnet.h"
#include "nnet2/nnet-dp.h"

using namespace kaldi;
using namespace kaldi::nnet2;
using cnsl::nnet0::ConvolutionComponent;
using cnsl::nnet0::FullyConnectedComponent;
using cnsl::nnet0::MaxpoolComponent;

namespace {

thread_local std::string g_last_error;

struct ThrowGuard {      // KALDI_ASSERT throws instead of aborting while inside the C API
  bool prev;
  ThrowGuard() : prev(kaldi::g_assert_throws) { kaldi::g_assert_throws = true; }
  ~ThrowGuard() { kaldi::g_assert_throws = prev; }
};

#define KCNN_TRY ThrowGuard guard_; try {
#define KCNN_CATCH(ret)                                              \
  } catch (const std::exception &e) { g_last_error = e.what(); return ret; } \
    catch (...) { g_last_error = "unknown C++ exception"; return ret; }

typedef CuSubMatrix<BaseFloat> View;

inline Component *C(kcnn_component *c) { return reinterpret_cast<Component *>(c); }
inline const Component *C(const kcnn_component *c) { return reinterpret_cast<const Component *>(c); }

struct NnetHandle {
  Nnet nnet;
  NnetMinibatchUpdater *updater;
  CuMatrix<BaseFloat> host_feats;      // staging for kcnn_nnet_train_minibatch_host
  int32 *host_labels_dev;
  int32 host_labels_rows;
  // kcnn_nnet_train_minibatch_host_async: two slots of (pinned host, device) buffers, a copy
  // stream, and events -- batch k+1 is staged and copied while batch k computes.
  struct Pipe {
    float *pin_feats[2]; int32 *pin_labels[2]; double *pin_objf;
    CuMatrix<BaseFloat> dev_feats[2]; int32 *dev_labels[2];
    cudaStream_t copy_stream; cudaEvent_t copied[2], done[2];
    int rows, dim; unsigned long long step;
    Pipe() : pin_objf(NULL), copy_stream(NULL), rows(0), dim(0), step(0) {
      for (int i = 0; i < 2; i++) { pin_feats[i] = NULL; pin_labels[i] = NULL; dev_labels[i] = NULL; copied[i] = NULL; done[i] = NULL; }
    }
    void Release() {
      if (copy_stream) cudaStreamSynchronize(copy_stream);
      for (int i = 0; i < 2; i++) {
        if (pin_feats[i]) cudaFreeHost(pin_feats[i]);
        if (pin_labels[i]) cudaFreeHost(pin_labels[i]);
        if (dev_labels[i]) CuDevice::Instantiate().Free(dev_labels[i]);
        if (copied[i]) cudaEventDestroy(copied[i]);
        if (done[i]) cudaEventDestroy(done[i]);
        pin_feats[i] = NULL; pin_labels[i] = NULL; dev_labels[i] = NULL; copied[i] = NULL; done[i] = NULL;
        dev_feats[i].Resize(0, 0);
      }
      if (pin_objf) cudaFreeHost(pin_objf);
      if (copy_stream) cudaStreamDestroy(copy_stream);
      pin_objf = NULL; copy_stream = NULL; rows = 0; dim = 0;
    }
    void Ensure(int r, int d, int labels) {      // r feature rows (frames), d columns, `labels` examples
      if (r == rows && d == dim && copy_stream != NULL) return;
      Release();
      CU_SAFE_CALL(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
      CU_SAFE_CALL(cudaMallocHost(reinterpret_cast<void **>(&pin_objf), 2 * sizeof(double)));
      pin_objf[0] = pin_objf[1] = 0.0;
      for (int i = 0; i < 2; i++) {
        CU_SAFE_CALL(cudaMallocHost(reinterpret_cast<void **>(&pin_feats[i]), sizeof(float) * (size_t)r * d));
        CU_SAFE_CALL(cudaMallocHost(reinterpret_cast<void **>(&pin_labels[i]), sizeof(int32) * (size_t)labels));
        dev_feats[i].Resize(r, d, kUndefined);
        dev_labels[i] = static_cast<int32 *>(CuDevice::Instantiate().Malloc(sizeof(int32) * (size_t)labels));
        CU_SAFE_CALL(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        CU_SAFE_CALL(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
      }
      rows = r; dim = d;
    }
  } pipe;
  NnetHandle() : updater(NULL), host_labels_dev(NULL), host_labels_rows(0) {}
  ~NnetHandle() {
    if (CuDevice::Instantiate().Enabled()) cudaStreamSynchronize(CuDevice::Instantiate().Stream());
    pipe.Release();
    delete updater;
    if (host_labels_dev) CuDevice::Instantiate().Free(host_labels_dev);
  }
  NnetMinibatchUpdater &U() {
    if (!updater) updater = new NnetMinibatchUpdater(&nnet);
    return *updater;
  }
};
inline NnetHandle *N(kcnn_nnet *n) { return reinterpret_cast<NnetHandle *>(n); }
inline const NnetHandle *N(const kcnn_nnet *n) { return reinterpret_cast<const NnetHandle *>(n); }

int WriteToBuffer(const std::string &s, char **data, size_t *len) {
  *data = static_cast<char *>(malloc(s.size() ? s.size() : 1));
  if (!*data) { g_last_error = "out of memory"; return -1; }
  memcpy(*data, s.data(), s.size());
  *len = s.size();
  return 0;
}

// A CuMatrix<float> that views a foreign buffer of the FINAL shape (the members resize
// their CuMatrix<Real>* outputs only when the shape differs, so this one is left alone).
struct Borrowed {
  CuMatrix<BaseFloat> m;
  Borrowed(float *p, int r, int c, int s) { m.Borrow(p, r, c, s); }
};

// Host -> device copy of a dense [rows x cols] host matrix into a pitched device matrix: one linear copy
// when the device pitch equals the row length (a 2-D copy of ten thousand 160-byte rows is a descriptor
// per row), else the pitched form.
void CopyRowsToDevice(CuMatrix<BaseFloat> &dst, const float *src, int rows, int cols, cudaStream_t stream) {
  if (dst.Stride() == cols) {
    CU_SAFE_CALL(cudaMemcpyAsync(dst.Data(), src, sizeof(float) * (size_t)rows * cols, cudaMemcpyHostToDevice, stream));
  } else {
    CU_SAFE_CALL(cudaMemcpy2DAsync(dst.Data(), sizeof(float) * dst.Stride(), src, sizeof(float) * cols,
                                   sizeof(float) * cols, rows, cudaMemcpyHostToDevice, stream));
  }
}

// KCNN_HOST_TIMING=1: where the host thread spends its time inside the pipelined entry points, printed
// every 256 calls (diagnostics; off by default, two clock reads per phase when on).
struct HostPhaseTimer {
  enum { kPhases = 8 };
  static bool On() { static const bool on = getenv("KCNN_HOST_TIMING") != NULL; return on; }
  struct Acc { double us[kPhases]; const char *name[kPhases]; long calls; };
  static Acc &Get() { static Acc a = {}; return a; }
  const char *what;
  std::chrono::steady_clock::time_point last;
  explicit HostPhaseTimer(const char *w) : what(w) { if (On()) last = std::chrono::steady_clock::now(); }
  void Mark(int phase, const char *name) {
    if (!On()) return;
    std::chrono::steady_clock::time_point now = std::chrono::steady_clock::now();
    Acc &a = Get();
    a.us[phase] += std::chrono::duration<double, std::micro>(now - last).count();
    a.name[phase] = name;
    last = now;
  }
  ~HostPhaseTimer() {
    if (!On()) return;
    Acc &a = Get();
    if (++a.calls % 256 == 0) {
      fprintf(stderr, "%s, host us per call:", what);
      for (int i = 0; i < kPhases; i++)
        if (a.name[i]) { fprintf(stderr, "  %s %.1f", a.name[i], a.us[i] / 256.0); a.us[i] = 0.0; }
      fprintf(stderr, "\n");
    }
  }
};

void CheckShape(const CuMatrix<BaseFloat> &m, float *p, const char *what) {
  if (m.Data() != p) KALDI_ERR << what << ": the output buffer does not have the shape the member "
                               << "produces (it would have been reallocated)";
}

}  // namespace

extern "C" {

int kcnn_select_gpu(const char *use_gpu) {
  KCNN_TRY
  CuDevice::Instantiate().SelectGpuId(use_gpu ? use_gpu : "yes");
  return CuDevice::Instantiate().Enabled() ? 0 : 1;
  KCNN_CATCH(-1)
}
void kcnn_set_compute_stream(void *s) { CuDevice::Instantiate().SetStream(static_cast<cudaStream_t>(s)); }
void kcnn_set_math_mode(int mode) { CuDevice::Instantiate().SetMathMode(mode); }
int kcnn_get_math_mode(void) { return CuDevice::Instantiate().MathMode(); }
void kcnn_set_rand_seed(unsigned long long seed) { CuDevice::Instantiate().SetRandSeed(seed); }
size_t kcnn_device_bytes_pinned_by_graphs(void) { return CuDevice::Instantiate().BytesPinnedByGraphs(); }
const char *kcnn_last_error(void) { return g_last_error.c_str(); }
void kcnn_enable_profile(int on) { CuDevice::Instantiate().EnableProfile(on != 0); }
void kcnn_print_profile(void) { CuDevice::Instantiate().PrintProfile(); }
size_t kcnn_device_bytes_allocated(void) { return CuDevice::Instantiate().BytesAllocated(); }
void kcnn_free(void *p) { free(p); }

// ---- L1 ------------------------------------------------------------------------------

int kcnn_mat_conv2d(const float *a, int ar, int ac, int as, const float *k, int kr, int kc, int ks,
                    int H, int W, int C_, int KH, int KW, int G, float *out, int orows, int ocols,
                    int os, int concat) {
  KCNN_TRY
  View A(a, ar, ac, as), K(k, kr, kc, ks), O(out, orows, ocols, os);
  A.Conv2D(K, H, W, C_, KH, KW, G, &O, concat != 0);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_mat_add_mat_rep_vec(float *a, int ar, int ac, int as, const float *vec, int vec_dim, int rep) {
  KCNN_TRY
  View A(a, ar, ac, as);
  CuSubVector<BaseFloat> v(vec, vec_dim);
  A.AddMatRepVec(v, rep);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_mat_flip_mat(const float *a, int ar, int ac, int as, int KH, int KW, int C_, int G,
                      float *flip, int fr, int fc, int fs) {
  KCNN_TRY
  View A(a, ar, ac, as);
  Borrowed F(flip, fr, fc, fs);
  A.FlipMat(KH, KW, C_, G, &F.m);
  CheckShape(F.m, flip, "FlipMat");
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_mat_padding_zero(const float *a, int ar, int ac, int as, int H, int W, int C_, int KH, int KW,
                          float *pad, int pr, int pc, int ps) {
  KCNN_TRY
  View A(a, ar, ac, as);
  Borrowed P(pad, pr, pc, ps);
  A.PaddingZero(H, W, C_, KH, KW, &P.m);
  CheckShape(P.m, pad, "PaddingZero");
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_mat_tp_block(const float *a, int ar, int ac, int as, int C_, int bs, float *out, int orows,
                      int ocols, int os) {
  KCNN_TRY
  View A(a, ar, ac, as);
  Borrowed O(out, orows, ocols, os);
  A.TpBlock(C_, bs, &O.m);
  CheckShape(O.m, out, "TpBlock");
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_mat_tp_inside_block(const float *a, int ar, int ac, int as, int G, int bs, float *out,
                             int orows, int ocols, int os) {
  KCNN_TRY
  View A(a, ar, ac, as);
  Borrowed O(out, orows, ocols, os);
  A.TpInsideBlock(G, bs, &O.m);
  CheckShape(O.m, out, "TpInsideBlock");
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_mat_mod_permute_row(const float *a, int ar, int ac, int as, int C_, int bs, float *out,
                             int orows, int ocols, int os) {
  KCNN_TRY
  View A(a, ar, ac, as);
  Borrowed O(out, orows, ocols, os);
  A.ModPermuteRow(C_, bs, &O.m);
  CheckShape(O.m, out, "ModPermuteRow");
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_mat_maxpool_prop(const float *a, int ar, int ac, int as, int H, int W, int ph, int pw, int pc,
                          int overlap, int overlap2D, float *out, int orows, int ocols, int os) {
  KCNN_TRY
  View A(a, ar, ac, as), O(out, orows, ocols, os);
  A.Maxpool_prop(H, W, ph, pw, pc, overlap != 0, overlap2D != 0, &O);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_mat_maxpool_backprop(const float *a, int ar, int ac, int as, const float *ov, int ovr,
                              int ovc, int ovs, const float *od, int odr, int odc, int ods,
                              float *id, int idr, int idc, int ids, int H, int W, int ph, int pw,
                              int pc, int overlap, int overlap2D) {
  KCNN_TRY
  View A(a, ar, ac, as), OV(ov, ovr, ovc, ovs), OD(od, odr, odc, ods);
  Borrowed I(id, idr, idc, ids);
  A.Maxpool_backprop(OV, OD, &I.m, H, W, ph, pw, pc, overlap != 0, overlap2D != 0);
  CheckShape(I.m, id, "Maxpool_backprop");
  return 0;
  KCNN_CATCH(-1)
}

// ---- L2 ------------------------------------------------------------------------------

kcnn_component *kcnn_component_new_from_string(const char *line) {
  KCNN_TRY
  return reinterpret_cast<kcnn_component *>(Component::NewFromString(line));
  KCNN_CATCH(NULL)
}

kcnn_component *kcnn_component_read(const char *data, size_t len, int binary) {
  KCNN_TRY
  std::istringstream is(std::string(data, len));
  return reinterpret_cast<kcnn_component *>(Component::ReadNew(is, binary != 0));
  KCNN_CATCH(NULL)
}

int kcnn_component_write(const kcnn_component *c, int binary, char **data, size_t *len) {
  KCNN_TRY
  std::ostringstream os;
  os.precision(7);
  C(c)->Write(os, binary != 0);
  return WriteToBuffer(os.str(), data, len);
  KCNN_CATCH(-1)
}

kcnn_component *kcnn_component_copy(const kcnn_component *c) {
  KCNN_TRY
  return reinterpret_cast<kcnn_component *>(C(c)->Copy());
  KCNN_CATCH(NULL)
}

void kcnn_component_delete(kcnn_component *c) { delete C(c); }

const char *kcnn_component_type(const kcnn_component *c) {
  static thread_local std::string t;
  t = C(c)->Type();
  return t.c_str();
}

int kcnn_component_info(const kcnn_component *c, char *buf, size_t buf_len) {
  KCNN_TRY
  std::string s = C(c)->Info();
  if (buf && buf_len) {
    size_t n = s.size() < buf_len - 1 ? s.size() : buf_len - 1;
    memcpy(buf, s.data(), n);
    buf[n] = '\0';
  }
  return (int)s.size();
  KCNN_CATCH(-1)
}

int kcnn_component_input_dim(const kcnn_component *c) { return C(c)->InputDim(); }
int kcnn_component_output_dim(const kcnn_component *c) { return C(c)->OutputDim(); }
int kcnn_component_backprop_needs_input(const kcnn_component *c) { return C(c)->BackpropNeedsInput(); }
int kcnn_component_backprop_needs_output(const kcnn_component *c) { return C(c)->BackpropNeedsOutput(); }

int kcnn_component_propagate(const kcnn_component *c, int num_chunks, const float *in, int ir, int ic,
                             int is, float *out, int orows, int ocols, int os) {
  KCNN_TRY
  ChunkInfo in_info(ic, num_chunks, 0, ir / (num_chunks > 0 ? num_chunks : 1) - 1),
            out_info(ocols, num_chunks, 0, orows / (num_chunks > 0 ? num_chunks : 1) - 1);
  View I(in, ir, ic, is), O(out, orows, ocols, os);
  C(c)->Propagate(in_info, out_info, I, static_cast<CuMatrixBase<BaseFloat> *>(&O));
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_propagate_chunks(const kcnn_component *c, int num_chunks, int in_first_offset, int in_last_offset,
                                    int out_first_offset, int out_last_offset, const float *in, int ir, int ic,
                                    int is, float *out, int orows, int ocols, int os) {
  KCNN_TRY
  ChunkInfo in_info(ic, num_chunks, in_first_offset, in_last_offset),
            out_info(ocols, num_chunks, out_first_offset, out_last_offset);
  View I(in, ir, ic, is), O(out, orows, ocols, os);
  C(c)->Propagate(in_info, out_info, I, static_cast<CuMatrixBase<BaseFloat> *>(&O));
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_backprop_chunks(const kcnn_component *c, int num_chunks, int in_first_offset, int in_last_offset,
                                   int out_first_offset, int out_last_offset, const float *od, int od_rows, int ods,
                                   float *id, int ids) {
  KCNN_TRY
  const Component *comp = C(c);
  ChunkInfo in_info(comp->InputDim(), num_chunks, in_first_offset, in_last_offset),
            out_info(comp->OutputDim(), num_chunks, out_first_offset, out_last_offset);
  View none(NULL, 0, 0, 0), OD(od, od_rows, comp->OutputDim(), ods);
  Borrowed ID(id, in_info.NumRows(), comp->InputDim(), ids);
  comp->Backprop(in_info, out_info, none, none, OD, NULL, &ID.m);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_backprop(const kcnn_component *c, int num_chunks, const float *iv, int ivs,
                            const float *ov, int ovs, const float *od, int od_rows, int ods,
                            kcnn_component *to_update, float *id, int ids) {
  KCNN_TRY
  const Component *comp = C(c);
  const int in_dim = comp->InputDim(), out_dim = comp->OutputDim();
  int per = od_rows / (num_chunks > 0 ? num_chunks : 1);
  ChunkInfo in_info(in_dim, num_chunks, 0, per - 1), out_info(out_dim, num_chunks, 0, per - 1);
  View IV(iv, iv ? od_rows : 0, iv ? in_dim : 0, iv ? ivs : 0),
       OV(ov, ov ? od_rows : 0, ov ? out_dim : 0, ov ? ovs : 0), OD(od, od_rows, out_dim, ods);
  Borrowed ID(id, od_rows, in_dim, ids);
  comp->Backprop(in_info, out_info, IV, OV, OD, C(to_update), &ID.m);
  CheckShape(ID.m, id, "Backprop in_deriv");
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_params(kcnn_component *c, int which, float **data, int *rows, int *cols, int *stride) {
  KCNN_TRY
  const CuMatrix<BaseFloat> *m = NULL;
  const CuVector<BaseFloat> *v = NULL;
  if (ConvolutionComponent *cc = dynamic_cast<ConvolutionComponent *>(C(c))) {
    if (which == 0) m = &cc->LinearParams(); else if (which == 1) v = &cc->BiasParams(); else m = &cc->PrevGrad();
  } else if (FullyConnectedComponent *fc = dynamic_cast<FullyConnectedComponent *>(C(c))) {
    if (which == 0) m = &fc->LinearParams(); else if (which == 1) v = &fc->BiasParams(); else m = &fc->PrevGrad();
  } else if (AffineComponent *ac = dynamic_cast<AffineComponent *>(C(c))) {
    if (which == 0) m = &ac->LinearParams(); else if (which == 1) v = &ac->BiasParams();
    else KALDI_ERR << "AffineComponent has no prev_grad_";
  } else {
    KALDI_ERR << C(c)->Type() << " has no parameters";
  }
  if (m) { *data = const_cast<float *>(m->Data()); *rows = m->NumRows(); *cols = m->NumCols(); *stride = m->Stride(); }
  else { *data = const_cast<float *>(v->Data()); *rows = 1; *cols = v->Dim(); *stride = v->Dim(); }
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_set_learning_rate(kcnn_component *c, float lr) {
  KCNN_TRY
  UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(C(c));
  if (!u) KALDI_ERR << "not an UpdatableComponent";
  u->SetLearningRate(lr);
  return 0;
  KCNN_CATCH(-1)
}

float kcnn_component_learning_rate(const kcnn_component *c) {
  const UpdatableComponent *u = dynamic_cast<const UpdatableComponent *>(C(c));
  return u ? u->LearningRate() : 0.0f;
}

int kcnn_component_set_weight_decay_momentum(kcnn_component *c, float wd, float mom) {
  KCNN_TRY
  if (ConvolutionComponent *cc = dynamic_cast<ConvolutionComponent *>(C(c))) { cc->SetWeightDecay(wd); cc->SetMomentum(mom); }
  else if (FullyConnectedComponent *fc = dynamic_cast<FullyConnectedComponent *>(C(c))) { fc->SetWeightDecay(wd); fc->SetMomentum(mom); }
  else KALDI_ERR << C(c)->Type() << " has no weight decay / momentum";
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_get_weight_decay_momentum(const kcnn_component *c, float *wd, float *mom) {
  KCNN_TRY
  if (const ConvolutionComponent *cc = dynamic_cast<const ConvolutionComponent *>(C(c))) { *wd = cc->WeightDecay(); *mom = cc->Momentum(); }
  else if (const FullyConnectedComponent *fc = dynamic_cast<const FullyConnectedComponent *>(C(c))) { *wd = fc->WeightDecay(); *mom = fc->Momentum(); }
  else KALDI_ERR << C(c)->Type() << " has no weight decay / momentum";
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_set_index_routing(kcnn_component *c, int on) {
  KCNN_TRY
  MaxpoolComponent *m = dynamic_cast<MaxpoolComponent *>(C(c));
  if (!m) KALDI_ERR << "not a MaxpoolComponent";
  m->SetIndexRouting(on != 0);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_set_deferred_update(kcnn_component *c, int deferred) {
  KCNN_TRY
  UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(C(c));
  if (!u) KALDI_ERR << "not an UpdatableComponent";
  u->SetDeferredUpdate(deferred != 0);
  return 0;
  KCNN_CATCH(-1)
}

size_t kcnn_component_gradient_floats(const kcnn_component *c) {
  const UpdatableComponent *u = dynamic_cast<const UpdatableComponent *>(C(c));
  return u ? u->GradientFloats() : 0;
}

int kcnn_component_set_gradient_storage(kcnn_component *c, float *base) {
  KCNN_TRY
  UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(C(c));
  if (!u) KALDI_ERR << "not an UpdatableComponent";
  u->SetGradientStorage(base);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_gradient(kcnn_component *c, int which, float **data, int *rows, int *cols, int *stride) {
  KCNN_TRY
  UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(C(c));
  if (!u) KALDI_ERR << "not an UpdatableComponent";
  std::vector<UpdatableComponent::GradBuffer> g = u->GradientBuffers();
  if (which < 0 || which >= (int)g.size()) KALDI_ERR << "no such gradient buffer";
  *data = g[which].data; *rows = g[which].rows; *cols = g[which].cols; *stride = g[which].stride;
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_component_apply_gradient(kcnn_component *c, int total_rows) {
  KCNN_TRY
  UpdatableComponent *u = dynamic_cast<UpdatableComponent *>(C(c));
  if (!u) KALDI_ERR << "not an UpdatableComponent";
  u->ApplyGradient(total_rows);
  return 0;
  KCNN_CATCH(-1)
}

// ---- nnet ----------------------------------------------------------------------------

kcnn_nnet *kcnn_nnet_new_from_config(const char *config_text, int skip_splice) {
  KCNN_TRY
  NnetHandle *h = new NnetHandle();
  try {
    std::istringstream is(config_text);
    h->nnet.Init(is, skip_splice != 0);
  } catch (...) { delete h; throw; }
  return reinterpret_cast<kcnn_nnet *>(h);
  KCNN_CATCH(NULL)
}

kcnn_nnet *kcnn_nnet_read(const char *data, size_t len, int binary) {
  KCNN_TRY
  NnetHandle *h = new NnetHandle();
  try {
    std::istringstream is(std::string(data, len));
    h->nnet.Read(is, binary != 0);
  } catch (...) { delete h; throw; }
  return reinterpret_cast<kcnn_nnet *>(h);
  KCNN_CATCH(NULL)
}

int kcnn_nnet_write(const kcnn_nnet *n, int binary, char **data, size_t *len) {
  KCNN_TRY
  std::ostringstream os;
  os.precision(7);
  N(n)->nnet.Write(os, binary != 0);
  return WriteToBuffer(os.str(), data, len);
  KCNN_CATCH(-1)
}

void kcnn_nnet_delete(kcnn_nnet *n) { delete N(n); }
int kcnn_nnet_num_components(const kcnn_nnet *n) { return N(n)->nnet.NumComponents(); }
kcnn_component *kcnn_nnet_component(kcnn_nnet *n, int i) {
  if (i < 0 || i >= N(n)->nnet.NumComponents()) return NULL;
  return reinterpret_cast<kcnn_component *>(&N(n)->nnet.GetComponent(i));
}
int kcnn_nnet_input_dim(const kcnn_nnet *n) { return N(n)->nnet.InputDim(); }
int kcnn_nnet_output_dim(const kcnn_nnet *n) { return N(n)->nnet.OutputDim(); }

int kcnn_nnet_forward(kcnn_nnet *n, const float *feats, int rows, int stride) {
  KCNN_TRY
  View F(feats, rows, N(n)->nnet.InputDim(), stride);
  N(n)->U().Forward(F);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_forward_range(kcnn_nnet *n, const float *feats, int rows, int stride, int first, int last) {
  KCNN_TRY
  View F(feats, rows, N(n)->nnet.InputDim(), stride);
  if (first < 0 || last >= N(n)->nnet.NumComponents()) KALDI_ERR << "bad component range";
  N(n)->U().ForwardRange(F, first, last);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_objf_and_deriv(kcnn_nnet *n, const int *labels) {
  KCNN_TRY
  N(n)->U().ComputeObjfAndDeriv(labels);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_backward(kcnn_nnet *n, int last, int first) {
  KCNN_TRY
  N(n)->U().Backward(last, first);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_activation(kcnn_nnet *n, int index, const float **data, int *rows, int *cols, int *stride) {
  KCNN_TRY
  if (index < 0 || index > N(n)->nnet.NumComponents()) KALDI_ERR << "bad activation index";
  const CuMatrix<BaseFloat> &m = N(n)->U().Activation(index);
  *data = m.Data(); *rows = m.NumRows(); *cols = m.NumCols(); *stride = m.Stride();
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_input_deriv(kcnn_nnet *n, const float **data, int *rows, int *cols, int *stride) {
  KCNN_TRY
  const CuMatrix<BaseFloat> &m = N(n)->U().InputDeriv();
  *data = m.Data(); *rows = m.NumRows(); *cols = m.NumCols(); *stride = m.Stride();
  return 0;
  KCNN_CATCH(-1)
}

double kcnn_nnet_objf_and_reset(kcnn_nnet *n) {
  KCNN_TRY
  return N(n)->U().GetObjfAndReset();
  KCNN_CATCH(0.0)
}

int kcnn_nnet_set_deferred_update(kcnn_nnet *n, int deferred) {
  KCNN_TRY
  N(n)->U().SetDeferredUpdate(deferred != 0);
  return 0;
  KCNN_CATCH(-1)
}

size_t kcnn_nnet_gradient_floats(const kcnn_nnet *n) {
  return const_cast<NnetHandle *>(N(n))->U().GradientFloats();
}

int kcnn_nnet_set_gradient_arena(kcnn_nnet *n, float *base) {
  KCNN_TRY
  N(n)->U().SetGradientArena(base);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_gradient_bucket(const kcnn_nnet *n, int component, size_t *offset, size_t *length) {
  KCNN_TRY
  const_cast<NnetHandle *>(N(n))->U().GradientBucket(component, offset, length);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_apply_gradients(kcnn_nnet *n, int total_rows) {
  KCNN_TRY
  N(n)->U().ApplyGradients(total_rows);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_frames_per_example(kcnn_nnet *n) {
  KCNN_TRY
  return N(n)->U().FramesPerExample();
  KCNN_CATCH(-1)
}

int kcnn_nnet_train_minibatch_host(kcnn_nnet *n, const float *feats_host, const int *labels_host,
                                   int rows, double *objf) {
  KCNN_TRY
  NnetHandle *h = N(n);
  const int dim = h->nnet.InputDim();
  const int frames = rows * h->U().FramesPerExample();
  cudaStream_t st = CuDevice::Instantiate().Stream();
  if (h->host_feats.NumRows() != frames || h->host_feats.NumCols() != dim)
    h->host_feats.Resize(frames, dim, kUndefined);
  if (h->host_labels_rows != rows) {
    if (h->host_labels_dev) CuDevice::Instantiate().Free(h->host_labels_dev);
    h->host_labels_dev = static_cast<int32 *>(CuDevice::Instantiate().Malloc(sizeof(int32) * rows));
    h->host_labels_rows = rows;
  }
  CopyRowsToDevice(h->host_feats, feats_host, frames, dim, st);
  CU_SAFE_CALL(cudaMemcpyAsync(h->host_labels_dev, labels_host, sizeof(int32) * rows,
                               cudaMemcpyHostToDevice, st));
  h->U().TrainStep(h->host_feats, h->host_labels_dev);
  double v = h->U().GetObjfAndReset();
  if (objf) *objf = v;
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_train_minibatch_host_async(kcnn_nnet *n, const float *feats_host, const int *labels_host,
                                         int rows) {
  KCNN_TRY
  NnetHandle *h = N(n);
  const int dim = h->nnet.InputDim();
  cudaStream_t st = CuDevice::Instantiate().Stream();
  NnetHandle::Pipe &p = h->pipe;
  const int frames = rows * h->U().FramesPerExample();
  p.Ensure(frames, dim, rows);
  const int s = (int)(p.step & 1ull);
  HostPhaseTimer tm("kcnn_nnet_train_minibatch_host_async");
  // the pinned slot is free once ITS previous copy (two calls ago) has run; then the caller's
  // buffers are staged and belong to the caller again as soon as this call returns
  CU_SAFE_CALL(cudaEventSynchronize(p.copied[s]));
  tm.Mark(0, "wait for the staging slot");
  memcpy(p.pin_feats[s], feats_host, sizeof(float) * (size_t)frames * dim);
  memcpy(p.pin_labels[s], labels_host, sizeof(int32) * (size_t)rows);
  tm.Mark(1, "stage into pinned memory");
  // the device slot is free once the step that read it (two calls ago) has finished
  CU_SAFE_CALL(cudaStreamWaitEvent(p.copy_stream, p.done[s], 0));
  CopyRowsToDevice(p.dev_feats[s], p.pin_feats[s], frames, dim, p.copy_stream);
  CU_SAFE_CALL(cudaMemcpyAsync(p.dev_labels[s], p.pin_labels[s], sizeof(int32) * rows, cudaMemcpyHostToDevice,
                               p.copy_stream));
  CU_SAFE_CALL(cudaEventRecord(p.copied[s], p.copy_stream));
  CU_SAFE_CALL(cudaStreamWaitEvent(st, p.copied[s], 0));
  tm.Mark(2, "enqueue the copies");
  h->U().TrainStep(p.dev_feats[s], p.dev_labels[s]);
  tm.Mark(3, "TrainStep (graph launch)");
  // the running objective comes back to the host after every step, without a synchronisation
  CU_SAFE_CALL(cudaMemcpyAsync(p.pin_objf + s, h->U().ObjfDevice(), sizeof(double), cudaMemcpyDeviceToHost, st));
  CU_SAFE_CALL(cudaEventRecord(p.done[s], st));
  tm.Mark(4, "objective read-back + event");
  p.step++;
  return 0;
  KCNN_CATCH(-1)
}

double kcnn_nnet_running_objf(kcnn_nnet *n) {
  NnetHandle::Pipe &p = N(n)->pipe;
  if (p.copy_stream == NULL || p.step == 0) return 0.0;
  // newest step whose objective has already landed in host memory (no waiting)
  for (unsigned long long back = 1; back <= 2 && back <= p.step; back++) {
    const int s = (int)((p.step - back) & 1ull);
    if (cudaEventQuery(p.done[s]) == cudaSuccess) return p.pin_objf[s];
  }
  cudaGetLastError();
  return 0.0;
}

int kcnn_nnet_train_step(kcnn_nnet *n, const float *feats, int rows, int stride, const int *labels) {
  KCNN_TRY
  NnetHandle *h = N(n);
  View F(feats, rows, h->nnet.InputDim(), stride);
  h->U().TrainStep(F, labels);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_set_fusion(kcnn_nnet *n, int on) {
  KCNN_TRY
  N(n)->U().SetFusion(on != 0);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_last_step_replayed(const kcnn_nnet *n) {
  const NnetHandle *h = N(n);
  return (h->updater != NULL && h->updater->LastStepReplayed()) ? 1 : 0;
}

int kcnn_nnet_set_graphs(kcnn_nnet *n, int on) {
  KCNN_TRY
  N(n)->U().SetGraphs(on != 0);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_fused_active(const kcnn_nnet *n) {
  const NnetHandle *h = N(n);
  return (h->updater != NULL && h->updater->FusedActive()) ? 1 : 0;
}

}  // extern "C"

// ---- data-parallel trainer -------------------------------------------------------------------

namespace {
// Host-buffer pipeline of the data-parallel step.  Three slots: while rotation j runs -- backward of
// batch j-1 (which still reads that batch's input for the first layer's weight gradient) and forward
// of batch j -- batch j+1 is being staged and copied.
struct DpHandle {
  NnetHandle *net;
  NnetDataParallel *dp;
  enum { kSlots = 3 };
  float *pin_feats[kSlots]; int32 *pin_labels[kSlots]; double *pin_objf;
  CuMatrix<BaseFloat> dev_feats[kSlots]; int32 *dev_labels[kSlots];
  cudaStream_t copy_stream; cudaEvent_t copied[kSlots], done[kSlots];
  int frames, dim, labels; unsigned long long call;
  DpHandle() : net(NULL), dp(NULL), pin_objf(NULL), copy_stream(NULL), frames(0), dim(0), labels(0), call(0) {
    for (int i = 0; i < kSlots; i++) { pin_feats[i] = NULL; pin_labels[i] = NULL; dev_labels[i] = NULL; copied[i] = NULL; done[i] = NULL; }
  }
  void Release() {
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    for (int i = 0; i < kSlots; i++) {
      if (pin_feats[i]) cudaFreeHost(pin_feats[i]);
      if (pin_labels[i]) cudaFreeHost(pin_labels[i]);
      if (dev_labels[i]) CuDevice::Instantiate().Free(dev_labels[i]);
      if (copied[i]) cudaEventDestroy(copied[i]);
      if (done[i]) cudaEventDestroy(done[i]);
      pin_feats[i] = NULL; pin_labels[i] = NULL; dev_labels[i] = NULL; copied[i] = NULL; done[i] = NULL;
      dev_feats[i].Resize(0, 0);
    }
    if (pin_objf) cudaFreeHost(pin_objf);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    pin_objf = NULL; copy_stream = NULL; frames = 0; dim = 0; labels = 0;
  }
  void Ensure(int f, int d, int l) {
    if (f == frames && d == dim && l == labels && copy_stream != NULL) return;
    if (call != 0) KALDI_ERR << "the minibatch shape must not change while a batch is in the pipeline";
    Release();
    CU_SAFE_CALL(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    CU_SAFE_CALL(cudaMallocHost(reinterpret_cast<void **>(&pin_objf), sizeof(double)));
    pin_objf[0] = 0.0;
    for (int i = 0; i < kSlots; i++) {
      CU_SAFE_CALL(cudaMallocHost(reinterpret_cast<void **>(&pin_feats[i]), sizeof(float) * (size_t)f * d));
      CU_SAFE_CALL(cudaMallocHost(reinterpret_cast<void **>(&pin_labels[i]), sizeof(int32) * (size_t)l));
      dev_feats[i].Resize(f, d, kUndefined);
      dev_labels[i] = static_cast<int32 *>(CuDevice::Instantiate().Malloc(sizeof(int32) * (size_t)l));
      CU_SAFE_CALL(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
      CU_SAFE_CALL(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    frames = f; dim = d; labels = l;
  }
  ~DpHandle() {
    if (CuDevice::Instantiate().Enabled()) cudaStreamSynchronize(CuDevice::Instantiate().Stream());
    delete dp;
    Release();
  }
};
inline DpHandle *D(kcnn_nnet_dp *d) { return reinterpret_cast<DpHandle *>(d); }
}  // namespace

extern "C" {

size_t kcnn_nnet_dp_arena_floats(kcnn_nnet *n) {
  KCNN_TRY
  return NnetDataParallel::ArenaFloats(&N(n)->U());
  KCNN_CATCH(0)
}

kcnn_nnet_dp *kcnn_nnet_dp_create(kcnn_nnet *n, int rank, int world, float *local_base,
                                  const unsigned long long *peer_bases, unsigned long long multicast_base) {
  KCNN_TRY
  DpHandle *h = new DpHandle();
  try {
    h->net = N(n);
    h->dp = new NnetDataParallel(&h->net->nnet, &h->net->U(), rank, world, local_base, peer_bases, multicast_base);
  } catch (...) { delete h; throw; }
  return reinterpret_cast<kcnn_nnet_dp *>(h);
  KCNN_CATCH(NULL)
}

void kcnn_nnet_dp_delete(kcnn_nnet_dp *dp) { delete D(dp); }

int kcnn_nnet_dp_prime(kcnn_nnet_dp *dp, const float *feats, int rows, int stride, const int *labels) {
  KCNN_TRY
  View F(feats, rows, D(dp)->net->nnet.InputDim(), stride);
  D(dp)->dp->Prime(F, labels);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_dp_rotate(kcnn_nnet_dp *dp, const float *feats_next, int rows, int stride, const int *labels_next,
                        int rows_global) {
  KCNN_TRY
  View F(feats_next, rows, D(dp)->net->nnet.InputDim(), stride);
  D(dp)->dp->Rotate(F, labels_next, rows_global);
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_dp_finish(kcnn_nnet_dp *dp, int rows_global) {
  KCNN_TRY
  D(dp)->dp->Finish(rows_global);
  D(dp)->call = 0;
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_dp_train_minibatch_host_async(kcnn_nnet_dp *dp, const float *feats_host, const int *labels_host,
                                            int rows_local, int rows_global) {
  KCNN_TRY
  DpHandle *h = D(dp);
  const int dim = h->net->nnet.InputDim();
  const int frames = rows_local * h->net->U().FramesPerExample();
  cudaStream_t st = CuDevice::Instantiate().Stream();
  h->Ensure(frames, dim, rows_local);
  const int s = (int)(h->call % DpHandle::kSlots);
  CU_SAFE_CALL(cudaEventSynchronize(h->copied[s]));                 // the pinned slot's previous copy has run
  memcpy(h->pin_feats[s], feats_host, sizeof(float) * (size_t)frames * dim);
  memcpy(h->pin_labels[s], labels_host, sizeof(int32) * (size_t)rows_local);
  // the device slot was last read by the backward pass of the rotation two calls ago
  if (h->call >= 2) CU_SAFE_CALL(cudaStreamWaitEvent(h->copy_stream, h->done[(h->call - 2) % DpHandle::kSlots], 0));
  CopyRowsToDevice(h->dev_feats[s], h->pin_feats[s], frames, dim, h->copy_stream);
  CU_SAFE_CALL(cudaMemcpyAsync(h->dev_labels[s], h->pin_labels[s], sizeof(int32) * rows_local, cudaMemcpyHostToDevice,
                               h->copy_stream));
  CU_SAFE_CALL(cudaEventRecord(h->copied[s], h->copy_stream));
  CU_SAFE_CALL(cudaStreamWaitEvent(st, h->copied[s], 0));
  if (h->call == 0) h->dp->Prime(h->dev_feats[s], h->dev_labels[s]);
  else h->dp->Rotate(h->dev_feats[s], h->dev_labels[s], rows_global);
  CU_SAFE_CALL(cudaMemcpyAsync(h->pin_objf, h->net->U().ObjfDevice(), sizeof(double), cudaMemcpyDeviceToHost, st));
  CU_SAFE_CALL(cudaEventRecord(h->done[s], st));
  h->call++;
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_dp_failed(kcnn_nnet_dp *dp, int synchronise) {
  KCNN_TRY
  return D(dp)->dp->Failed(synchronise != 0) ? 1 : 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_dp_gather_momentum(kcnn_nnet_dp *dp) {
  KCNN_TRY
  D(dp)->dp->GatherMomentum();
  return 0;
  KCNN_CATCH(-1)
}

int kcnn_nnet_dp_last_rotate_replayed(const kcnn_nnet_dp *dp) {
  return reinterpret_cast<const DpHandle *>(dp)->dp->LastRotateReplayed() ? 1 : 0;
}

}  // extern "C"
