// kaldi-cnn_b200/csrc/cnslmat/kcnn_lib.cu -- library state of libkaldicnn_b200.so.

#include <mutex>
#include "kcnn_common.cuh"

#include <cxxabi.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

namespace kcnn {
unsigned long long g_launch_count = 0;
cudaStream_t g_legacy_stream = 0;
bool g_profile_on = false;

namespace {
struct LaunchRec {
  const void *kernel;
  dim3 grid;
  std::string label;
  double flops, bytes;
  cudaEvent_t e0, e1;
  bool open;
};
std::vector<LaunchRec> g_recs;
std::string g_label;
double g_label_flops = 0.0, g_label_bytes = 0.0;
}  // namespace

void profile_before(const void *kernel, dim3 grid, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
  LaunchRec r;
  r.kernel = kernel; r.grid = grid; r.label = g_label; r.flops = g_label_flops; r.bytes = g_label_bytes;
  r.open = false;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) { cudaGetLastError(); return; }
  cudaEventRecord(r.e0, st);
  r.open = true;
  g_recs.push_back(r);
  g_label_flops = 0.0; g_label_bytes = 0.0;        // the work belongs to the FIRST launch under a label
}

void profile_after(cudaStream_t st) {
  if (g_recs.empty() || !g_recs.back().open) return;
  cudaEventRecord(g_recs.back().e1, st);
  g_recs.back().open = false;
}

// On by default (KCNN_PDL=0 turns it off for good; kcnn_set_pdl() switches it at run time -- the setting is
// read at launch, so a recorded graph keeps what it was recorded with).  With the 76-launch step of round 1
// it LOST 2.5 % (0.859 vs 0.837 ms: the early-scheduled dependents took SM slots from the tail of the running
// kernel); with the 33 fatter launches of the fused plan it gains 2.6 % (0.641 -> 0.624 ms, identical
// parameters): the prologue of the next GEMM (tensor-map fetch, barrier init, TMEM allocation) hides under
// the epilogue of the current one.  The data-parallel trainer switches it OFF for its launches: dependents
// parked on the SMs get in the way of the communication kernels and of the weight-gradient branch (2 GPUs:
// 0.823 ms without, 0.969 ms with).
static int g_pdl = -1;
static bool pdl_allowed() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
bool pdl_enabled() {
  if (g_pdl < 0) g_pdl = pdl_allowed() ? 1 : 0;
  return g_pdl == 1;
}

namespace {
// One side stream and one fork / join event pair per DEVICE, created on first use under a mutex.  A
// ForkJoin region belongs to the stream that opened it; two streams of one device forking at the same time
// would share the side stream (correct -- it is ordered by the events -- but serialised).
struct SideState { cudaStream_t stream; cudaEvent_t fork, join; bool tried; };
constexpr int kMaxDevices = 32;
SideState g_side[kMaxDevices] = {};
std::mutex g_side_mu;

SideState *side_state(cudaStream_t main) {
  static int enabled = -1;
  if (enabled < 0) {
    const char *e = getenv("KCNN_SIDE_STREAM");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lock(g_side_mu);
  SideState &g = g_side[dev];
  if (!g.tried) {
    // created on first use, never while the caller's stream is being captured (the warm-up
    // steps that precede any capture get here first)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(main, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return nullptr;
    }
    g.tried = true;
    if (cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&g.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&g.join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      g.stream = nullptr;
    }
  }
  return g.stream ? &g : nullptr;
}
}  // namespace

ForkJoin::ForkJoin(cudaStream_t main, bool want) : main_(main), side_(nullptr), join_ev_(nullptr) {
  if (!want) return;
  SideState *s = side_state(main);
  if (!s) return;
  if (cudaEventRecord(s->fork, main) != cudaSuccess || cudaStreamWaitEvent(s->stream, s->fork, 0) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  side_ = s->stream;
  join_ev_ = s->join;
}

void ForkJoin::join() {
  if (!side_) return;
  cudaEventRecord(join_ev_, side_);
  cudaStreamWaitEvent(main_, join_ev_, 0);
  side_ = nullptr;
}
}  // namespace kcnn

extern "C" {

void kcnn_set_stream(cudaStream_t stream) { kcnn::g_legacy_stream = stream; }
cudaStream_t kcnn_get_stream(void) { return kcnn::g_legacy_stream; }
unsigned long long kcnn_launch_count(void) { return kcnn::g_launch_count; }
void kcnn_reset_launch_count(void) { kcnn::g_launch_count = 0; }
/* ---- per-launch timing (bench.py) ---- */
void kcnn_profile_start(void) {
  for (size_t i = 0; i < kcnn::g_recs.size(); i++) { cudaEventDestroy(kcnn::g_recs[i].e0); cudaEventDestroy(kcnn::g_recs[i].e1); }
  kcnn::g_recs.clear();
  kcnn::g_label.clear();
  kcnn::g_profile_on = true;
}
int kcnn_set_pdl(int on) {
  const int before = kcnn::pdl_enabled() ? 1 : 0;
  kcnn::g_pdl = (on && kcnn::pdl_allowed()) ? 1 : 0;
  return before;
}
int kcnn_profile_active(void) { return kcnn::g_profile_on ? 1 : 0; }
int kcnn_profile_stop(void) {
  kcnn::g_profile_on = false;
  cudaDeviceSynchronize();
  return (int)kcnn::g_recs.size();
}
void kcnn_profile_label(const char *label, double flops, double bytes) {
  if (!kcnn::g_profile_on) return;
  kcnn::g_label = label ? label : "";
  kcnn::g_label_flops = flops;
  kcnn::g_label_bytes = bytes;
}
int kcnn_profile_get(int i, char *kernel, int kernel_len, char *label, int label_len, float *ms, double *flops,
                     double *bytes, unsigned int *grid3) {
  if (i < 0 || i >= (int)kcnn::g_recs.size()) return -1;
  const kcnn::LaunchRec &r = kcnn::g_recs[i];
  float t = 0.f;
  if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); t = -1.f; }
  if (ms) *ms = t;
  if (flops) *flops = r.flops;
  if (bytes) *bytes = r.bytes;
  if (grid3) { grid3[0] = r.grid.x; grid3[1] = r.grid.y; grid3[2] = r.grid.z; }
  if (label && label_len > 0) { strncpy(label, r.label.c_str(), label_len - 1); label[label_len - 1] = 0; }
  if (kernel && kernel_len > 0) {
    std::string name = "?";
    Dl_info info;
    if (dladdr(r.kernel, &info) && info.dli_sname) {
      int status = 0;
      char *dem = abi::__cxa_demangle(info.dli_sname, NULL, NULL, &status);
      name = (status == 0 && dem) ? dem : info.dli_sname;
      free(dem);
    }
    strncpy(kernel, name.c_str(), kernel_len - 1);
    kernel[kernel_len - 1] = 0;
  }
  return 0;
}

const char *kcnn_build_info(void) {
  return "kaldi-cnn_b200 sm_100a (compute_100a) nvcc " __DATE__;
}
int kcnn_abi_version(void) { return 1; }

}  // extern "C"
