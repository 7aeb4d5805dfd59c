// cudamatrix/cu-device.h -- shim: the CuDevice singleton.
//
// Kaldi's CuDevice selects the GPU, owns a caching allocator and accumulates a
// per-function host-side profile (AccuProfile, printed by PrintProfile; used
// after every launch in the reference, e.g. cnslmat/conv2D.cc:110).  This shim
// keeps those entry points and adds the one thing the B200 build needs: the
// CUDA stream every member launches on (default: the legacy default stream, as
// in Kaldi; bench.py / the C API point it at the caller's stream so the whole
// step is stream-ordered and CUDA-graph capturable).
#ifndef KALDI_CUDAMATRIX_CU_DEVICE_H_
#define KALDI_CUDAMATRIX_CU_DEVICE_H_

#include <cuda_runtime_api.h>
#include <map>
#include <string>
#include <vector>

#include "base/kaldi-common.h"

namespace kaldi {

class CuDevice {
 public:
  static CuDevice &Instantiate() { static CuDevice d; return d; }

  /// "yes" / "no" / "optional" as in Kaldi.  "no" leaves the device disabled,
  /// and every CuMatrix operation then fails loudly: this build has no CPU path.
  void SelectGpuId(std::string use_gpu);
  bool Enabled() const { return enabled_; }

  void AccuProfile(const std::string &key, double time) {
    if (profile_) profile_map_[key] += time;
  }
  void PrintProfile();
  void EnableProfile(bool on) { profile_ = on; }

  cudaStream_t Stream() const { return stream_; }
  void SetStream(cudaStream_t s);

  /// Arithmetic of the GEMM-shaped members (KCNN_MATH_FP32_SIMT = 0: FP32 FMA, the
  /// accuracy class of the reference's SGEMM; KCNN_MATH_TF32_TC = 1: tcgen05 TF32).
  /// Default FP32; the environment variable KCNN_MATH=tf32 or SetMathMode() selects TF32.
  int MathMode() const { return math_mode_; }
  void SetMathMode(int m) { math_mode_ = m; }

  /// Seed of SetRandn() (host generator; parameters are initialised off the hot path).
  void SetRandSeed(unsigned long long s) { rand_seed_ = s; }
  unsigned long long NextRandSeed() { return rand_seed_++; }

  /// Caching allocator (device memory is reused across Resize calls so the
  /// per-minibatch temporaries of the host code cost no cudaMalloc in steady state).
  void *Malloc(size_t bytes);
  void Free(void *ptr);
  void ReleaseCache();
  size_t BytesAllocated() const { return bytes_allocated_; }
  /// Recorded CUDA graphs bake device addresses in.  A block that is handed out or returned WHILE the
  /// compute stream is being captured is therefore pinned: when it is freed it does not go back to the
  /// cache (where the next Malloc of that size class -- a held-out batch, a second network, a staging
  /// slot -- would receive memory a later replay still writes to) until every recorded graph has been
  /// destroyed.  NnetMinibatchUpdater / NnetDataParallel report their graph executables here; their own
  /// buffers are sized before the capture, so in steady state nothing is pinned.
  void GraphRecorded() { live_graphs_++; }
  void GraphDestroyed();
  void CaptureAbandoned() { if (live_graphs_ == 0) { live_graphs_ = 1; GraphDestroyed(); } }
  size_t BytesPinnedByGraphs() const;

  /// Rows are pitched to a multiple of 16 bytes so 128-bit and TMA paths apply.
  static int32 PitchInElements(int32 cols, size_t elem_size) {
    int32 q = 16 / (int32)elem_size;
    return ((cols + q - 1) / q) * q;
  }

  void RequireEnabled(const char *what) const {
    if (!enabled_) KALDI_ERR << what << ": no CUDA device is selected and this is the GPU build "
                             << "(no CPU fallback). Call CuDevice::Instantiate().SelectGpuId(\"yes\").";
  }

 private:
  CuDevice();
  ~CuDevice() {}
  bool enabled_, profile_;
  int math_mode_;
  unsigned long long rand_seed_;
  cudaStream_t stream_;
  int live_graphs_;
  std::map<void *, size_t> graph_blocks_;                  // live blocks handed out during a capture
  std::vector<std::pair<void *, size_t> > graph_pinned_;   // freed blocks a graph may still address
  bool Capturing() const;
  std::map<std::string, double> profile_map_;
  std::map<size_t, std::vector<void *> > free_;
  std::map<void *, size_t> live_;
  size_t bytes_allocated_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(CuDevice);
};

}  // namespace kaldi
#endif
