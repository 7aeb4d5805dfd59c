// cudamatrix/cu-device.cc -- shim: CuDevice.
#include <algorithm>
#include "cudamatrix/cu-device.h"
#include "cudamatrix/cu-common.h"
#include "cnsl-cu-kernels.h"

namespace kaldi {

CuDevice::CuDevice()
    : enabled_(false), profile_(false), math_mode_(KCNN_MATH_FP32_SIMT), rand_seed_(5489),
      stream_(0), live_graphs_(0), bytes_allocated_(0) {
  const char *m = getenv("KCNN_MATH");
  if (m && (std::string(m) == "tf32" || std::string(m) == "1")) math_mode_ = KCNN_MATH_TF32_TC;
}

void CuDevice::SelectGpuId(std::string use_gpu) {
  if (use_gpu == "no") { enabled_ = false; return; }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    if (use_gpu == "optional") { enabled_ = false; return; }
    KALDI_ERR << "No CUDA device found (" << cudaGetErrorString(e) << ") and use_gpu=" << use_gpu;
  }
  int dev = 0;
  CU_SAFE_CALL(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_SAFE_CALL(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    KALDI_WARN << "This library is built for sm_100a; device " << dev << " is sm_" << prop.major
               << prop.minor;
  enabled_ = true;
}

void CuDevice::SetStream(cudaStream_t s) {
  stream_ = s;
  kcnn_set_stream(s);
}

static size_t RoundUp(size_t bytes) {
  if (bytes < 512) return 512;
  // next multiple of 1/8 of the enclosing power of two: <= 12.5 % slack, few classes
  size_t p = 512;
  while (p < bytes) p <<= 1;
  size_t step = p >> 3;
  return ((bytes + step - 1) / step) * step;
}

void *CuDevice::Malloc(size_t bytes) {
  RequireEnabled("CuDevice::Malloc");
  if (bytes == 0) return NULL;
  size_t sz = RoundUp(bytes);
  std::vector<void *> &bucket = free_[sz];
  void *p = NULL;
  if (!bucket.empty()) {
    p = bucket.back();
    bucket.pop_back();
  } else {
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {
      cudaGetLastError();
      ReleaseCache();
      CU_SAFE_CALL(cudaMalloc(&p, sz));
    }
    bytes_allocated_ += sz;
  }
  live_[p] = sz;
  if (Capturing()) graph_blocks_[p] = sz;
  return p;
}

bool CuDevice::Capturing() const {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream_, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs != cudaStreamCaptureStatusNone;
}

void CuDevice::GraphDestroyed() {
  if (live_graphs_ > 0) live_graphs_--;
  if (live_graphs_ > 0) return;
  for (size_t i = 0; i < graph_pinned_.size(); i++) free_[graph_pinned_[i].second].push_back(graph_pinned_[i].first);
  graph_pinned_.clear();
  graph_blocks_.clear();          // blocks still live are ordinary blocks again: no graph is left to address them
}

size_t CuDevice::BytesPinnedByGraphs() const {
  size_t b = 0;
  for (size_t i = 0; i < graph_pinned_.size(); i++) b += graph_pinned_[i].second;
  return b;
}

void CuDevice::Free(void *ptr) {
  if (!ptr) return;
  std::map<void *, size_t>::iterator it = live_.find(ptr);
  if (it == live_.end()) { cudaFree(ptr); return; }
  std::map<void *, size_t>::iterator gb = graph_blocks_.find(ptr);
  if (gb != graph_blocks_.end() || Capturing()) {
    // a recorded (or recording) step addresses this block: keep it out of circulation (cu-device.h)
    graph_pinned_.push_back(std::make_pair(ptr, it->second));
    if (gb != graph_blocks_.end()) graph_blocks_.erase(gb);
    live_.erase(it);
    return;
  }
  // Stream-ordered reuse: all work is issued on one stream (Stream()), so a block
  // handed out again is only touched by later work on that same stream.
  free_[it->second].push_back(ptr);
  live_.erase(it);
}

void CuDevice::ReleaseCache() {
  cudaDeviceSynchronize();
  for (std::map<size_t, std::vector<void *> >::iterator it = free_.begin(); it != free_.end(); ++it) {
    for (size_t i = 0; i < it->second.size(); i++) {
      cudaFree(it->second[i]);
      bytes_allocated_ -= it->first;
    }
    it->second.clear();
  }
}

void CuDevice::PrintProfile() {
  if (profile_map_.empty()) return;
  std::vector<std::pair<double, std::string> > v;
  double total = 0;
  for (std::map<std::string, double>::iterator it = profile_map_.begin(); it != profile_map_.end(); ++it) {
    v.push_back(std::make_pair(it->second, it->first));
    total += it->second;
  }
  std::sort(v.begin(), v.end());
  std::ostringstream os;
  os << "-----\n[cudevice profile]\n";
  for (size_t i = 0; i < v.size(); i++) os << v[i].second << "\t" << v[i].first << "s\n";
  os << "Total GPU time:\t" << total << "s (may involve some double-counting)\n-----";
  KALDI_LOG << os.str();
}

}  // namespace kaldi
