// kaldi-cnn_b200/csrc/cnslmat/kcnn_common.cuh
//
// Shared device/host helpers of the sm_100a kernel library.

#ifndef KCNN_COMMON_CUH_
#define KCNN_COMMON_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "cnsl-cu-kernels.h"

namespace kcnn {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Launch bookkeeping: every kernel launch of the library goes through
// KCNN_LAUNCH so bench.py can report how many of OUR kernels ran.
extern unsigned long long g_launch_count;
extern cudaStream_t g_legacy_stream;

inline void count_launch() { ++g_launch_count; }

// Programmatic dependent launch (PDL): every kernel of the library starts with
// pdl_prologue() -- "let the NEXT kernel of the stream be scheduled as soon as all my CTAs
// have started; wait until the PREVIOUS kernel has completed and flushed" -- and is launched
// with the programmatic-stream-serialization attribute, so the launch latency and the
// prologue of kernel i+1 (barrier init, TMEM allocation, tensor-map fetch) hide under the
// tail of kernel i instead of adding ~90 serial gaps to a training step.  Without the
// attribute (the default; KCNN_PDL=1 turns it on) both instructions are no-ops.
__device__ __forceinline__ void pdl_prologue() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// The two halves separately, for kernels whose set-up (barrier init, TMEM allocation) may
// run before the previous kernel has finished: trigger first, wait right before the first
// access to memory another kernel may have written.
__device__ __forceinline__ void pdl_trigger() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

bool pdl_enabled();      // kcnn_lib.cu: KCNN_PDL=1 turns the launch attribute on

// Fork / join of independent launches inside one library call (e.g. the input-gradient and
// the weight-gradient GEMM of a convolution: small grids that leave most SMs idle when run
// back to back).  The work between fork() and join() on side() runs concurrently with the
// caller's stream and is ordered before everything the caller enqueues after join(); both
// are event waits, so the pair is captured into CUDA graphs as two parallel branches.
// KCNN_SIDE_STREAM=0 disables it (active() is then false and callers stay on one stream).
class ForkJoin {
 public:
  explicit ForkJoin(cudaStream_t main, bool want);
  bool active() const { return side_ != nullptr; }
  cudaStream_t side() const { return side_; }
  void join();                       // idempotent
  ~ForkJoin() { join(); }
 private:
  cudaStream_t main_, side_;
  cudaEvent_t join_ev_;
};

template <class... KArgs, class... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          unsigned cluster_x, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    n++;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    n++;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  count_launch();
}

#define KCNN_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::kcnn::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), 1u, __VA_ARGS__)

inline unsigned int ceil_div_u(long long a, long long b) {
  return (unsigned int)((a + b - 1) / b);
}

// Division by a runtime constant without the integer divider: the kernels
// decompose flat indices into (row, channel, w, h) many times per element.
struct FastDiv {
  uint32_t d, mul, shr;
  FastDiv() : d(1), mul(0), shr(0) {}
  explicit FastDiv(uint32_t divisor) : d(divisor) {
    if (d == 1) { mul = 0; shr = 0; return; }
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                 // ceil(log2 d)
    uint64_t m = ((1ull << 32) * ((1ull << l) - d)) / d + 1;
    mul = (uint32_t)m;
    shr = l - 1;
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    if (d == 1) return n;
    uint32_t t = __umulhi(n, mul);
    return (t + ((n - t) >> 1)) >> shr;
#else
    return n / d;
#endif
  }
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t &q, uint32_t &r) const {
    q = div(n);
    r = n - q * d;
  }
};

__device__ __forceinline__ bool aligned16(const void *p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

inline bool host_aligned16(const void *p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

}  // namespace kcnn

#endif
