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

// Per-launch timing for bench.py's roofline table (kcnn_profile_start / _stop / _get in kcnn_lib.cu):
// while recording, every launch of the library is bracketed by a pair of CUDA events on its own stream
// and tagged with the label / algorithmic work the caller announced last (kcnn_profile_label).  Off the
// hot path: one predictable branch per launch when not recording; never records inside a stream capture.
extern bool g_profile_on;
void profile_before(const void *kernel, dim3 grid, cudaStream_t st);
void profile_after(cudaStream_t st);

// Programmatic dependent launch (PDL): every kernel of the library starts with
// pdl_prologue() -- "let the NEXT kernel of the stream be scheduled as soon as all my CTAs
// have started; wait until the PREVIOUS kernel has completed and flushed" -- and is launched
// with the programmatic-stream-serialization attribute, so the launch latency and the
// prologue of kernel i+1 (barrier init, TMEM allocation, tensor-map fetch) hide under the
// tail of kernel i instead of adding serial gaps to a training step.  Without the attribute
// (KCNN_PDL=0) both instructions are no-ops.
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

bool pdl_enabled();      // kcnn_lib.cu: the launch attribute; on unless KCNN_PDL=0

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

// cluster: thread-block cluster shape ((1,1,1) = none).  x = 2 pairs two CTAs of a TPC
// (cta_group::2 tiles); z = S makes the S K-splits of one output tile a cluster, reduced
// through distributed shared memory (gemm_tma.cuh).
template <class... KArgs, class... Args>
inline void launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  dim3 cluster, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster.x * cluster.y * cluster.z > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster.x; at[n].val.clusterDim.y = cluster.y; at[n].val.clusterDim.z = cluster.z;
    n++;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    n++;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  if (g_profile_on) profile_before(reinterpret_cast<const void *>(kernel), grid, stream);
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (g_profile_on) profile_after(stream);
  count_launch();
}

template <class... KArgs, class... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          unsigned cluster_x, Args &&...args) {
  launch_kernel_cluster(kernel, grid, block, smem, stream, dim3(cluster_x, 1, 1), static_cast<Args &&>(args)...);
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
  __host__ __device__ FastDiv() : d(1), mul(0), shr(0) {}
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

// Counter-based uniform of the dropout mask (splitmix64 finaliser): a pure function of
// (seed, element index), so the forward kernel, a fused GEMM epilogue and a replayed CUDA
// graph all draw the same mask from the same device-resident seed.
__host__ __device__ __forceinline__ uint32_t mix32(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 32);
}
// scale of element (row, col) of a [rows x cols] matrix: `low` with probability dp, else `high`
__device__ __forceinline__ float dropout_scale_at(unsigned long long seed, unsigned long long row, int cols, int col,
                                                  float dp, float low, float high) {
  const unsigned long long base = seed * 0x100000001B3ull + row * (unsigned long long)cols;
  const float r = (mix32(base + (unsigned long long)col) >> 8) * (1.0f / 16777216.0f);
  return (r - dp > 0.0f) ? high : low;
}

}  // namespace kcnn

#endif
