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

#define KCNN_LAUNCH(kernel, grid, block, smem, stream, ...)        \
  do {                                                             \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
    ::kcnn::count_launch();                                        \
  } while (0)

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
