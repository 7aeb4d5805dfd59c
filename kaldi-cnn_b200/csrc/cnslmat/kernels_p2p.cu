// kaldi-cnn_b200/csrc/cnslmat/kernels_p2p.cu
//
// Gradient all-reduce over NVLink / NVSwitch peer memory, written for the data-parallel step
// (SURVEY 8e): it replaces the reference's file-based nnet-am-average
// (egs/steps/nnet0/train_conv_dropout.sh:323-341) and the NCCL all-reduce of the first
// version of this path.
//
// Why not NCCL here: on this step a 147 MB NCCL all-reduce runs on 32 channel CTAs that each
// hold most of an SM's shared memory, so the concurrent GEMMs (one ~200 KB CTA per SM, grids of
// 128 / 148 CTAs) lose SMs and fall into a second wave: measured at 2 GPUs, 0.83 ms for the
// step without the reduction vs 1.04 ms with it, and fewer channels only make the reduction the
// critical path (NCCL_MAX_CTAS = 8: 1.24 ms).  The kernel below needs no shared memory and 32
// registers per thread, so its CTAs co-reside with the GEMM CTAs instead of displacing them.
//
// Algorithm (two-shot, every byte crosses NVLink once in each direction):
//   every rank's gradient arena lives at the same offset of a symmetric allocation whose
//   peer addresses all ranks know (torch.distributed._symmetric_memory / cudaIpc);
//   barrier A   "my gradients are written"  (flag store to each peer, spin on own flags)
//   phase 1     rank r owns slice r of the bucket: 128-bit loads of that slice from every
//               rank (peers: volatile, they bypass the local L2), summed in rank order,
//               and the sum is stored to EVERY rank's arena (peer stores)
//   barrier B   "my stores to you are done" (system fence, flag store, spin)
//   after which each rank holds the identical, complete sum and runs its own SGD step.
// Flags are per (CTA slot, source rank) words in the same symmetric allocation, compared
// against an epoch the kernel keeps in device memory, so CUDA-graph replays need no host
// bookkeeping.  CTA b only ever waits for CTA b of its peers: no co-residency assumption.

#include "kcnn_common.cuh"

#include <stdlib.h>

namespace kcnn {
namespace p2p {

constexpr int kMaxRanks = 8;
constexpr int kMaxCtas = 64;
constexpr int kThreads = 512;
// uint32 words of one flag channel
constexpr int kReady = 0;                            // [kMaxCtas][kMaxRanks]
constexpr int kDone = kMaxCtas * kMaxRanks;          // [kMaxCtas][kMaxRanks]
constexpr int kEpoch = 2 * kMaxCtas * kMaxRanks;     // [kMaxCtas]   (local)
constexpr int kError = kEpoch + kMaxCtas;            // [1]          (local): spin limit hit
constexpr int kChannelWords = 2048;

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float *p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

struct Peers {
  float *buf[kMaxRanks];          // base of each rank's symmetric allocation (this process's mapping)
};

// One lane per peer: publish `epoch` in the peer's flag word for (slot, me), then wait until the
// peer has published it in mine.  A bounded spin: a rank that never arrives leaves an error
// mark instead of a hung GPU.
__device__ __forceinline__ void cross_barrier(const Peers &pr, size_t flag_off, int which, int slot, int rank,
                                              int world, uint32_t epoch) {
  const int t = threadIdx.x;
  if (t < world && t != rank) {
    uint32_t *theirs = reinterpret_cast<uint32_t *>(pr.buf[t] + flag_off) + which + slot * kMaxRanks + rank;
    const uint32_t *mine = reinterpret_cast<const uint32_t *>(pr.buf[rank] + flag_off) + which + slot * kMaxRanks + t;
    st_release_sys(theirs, epoch);
    unsigned long long spins = 0;
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spins > (1ull << 25)) {          // seconds, not forever
        reinterpret_cast<uint32_t *>(pr.buf[rank] + flag_off)[kError] = 1u;
        break;
      }
    }
  }
  __syncthreads();
}

// In-switch reduction (NVLS): one 128-bit multimem.ld_reduce on the MULTICAST address of an
// element makes the NVSwitch fetch it from every rank's replica and return the sum; multimem.st
// writes a value to every replica.  Per GPU the links then carry (1 + 1/world) x the bucket in
// each direction instead of 2 (world-1)/world x for the two-shot above, and the SMs do no adds.
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float *mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float *mc, const float4 &v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// mc: multicast mapping of the same symmetric allocation (nullptr: two-shot over unicast peers)
__global__ void __launch_bounds__(kThreads)
p2p_allreduce_kernel(Peers pr, float *mc, int rank, int world, size_t off, size_t n4, size_t flag_off) {
  const int b = blockIdx.x, G = gridDim.x, t = threadIdx.x;
  __shared__ uint32_t s_epoch;
  uint32_t *my_flags = reinterpret_cast<uint32_t *>(pr.buf[rank] + flag_off);
  if (t == 0) s_epoch = my_flags[kEpoch + b] + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;

  cross_barrier(pr, flag_off, kReady, b, rank, world, epoch);

  const size_t per = (n4 + world - 1) / world;
  const size_t lo = (size_t)rank * per;
  const size_t hi = lo + per < n4 ? lo + per : n4;
  const size_t step = (size_t)G * kThreads;
  constexpr int U = 4;                                   // independent 128-bit loads per peer in flight
  if (mc != nullptr) {
    float *m = mc + off;
    for (size_t i0 = lo + (size_t)b * kThreads + t; i0 < hi; i0 += U * step) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + u * step;
        if (i < hi) v[u] = multimem_ld_reduce_f4(m + 4 * i);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + u * step;
        if (i < hi) multimem_st_f4(m + 4 * i, v[u]);
      }
    }
  } else
  for (size_t i0 = lo + (size_t)b * kThreads + t; i0 < hi; i0 += U * step) {
    float4 acc[U];
#pragma unroll
    for (int u = 0; u < U; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < world; p++) {                    // rank order: every owner sums the same way
      const float *src = pr.buf[p] + off;
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + u * step;
        if (i < hi) v[u] = p == rank ? *reinterpret_cast<const float4 *>(src + 4 * i) : ld_volatile_f4(src + 4 * i);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + u * step;
        if (i < hi) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
      }
    }
    for (int p = 0; p < world; p++) {
      float *dst = pr.buf[p] + off;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + u * step;
        if (i < hi) *reinterpret_cast<float4 *>(dst + 4 * i) = acc[u];
      }
    }
  }
  __threadfence_system();                                // my stores, before anyone is told
  __syncthreads();
  cross_barrier(pr, flag_off, kDone, b, rank, world, epoch);
  if (t == 0) my_flags[kEpoch + b] = epoch;
}

}  // namespace p2p
}  // namespace kcnn

using namespace kcnn;

extern "C" {

size_t kcnn_p2p_flag_floats(void) { return (size_t)p2p::kChannelWords * 2; }   // two channels

static int p2p_launch(void *stream, const unsigned long long *peer_bases, unsigned long long mc_base, int rank,
                      int world, size_t offset_floats, size_t count_floats, size_t flag_offset_floats, int channel) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (world < 1 || world > p2p::kMaxRanks || rank < 0 || rank >= world) return -1;
  if (channel < 0 || channel > 1) return -1;
  if ((count_floats & 3) != 0 || (offset_floats & 3) != 0 || (flag_offset_floats & 3) != 0) return -1;
  if (count_floats == 0 || world == 1) return 0;
  p2p::Peers pr;
  for (int p = 0; p < p2p::kMaxRanks; p++)
    pr.buf[p] = p < world ? reinterpret_cast<float *>(static_cast<uintptr_t>(peer_bases[p])) : nullptr;
  static int max_ctas = -1;
  if (max_ctas < 0) {
    const char *e = getenv("KCNN_P2P_CTAS");
    max_ctas = e ? atoi(e) : 32;
    if (max_ctas < 1) max_ctas = 1;
    if (max_ctas > p2p::kMaxCtas) max_ctas = p2p::kMaxCtas;
  }
  const size_t n4 = count_floats >> 2;
  const size_t per = (n4 + world - 1) / world;
  size_t want = (per + p2p::kThreads * 4 - 1) / (p2p::kThreads * 4);       // one pass of 4 units per thread
  if (want < 1) want = 1;
  const unsigned grid = (unsigned)(want < (size_t)max_ctas ? want : (size_t)max_ctas);
  KCNN_LAUNCH(p2p::p2p_allreduce_kernel, grid, p2p::kThreads, 0, st, pr,
              reinterpret_cast<float *>(static_cast<uintptr_t>(mc_base)), rank, world, offset_floats, n4,
              flag_offset_floats + (size_t)channel * p2p::kChannelWords);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int kcnn_p2p_allreduce_f32(void *stream, const unsigned long long *peer_bases, int rank, int world,
                           size_t offset_floats, size_t count_floats, size_t flag_offset_floats, int channel) {
  return p2p_launch(stream, peer_bases, 0ull, rank, world, offset_floats, count_floats, flag_offset_floats, channel);
}

int kcnn_p2p_allreduce_multicast_f32(void *stream, const unsigned long long *peer_bases,
                                     unsigned long long multicast_base, int rank, int world, size_t offset_floats,
                                     size_t count_floats, size_t flag_offset_floats, int channel) {
  if (multicast_base == 0ull) return -1;
  return p2p_launch(stream, peer_bases, multicast_base, rank, world, offset_floats, count_floats,
                    flag_offset_floats, channel);
}

/* 1 when a barrier of this rank gave up waiting for a peer (the arena contents are then undefined). */
int kcnn_p2p_error(const float *local_base, size_t flag_offset_floats) {
  unsigned int v[2] = {0, 0};
  for (int c = 0; c < 2; c++)
    cudaMemcpy(&v[c], reinterpret_cast<const unsigned int *>(local_base + flag_offset_floats) +
                          c * p2p::kChannelWords + p2p::kError, sizeof(unsigned int), cudaMemcpyDeviceToHost);
  return (v[0] | v[1]) ? 1 : 0;
}

}  // extern "C"
