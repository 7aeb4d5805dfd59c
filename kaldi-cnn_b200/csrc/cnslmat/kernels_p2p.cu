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
// critical path (NCCL_MAX_CTAS = 8: 1.24 ms).  The kernels below need (next to) no shared memory and
// at most 64 registers per thread, so their CTAs co-reside with the GEMM CTAs instead of displacing them
// (the one variant that did not -- 124 registers x 512 threads -- is kept behind KCNN_P2P_LIGHT=0 and is
// 3 % slower per step at 8 GPUs).
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
#include "kcnn_capi.h"

#include <stdlib.h>
#include <string.h>

namespace kcnn {
namespace p2p {

constexpr int kMaxRanks = 8;
constexpr int kMaxCtas = 128;
constexpr int kThreads = 512;
// uint32 words of one flag channel
constexpr int kReady = 0;                            // [kMaxCtas][kMaxRanks]
constexpr int kDone = kMaxCtas * kMaxRanks;          // [kMaxCtas][kMaxRanks]
constexpr int kEpoch = 2 * kMaxCtas * kMaxRanks;     // [kMaxCtas]   (local)
constexpr int kError = kEpoch + kMaxCtas;            // [1]          (local): spin limit hit
constexpr int kChannelWords = 4096;

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

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One lane per peer: publish `epoch` in the peer's flag word for (slot, me), then wait until the
// peer has published it in mine.  The wait is bounded by WALL-CLOCK time (timeout_ns, default 20 s:
// a first-step cudaMalloc, a graph capture or a checkpoint write on a peer must not trip it): a rank
// that never arrives leaves an error mark instead of a hung GPU, and the barrier then returns false
// in every thread of the CTA -- the caller skips its stores, so that neither gradients nor weights are
// overwritten with a partial sum.  kcnn_p2p_error() / NnetDataParallel report the mark to the host.
__device__ __forceinline__ bool cross_barrier(const Peers &pr, size_t flag_off, int which, int slot, int rank,
                                              int world, uint32_t epoch, unsigned long long timeout_ns) {
  __shared__ int s_failed;
  const int t = threadIdx.x;
  if (t == 0) s_failed = 0;
  __syncthreads();
  if (t < world && t != rank) {
    uint32_t *theirs = reinterpret_cast<uint32_t *>(pr.buf[t] + flag_off) + which + slot * kMaxRanks + rank;
    const uint32_t *mine = reinterpret_cast<const uint32_t *>(pr.buf[rank] + flag_off) + which + slot * kMaxRanks + t;
    st_release_sys(theirs, epoch);
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) {
        reinterpret_cast<uint32_t *>(pr.buf[rank] + flag_off)[kError] = 1u;
        s_failed = 1;
        break;
      }
    }
  }
  __syncthreads();
  return s_failed == 0;
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
p2p_allreduce_kernel(Peers pr, float *mc, int rank, int world, size_t off, size_t n4, size_t flag_off,
                     unsigned long long timeout_ns) {
  kcnn::pdl_wait();            // (no early trigger: what follows a collective waits for all of it)
  const int b = blockIdx.x, G = gridDim.x, t = threadIdx.x;
  __shared__ uint32_t s_epoch;
  uint32_t *my_flags = reinterpret_cast<uint32_t *>(pr.buf[rank] + flag_off);
  if (t == 0) s_epoch = my_flags[kEpoch + b] + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;

  if (!cross_barrier(pr, flag_off, kReady, b, rank, world, epoch, timeout_ns)) {
    if (t == 0) my_flags[kEpoch + b] = epoch;
    return;                                              // a peer is missing: leave the arena as it is
  }

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
  cross_barrier(pr, flag_off, kDone, b, rank, world, epoch, timeout_ns);
  if (t == 0) my_flags[kEpoch + b] = epoch;
}

// ---- reduce-scatter + momentum SGD + all-gather in ONE kernel -------------------------------------
//
// The data-parallel step of round 1 all-reduced every gradient bucket and then ran the SGD pass on
// every rank: each rank read 147 MB of reduced gradients back and re-did the same 733 MB
// read-modify-write of (W, prev_grad) that all other ranks did too.  Here the OWNER of a slice does
// the update once:
//   the symmetric allocation holds [gradient arena | parameter arena | flags]; the parameter arena
//   is laid out exactly like the gradient arena (per layer: W rows x pitch, then the bias), so element
//   i of a bucket has its gradient at G[i] and its weight at G[i + param_delta];
//   barrier A   every rank has finished this layer's backward (its gradients are written AND its
//               input-gradient GEMM no longer reads the weights)
//   rank r, for its slice of the bucket:  g = sum over ranks of G_p[i]   (peer loads in rank order, or
//               one multimem.ld_reduce inside the NVSwitch)
//               weights:  prev = m prev + a_decay W + a_grad g ; W += prev     (reference
//               nnet0/nnet-component-nnet0.cc:767-773, 1138-1142 -- the same sgd_apply as everywhere;
//               prev_grad is touched by the owner only, so momentum is SHARDED over the ranks)
//               bias:     b += a_grad g                                          (:775, :1137)
//               the NEW weights are stored to every rank's parameter arena (peer stores / multimem.st)
//   barrier B   all replicas hold the new weights.
// NVLink traffic is that of the all-reduce; HBM traffic per rank drops from ~1.0 GB to ~0.37 GB per
// step and the separate apply pass (0.12 ms) disappears.  Bit-identical to all-reduce + apply (same
// summation order, same sgd_apply, one owner per element).
struct SgdBucket {
  size_t off;              // floats: start of the bucket in the gradient arena
  size_t n4;               // float4 units in the bucket
  size_t w4;               // float4 units of the weight matrix (the rest is the bias)
  size_t param_delta;      // floats from a gradient element to its parameter
  float *prev;             // LOCAL momentum matrix, indexed like the weight part of the bucket
  float momentum, a_decay, a_grad;
};

__device__ __forceinline__ void sgd4(float4 &w, float4 &p, const float4 &g, const SgdBucket &k) {
  float *wv = &w.x, *pv = &p.x;
  const float *gv = &g.x;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    pv[j] = pv[j] * k.momentum;
    pv[j] = fmaf(k.a_decay, wv[j], pv[j]);
    pv[j] = fmaf(k.a_grad, gv[j], pv[j]);
    wv[j] = wv[j] + pv[j];
  }
}

// One pass over this rank's slice [lo, hi) of a bucket.  kWorld ranks (compile time, so that the peer
// loads of a unit are independent instructions), U units per thread: all U * kWorld gradient loads and the
// 2 U loads of (W, prev) are issued before the first add -- a peer load over NVLink takes ~2 us, so the
// bytes in flight per SM decide the link throughput (measured at 2 GPUs: U = 2 with a serial loop over the
// ranks kept ~0.5 MB in flight per GPU and moved ~260 GB/s).  Sum order = rank order, as in the all-reduce.
template <int kWorld, int U, int kT = kThreads>
__device__ __forceinline__ void reduce_sgd_slice(const Peers &pr, float *mc, int rank, int world, const SgdBucket &k,
                                                 size_t lo, size_t hi, int b, int G, int t) {
  const size_t step = (size_t)G * kT;
  float *wmine = pr.buf[rank] + k.off + k.param_delta;
  for (size_t i0 = lo + (size_t)b * kT + t; i0 < hi; i0 += U * step) {
    float4 v[U][kWorld], w[U], p[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = i0 + u * step;
      if (i >= hi) continue;
      if (mc != nullptr) {
        v[u][0] = multimem_ld_reduce_f4(mc + k.off + 4 * i);
      } else {
#pragma unroll
        for (int q = 0; q < kWorld; q++) {
          if (q >= world) continue;
          const float *src = pr.buf[q] + k.off + 4 * i;
          v[u][q] = q == rank ? *reinterpret_cast<const float4 *>(src) : ld_volatile_f4(src);
        }
      }
      w[u] = *reinterpret_cast<const float4 *>(wmine + 4 * i);
      if (i < k.w4) p[u] = *reinterpret_cast<const float4 *>(k.prev + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = i0 + u * step;
      if (i >= hi) continue;
      float4 g;
      if (mc != nullptr) {
        g = v[u][0];
      } else {
        g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < kWorld; q++)
          if (q < world) { g.x += v[u][q].x; g.y += v[u][q].y; g.z += v[u][q].z; g.w += v[u][q].w; }
      }
      if (i < k.w4) {
        sgd4(w[u], p[u], g, k);
        *reinterpret_cast<float4 *>(k.prev + 4 * i) = p[u];
      } else {
        w[u].x = fmaf(k.a_grad, g.x, w[u].x); w[u].y = fmaf(k.a_grad, g.y, w[u].y);
        w[u].z = fmaf(k.a_grad, g.z, w[u].z); w[u].w = fmaf(k.a_grad, g.w, w[u].w);
      }
      if (mc != nullptr) {
        multimem_st_f4(mc + k.off + k.param_delta + 4 * i, w[u]);
      } else {
#pragma unroll
        for (int q = 0; q < kWorld; q++)
          if (q < world) *reinterpret_cast<float4 *>(pr.buf[q] + k.off + k.param_delta + 4 * i) = w[u];
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
p2p_reduce_sgd_kernel(Peers pr, float *mc, int rank, int world, SgdBucket k, size_t flag_off,
                      unsigned long long timeout_ns) {
  kcnn::pdl_wait();
  const int b = blockIdx.x, G = gridDim.x, t = threadIdx.x;
  __shared__ uint32_t s_epoch;
  uint32_t *my_flags = reinterpret_cast<uint32_t *>(pr.buf[rank] + flag_off);
  if (t == 0) s_epoch = my_flags[kEpoch + b] + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  if (!cross_barrier(pr, flag_off, kReady, b, rank, world, epoch, timeout_ns)) {
    if (t == 0) my_flags[kEpoch + b] = epoch;
    return;                                              // no partial update: weights stay as they are
  }
  const size_t per = (k.n4 + world - 1) / world;
  const size_t lo = (size_t)rank * per;
  const size_t hi = lo + per < k.n4 ? lo + per : k.n4;
  if (mc != nullptr)   reduce_sgd_slice<1, 8>(pr, mc, rank, world, k, lo, hi, b, G, t);
  else if (world <= 2) reduce_sgd_slice<2, 4>(pr, mc, rank, world, k, lo, hi, b, G, t);
  else if (world <= 4) reduce_sgd_slice<4, 2>(pr, mc, rank, world, k, lo, hi, b, G, t);
  else                 reduce_sgd_slice<8, 1>(pr, mc, rank, world, k, lo, hi, b, G, t);
  __threadfence_system();
  __syncthreads();
  cross_barrier(pr, flag_off, kDone, b, rank, world, epoch, timeout_ns);
  if (t == 0) my_flags[kEpoch + b] = epoch;
}

// Several (small) layers' buckets between ONE pair of barriers: the convolution stack of the benchmarked
// model is six buckets of 20 K .. 790 K floats whose gradients all exist when the backward pass reaches the
// bottom; one launch and two peer barriers instead of six and twelve.  Same per-bucket slices and arithmetic
// as p2p_reduce_sgd_kernel, bucket after bucket.
constexpr int kMaxBuckets = 8;
template <int V> struct IntTag { static constexpr int value = V; };
struct SgdBucketList {
  SgdBucket b[kMaxBuckets];
  int n;
};

__global__ void __launch_bounds__(kThreads)
p2p_reduce_sgd_multi_kernel(Peers pr, float *mc, int rank, int world, SgdBucketList list, size_t flag_off,
                            unsigned long long timeout_ns) {
  kcnn::pdl_wait();
  const int b = blockIdx.x, G = gridDim.x, t = threadIdx.x;
  __shared__ uint32_t s_epoch;
  uint32_t *my_flags = reinterpret_cast<uint32_t *>(pr.buf[rank] + flag_off);
  if (t == 0) s_epoch = my_flags[kEpoch + b] + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  if (!cross_barrier(pr, flag_off, kReady, b, rank, world, epoch, timeout_ns)) {
    if (t == 0) my_flags[kEpoch + b] = epoch;
    return;
  }
  // (the list is indexed with a run-time j: from shared memory, not from a local copy of the parameters)
  __shared__ SgdBucket sb[kMaxBuckets];
  if (t < list.n) sb[t] = list.b[t];
  __syncthreads();
  auto all_buckets = [&](auto tag_world, auto tag_u) {
    for (int j = 0; j < list.n; j++) {
      const SgdBucket k = sb[j];
      const size_t per = (k.n4 + world - 1) / world;
      const size_t lo = (size_t)rank * per;
      const size_t hi = lo + per < k.n4 ? lo + per : k.n4;
      reduce_sgd_slice<decltype(tag_world)::value, decltype(tag_u)::value>(pr, mc, rank, world, k, lo, hi, b, G, t);
    }
  };
  if (mc != nullptr)   all_buckets(IntTag<1>(), IntTag<8>());
  else if (world <= 2) all_buckets(IntTag<2>(), IntTag<4>());
  else if (world <= 4) all_buckets(IntTag<4>(), IntTag<2>());
  else                 all_buckets(IntTag<8>(), IntTag<1>());
  __threadfence_system();
  __syncthreads();
  cross_barrier(pr, flag_off, kDone, b, rank, world, epoch, timeout_ns);
  if (t == 0) my_flags[kEpoch + b] = epoch;
}

// "Light" form of the two kernels above: 256 threads and at most 64 registers per thread (16 K registers per
// CTA, no shared memory to speak of), so that a communication CTA fits on an SM NEXT TO a GEMM CTA (192 threads
// x 118 registers, ~200 KB of shared memory).  The 512-thread form needs 124 registers per thread -- a whole
// SM's register file per CTA -- and can only run where no GEMM CTA is resident.  Fewer loads in flight per
// thread (U halved); twice the CTAs.  Measured (step, ms): 2 GPUs 0.780 -> 0.766, 8 GPUs 0.870 -> 0.845.
// The default; KCNN_P2P_LIGHT=0 selects the 512-thread form.
constexpr int kLightThreads = 256;

template <bool kMulti>
__global__ void __launch_bounds__(kLightThreads, 4)
p2p_reduce_sgd_light_kernel(Peers pr, float *mc, int rank, int world, SgdBucketList list, size_t flag_off,
                            unsigned long long timeout_ns) {
  kcnn::pdl_wait();
  const int b = blockIdx.x, G = gridDim.x, t = threadIdx.x;
  __shared__ uint32_t s_epoch;
  __shared__ SgdBucket sb[kMaxBuckets];
  uint32_t *my_flags = reinterpret_cast<uint32_t *>(pr.buf[rank] + flag_off);
  if (t == 0) s_epoch = my_flags[kEpoch + b] + 1u;
  if (t < list.n) sb[t] = list.b[t];
  __syncthreads();
  const uint32_t epoch = s_epoch;
  if (!cross_barrier(pr, flag_off, kReady, b, rank, world, epoch, timeout_ns)) {
    if (t == 0) my_flags[kEpoch + b] = epoch;
    return;
  }
  auto all_buckets = [&](auto tag_world, auto tag_u) {
    const int n = kMulti ? list.n : 1;
    for (int j = 0; j < n; j++) {
      const SgdBucket k = sb[j];
      const size_t per = (k.n4 + world - 1) / world;
      const size_t lo = (size_t)rank * per;
      const size_t hi = lo + per < k.n4 ? lo + per : k.n4;
      reduce_sgd_slice<decltype(tag_world)::value, decltype(tag_u)::value, kLightThreads>(pr, mc, rank, world, k, lo, hi,
                                                                                          b, G, t);
    }
  };
  if (mc != nullptr)   all_buckets(IntTag<1>(), IntTag<4>());
  else if (world <= 2) all_buckets(IntTag<2>(), IntTag<2>());
  else if (world <= 4) all_buckets(IntTag<4>(), IntTag<1>());
  else                 all_buckets(IntTag<8>(), IntTag<1>());
  __threadfence_system();
  __syncthreads();
  cross_barrier(pr, flag_off, kDone, b, rank, world, epoch, timeout_ns);
  if (t == 0) my_flags[kEpoch + b] = epoch;
}

}  // namespace p2p
}  // namespace kcnn

using namespace kcnn;

extern "C" {

size_t kcnn_p2p_flag_floats(void) { return (size_t)p2p::kChannelWords * 2; }   // two channels

static unsigned long long p2p_timeout_ns() {
  static long long v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_P2P_TIMEOUT_MS");
    v = e ? atoll(e) : 20000;
    if (v < 1) v = 1;
  }
  return (unsigned long long)v * 1000000ull;
}

static bool p2p_light() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("KCNN_P2P_LIGHT");          // default on; 0 selects the 512-thread form
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static int p2p_max_ctas() {
  static int max_ctas = -1;
  if (max_ctas < 0) {
    const char *e = getenv("KCNN_P2P_CTAS");
    max_ctas = e ? atoi(e) : 32;
    if (max_ctas < 1) max_ctas = 1;
    if (max_ctas > p2p::kMaxCtas) max_ctas = p2p::kMaxCtas;
  }
  return max_ctas;
}

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
  const int max_ctas = p2p_max_ctas();
  const size_t n4 = count_floats >> 2;
  const size_t per = (n4 + world - 1) / world;
  size_t want = (per + p2p::kThreads * 4 - 1) / (p2p::kThreads * 4);       // one pass of 4 units per thread
  if (want < 1) want = 1;
  const unsigned grid = (unsigned)(want < (size_t)max_ctas ? want : (size_t)max_ctas);
  KCNN_LAUNCH(p2p::p2p_allreduce_kernel, grid, p2p::kThreads, 0, st, pr,
              reinterpret_cast<float *>(static_cast<uintptr_t>(mc_base)), rank, world, offset_floats, n4,
              flag_offset_floats + (size_t)channel * p2p::kChannelWords, p2p_timeout_ns());
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int kcnn_p2p_reduce_sgd_f32(void *stream, const unsigned long long *peer_bases, unsigned long long multicast_base,
                            int rank, int world, size_t offset_floats, size_t count_floats, size_t weight_floats,
                            size_t param_delta_floats, float *prev_grad, float momentum, float decay_alpha,
                            float grad_alpha, size_t flag_offset_floats, int channel) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (world < 1 || world > p2p::kMaxRanks || rank < 0 || rank >= world) return -1;
  if (channel < 0 || channel > 1) return -1;
  if ((count_floats & 3) || (offset_floats & 3) || (flag_offset_floats & 3) || (weight_floats & 3) ||
      (param_delta_floats & 3) || weight_floats > count_floats)
    return -1;
  if (count_floats == 0) return 0;
  p2p::Peers pr;
  for (int p = 0; p < p2p::kMaxRanks; p++)
    pr.buf[p] = p < world ? reinterpret_cast<float *>(static_cast<uintptr_t>(peer_bases[p])) : nullptr;
  p2p::SgdBucket k;
  k.off = offset_floats; k.n4 = count_floats >> 2; k.w4 = weight_floats >> 2; k.param_delta = param_delta_floats;
  k.prev = prev_grad; k.momentum = momentum; k.a_decay = decay_alpha; k.a_grad = grad_alpha;
  const size_t per = (k.n4 + world - 1) / world;
  const int max_ctas = p2p_max_ctas();
  if (p2p_light()) {
    p2p::SgdBucketList list;
    list.b[0] = k;
    list.n = 1;
    size_t want = (per + p2p::kLightThreads * 2 - 1) / (p2p::kLightThreads * 2);
    if (want < 1) want = 1;
    const size_t cap = (size_t)(2 * max_ctas < p2p::kMaxCtas ? 2 * max_ctas : p2p::kMaxCtas);
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    KCNN_LAUNCH(p2p::p2p_reduce_sgd_light_kernel<false>, grid, p2p::kLightThreads, 0, st, pr,
                reinterpret_cast<float *>(static_cast<uintptr_t>(multicast_base)), rank, world, list,
                flag_offset_floats + (size_t)channel * p2p::kChannelWords, p2p_timeout_ns());
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  size_t want = (per + p2p::kThreads * 2 - 1) / (p2p::kThreads * 2);
  if (want < 1) want = 1;
  const unsigned grid = (unsigned)(want < (size_t)max_ctas ? want : (size_t)max_ctas);
  KCNN_LAUNCH(p2p::p2p_reduce_sgd_kernel, grid, p2p::kThreads, 0, st, pr,
              reinterpret_cast<float *>(static_cast<uintptr_t>(multicast_base)), rank, world, k,
              flag_offset_floats + (size_t)channel * p2p::kChannelWords, p2p_timeout_ns());
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int kcnn_p2p_reduce_sgd_multi_f32(void *stream, const unsigned long long *peer_bases, unsigned long long multicast_base,
                                  int rank, int world, int num_buckets, const KcnnSgdBucket *buckets,
                                  size_t param_delta_floats, size_t flag_offset_floats, int channel) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (world < 1 || world > p2p::kMaxRanks || rank < 0 || rank >= world) return -1;
  if (channel < 0 || channel > 1 || num_buckets < 1 || num_buckets > p2p::kMaxBuckets || buckets == nullptr) return -1;
  if ((flag_offset_floats & 3) || (param_delta_floats & 3)) return -1;
  p2p::Peers pr;
  for (int p = 0; p < p2p::kMaxRanks; p++)
    pr.buf[p] = p < world ? reinterpret_cast<float *>(static_cast<uintptr_t>(peer_bases[p])) : nullptr;
  p2p::SgdBucketList list;
  list.n = 0;
  size_t largest = 0;
  for (int j = 0; j < num_buckets; j++) {
    const KcnnSgdBucket &q = buckets[j];
    if ((q.count_floats & 3) || (q.offset_floats & 3) || (q.weight_floats & 3) || q.weight_floats > q.count_floats)
      return -1;
    if (q.count_floats == 0) continue;
    p2p::SgdBucket &k = list.b[list.n++];
    k.off = q.offset_floats; k.n4 = q.count_floats >> 2; k.w4 = q.weight_floats >> 2; k.param_delta = param_delta_floats;
    k.prev = q.prev_grad; k.momentum = q.momentum; k.a_decay = q.decay_alpha; k.a_grad = q.grad_alpha;
    const size_t per = (k.n4 + world - 1) / world;
    if (per > largest) largest = per;
  }
  if (list.n == 0) return 0;
  const int max_ctas = p2p_max_ctas();
  if (p2p_light()) {
    size_t want = (largest + p2p::kLightThreads * 2 - 1) / (p2p::kLightThreads * 2);
    if (want < 1) want = 1;
    const size_t cap = (size_t)(2 * max_ctas < p2p::kMaxCtas ? 2 * max_ctas : p2p::kMaxCtas);
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    KCNN_LAUNCH(p2p::p2p_reduce_sgd_light_kernel<true>, grid, p2p::kLightThreads, 0, st, pr,
                reinterpret_cast<float *>(static_cast<uintptr_t>(multicast_base)), rank, world, list,
                flag_offset_floats + (size_t)channel * p2p::kChannelWords, p2p_timeout_ns());
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  size_t want = (largest + p2p::kThreads * 2 - 1) / (p2p::kThreads * 2);
  if (want < 1) want = 1;
  const unsigned grid = (unsigned)(want < (size_t)max_ctas ? want : (size_t)max_ctas);
  KCNN_LAUNCH(p2p::p2p_reduce_sgd_multi_kernel, grid, p2p::kThreads, 0, st, pr,
              reinterpret_cast<float *>(static_cast<uintptr_t>(multicast_base)), rank, world, list,
              flag_offset_floats + (size_t)channel * p2p::kChannelWords, p2p_timeout_ns());
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

/* ---- symmetric memory from CUDA IPC handles (no framework involved) -------------------------------
 * Each rank allocates its arena with kcnn_ipc_alloc, sends the 64-byte handle to its peers by whatever
 * means the host has (MPI, a socket, torch.distributed.all_gather_object ...), and maps the peers'
 * arenas with kcnn_ipc_open (peer access is enabled on first use). */
int kcnn_ipc_alloc(size_t bytes, void **ptr, unsigned char *handle64) {
  if (!ptr || !handle64 || bytes == 0) return -1;
  if (cudaMalloc(ptr, bytes) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (cudaMemset(*ptr, 0, bytes) != cudaSuccess) { cudaGetLastError(); return -1; }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, *ptr) != cudaSuccess) { cudaGetLastError(); cudaFree(*ptr); *ptr = nullptr; return -1; }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
int kcnn_ipc_open(const unsigned char *handle64, void **ptr) {
  if (!ptr || !handle64) return -1;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  if (cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return -1; }
  return 0;
}
int kcnn_ipc_close(void *ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? 0 : (cudaGetLastError(), -1); }
int kcnn_ipc_free(void *ptr) { return cudaFree(ptr) == cudaSuccess ? 0 : (cudaGetLastError(), -1); }

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

/* Device address of the error word of one flag channel (for an asynchronous copy to pinned memory). */
const unsigned int *kcnn_p2p_error_word(const float *local_base, size_t flag_offset_floats, int channel) {
  return reinterpret_cast<const unsigned int *>(local_base + flag_offset_floats) + channel * p2p::kChannelWords +
         p2p::kError;
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
