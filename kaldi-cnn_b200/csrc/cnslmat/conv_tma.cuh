// kaldi-cnn_b200/csrc/cnslmat/conv_tma.cuh
//
// Convolution forward / input-gradient / weight-gradient of the time-axis layers
// (in_height = 1: every layer of egs/exp/nnet/nnet.config after conv1) on the TMA-fed
// tcgen05 pipeline of gemm_tma.cuh.
//
// The reference layout [C][W] (cnsl-cu-kernels.cu:28-32) puts the reduction axis c at a
// pitch of W floats -- 72 / 56 / 24 bytes for W = 18 / 14 / 6 -- which no tensor map can
// express (pitches must be multiples of 16 bytes).  So each call first transposes the
// activation it reduces over into a channels-last staging copy [N][W][C] (one
// bandwidth-bound pass over a tensor of a few MB that stays in L2; the reference's im2col
// writes KH*KW times as much), and every operand box after that is pure TMA addressing:
//
//   fprop  Y[(n,ow), g]  = sum_{kw,c} Xcl[n, ow+kw-pw, c] K[(c,kw), g]
//          A box {32 c, OW, NB samples} at (c0, kw-pw, n0)   K-major   (zero padding = OOB fill)
//          B box {32 g, 1, 32 c}        at (g0, kw, c0)      MN-major  (3-D view of the kernel matrix)
//   dgrad  dX[(n,w), c]  = sum_{kw,g} dYcl[n, w+pw-kw, g] K[(c,kw), g]
//          A box {32 g, W, NB}          at (g0, pw-kw, n0)   K-major   (flip + padding = coordinates)
//          B box {32 g, 1, 128 c}       at (g0, kw, c0)      K-major
//   wgrad  dK[(c,kw), g] = sum_{ow,n} Xcl[n, ow+kw-pw, c] dYcl[n, ow, g]
//          A box {32 c, 1, 32 n}        at (c, ow+kw-pw, n0) MN-major, rows m = kw*C + c
//          B box {32 g, 1, 32 n}        at (g0, ow, n0)      MN-major
//
// The epilogues write the reference layouts directly: fprop / dgrad transpose the
// [(n,pos)][map] tile into contiguous [map][pos] runs of each sample (folds
// _convmat_to_out, TpBlock and the bias AddMatRepVec); wgrad writes kernel rows
// (c*KW + kw) (folds ModPermuteRow) or split-K partials whose reduction also applies the
// momentum / weight-decay SGD step.

#ifndef KCNN_CONV_TMA_CUH_
#define KCNN_CONV_TMA_CUH_

#include "gemm_tma.cuh"

namespace kcnn {
namespace tma {

// ------------------------------------------------------------- channels-last pack --

// in: [N][ld] rows holding [C][R] (R fastest)  ->  out: [N][R][C] (C fastest), dense.
// One CTA moves kSamples samples through shared memory (row pitch R|1: conflict-free both
// ways).  With kColSum it also writes partial[blockIdx.x][c] = sum over its samples and r
// (the bias gradient's first stage, reference nnet0/nnet-component-nnet0.cc:775).
template <bool kColSum>
__global__ void __launch_bounds__(256)
pack_channels_last_kernel(const float *__restrict__ in, int ld, int N, int C, int R, float *__restrict__ out,
                          float *__restrict__ partial, int samples_per_cta, FastDiv div_r, FastDiv div_c) {
  extern __shared__ float tile[];
  const int rp = R | 1;
  const int per = C * R;
  const int n_begin = blockIdx.x * samples_per_cta;
  const int n_end = min(N, n_begin + samples_per_cta);
  float *colacc = tile + C * rp;                // kColSum: per-channel sums over this CTA's samples
  if (kColSum)
    for (int c = threadIdx.x; c < C; c += 256) colacc[c] = 0.f;     // channel c stays with one thread
  for (int n = n_begin; n < n_end; n++) {
    const float *src = in + (size_t)n * ld;
    for (int i = threadIdx.x; i < per; i += 256) {
      uint32_t c, r;
      div_r.divmod((uint32_t)i, c, r);
      tile[c * rp + r] = __ldg(src + i);
    }
    __syncthreads();
    float *dst = out + (size_t)n * per;
    for (int i = threadIdx.x; i < per; i += 256) {
      uint32_t r, c;
      div_c.divmod((uint32_t)i, r, c);
      dst[i] = tile[c * rp + r];
    }
    if (kColSum) {
      for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int r = 0; r < R; r++) s += tile[c * rp + r];
        colacc[c] += s;
      }
    }
    __syncthreads();
  }
  if (kColSum) {
    float *prow = partial + (size_t)blockIdx.x * C;
    for (int c = threadIdx.x; c < C; c += 256) prow[c] = colacc[c];
  }
}

constexpr int kPackSamples = 4;
constexpr int kPackMaxSmem = 96 * 1024;

inline size_t pack_smem_bytes(int C, int R) { return (size_t)C * ((R | 1) + 1) * sizeof(float); }

// Returns the number of partial rows written when colsum_partial != nullptr.
inline int launch_pack(cudaStream_t st, const float *in, int ld, int N, int C, int R, float *out,
                       float *colsum_partial) {
  const size_t smem = pack_smem_bytes(C, R);
  const int ctas = (N + kPackSamples - 1) / kPackSamples;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(pack_channels_last_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPackMaxSmem);
    cudaFuncSetAttribute(pack_channels_last_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPackMaxSmem);
    attr_set = true;
  }
  if (colsum_partial)
    KCNN_LAUNCH(pack_channels_last_kernel<true>, ctas, 256, smem, st, in, ld, N, C, R, out, colsum_partial,
                kPackSamples, FastDiv((uint32_t)R), FastDiv((uint32_t)C));
  else
    KCNN_LAUNCH(pack_channels_last_kernel<false>, ctas, 256, smem, st, in, ld, N, C, R, out, nullptr,
                kPackSamples, FastDiv((uint32_t)R), FastDiv((uint32_t)C));
  return ctas;
}

// ------------------------------------------------------- fprop / dgrad problem --

// Rows of the GEMM are (sample, position) with R positions per sample and NB = 128 / R
// samples per tile; the K-blocks walk taps t (outer) x 32-wide slices of the reduced
// channel axis (inner).  kBMn: fprop (B = kernel as [k rows][g]) ; !kBMn: dgrad (B rows c).
template <bool kBMn_>
struct ConvRowsProb {
  static constexpr bool kAMn = false, kBMn = kBMn_;
  int num_samples;         // N
  int R;                   // positions per sample in the OUTPUT (OW for fprop, W for dgrad)
  int nb;                  // samples per tile
  int taps;                // KW
  int inner_blocks;        // ceil(reduced channels / 32)
  int a_w0, a_wstep;       // A position coordinate = a_w0 + a_wstep * t
  int out_maps;            // GEMM N: G (fprop) or C (dgrad)
  float *out;              // [N][ldo], sample rows hold [map][R]
  int ldo;
  const float *bias;       // per map or nullptr
  FastDiv div_inner, div_r;

  __device__ __forceinline__ void kb_range(int &b, int &e) const { b = 0; e = taps * inner_blocks; }
  __device__ __forceinline__ uint32_t tx_bytes() const { return (uint32_t)(nb * R * 128 + B_STAGE_BYTES); }
  __device__ __forceinline__ void load(int kb, uint32_t a_addr, uint32_t b_addr, uint32_t bar,
                                       const CUtensorMap *ma, const CUtensorMap *mb) const {
    uint32_t t, ib;
    div_inner.divmod((uint32_t)kb, t, ib);
    const int i0 = (int)ib * 32;
    const int n0 = blockIdx.x * nb, col0 = blockIdx.y * BN;
    tma_load_3d(a_addr, ma, i0, a_w0 + a_wstep * (int)t, n0, bar);
    if (kBMn) {
#pragma unroll
      for (int i = 0; i < BN / 32; i++) tma_load_3d(b_addr + i * ATOM_BYTES, mb, col0 + 32 * i, (int)t, i0, bar);
    } else {
      tma_load_3d(b_addr, mb, i0, (int)t, col0, bar);
    }
  }
  // stage[(s*R + pos)][map] -> out[n0 + s][(col0 + map) * R + pos]: per sample one contiguous run
  __device__ __forceinline__ void store(const float *stage, int tid) const {
    const int n0 = blockIdx.x * nb, col0 = blockIdx.y * BN;
    const int maps = min(BN, out_maps - col0);
    const int total = maps * R;
    for (int s = 0; s < nb; s++) {
      const int n = n0 + s;
      if (n >= num_samples) break;
      float *orow = out + (size_t)n * ldo + (size_t)col0 * R;
      const float *srow = stage + s * R * PITCH;
      for (int idx = tid; idx < total; idx += 128) {
        uint32_t g, pos;
        div_r.divmod((uint32_t)idx, g, pos);
        float v = srow[pos * PITCH + g];
        if (bias) v += __ldg(bias + col0 + g);
        orow[idx] = v;
      }
    }
  }
};

// --------------------------------------------------------------- wgrad problem --

// GEMM row m = kw * C + c  ->  kernel-matrix row c * KW + kw (folds ModPermuteRow)
struct KernelRow {
  FastDiv div_c;
  int KW;
  __device__ __forceinline__ int operator()(int m) const {
    uint32_t kw, c;
    div_c.divmod((uint32_t)m, kw, c);
    return (int)c * KW + (int)kw;
  }
};

template <int kEpi>
struct ConvWgradProb {
  static constexpr bool kAMn = true, kBMn = true;
  int C, KW, G, M;         // M = KW * C, rows m = kw * C + c
  int pw;
  int n_blocks;            // ceil(N / 32): K-blocks per output position
  int total_kb;            // OW * n_blocks
  int kb_per_split;
  float *out;              // kernel-shaped [C*KW][ldo]: row c*KW + kw
  int ldo;
  float *workspace;        // [splits][M][G]
  float *aux;              // EPI_SGD: prev_grad (same shape as out)
  SgdCoef sgd;
  FastDiv div_nb, div_c;

  __device__ __forceinline__ void kb_range(int &b, int &e) const {
    b = blockIdx.z * kb_per_split;
    e = min(total_kb, b + kb_per_split);
    if (e < b) e = b;
  }
  __device__ __forceinline__ uint32_t tx_bytes() const { return STAGE_BYTES; }
  __device__ __forceinline__ void load(int kb, uint32_t a_addr, uint32_t b_addr, uint32_t bar,
                                       const CUtensorMap *ma, const CUtensorMap *mb) const {
    uint32_t ow, nblk;
    div_nb.divmod((uint32_t)kb, ow, nblk);
    const int n0 = (int)nblk * 32;
    const int m0 = blockIdx.x * BM, g0 = blockIdx.y * BN;
#pragma unroll
    for (int i = 0; i < BM / 32; i++) {
      uint32_t kw, c;
      div_c.divmod((uint32_t)(m0 + 32 * i), kw, c);
      tma_load_3d(a_addr + i * ATOM_BYTES, ma, (int)c, (int)ow + (int)kw - pw, n0, bar);
    }
#pragma unroll
    for (int i = 0; i < BN / 32; i++) tma_load_3d(b_addr + i * ATOM_BYTES, mb, g0 + 32 * i, (int)ow, n0, bar);
  }
  __device__ __forceinline__ void store(const float *stage, int tid) const {
    const int m0 = blockIdx.x * BM, g0 = blockIdx.y * BN;
    if (kEpi == EPI_PARTIAL) {
      store_rows<EPI_STORE>(stage, tid, m0, g0, M, G, workspace + (size_t)blockIdx.z * M * G, G, nullptr,
                            nullptr, sgd, IdentityRow());
    } else {
      KernelRow rm; rm.div_c = div_c; rm.KW = KW;
      store_rows<kEpi>(stage, tid, m0, g0, M, G, out, ldo, nullptr, aux, sgd, rm);
    }
  }
};

// ------------------------------------------------------------------- host side --

struct ConvShape {
  int N, W, C, pw, KW, G, OW;     // in_height = kernel_height = 1
};

// Shapes the TMA convolution path takes.  Everything else (2-D kernels, tiny channel
// counts, unaligned views) stays on the software-producer kernel.
inline bool conv_tma_shape_ok(const ConvShape &q) {
  if (!enabled()) return false;
  if (q.N <= 0 || q.OW <= 0 || q.OW > 128 || q.W > 128) return false;
  if (q.C < 32 || (q.C & 31) != 0) return false;          // 32-channel K slices / M atoms
  if (q.G < 32 || (q.G & 3) != 0) return false;
  if (pack_smem_bytes(q.C, q.W) > (size_t)kPackMaxSmem || pack_smem_bytes(q.G, q.OW) > (size_t)kPackMaxSmem)
    return false;
  return true;
}

// 3-D view (inner, taps, outer) of the kernel matrix [C*KW rows][G], row = c*KW + kw:
// dims (g, kw, c).
inline bool encode_kernel_map(CUtensorMap *map, const float *kernel, int ld, int C, int KW, int G,
                              unsigned box_c, bool mn_major) {
  unsigned long long dims[3] = {(unsigned long long)G, (unsigned long long)KW, (unsigned long long)C};
  unsigned long long str[2] = {(unsigned long long)ld * 4, (unsigned long long)ld * 4 * KW};
  unsigned box[3] = {32, 1, box_c};
  return encode_map(map, kernel, 3, dims, str, box, mn_major);
}

// channels-last activation [N][R][Cn] as dims (c, r, n)
inline bool encode_act_map(CUtensorMap *map, const float *act, int N, int R, int Cn, unsigned box_r,
                           unsigned box_n, bool mn_major) {
  unsigned long long dims[3] = {(unsigned long long)Cn, (unsigned long long)R, (unsigned long long)N};
  unsigned long long str[2] = {(unsigned long long)Cn * 4, (unsigned long long)Cn * 4 * R};
  unsigned box[3] = {32, box_r, box_n};
  return encode_map(map, act, 3, dims, str, box, mn_major);
}

inline bool conv_fprop(cudaStream_t st, const ConvShape &q, const float *in, int ld_in, const float *kernel,
                       int ld_k, const float *bias, float *out, int ldo) {
  if (!conv_tma_shape_ok(q)) return false;
  float *xcl = scratch(SCRATCH_XCL, (size_t)q.N * q.W * q.C * sizeof(float));
  if (!xcl) return false;
  const int nb = 128 / q.OW;
  CUtensorMap ma, mb;
  if (!encode_act_map(&ma, xcl, q.N, q.W, q.C, (unsigned)q.OW, (unsigned)nb, false)) return false;
  if (!encode_kernel_map(&mb, kernel, ld_k, q.C, q.KW, q.G, 32, true)) return false;
  launch_pack(st, in, ld_in, q.N, q.C, q.W, xcl, nullptr);
  ConvRowsProb<true> p;
  p.num_samples = q.N; p.R = q.OW; p.nb = nb; p.taps = q.KW; p.inner_blocks = (q.C + 31) / 32;
  p.a_w0 = -q.pw; p.a_wstep = 1; p.out_maps = q.G; p.out = out; p.ldo = ldo; p.bias = bias;
  p.div_inner = FastDiv((uint32_t)p.inner_blocks); p.div_r = FastDiv((uint32_t)q.OW);
  launch_prob(st, ma, mb, p, dim3(ceil_div_u(q.N, nb), ceil_div_u(q.G, BN), 1));
  return true;
}

inline bool conv_dgrad(cudaStream_t st, const ConvShape &q, const float *out_deriv, int ld_od,
                       const float *kernel, int ld_k, float *in_deriv, int ld_id) {
  if (!conv_tma_shape_ok(q)) return false;
  float *dycl = scratch(SCRATCH_DYCL, (size_t)q.N * q.OW * q.G * sizeof(float));
  if (!dycl) return false;
  const int nb = 128 / q.W;
  CUtensorMap ma, mb;
  if (!encode_act_map(&ma, dycl, q.N, q.OW, q.G, (unsigned)q.W, (unsigned)nb, false)) return false;
  if (!encode_kernel_map(&mb, kernel, ld_k, q.C, q.KW, q.G, 128, false)) return false;
  launch_pack(st, out_deriv, ld_od, q.N, q.G, q.OW, dycl, nullptr);
  ConvRowsProb<false> p;
  p.num_samples = q.N; p.R = q.W; p.nb = nb; p.taps = q.KW; p.inner_blocks = (q.G + 31) / 32;
  p.a_w0 = q.pw; p.a_wstep = -1; p.out_maps = q.C; p.out = in_deriv; p.ldo = ld_id; p.bias = nullptr;
  p.div_inner = FastDiv((uint32_t)p.inner_blocks); p.div_r = FastDiv((uint32_t)q.W);
  launch_prob(st, ma, mb, p, dim3(ceil_div_u(q.N, nb), ceil_div_u(q.C, BN), 1));
  return true;
}

// Weight gradient into kernel_grad (sgd == nullptr) or, with sgd, straight into the
// update prev = m prev - lr wd K + lr dK ; K += prev (kernel_grad = K, prev = aux).
// bias_partial (optional): [ceil(N / kPackSamples)][G] first-stage column sums of dY; the
// number of rows is returned through bias_rows.
inline bool conv_wgrad(cudaStream_t st, const ConvShape &q, const float *in_value, int ld_iv,
                       const float *out_deriv, int ld_od, float *kernel_grad, int ld_kg, float *prev,
                       const SgdCoef *sgd, float **bias_partial, int *bias_rows) {
  if (!conv_tma_shape_ok(q)) return false;
  if ((ld_kg & 3) != 0 || !host_aligned16(kernel_grad) || (prev && !host_aligned16(prev))) return false;
  float *xcl = scratch(SCRATCH_XCL, (size_t)q.N * q.W * q.C * sizeof(float));
  float *dycl = scratch(SCRATCH_DYCL, (size_t)q.N * q.OW * q.G * sizeof(float));
  const int pack_ctas = (q.N + kPackSamples - 1) / kPackSamples;
  float *bpart = bias_partial ? scratch(SCRATCH_BIAS, (size_t)pack_ctas * q.G * sizeof(float)) : nullptr;
  if (!xcl || !dycl || (bias_partial && !bpart)) return false;
  const int M = q.KW * q.C;
  const int n_blocks = (q.N + 31) / 32;
  const int total_kb = q.OW * n_blocks;
  const long long tiles = (long long)ceil_div_u(M, BM) * ceil_div_u(q.G, BN);
  int splits = pick_splits(tiles, total_kb);
  int per = (total_kb + splits - 1) / splits;
  splits = (total_kb + per - 1) / per;
  float *ws = nullptr;
  if (splits > 1) {
    ws = scratch(SCRATCH_SPLITK, (size_t)splits * M * q.G * sizeof(float));
    if (!ws) return false;
  }
  CUtensorMap ma, mb;
  if (!encode_act_map(&ma, xcl, q.N, q.W, q.C, 1, 32, true)) return false;
  if (!encode_act_map(&mb, dycl, q.N, q.OW, q.G, 1, 32, true)) return false;
  launch_pack(st, in_value, ld_iv, q.N, q.C, q.W, xcl, nullptr);
  launch_pack(st, out_deriv, ld_od, q.N, q.G, q.OW, dycl, bpart);
  if (bias_partial) { *bias_partial = bpart; *bias_rows = pack_ctas; }
  dim3 grid(ceil_div_u(M, BM), ceil_div_u(q.G, BN), splits);
  SgdCoef none = {0.f, 0.f, 0.f};
  auto fill = [&](auto &p) {
    p.C = q.C; p.KW = q.KW; p.G = q.G; p.M = M; p.pw = q.pw; p.n_blocks = n_blocks; p.total_kb = total_kb;
    p.kb_per_split = per; p.out = kernel_grad; p.ldo = ld_kg; p.workspace = ws; p.aux = prev;
    p.sgd = sgd ? *sgd : none;
    p.div_nb = FastDiv((uint32_t)n_blocks); p.div_c = FastDiv((uint32_t)q.C);
  };
  KernelRow rm;
  rm.div_c = FastDiv((uint32_t)q.C); rm.KW = q.KW;
  if (splits > 1) {
    ConvWgradProb<EPI_PARTIAL> p; fill(p);
    launch_prob(st, ma, mb, p, grid);
    const unsigned blocks = ceil_div_u(((long long)M * q.G) >> 2, 256);
    if (sgd)
      KCNN_LAUNCH((splitk_reduce_kernel<EPI_SGD, KernelRow>), blocks, 256, 0,
                  st, ws, splits, M, q.G, kernel_grad, ld_kg, nullptr, prev, *sgd, rm);
    else
      KCNN_LAUNCH((splitk_reduce_kernel<EPI_STORE, KernelRow>), blocks, 256, 0,
                  st, ws, splits, M, q.G, kernel_grad, ld_kg, nullptr, nullptr, none, rm);
  } else if (sgd) {
    ConvWgradProb<EPI_SGD> p; fill(p);
    launch_prob(st, ma, mb, p, grid);
  } else {
    ConvWgradProb<EPI_STORE> p; fill(p);
    launch_prob(st, ma, mb, p, grid);
  }
  return true;
}

}  // namespace tma
}  // namespace kcnn

#endif
