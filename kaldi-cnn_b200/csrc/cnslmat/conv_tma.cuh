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
// CTA (bx, by) moves channels [by*64, by*64+64) of samples [bx*spc, bx*spc+spc) through
// shared memory in one pass (row pitch R|1: conflict-free both ways); reads are the
// contiguous 64*R-float run of each sample, writes are 256-byte channel runs.  With kColSum
// it also writes partial[bx][c] = sum over its samples and r of in[n][c][r] -- the first stage
// of the bias gradient (reference nnet0/nnet-component-nnet0.cc:775), deterministic.
constexpr int kPackCh = 64;
constexpr int kPackMaxSamples = 4;
constexpr int kPackSmemBudget = 44 * 1024;

template <bool kColSum, bool kVec4>
__global__ void __launch_bounds__(256)
pack_channels_last_kernel(const float *__restrict__ in, int ld, int N, int C, int R, float *__restrict__ out,
                          float *__restrict__ partial, int spc, FastDiv div_r, FastDiv div_ch) {
  kcnn::pdl_prologue();
  extern __shared__ float tile[];
  const int rp = R | 1;
  const int c0 = blockIdx.y * kPackCh;
  const int ch = min(kPackCh, C - c0);
  const int n_begin = blockIdx.x * spc;
  const int ns = min(spc, N - n_begin);
  const int per = ch * R;                       // elements per sample in this CTA
  const int sample_tile = kPackCh * rp;
  if (kVec4) {
    // 128-bit global accesses both ways (host checked: ch == 64, 16-byte aligned rows)
    const int per4 = per >> 2;
    for (int s = 0; s < ns; s++) {
      const float4 *src = reinterpret_cast<const float4 *>(in + (size_t)(n_begin + s) * ld + (size_t)c0 * R);
      float *ts = tile + s * sample_tile;
      for (int i4 = threadIdx.x; i4 < per4; i4 += 256) {
        const float4 v = __ldg(src + i4);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          uint32_t c, r;
          div_r.divmod((uint32_t)(4 * i4 + k), c, r);
          ts[c * rp + r] = e[k];
        }
      }
    }
    __syncthreads();
    for (int s = 0; s < ns; s++) {
      float *dst = out + (size_t)(n_begin + s) * R * C + c0;
      const float *ts = tile + s * sample_tile;
      for (int i4 = threadIdx.x; i4 < R * (kPackCh / 4); i4 += 256) {
        const int r = i4 >> 4, c = (i4 & 15) << 2;           // 16 float4 per position
        const float *t4 = ts + c * rp + r;
        *reinterpret_cast<float4 *>(dst + (size_t)r * C + c) = make_float4(t4[0], t4[rp], t4[2 * rp], t4[3 * rp]);
      }
    }
  } else {
    for (int s = 0; s < ns; s++) {
      const float *src = in + (size_t)(n_begin + s) * ld + (size_t)c0 * R;
      float *ts = tile + s * sample_tile;
      for (int i = threadIdx.x; i < per; i += 256) {
        uint32_t c, r;
        div_r.divmod((uint32_t)i, c, r);
        ts[c * rp + r] = __ldg(src + i);
      }
    }
    __syncthreads();
    for (int s = 0; s < ns; s++) {
      float *dst = out + (size_t)(n_begin + s) * R * C + c0;
      const float *ts = tile + s * sample_tile;
      for (int i = threadIdx.x; i < R * kPackCh; i += 256) {
        uint32_t r, c;
        div_ch.divmod((uint32_t)i, r, c);
        if ((int)c < ch) dst[(size_t)r * C + c] = ts[c * rp + r];
      }
    }
  }
  if (kColSum && (int)threadIdx.x < ch) {
    float acc = 0.f;
    for (int s = 0; s < ns; s++) {
      const float *tc = tile + s * sample_tile + threadIdx.x * rp;
      for (int r = 0; r < R; r++) acc += tc[r];
    }
    partial[(size_t)blockIdx.x * C + c0 + threadIdx.x] = acc;
  }
}

inline int pack_samples_per_cta(int R) {
  int spc = kPackSmemBudget / (kPackCh * (R | 1) * (int)sizeof(float));
  if (spc > kPackMaxSamples) spc = kPackMaxSamples;
  return spc;                                   // 0: R too large for the pack kernel
}

// Rows of partial sums the pack writes for `N` samples with `R` positions.
inline int pack_partial_rows(int N, int R) {
  int spc = pack_samples_per_cta(R);
  return spc > 0 ? (N + spc - 1) / spc : 0;
}

inline void launch_pack(cudaStream_t st, const float *in, int ld, int N, int C, int R, float *out,
                        float *colsum_partial) {
  const int spc = pack_samples_per_cta(R);
  const size_t smem = (size_t)spc * kPackCh * (R | 1) * sizeof(float);
  dim3 grid((N + spc - 1) / spc, (C + kPackCh - 1) / kPackCh);
  const bool v4 = (C % kPackCh) == 0 && (ld & 3) == 0 && host_aligned16(in) && host_aligned16(out);
  const FastDiv dr((uint32_t)R), dc((uint32_t)kPackCh);
  if (colsum_partial) {
    if (v4) KCNN_LAUNCH((pack_channels_last_kernel<true, true>), grid, 256, smem, st, in, ld, N, C, R, out,
                        colsum_partial, spc, dr, dc);
    else    KCNN_LAUNCH((pack_channels_last_kernel<true, false>), grid, 256, smem, st, in, ld, N, C, R, out,
                        colsum_partial, spc, dr, dc);
  } else {
    if (v4) KCNN_LAUNCH((pack_channels_last_kernel<false, true>), grid, 256, smem, st, in, ld, N, C, R, out,
                        nullptr, spc, dr, dc);
    else    KCNN_LAUNCH((pack_channels_last_kernel<false, false>), grid, 256, smem, st, in, ld, N, C, R, out,
                        nullptr, spc, dr, dc);
  }
}

// ------------------------------------------------------- fprop / dgrad problem --

// stage[(s*R + pos)][map] -> out[n0 + s][(col0 + map) * R + pos]: per sample ONE contiguous run of
// maps * R floats (the reference's [map][pos] layout), written 512 bytes per warp instruction.
// Only tile rows [rb, re) are written (the whole tile, or a cluster split-K slice).
__device__ __forceinline__ void store_maps_transposed(const float *stage, int tid, int n0, int nb, int R,
                                                      int col0, int maps, int num_samples, float *out, int ldo,
                                                      const float *bias, const FastDiv &div_r, bool relu,
                                                      int rb, int re) {
  const int total = maps * R;
  for (int s = 0; s < nb; s++) {
    const int n = n0 + s;
    if (n >= num_samples) break;
    const int p0 = max(0, rb - s * R), p1 = min(R, re - s * R);
    if (p0 >= p1) continue;
    float *orow = out + (size_t)n * ldo + (size_t)col0 * R;
    const float *srow = stage + s * R * PITCH;
    for (int idx = tid; idx < total; idx += 128) {
      uint32_t g, pos;
      div_r.divmod((uint32_t)idx, g, pos);
      if ((int)pos < p0 || (int)pos >= p1) continue;
      float v = srow[pos * PITCH + g];
      if (bias) v += __ldg(bias + col0 + g);
      if (relu) v = v > 0.0f ? v : 0.0f;
      orow[idx] = v;
    }
  }
}

// Rows of the GEMM are (sample, position) with R positions per sample and NB = 128 / R
// samples per tile; the K-blocks walk taps t (outer) x 32-wide slices of the reduced
// channel axis (inner).  kBMn: fprop (B = kernel as [k rows][g]) ; !kBMn: dgrad (B rows c).
// blockIdx.z owns K-blocks [z*kb_per_split, ...): the splits of a tile are one cluster (gemm_tma.cuh,
// kClusterK), used when the tile grid alone would leave most SMs idle (conv5 / conv6 of nnet.config:
// 32 - 64 tiles with 24 - 48 K-blocks each).
// Output layout: out_cl = 0 the reference's rows [map][R] (the tile is transposed on the way out: folds
// _convmat_to_out / TpBlock); out_cl = 1 channels-last [N][R][maps], which IS the tile, row-major --
// the layout the next convolution's tensor maps read, so activations and derivatives flow through a
// stack of time-axis layers without any staging copy (NnetMinibatchUpdater's fused step).
template <bool kBMn_>
struct ConvRowsProb {
  static constexpr bool kAMn = false, kBMn = kBMn_;
  int kb_per_split;
  int num_samples;         // N
  int R;                   // positions per sample in the OUTPUT (OW for fprop, W for dgrad)
  int nb;                  // samples per tile
  int taps;                // KW
  int inner_blocks;        // ceil(reduced channels / 32)
  int a_w0, a_wstep;       // A position coordinate = a_w0 + a_wstep * t
  int out_maps;            // GEMM N: G (fprop) or C (dgrad)
  float *out;              // out_cl ? [N][R][out_maps] : [N][ldo] rows holding [map][R]
  int ldo;
  int out_cl;
  RowsEpi epi;             // bias / ReLU (fprop), ReLU mask of the producer (dgrad)
  FastDiv div_inner, div_r;

  __device__ __forceinline__ void kb_range(int z, int &b, int &e) const {
    const int total = taps * inner_blocks;
    b = z * kb_per_split;
    e = min(total, b + kb_per_split);
    if (e < b) e = b;
  }
  __device__ __forceinline__ uint32_t tx_bytes() const { return (uint32_t)(nb * R * 128 + B_STAGE_BYTES); }
  template <bool kPair>
  __device__ __forceinline__ void load(int kb, uint32_t a_addr, uint32_t b_addr, uint32_t bar,
                                       const CUtensorMap *ma, const CUtensorMap *mb, int mt, int nt) const {
    uint32_t t, ib;
    div_inner.divmod((uint32_t)kb, t, ib);
    const int i0 = (int)ib * 32;
    const int n0 = mt * nb, col0 = nt * BN;
    tma_load_3d<kPair>(a_addr, ma, i0, a_w0 + a_wstep * (int)t, n0, bar);
    if (kBMn) {
#pragma unroll
      for (int i = 0; i < BN / 32; i++) tma_load_3d<kPair>(b_addr + i * ATOM_BYTES, mb, col0 + 32 * i, (int)t, i0, bar);
    } else {
      tma_load_3d<kPair>(b_addr, mb, i0, (int)t, col0, bar);
    }
  }
  __device__ __forceinline__ void prefetch(int tid, int mt, int nt) const {
    if (out_cl && epi.mask_x != nullptr)
      prefetch_tile_l2(tid, mt * nb * R, nt * BN, min(num_samples * R, (mt + 1) * nb * R), out_maps, epi.mask_x,
                       epi.mask_x, epi.ld_mx, IdentityRow());
  }
  template <int kRows>
  __device__ __forceinline__ void store(const float *stage, int tid, int mt, int nt, int z, int rb, int re) const {
    const int col0 = nt * BN;
    if (col0 >= out_maps) return;
    if (out_cl) {
      SgdCoef none = {0.f, 0.f, 0.f};
      store_rows<EPI_STORE, kRows>(stage, tid, mt * nb * R, col0, num_samples * R, out_maps, out, ldo, nullptr, none,
                                   IdentityRow(), rb, min(re, nb * R), epi);
    } else {
      store_maps_transposed(stage, tid, mt * nb, nb, R, col0, min(BN, out_maps - col0), num_samples, out, ldo,
                            epi.bias_n, div_r, epi.relu != 0, rb, re);
    }
  }
};

// ------------------------------------------- full-height kernels (KH = H, OH = 1) --
//
// conv1 of nnet.config (40 x 21 x 1, kernel 40 x 4) and C1a: the (kw, kh) window of output
// column ow is the CONTIGUOUS run X[n, c*H*W + ow*H ... + KW*H), so the im2col matrix is a
// 4-D tensor map over the input itself -- dims (j < KW*KH, ow [pitch H], c [pitch H*W],
// n [row pitch]) with overlapping rows -- and the forward pass needs no staging copy at all.
// (No zero padding on this path: the window position is folded into the address.)

// fprop: A box {32 j, OW, 1, NB} at (j0, 0, c, n0) K-major; B = kernel rows c*ks + j, MN-major.
struct ConvFullFpropProb {
  static constexpr bool kAMn = false, kBMn = true;
  int num_samples, OW, nb, ks, j_blocks, G;
  float *out;
  int ldo;
  int out_cl;              // as ConvRowsProb
  RowsEpi epi;
  FastDiv div_jb, div_ow;

  int total_kb;            // C * j_blocks
  __device__ __forceinline__ void kb_range(int, int &b, int &e) const { b = 0; e = total_kb; }
  __device__ __forceinline__ uint32_t tx_bytes() const { return (uint32_t)(nb * OW * 128 + B_STAGE_BYTES); }
  template <bool kPair>
  __device__ __forceinline__ void load(int kb, uint32_t a_addr, uint32_t b_addr, uint32_t bar,
                                       const CUtensorMap *ma, const CUtensorMap *mb, int mt, int nt) const {
    uint32_t c, jb;
    div_jb.divmod((uint32_t)kb, c, jb);
    const int j0 = (int)jb * 32;
    const int n0 = mt * nb, g0 = nt * BN;
    tma_load_4d<kPair>(a_addr, ma, j0, 0, (int)c, n0, bar);
#pragma unroll
    for (int i = 0; i < BN / 32; i++) tma_load_2d<kPair>(b_addr + i * ATOM_BYTES, mb, g0 + 32 * i, (int)c * ks + j0, bar);
  }
  __device__ __forceinline__ void prefetch(int, int, int) const {}
  template <int kRows>
  __device__ __forceinline__ void store(const float *stage, int tid, int mt, int nt, int z, int rb, int re) const {
    const int g0 = nt * BN;
    if (g0 >= G) return;
    if (out_cl) {
      SgdCoef none = {0.f, 0.f, 0.f};
      store_rows<EPI_STORE, kRows>(stage, tid, mt * nb * OW, g0, num_samples * OW, G, out, ldo, nullptr, none,
                                   IdentityRow(), rb, min(re, nb * OW), epi);
    } else {
      store_maps_transposed(stage, tid, mt * nb, nb, OW, g0, min(BN, G - g0), num_samples, out, ldo, epi.bias_n,
                            div_ow, epi.relu != 0, rb, re);
    }
  }
};

// dgrad: dX[(n,w),(c,h)] = sum_{kw,g} dYcl[n, w-kw, g] K[c*ks + kw*KH + h, g]  (kh = h is fixed by
// the row of the image, so it moves from the reduction into the GEMM's N axis).
// A box {32 g, W, NB} of dYcl at (g0, -kw, n0); B box {32 g, KH, NBC channels} of the kernel
// viewed as (g, r < ks, c) at (g0, kw*KH, c0), K-major.  The tile rows (s, w) x columns
// (cb, h) ARE the reference layout [c][w][h]: each (sample, channel) is one contiguous run.
struct ConvFullDgradProb {
  static constexpr bool kAMn = false, kBMn = false;
  int num_samples, W, H, C, nb, nbc, g_blocks, total_kb;
  float *out;
  int ldo;
  FastDiv div_gb, div_h;

  __device__ __forceinline__ void kb_range(int, int &b, int &e) const { b = 0; e = total_kb; }
  __device__ __forceinline__ uint32_t tx_bytes() const { return (uint32_t)((nb * W + nbc * H) * 128); }
  template <bool kPair>
  __device__ __forceinline__ void load(int kb, uint32_t a_addr, uint32_t b_addr, uint32_t bar,
                                       const CUtensorMap *ma, const CUtensorMap *mb, int mt, int nt) const {
    uint32_t kw, gb;
    div_gb.divmod((uint32_t)kb, kw, gb);
    const int g0 = (int)gb * 32;
    tma_load_3d<kPair>(a_addr, ma, g0, -(int)kw, mt * nb, bar);
    tma_load_3d<kPair>(b_addr, mb, g0, (int)kw * H, nt * nbc, bar);
  }
  __device__ __forceinline__ void prefetch(int, int, int) const {}
  template <int kRows>
  __device__ __forceinline__ void store(const float *stage, int tid, int mt, int nt, int z, int, int) const {
    const int n0 = mt * nb, c0 = nt * nbc;
    const int chans = min(nbc, C - c0);
    const int hw = H * W;
    for (int s = 0; s < nb; s++) {
      const int n = n0 + s;
      if (n >= num_samples) break;
      for (int cb = 0; cb < chans; cb++) {
        float *orow = out + (size_t)n * ldo + (size_t)(c0 + cb) * hw;
        const float *sbase = stage + s * W * PITCH + cb * H;
        for (int idx = tid; idx < hw; idx += 128) {
          uint32_t w, h;
          div_h.divmod((uint32_t)idx, w, h);
          orow[idx] = sbase[w * PITCH + h];
        }
      }
    }
  }
};

// --------------------------------------------------------------- wgrad problem --

// GEMM row m = kw * Cp + c (Cp = C rounded up to 32, so a 32-row operand atom never straddles
// two taps)  ->  kernel-matrix row c * KW + kw (folds ModPermuteRow); rows c >= C are padding.
struct KernelRow {
  FastDiv div_c;        // by Cp
  int KW, C;
  __device__ __forceinline__ int operator()(int m) const {
    uint32_t kw, c;
    div_c.divmod((uint32_t)m, kw, c);
    return (int)c < C ? (int)c * KW + (int)kw : -1;
  }
};

// kFull: full-height kernels, A straight from the input through the 4-D window map, rows m in
// kernel order c*ks + j (ks % 32 == 0); div_c then divides by ks.
// kEpi = EPI_SGD: the momentum / weight-decay step runs on the (cluster-reduced) tile, the gradient
// is never written; EPI_STORE: dK goes to `out` (data-parallel mode, gradient checks).
template <int kEpi, bool kFull = false>
struct ConvWgradProb {
  static constexpr bool kAMn = true, kBMn = true;
  int C, KW, G, M;         // M = KW * Cp, rows m = kw * Cp + c   (kFull: M = C * ks, kernel order)
  int pw;
  int n_blocks;            // ceil(N / 32): K-blocks per output position
  int total_kb;            // OW * n_blocks
  int kb_per_split;
  float *out;              // kernel-shaped [C*KW][ldo]: row c*KW + kw
  int ldo;
  float *aux;              // EPI_SGD: prev_grad (same shape as out)
  SgdCoef sgd;
  FastDiv div_nb, div_c;

  __device__ __forceinline__ void kb_range(int z, int &b, int &e) const {
    b = z * kb_per_split;
    e = min(total_kb, b + kb_per_split);
    if (e < b) e = b;
  }
  __device__ __forceinline__ uint32_t tx_bytes() const { return STAGE_BYTES; }
  template <bool kPair>
  __device__ __forceinline__ void load(int kb, uint32_t a_addr, uint32_t b_addr, uint32_t bar,
                                       const CUtensorMap *ma, const CUtensorMap *mb, int mt, int nt) const {
    uint32_t ow, nblk;
    div_nb.divmod((uint32_t)kb, ow, nblk);
    const int n0 = (int)nblk * 32;
    const int m0 = mt * BM, g0 = nt * BN;
#pragma unroll
    for (int i = 0; i < BM / 32; i++) {
      uint32_t q, r;
      div_c.divmod((uint32_t)(m0 + 32 * i), q, r);
      if (kFull) tma_load_4d<kPair>(a_addr + i * ATOM_BYTES, ma, (int)r, (int)ow, (int)q, n0, bar);      // (j, ow, c, n)
      else       tma_load_3d<kPair>(a_addr + i * ATOM_BYTES, ma, (int)r, (int)ow + (int)q - pw, n0, bar); // (c, w, n)
    }
#pragma unroll
    for (int i = 0; i < BN / 32; i++) tma_load_3d<kPair>(b_addr + i * ATOM_BYTES, mb, g0 + 32 * i, (int)ow, n0, bar);
  }
  __device__ __forceinline__ void prefetch(int tid, int mt, int nt) const {
    if (kEpi == EPI_SGD) {
      if (kFull) {
        prefetch_tile_l2(tid, mt * BM, nt * BN, M, G, out, aux, ldo, IdentityRow());
      } else {
        KernelRow rm; rm.div_c = div_c; rm.KW = KW; rm.C = C;
        prefetch_tile_l2(tid, mt * BM, nt * BN, M, G, out, aux, ldo, rm);
      }
    }
  }
  template <int kRows>
  __device__ __forceinline__ void store(const float *stage, int tid, int mt, int nt, int z, int rb, int re) const {
    const int m0 = mt * BM, g0 = nt * BN;
    RowsEpi none = {};
    none.vec = 1;
    if (kFull) {
      store_rows<kEpi, kRows>(stage, tid, m0, g0, M, G, out, ldo, aux, sgd, IdentityRow(), rb, re, none);
    } else {
      KernelRow rm; rm.div_c = div_c; rm.KW = KW; rm.C = C;
      store_rows<kEpi, kRows>(stage, tid, m0, g0, M, G, out, ldo, aux, sgd, rm, rb, re, none);
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
  if (q.C < 8 || (q.C & 3) != 0) return false;            // 16-byte pitch of the channels-last copy
  if (q.G < 32 || (q.G & 3) != 0) return false;
  if (pack_samples_per_cta(q.W) < 1 || pack_samples_per_cta(q.OW) < 1) return false;
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

// Where a forward / input-gradient GEMM puts its result.
struct ConvOut {
  float *out;
  int ldo;                 // cl: unused (dense); else the row pitch of the reference-layout matrix
  bool cl;                 // channels-last [N][R][maps] instead of rows [map][R]
  const float *bias;       // per map or nullptr
  bool relu;
  const float *mask;       // cl only: same shape as out; out = mask > 0 ? out : 0 (the producer's ReLU backward)
};

template <bool kBMn>
inline void launch_conv_rows(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, ConvRowsProb<kBMn> p,
                             dim3 grid) {
  const int num_kb = p.taps * p.inner_blocks;
  int splits = pick_splits((long long)grid.x * grid.y, num_kb);
  const int per = (num_kb + splits - 1) / splits;
  splits = (num_kb + per - 1) / per;
  p.kb_per_split = per;
  launch_prob(st, ma, mb, p, dim3(grid.x, grid.y, splits), per);
}

template <class Prob>
inline void fill_conv_out(Prob &p, const ConvOut &o, int maps) {
  p.out = o.out; p.out_cl = o.cl ? 1 : 0; p.ldo = o.cl ? maps : o.ldo;
  p.epi = rows_epi_none();
  p.epi.bias_n = o.bias; p.epi.relu = o.relu ? 1 : 0;
  if (o.cl && o.mask) { p.epi.mask_x = o.mask; p.epi.ld_mx = maps; }
  p.epi.vec = rows_epi_vec_ok(p.epi) ? 1 : 0;
}

// Y[(n,ow), g] = sum_{kw,c} xcl[n, ow+kw-pw, c] K[(c,kw), g]   from a channels-last input [N][W][C]
inline bool conv_rows_fprop(cudaStream_t st, const ConvShape &q, const float *xcl, const float *kernel, int ld_k,
                            const ConvOut &o) {
  if (!conv_tma_shape_ok(q) || !host_aligned16(xcl)) return false;
  if (o.cl && !host_aligned16(o.out)) return false;
  const int nb = 128 / q.OW;
  CUtensorMap ma, mb;
  if (!encode_act_map(&ma, xcl, q.N, q.W, q.C, (unsigned)q.OW, (unsigned)nb, false)) return false;
  if (!encode_kernel_map(&mb, kernel, ld_k, q.C, q.KW, q.G, 32, true)) return false;
  ConvRowsProb<true> p;
  p.num_samples = q.N; p.R = q.OW; p.nb = nb; p.taps = q.KW; p.inner_blocks = (q.C + 31) / 32;
  p.a_w0 = -q.pw; p.a_wstep = 1; p.out_maps = q.G;
  fill_conv_out(p, o, q.G);
  p.div_inner = FastDiv((uint32_t)p.inner_blocks); p.div_r = FastDiv((uint32_t)q.OW);
  launch_conv_rows(st, ma, mb, p, dim3(ceil_div_u(q.N, nb), ceil_div_u(q.G, BN), 1));
  return true;
}

// dX[(n,w), c] = sum_{kw,g} dycl[n, w+pw-kw, g] K[(c,kw), g]   from a channels-last out_deriv [N][OW][G]
inline bool conv_rows_dgrad(cudaStream_t st, const ConvShape &q, const float *dycl, const float *kernel, int ld_k,
                            const ConvOut &o) {
  if (!conv_tma_shape_ok(q) || !host_aligned16(dycl)) return false;
  if (o.cl && !host_aligned16(o.out)) return false;
  const int nb = 128 / q.W;
  CUtensorMap da, db;
  if (!encode_act_map(&da, dycl, q.N, q.OW, q.G, (unsigned)q.W, (unsigned)nb, false)) return false;
  if (!encode_kernel_map(&db, kernel, ld_k, q.C, q.KW, q.G, 128, false)) return false;
  ConvRowsProb<false> p;
  p.num_samples = q.N; p.R = q.W; p.nb = nb; p.taps = q.KW; p.inner_blocks = (q.G + 31) / 32;
  p.a_w0 = q.pw; p.a_wstep = -1; p.out_maps = q.C;
  fill_conv_out(p, o, q.C);
  p.div_inner = FastDiv((uint32_t)p.inner_blocks); p.div_r = FastDiv((uint32_t)q.W);
  launch_conv_rows(st, da, db, p, dim3(ceil_div_u(q.N, nb), ceil_div_u(q.C, BN), 1));
  return true;
}

// Where a weight-gradient GEMM puts its result: sgd != nullptr -> the momentum / weight-decay step on
// (w = kernel, prev) in the epilogue, the gradient is never written; else w receives dK.
struct ConvWgradOut {
  float *w; int ld_w;
  float *prev; int ld_p;
  const SgdCoef *sgd;
};
inline bool conv_wgrad_out_ok(const ConvWgradOut &o) {
  if ((o.ld_w & 3) != 0 || !host_aligned16(o.w)) return false;
  if (o.sgd && (o.ld_p != o.ld_w || !host_aligned16(o.prev))) return false;
  return true;
}

template <bool kFull>
inline void launch_conv_wgrad(cudaStream_t st, const CUtensorMap &wa, const CUtensorMap &wb, const ConvWgradOut &o,
                              int C, int KW, int G, int M, int pw, int N, int OW, uint32_t div_c) {
  const int n_blocks = (N + 31) / 32;
  const int total_kb = OW * n_blocks;
  int splits = pick_splits((long long)ceil_div_u(M, BM) * ceil_div_u(G, BN), total_kb);
  const int per = (total_kb + splits - 1) / splits;
  splits = (total_kb + per - 1) / per;
  dim3 grid(ceil_div_u(M, BM), ceil_div_u(G, BN), splits);
  SgdCoef none = {0.f, 0.f, 0.f};
  auto fill = [&](auto &p) {
    p.C = C; p.KW = KW; p.G = G; p.M = M; p.pw = pw; p.n_blocks = n_blocks; p.total_kb = total_kb;
    p.kb_per_split = per; p.out = o.w; p.ldo = o.ld_w; p.aux = o.prev;
    p.sgd = o.sgd ? *o.sgd : none;
    p.div_nb = FastDiv((uint32_t)n_blocks); p.div_c = FastDiv(div_c);
  };
  if (o.sgd) {
    ConvWgradProb<EPI_SGD, kFull> p; fill(p);
    launch_prob(st, wa, wb, p, grid, per);
  } else {
    ConvWgradProb<EPI_STORE, kFull> p; fill(p);
    launch_prob(st, wa, wb, p, grid, per);
  }
}

// dK[(c,kw), g] = sum_{ow,n} xcl[n, ow+kw-pw, c] dycl[n, ow, g]
inline bool conv_rows_wgrad(cudaStream_t st, const ConvShape &q, const float *xcl, const float *dycl,
                            const ConvWgradOut &o) {
  if (!conv_tma_shape_ok(q) || !conv_wgrad_out_ok(o)) return false;
  const int Cp = (q.C + 31) & ~31;
  CUtensorMap wa, wb;
  if (!encode_act_map(&wa, xcl, q.N, q.W, q.C, 1, 32, true)) return false;
  if (!encode_act_map(&wb, dycl, q.N, q.OW, q.G, 1, 32, true)) return false;
  launch_conv_wgrad<false>(st, wa, wb, o, q.C, q.KW, q.G, q.KW * Cp, q.pw, q.N, q.OW, (uint32_t)Cp);
  return true;
}

// ---- the bare-component path: reference-layout matrices in and out, staging copies per call ----

// staging (optional): caller-owned [N*W*C] floats that receive the channels-last copy of `in`
// and stay valid after the call, so the matching Backprop can skip its own pack of in_value.
inline bool conv_fprop(cudaStream_t st, const ConvShape &q, const float *in, int ld_in, const float *kernel,
                       int ld_k, const float *bias, float *out, int ldo, float *staging = nullptr,
                       bool relu = false) {
  if (!conv_tma_shape_ok(q)) return false;
  float *xcl = staging ? staging : scratch(SCRATCH_XCL, (size_t)q.N * q.W * q.C * sizeof(float), st);
  if (!xcl || !host_aligned16(xcl)) return false;
  CUtensorMap probe;                              // nothing may fail after the pack has been launched
  if (!encode_kernel_map(&probe, kernel, ld_k, q.C, q.KW, q.G, 32, true)) return false;
  launch_pack(st, in, ld_in, q.N, q.C, q.W, xcl, nullptr);
  ConvOut o = {out, ldo, false, bias, relu, nullptr};
  return conv_rows_fprop(st, q, xcl, kernel, ld_k, o);
}

// Everything ConvolutionComponent::Backprop needs (reference nnet0/nnet-component-nnet0.cc:
// 461-544 + 738-777) from ONE channels-last copy of out_deriv and one of in_value:
//   in_deriv (optional)            dgrad GEMM
//   kernel gradient                wgrad GEMM (cluster split-K)
//   bias gradient                  column sums, first stage inside the out_deriv pack
// With sgd != nullptr the weight step  prev = m prev + a_decay K + a_grad dK ; K += prev  runs in
// the wgrad epilogue on (kernel, prev) and the gradient is never written; otherwise kernel_grad
// receives dK.  The bias partial sums are returned for the caller's final column-sum stage.
struct ConvBackward {
  const float *in_value; int ld_iv;
  const float *out_deriv; int ld_od;
  float *kernel; int ld_k;                // weights (read by dgrad; updated in place when sgd)
  float *in_deriv; int ld_id;             // nullptr: skip dgrad
  float *kernel_grad; int ld_kg;          // !sgd: receives dK;  nullptr: skip wgrad
  float *prev; int ld_p;                  // sgd: momentum state, same shape / pitch as kernel
  const SgdCoef *sgd;
  const float *staged_x;                  // optional: channels-last copy of in_value left by conv_fprop
  bool want_bias;
  float *bias_partial; int bias_rows;     // out: partial sums (the caller reduces them)
};

inline bool conv_backward(cudaStream_t st, const ConvShape &q, ConvBackward &b) {
  if (!conv_tma_shape_ok(q)) return false;
  const bool do_dgrad = b.in_deriv != nullptr;
  const bool do_wgrad = b.sgd != nullptr || b.kernel_grad != nullptr;
  ConvWgradOut wo = {b.sgd ? b.kernel : b.kernel_grad, b.sgd ? b.ld_k : b.ld_kg, b.prev, b.ld_p, b.sgd};
  if (do_wgrad && !conv_wgrad_out_ok(wo)) return false;
  float *dycl = scratch(SCRATCH_DYCL, (size_t)q.N * q.OW * q.G * sizeof(float), st);
  const bool have_x = b.staged_x != nullptr && host_aligned16(b.staged_x);
  float *xcl = !do_wgrad ? nullptr
               : have_x ? const_cast<float *>(b.staged_x)
                        : scratch(SCRATCH_XCL, (size_t)q.N * q.W * q.C * sizeof(float), st);
  const int prow = pack_partial_rows(q.N, q.OW);
  float *bpart = b.want_bias ? scratch(SCRATCH_BIAS, (size_t)prow * q.G * sizeof(float), st) : nullptr;
  if (!dycl || !host_aligned16(dycl) || (do_wgrad && (!xcl || !host_aligned16(xcl))) || (b.want_bias && !bpart))
    return false;
  CUtensorMap probe;
  if (!encode_kernel_map(&probe, b.kernel ? b.kernel : wo.w, b.kernel ? b.ld_k : wo.ld_w, q.C, q.KW, q.G, 128, false))
    return false;
  // ---- nothing can fail past this point
  launch_pack(st, b.out_deriv, b.ld_od, q.N, q.G, q.OW, dycl, bpart);
  b.bias_partial = bpart; b.bias_rows = prow;
  // Without the in-place SGD step (data-parallel mode) the input-gradient and the weight-gradient GEMM
  // are independent readers of the staging copy: two branches.  With it the weight gradient WRITES the
  // kernel matrix dgrad reads, so it runs behind dgrad.
  ForkJoin fj(st, do_dgrad && do_wgrad && !b.sgd);
  cudaStream_t wst = fj.active() ? fj.side() : st;
  if (do_dgrad) {
    ConvOut o = {b.in_deriv, b.ld_id, false, nullptr, false, nullptr};
    conv_rows_dgrad(st, q, dycl, b.kernel, b.ld_k, o);
  }
  if (!do_wgrad) return true;
  if (!have_x) launch_pack(wst, b.in_value, b.ld_iv, q.N, q.C, q.W, xcl, nullptr);
  conv_rows_wgrad(wst, q, xcl, dycl, wo);
  return true;
}

// ---- full-height kernels (KH = H, no padding): conv1 of nnet.config, C1a -----------------

struct ConvFullShape {
  int N, H, W, C, KW, G, OW;      // kernel_height = in_height, pads = 0
};

inline bool conv_full_shape_ok(const ConvFullShape &q) {
  if (!enabled()) return false;
  if (q.N <= 0 || q.OW <= 0 || q.OW > 128 || q.W > 128 || q.H > 128) return false;
  if ((q.H & 3) != 0 || q.G < 32) return false;
  return true;
}

// (j < KW*H, ow [pitch H], c [pitch H*W], n [row pitch]) over the input matrix itself
inline bool encode_window_map(CUtensorMap *map, const float *in, int ld, const ConvFullShape &q, unsigned box_ow,
                              unsigned box_n, bool mn_major) {
  unsigned long long dims[4] = {(unsigned long long)q.KW * q.H, (unsigned long long)q.OW, (unsigned long long)q.C,
                                (unsigned long long)q.N};
  unsigned long long str[3] = {(unsigned long long)q.H * 4, (unsigned long long)q.H * q.W * 4,
                               (unsigned long long)ld * 4};
  unsigned box[4] = {32, box_ow, 1, box_n};
  return encode_map(map, in, 4, dims, str, box, mn_major);
}

inline bool conv_full_fprop(cudaStream_t st, const ConvFullShape &q, const float *in, int ld_in,
                            const float *kernel, int ld_k, const ConvOut &o) {
  if (!conv_full_shape_ok(q)) return false;
  if (o.cl && (!host_aligned16(o.out) || (q.G & 3) != 0)) return false;
  const int ks = q.KW * q.H, nb = 128 / q.OW;
  CUtensorMap ma, mb;
  if (!encode_window_map(&ma, in, ld_in, q, (unsigned)q.OW, (unsigned)nb, false)) return false;
  if (!encode_2d(&mb, Matrix{kernel, q.C * ks, q.G, ld_k}, 32, true)) return false;
  ConvFullFpropProb p;
  p.num_samples = q.N; p.OW = q.OW; p.nb = nb; p.ks = ks; p.j_blocks = (ks + 31) / 32; p.G = q.G;
  p.total_kb = q.C * p.j_blocks;
  fill_conv_out(p, o, q.G);
  p.div_jb = FastDiv((uint32_t)p.j_blocks); p.div_ow = FastDiv((uint32_t)q.OW);
  launch_prob(st, ma, mb, p, dim3(ceil_div_u(q.N, nb), ceil_div_u(q.G, BN), 1), p.total_kb);
  return true;
}

// in_deriv (reference layout) from a channels-last out_deriv [N][OW][G]
inline bool conv_full_dgrad(cudaStream_t st, const ConvFullShape &q, const float *dycl, const float *kernel,
                            int ld_k, float *in_deriv, int ld_id) {
  if (!conv_full_shape_ok(q) || !host_aligned16(dycl)) return false;
  const int ks = q.KW * q.H;
  const int nb = 128 / q.W, nbc = 128 / q.H;
  CUtensorMap da, db;
  if (!encode_act_map(&da, dycl, q.N, q.OW, q.G, (unsigned)q.W, (unsigned)nb, false)) return false;
  unsigned long long dims[3] = {(unsigned long long)q.G, (unsigned long long)ks, (unsigned long long)q.C};
  unsigned long long str[2] = {(unsigned long long)ld_k * 4, (unsigned long long)ld_k * 4 * ks};
  unsigned box[3] = {32, (unsigned)q.H, (unsigned)nbc};
  if (!encode_map(&db, kernel, 3, dims, str, box, false)) return false;
  ConvFullDgradProb p;
  p.num_samples = q.N; p.W = q.W; p.H = q.H; p.C = q.C; p.nb = nb; p.nbc = nbc;
  p.g_blocks = (q.G + 31) / 32; p.total_kb = q.KW * p.g_blocks; p.out = in_deriv; p.ldo = ld_id;
  p.div_gb = FastDiv((uint32_t)p.g_blocks); p.div_h = FastDiv((uint32_t)q.H);
  launch_prob(st, da, db, p, dim3(ceil_div_u(q.N, nb), ceil_div_u(q.C, nbc), 1), p.total_kb);
  return true;
}

// kernel gradient from the input itself (4-D window map) and a channels-last out_deriv
inline bool conv_full_wgrad(cudaStream_t st, const ConvFullShape &q, const float *in, int ld_in, const float *dycl,
                            const ConvWgradOut &o) {
  if (!conv_full_shape_ok(q) || !conv_wgrad_out_ok(o) || !host_aligned16(dycl)) return false;
  const int ks = q.KW * q.H;
  if ((ks & 31) != 0) return false;                         // M atoms must not straddle channels
  CUtensorMap wa, wb;
  if (!encode_window_map(&wa, in, ld_in, q, 1, 32, true)) return false;
  if (!encode_act_map(&wb, dycl, q.N, q.OW, q.G, 1, 32, true)) return false;
  launch_conv_wgrad<true>(st, wa, wb, o, q.C, q.KW, q.G, q.C * ks, 0, q.N, q.OW, (uint32_t)ks);
  return true;
}

// Same contract as conv_backward(); in_value must be TMA-addressable (16-byte aligned rows).
inline bool conv_full_backward(cudaStream_t st, const ConvFullShape &q, ConvBackward &b) {
  if (!conv_full_shape_ok(q)) return false;
  if (pack_samples_per_cta(q.OW) < 1) return false;
  const int ks = q.KW * q.H;
  const bool do_dgrad = b.in_deriv != nullptr;
  const bool do_wgrad = b.sgd != nullptr || b.kernel_grad != nullptr;
  ConvWgradOut wo = {b.sgd ? b.kernel : b.kernel_grad, b.sgd ? b.ld_k : b.ld_kg, b.prev, b.ld_p, b.sgd};
  if (do_wgrad && (!conv_wgrad_out_ok(wo) || (ks & 31) != 0)) return false;
  float *dycl = scratch(SCRATCH_DYCL, (size_t)q.N * q.OW * q.G * sizeof(float), st);
  const int prow = pack_partial_rows(q.N, q.OW);
  float *bpart = b.want_bias ? scratch(SCRATCH_BIAS, (size_t)prow * q.G * sizeof(float), st) : nullptr;
  if (!dycl || !host_aligned16(dycl) || (b.want_bias && !bpart)) return false;
  CUtensorMap probe;
  if (do_wgrad && !encode_window_map(&probe, b.in_value, b.ld_iv, q, 1, 32, true)) return false;
  if (do_dgrad) {
    unsigned long long dims[3] = {(unsigned long long)q.G, (unsigned long long)ks, (unsigned long long)q.C};
    unsigned long long str[2] = {(unsigned long long)b.ld_k * 4, (unsigned long long)b.ld_k * 4 * ks};
    unsigned box[3] = {32, (unsigned)q.H, (unsigned)(128 / q.H)};
    if (!encode_map(&probe, b.kernel, 3, dims, str, box, false)) return false;
  }
  // ---- nothing can fail past this point
  launch_pack(st, b.out_deriv, b.ld_od, q.N, q.G, q.OW, dycl, bpart);
  b.bias_partial = bpart; b.bias_rows = prow;
  ForkJoin fj(st, do_dgrad && do_wgrad && !b.sgd);           // see conv_backward()
  cudaStream_t wst = fj.active() ? fj.side() : st;
  if (do_dgrad) conv_full_dgrad(st, q, dycl, b.kernel, b.ld_k, b.in_deriv, b.ld_id);
  if (do_wgrad) conv_full_wgrad(wst, q, b.in_value, b.ld_iv, dycl, wo);
  return true;
}

}  // namespace tma
}  // namespace kcnn

#endif
