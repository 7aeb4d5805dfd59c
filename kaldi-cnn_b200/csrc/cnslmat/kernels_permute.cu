// kaldi-cnn_b200/csrc/cnslmat/kernels_permute.cu
//
// The bit-exact data-movement members: FlipMat, PaddingZero, TpBlock,
// TpInsideBlock, ModPermuteRow, AddMatRepVec, plus the legacy im2col / col2im /
// copy_rows_at launchers that keep the reference's own conv2D.cc linkable.
//
// They replace the per-element SIMT kernels of src/cnslmat/cnsl-cu-kernels.cu
// (:10-228), each of which maps threadIdx.x to the OUTPUT column only, so every
// transposing copy there (FlipMat, TpInsideBlock, _convmat_to_out) reads with a
// large stride.  All of these are HBM-bound copies; the roofline is
// 4 * (elements read + written) bytes over 6.5 TB/s.  Three shapes cover them:
//
//   swap_outer  out[b][a][x] = in[a][b][x]   (TpBlock, ModPermuteRow)
//               x is contiguous on both sides: direct coalesced copy when the
//               run is long, shared-memory tile when it is short.
//   swap_inner  out[z][q][r] = in[z][r][q]   (TpInsideBlock, FlipMat, col2im)
//               the contiguous axis changes: 32 x 32 shared-memory transpose,
//               coalesced on both the read and the write side.
//   map         out[j] = f(in, j)            (PaddingZero, AddMatRepVec, im2col)
//
// In the fused hot path (cudaF_conv2d_*) none of these copies is materialised:
// the same index algebra lives in the implicit-GEMM operand addressing.  These
// kernels exist because the members stay public (cudamatrix/cu-matrix.h:463-477).

#include "kcnn_common.cuh"

#include <type_traits>

namespace kcnn {

// ---------------------------------------------------------------- swap_outer --
// in  address: a*sa_in  + b*sb_in  + x
// out address: b*sb_out + a*sa_out + x          x in [0, L)
// One thread moves kIlp units (float or float4) that are a whole grid apart: the loads are all
// issued before the first store, so a thread keeps kIlp requests in flight (one 4-byte load per
// thread caps a 2048-thread SM at ~8 KB in flight, i.e. ~1.5 TB/s by Little's law).
constexpr int kIlp = 4;
template <int VEC>
__global__ void __launch_bounds__(256)
swap_outer_direct(const float *__restrict__ in, float *__restrict__ out, int A, int B, int units,
                  long long sa_in, long long sb_in, long long sb_out, long long sa_out,
                  FastDiv div_units, FastDiv div_A, int ab_limit, long long total) {
  kcnn::pdl_prologue();
  typedef typename std::conditional<VEC == 4, float4, float>::type T;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  T v[kIlp];
  long long dst[kIlp];
#pragma unroll
  for (int k = 0; k < kIlp; k++) {
    const long long t = t0 + k * step;
    dst[k] = -1;
    if (t >= total) continue;
    // order threads by OUTPUT address: (b, a, x)
    uint32_t ba, x, b, a;
    div_units.divmod((uint32_t)t, ba, x);
    div_A.divmod(ba, b, a);
    if ((long long)a * B + b >= ab_limit) continue;    // ragged last block of rows
    dst[k] = b * sb_out + a * sa_out + (long long)x * VEC;
    v[k] = __ldg(reinterpret_cast<const T *>(in + a * sa_in + b * sb_in + (long long)x * VEC));
  }
#pragma unroll
  for (int k = 0; k < kIlp; k++)
    if (dst[k] >= 0) *reinterpret_cast<T *>(out + dst[k]) = v[k];
}

static void launch_swap_outer_direct(cudaStream_t st, const float *in, float *out, int A, int B, int L,
                                     long long sa_in, long long sb_in, long long sb_out, long long sa_out,
                                     int ab_limit) {
  const bool v4 = (L % 4 == 0) && (sa_in % 4 == 0) && (sb_in % 4 == 0) && (sb_out % 4 == 0) && (sa_out % 4 == 0) &&
                  host_aligned16(in) && host_aligned16(out);
  const int units = v4 ? L / 4 : L;
  const long long total = (long long)A * B * units;
  if (total == 0) return;
  const unsigned grid = ceil_div_u(total, 256 * kIlp);
  if (v4)
    KCNN_LAUNCH(swap_outer_direct<4>, grid, 256, 0, st, in, out, A, B, units, sa_in, sb_in, sb_out, sa_out,
                FastDiv((uint32_t)units), FastDiv((uint32_t)A), ab_limit, total);
  else
    KCNN_LAUNCH(swap_outer_direct<1>, grid, 256, 0, st, in, out, A, B, units, sa_in, sb_in, sb_out, sa_out,
                FastDiv((uint32_t)units), FastDiv((uint32_t)A), ab_limit, total);
}

// Short runs (L < 32) with sb_in == L and sa_out == L (TpBlock): a block takes 32 values of a
// and KB whole values of b (KB*L ~ 128 floats), so it reads 32 spans of KB*L contiguous floats
// and writes KB spans of 32*L contiguous floats.
__global__ void __launch_bounds__(256)
swap_outer_tiled(const float *__restrict__ in, float *__restrict__ out, int A, int B, int L,
                 int KB, long long sa_in, long long sb_out, FastDiv div_span, FastDiv div_run, FastDiv div_L) {
  kcnn::pdl_prologue();
  extern __shared__ float tile[];
  const int a0 = blockIdx.x * 32, b0 = blockIdx.y * KB;
  const int na = min(32, A - a0), nb = min(KB, B - b0);
  const int span = KB * L;                 // contiguous input floats per a (nb*L of them valid)
  const int pitch = span | 1;              // odd pitch: conflict-free column reads
  const float *src = in + (long long)a0 * sa_in + (long long)b0 * L;
#pragma unroll 4
  for (int e = threadIdx.x; e < 32 * span; e += 256) {
    uint32_t al, k;
    div_span.divmod((uint32_t)e, al, k);
    if ((int)al < na && (int)k < nb * L) tile[al * pitch + k] = __ldg(src + (long long)al * sa_in + k);
  }
  __syncthreads();
  const int run = 32 * L;                  // contiguous output floats per b (na*L of them valid)
  float *dst = out + (long long)b0 * sb_out + (long long)a0 * L;
#pragma unroll 4
  for (int e = threadIdx.x; e < KB * run; e += 256) {
    uint32_t bl, r, al, x;
    div_run.divmod((uint32_t)e, bl, r);
    div_L.divmod(r, al, x);
    if ((int)bl < nb && (int)al < na) dst[(long long)bl * sb_out + r] = tile[al * pitch + bl * L + x];
  }
}

// Small-Q transpose (TpInsideBlock with a short block: Q = block_size): one CTA moves the whole
// [R][Q] block of one batch entry through shared memory -- the read is one contiguous run when
// sr_in == Q, the write Q runs of R floats.  The 32 x 32 tiles below would use Q / 32 of each
// warp on the read side.
__global__ void __launch_bounds__(256)
swap_inner_smallq(const float *__restrict__ in, float *__restrict__ out, int R, int Q, long long base_in,
                  long long sz_in, long long sr_in, long long sz_out, long long sq_out, FastDiv div_Q,
                  FastDiv div_R) {
  kcnn::pdl_prologue();
  extern __shared__ float tile[];
  const int pitch = Q | 1;
  const float *src = in + base_in + (long long)blockIdx.x * sz_in;
  float *dst = out + (long long)blockIdx.x * sz_out;
  const int total = R * Q;
#pragma unroll 4
  for (int e = threadIdx.x; e < total; e += 256) {
    uint32_t r, q;
    div_Q.divmod((uint32_t)e, r, q);
    tile[r * pitch + q] = __ldg(src + (long long)r * sr_in + q);
  }
  __syncthreads();
#pragma unroll 4
  for (int e = threadIdx.x; e < total; e += 256) {
    uint32_t q, r;
    div_R.divmod((uint32_t)e, q, r);
    dst[(long long)q * sq_out + r] = tile[r * pitch + q];
  }
}

// ---------------------------------------------------------------- swap_inner --
// in  address: base_in + z*sz_in + r*sr_in + q        q in [0, Q) contiguous
// out address:          z*sz_out + q*sq_out + r       r in [0, R) contiguous
__global__ void __launch_bounds__(256)
swap_inner_tiled(const float *__restrict__ in, float *__restrict__ out, int R, int Q,
                 long long base_in, long long sz_in, long long sr_in, long long sz_out,
                 long long sq_out) {
  kcnn::pdl_prologue();
  __shared__ float tile[32][33];
  const int z = blockIdx.z;
  const int q0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float *src = in + base_in + (long long)z * sz_in;
  float *dst = out + (long long)z * sz_out;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int r = r0 + ty + k, q = q0 + tx;
    if (r < R && q < Q) tile[ty + k][tx] = __ldg(src + (long long)r * sr_in + q);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int q = q0 + ty + k, r = r0 + tx;
    if (r < R && q < Q) dst[(long long)q * sq_out + r] = tile[tx][ty + k];
  }
}

static void launch_swap_inner(cudaStream_t st, const float *in, float *out, int Z, int R, int Q,
                              long long base_in, long long sz_in, long long sr_in,
                              long long sz_out, long long sq_out) {
  if (Z == 0 || R == 0 || Q == 0) return;
  if (Q < 32 && (long long)R * (Q | 1) <= 12288) {
    KCNN_LAUNCH(swap_inner_smallq, (unsigned)Z, 256, (size_t)R * (Q | 1) * sizeof(float), st, in, out, R, Q, base_in,
                sz_in, sr_in, sz_out, sq_out, FastDiv((uint32_t)Q), FastDiv((uint32_t)R));
    return;
  }
  // gridDim.z is limited to 65535: walk the batch in slabs.
  for (int z0 = 0; z0 < Z; z0 += 65535) {
    int nz = Z - z0 < 65535 ? Z - z0 : 65535;
    dim3 grid(ceil_div_u(Q, 32), ceil_div_u(R, 32), nz);
    KCNN_LAUNCH(swap_inner_tiled, grid, 256, 0, st, in, out + (long long)z0 * sz_out, R, Q,
                base_in + (long long)z0 * sz_in, sz_in, sr_in, sz_out, sq_out);
  }
}

// ----------------------------------------------------------------------- map --

__global__ void __launch_bounds__(256)
pad_zero_kernel(const float *__restrict__ orig, int orig_stride, float *__restrict__ pad,
                int pad_stride, int rows, int pad_cols, int H, int W, int KH, int KW, int PH,
                FastDiv div_cols, FastDiv div_ps, FastDiv div_ph) {
  kcnn::pdl_prologue();
  const long long total = (long long)rows * pad_cols;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  float v[kIlp];
  long long dst[kIlp];
#pragma unroll
  for (int k = 0; k < kIlp; k++) {                   // kIlp independent loads in flight per thread
    const long long t = t0 + k * step;
    dst[k] = -1;
    v[k] = 0.0f;
    if (t >= total) continue;
    uint32_t i, j, c, p, J, I;
    div_cols.divmod((uint32_t)t, i, j);
    div_ps.divmod(j, c, p);
    div_ph.divmod(p, J, I);
    int m = (int)I - (KH - 1), n = (int)J - (KW - 1);
    if (m >= 0 && m < H && n >= 0 && n < W)
      v[k] = __ldg(orig + (size_t)i * orig_stride + (size_t)c * (H * W) + n * H + m);
    dst[k] = (long long)i * pad_stride + j;
  }
#pragma unroll
  for (int k = 0; k < kIlp; k++)
    if (dst[k] >= 0) pad[dst[k]] = v[k];
}

template <bool kVec4>
__global__ void __launch_bounds__(256)
add_mat_rep_vec_kernel(const float *__restrict__ vec, float *__restrict__ out, int rows,
                       int cols, int stride, FastDiv div_units, FastDiv div_rep) {
  kcnn::pdl_prologue();
  const int units = kVec4 ? cols / 4 : cols;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * units) return;
  uint32_t i, u;
  div_units.divmod((uint32_t)t, i, u);
  if (kVec4) {
    float4 *p = reinterpret_cast<float4 *>(out + (size_t)i * stride) + u;
    float4 v = *p;
    uint32_t j = 4 * u;
    v.x += __ldg(vec + div_rep.div(j));
    v.y += __ldg(vec + div_rep.div(j + 1));
    v.z += __ldg(vec + div_rep.div(j + 2));
    v.w += __ldg(vec + div_rep.div(j + 3));
    *p = v;
  } else {
    out[(size_t)i * stride + u] += __ldg(vec + div_rep.div(u));
  }
}

__global__ void __launch_bounds__(256)
copy_rows_at_kernel(const float *__restrict__ src, int src_stride, float *__restrict__ dest,
                    int dest_stride, int rows, int cols, int row_offset, FastDiv div_cols) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * cols) return;
  uint32_t i, j;
  div_cols.divmod((uint32_t)t, i, j);
  dest[(size_t)(i + row_offset) * dest_stride + j] = __ldg(src + (size_t)i * src_stride + j);
}

// Legacy im2col (cnsl-cu-kernels.cu:10-38): rows position-major, sample-minor.
__global__ void __launch_bounds__(256)
span_row_to_convmat_kernel(const float *__restrict__ in, int in_rows, int in_stride,
                           float *__restrict__ span, int span_rows, int span_cols,
                           int span_stride, int H, int W, int KH, int KW, int row_offset,
                           FastDiv div_cols, FastDiv div_rows, FastDiv div_ks, FastDiv div_kh,
                           FastDiv div_q) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)span_rows * span_cols) return;
  uint32_t i, j, I, Ir, J, Jr, kw, kh, ow, oh;
  div_cols.divmod((uint32_t)t, i, j);
  div_rows.divmod(i + row_offset, I, Ir);
  div_ks.divmod(j, J, Jr);
  div_kh.divmod(Jr, kw, kh);
  div_q.divmod(I, ow, oh);
  size_t idx = (size_t)Ir * in_stride + (oh + ow * H) + (kh + kw * H) + (size_t)J * H * W;
  span[(size_t)i * span_stride + j] = __ldg(in + idx);
}

}  // namespace kcnn

using namespace kcnn;

extern "C" {

void cudaF_add_mat_rep_vec_s(cudaStream_t st, const float *vec, int rep, float *out,
                             MatrixDim d) {
  if (d.rows == 0 || d.cols == 0) return;
  bool vec4 = d.cols % 4 == 0 && d.stride % 4 == 0 && host_aligned16(out);
  int units = vec4 ? d.cols / 4 : d.cols;
  unsigned int grid = ceil_div_u((long long)d.rows * units, 256);
  FastDiv du((uint32_t)units), dr((uint32_t)rep);
  if (vec4)
    KCNN_LAUNCH(add_mat_rep_vec_kernel<true>, grid, 256, 0, st, vec, out, d.rows, d.cols,
                d.stride, du, dr);
  else
    KCNN_LAUNCH(add_mat_rep_vec_kernel<false>, grid, 256, 0, st, vec, out, d.rows, d.cols,
                d.stride, du, dr);
}

void cudaF_flip_mat_s(cudaStream_t st, const float *orig, MatrixDim od, int KH, int KW, int group,
                      float *flip, MatrixDim fd) {
  // flip[g*ks + r, c] = orig[c*ks + (ks-1-r), g]   batch z = r, r-axis = c, q-axis = g
  int ks = KH * KW, C = fd.cols;
  launch_swap_inner(st, orig, flip, /*Z=*/ks, /*R=*/C, /*Q=*/group,
                    /*base_in=*/(long long)(ks - 1) * od.stride, /*sz_in=*/-(long long)od.stride,
                    /*sr_in=*/(long long)ks * od.stride, /*sz_out=*/fd.stride,
                    /*sq_out=*/(long long)ks * fd.stride);
}

void cudaF_pad_zero_s(cudaStream_t st, const float *orig, MatrixDim od, int H, int W, int KH,
                      int KW, float *pad, MatrixDim pd) {
  if (pd.rows == 0 || pd.cols == 0) return;
  int PH = H + 2 * (KH - 1), PW = W + 2 * (KW - 1);
  unsigned int grid = ceil_div_u((long long)pd.rows * pd.cols, 256 * kIlp);
  KCNN_LAUNCH(pad_zero_kernel, grid, 256, 0, st, orig, od.stride, pad, pd.stride, pd.rows,
              pd.cols, H, W, KH, KW, PH, FastDiv((uint32_t)pd.cols), FastDiv((uint32_t)(PH * PW)),
              FastDiv((uint32_t)PH));
}

void cudaF_tp_block_s(cudaStream_t st, const float *in, MatrixDim id, float *out, MatrixDim od,
                      int bs) {
  // out[c, n*bs + p] = in[n, c*bs + p]   a = n, b = c, x = p
  int A = id.rows, B = od.rows, L = bs;
  if (A == 0 || B == 0 || L == 0) return;
  int KB = 128 / L; if (KB < 1) KB = 1;
  dim3 grid(ceil_div_u(A, 32), ceil_div_u(B, KB));
  if (L >= 32 || grid.y > 65535) {            // long runs: direct coalesced copy
    launch_swap_outer_direct(st, in, out, A, B, L, (long long)id.stride, (long long)bs, (long long)od.stride,
                             (long long)bs, A * B);
    return;
  }
  const int span = KB * L, pitch = span | 1;
  KCNN_LAUNCH(swap_outer_tiled, grid, 256, 32 * pitch * sizeof(float), st, in, out, A, B, L, KB,
              (long long)id.stride, (long long)od.stride, FastDiv((uint32_t)span), FastDiv((uint32_t)(32 * L)),
              FastDiv((uint32_t)L));
}

void cudaF_tp_inside_block_s(cudaStream_t st, const float *in, MatrixDim id, float *out,
                             MatrixDim od, int bs) {
  // out[n*bs + p, g] = in[n, g*bs + p]   batch z = n, r-axis = g, q-axis = p
  int G = od.cols;
  launch_swap_inner(st, in, out, /*Z=*/id.rows, /*R=*/G, /*Q=*/bs, 0, (long long)id.stride,
                    (long long)bs, (long long)bs * od.stride, (long long)od.stride);
}

void cudaF_mod_permute_row_s(cudaStream_t st, const float *in, MatrixDim id, float *out,
                             MatrixDim od, int bs, int C) {
  // out[c*bs + pos, :] = in[pos*C + c, :]   a = pos, b = c, x = column
  // (the reference walks i < rows with c = i % C, pos = i / C; rows need not be bs*C)
  int B = C, A = (id.rows + C - 1) / C, L = id.cols;
  if (id.rows == 0 || L == 0) return;
  launch_swap_outer_direct(st, in, out, A, B, L, (long long)C * id.stride, (long long)id.stride,
                           (long long)bs * od.stride, (long long)od.stride, id.rows);
}

void cudaF_copy_rows_at_s(cudaStream_t st, const float *src, MatrixDim sd, float *dest,
                          MatrixDim dd, int row_offset) {
  if (sd.rows == 0 || sd.cols == 0) return;
  unsigned int grid = ceil_div_u((long long)sd.rows * sd.cols, 256);
  KCNN_LAUNCH(copy_rows_at_kernel, grid, 256, 0, st, src, sd.stride, dest, dd.stride, sd.rows,
              sd.cols, row_offset, FastDiv((uint32_t)sd.cols));
}

// ---- legacy launchers (cnsl-cu-kernels.h:25-35); Gr / Bl ignored -----------

void cudaF_span_row_to_convmat(dim3, dim3, const float *in, MatrixDim id, float *span,
                               MatrixDim sd, int H, int W, int C, int KH, int KW,
                               int row_offset) {
  (void)C;
  if (sd.rows == 0 || sd.cols == 0) return;
  unsigned int grid = ceil_div_u((long long)sd.rows * sd.cols, 256);
  KCNN_LAUNCH(span_row_to_convmat_kernel, grid, 256, 0, g_legacy_stream, in, id.rows, id.stride,
              span, sd.rows, sd.cols, sd.stride, H, W, KH, KW, row_offset,
              FastDiv((uint32_t)sd.cols), FastDiv((uint32_t)id.rows), FastDiv((uint32_t)(KH * KW)),
              FastDiv((uint32_t)KH), FastDiv((uint32_t)(H - KH + 1)));
}

void cudaF_convmat_to_out(dim3, dim3, const float *conv, MatrixDim cd, float *out, MatrixDim od,
                          int OH, int OW, int num_sample) {
  // out[n, pos + g*OH*OW] = conv[pos*N + n, g]   batch z = n, r-axis = pos, q-axis = g
  launch_swap_inner(g_legacy_stream, conv, out, /*Z=*/num_sample, /*R=*/OH * OW, /*Q=*/cd.cols, 0,
                    (long long)cd.stride, (long long)num_sample * cd.stride, (long long)od.stride,
                    (long long)OH * OW);
}

void cudaF_add_mat_rep_vec(dim3, dim3, const float *vec, int rep, float *out, MatrixDim d) {
  cudaF_add_mat_rep_vec_s(g_legacy_stream, vec, rep, out, d);
}
void cudaF_flip_mat(dim3, dim3, const float *orig, MatrixDim od, int KH, int KW, int group,
                    float *flip, MatrixDim fd) {
  cudaF_flip_mat_s(g_legacy_stream, orig, od, KH, KW, group, flip, fd);
}
void cudaF_pad_zero(dim3, dim3, const float *orig, MatrixDim od, int H, int W, int KH, int KW,
                    float *pad, MatrixDim pd) {
  cudaF_pad_zero_s(g_legacy_stream, orig, od, H, W, KH, KW, pad, pd);
}
void cudaF_tp_block(dim3, dim3, const float *in, MatrixDim id, float *out, MatrixDim od, int bs) {
  cudaF_tp_block_s(g_legacy_stream, in, id, out, od, bs);
}
void cudaF_tp_inside_block(dim3, dim3, const float *in, MatrixDim id, float *out, MatrixDim od,
                           int bs) {
  cudaF_tp_inside_block_s(g_legacy_stream, in, id, out, od, bs);
}
void cudaF_mod_permute_row(dim3, dim3, const float *in, MatrixDim id, float *out, MatrixDim od,
                           int bs, int C) {
  cudaF_mod_permute_row_s(g_legacy_stream, in, id, out, od, bs, C);
}
void cudaF_copy_rows_at(dim3, dim3, const float *src, MatrixDim sd, float *dest, MatrixDim dd,
                        int row_offset) {
  cudaF_copy_rows_at_s(g_legacy_stream, src, sd, dest, dd, row_offset);
}

}  // extern "C"
