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

namespace kcnn {

// ---------------------------------------------------------------- swap_outer --
// in  address: a*sa_in  + b*sb_in  + x
// out address: b*sb_out + a*sa_out + x          x in [0, L)
__global__ void __launch_bounds__(256)
swap_outer_direct(const float *__restrict__ in, float *__restrict__ out, int A, int B, int L,
                  long long sa_in, long long sb_in, long long sb_out, long long sa_out,
                  FastDiv div_L, FastDiv div_A, int ab_limit) {
  kcnn::pdl_prologue();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)A * B * L) return;
  // order threads by OUTPUT address: (b, a, x)
  uint32_t ba, x, b, a;
  div_L.divmod((uint32_t)t, ba, x);
  div_A.divmod(ba, b, a);
  if ((long long)a * B + b >= ab_limit) return;      // ragged last block of rows
  out[b * sb_out + a * sa_out + x] = __ldg(in + a * sa_in + b * sb_in + x);
}

// Short runs (L < 32) with sb_in == L and sa_out == L (TpBlock): a block takes 32
// values of a and KB whole values of b, so it reads 32 spans of KB*L contiguous
// floats and writes KB spans of 32*L contiguous floats.
__global__ void __launch_bounds__(256)
swap_outer_tiled(const float *__restrict__ in, float *__restrict__ out, int A, int B, int L,
                 int KB, long long sa_in, long long sb_out) {
  kcnn::pdl_prologue();
  extern __shared__ float tile[];
  const int a0 = blockIdx.x * 32, b0 = blockIdx.y * KB;
  const int na = min(32, A - a0), nb = min(KB, B - b0);
  const int span = nb * L;                 // contiguous input floats per a
  const int pitch = (KB * L) | 1;          // odd pitch: conflict-free column reads
  for (int e = threadIdx.x; e < na * span; e += blockDim.x) {
    int al = e / span, k = e - al * span;
    tile[al * pitch + k] = __ldg(in + (long long)(a0 + al) * sa_in + (long long)b0 * L + k);
  }
  __syncthreads();
  const int run = na * L;                  // contiguous output floats per b
  for (int e = threadIdx.x; e < nb * run; e += blockDim.x) {
    int bl = e / run, r = e - bl * run;
    int al = r / L, x = r - al * L;
    out[(long long)(b0 + bl) * sb_out + (long long)a0 * L + r] = tile[al * pitch + bl * L + x];
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
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * pad_cols) return;
  uint32_t i, j, c, p, J, I;
  div_cols.divmod((uint32_t)t, i, j);
  div_ps.divmod(j, c, p);
  div_ph.divmod(p, J, I);
  int m = (int)I - (KH - 1), n = (int)J - (KW - 1);
  float v = 0.0f;
  if (m >= 0 && m < H && n >= 0 && n < W)
    v = __ldg(orig + (size_t)i * orig_stride + (size_t)c * (H * W) + n * H + m);
  pad[(size_t)i * pad_stride + j] = v;
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
  unsigned int grid = ceil_div_u((long long)pd.rows * pd.cols, 256);
  KCNN_LAUNCH(pad_zero_kernel, grid, 256, 0, st, orig, od.stride, pad, pd.stride, pd.rows,
              pd.cols, H, W, KH, KW, PH, FastDiv((uint32_t)pd.cols), FastDiv((uint32_t)(PH * PW)),
              FastDiv((uint32_t)PH));
}

void cudaF_tp_block_s(cudaStream_t st, const float *in, MatrixDim id, float *out, MatrixDim od,
                      int bs) {
  // out[c, n*bs + p] = in[n, c*bs + p]   a = n, b = c, x = p
  int A = id.rows, B = od.rows, L = bs;
  if (A == 0 || B == 0 || L == 0) return;
  if (L >= 32) {
    unsigned int grid = ceil_div_u((long long)A * B * L, 256);
    KCNN_LAUNCH(swap_outer_direct, grid, 256, 0, st, in, out, A, B, L, (long long)id.stride,
                (long long)bs, (long long)od.stride, (long long)bs, FastDiv((uint32_t)L),
                FastDiv((uint32_t)A), A * B);
  } else {
    int KB = 32 / L; if (KB < 1) KB = 1;
    int pitch = (KB * L) | 1;
    dim3 grid(ceil_div_u(A, 32), ceil_div_u(B, KB));
    // gridDim.y <= 65535
    if (grid.y > 65535) {
      unsigned int g1 = ceil_div_u((long long)A * B * L, 256);
      KCNN_LAUNCH(swap_outer_direct, g1, 256, 0, st, in, out, A, B, L, (long long)id.stride,
                  (long long)bs, (long long)od.stride, (long long)bs, FastDiv((uint32_t)L),
                  FastDiv((uint32_t)A), A * B);
    } else {
      KCNN_LAUNCH(swap_outer_tiled, grid, 256, 32 * pitch * sizeof(float), st, in, out, A, B, L,
                  KB, (long long)id.stride, (long long)od.stride);
    }
  }
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
  unsigned int grid = ceil_div_u((long long)A * B * L, 256);
  KCNN_LAUNCH(swap_outer_direct, grid, 256, 0, st, in, out, A, B, L, (long long)C * id.stride,
              (long long)id.stride, (long long)bs * od.stride, (long long)od.stride,
              FastDiv((uint32_t)L), FastDiv((uint32_t)A), id.rows);
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
