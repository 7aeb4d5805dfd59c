/*
 * include/cnsl-cu-kernels.h -- L0: the extern "C" kernel-launcher ABI of the
 * B200-native CNN-layer hot path (libkaldicnn_b200.so).
 *
 * This header replaces the reference's src/cnslmat/cnsl-cu-kernels.h (itself
 * styled after Kaldi's cu-kernels-ansi.h): plain C linkage, raw device
 * pointers, MatrixDim {rows, cols, stride} by value, ints; launchers return
 * void and errors surface through cudaGetLastError() in the caller, exactly as
 * the reference's host code expects (cnslmat/conv2D.cc:108).
 *
 * Three groups:
 *  (1) LEGACY launchers -- same names and parameter lists as
 *      cnsl-cu-kernels.h:25-45, so the reference's own conv2D.cc links against
 *      this library unchanged.  The Gr/Bl arguments are accepted and ignored
 *      (the sm_100a kernels choose their own geometry); work is issued on the
 *      stream set by kcnn_set_stream() (default: the legacy default stream).
 *  (2) STREAM-ORDERED launchers cudaF_*_s -- what the new CuMatrixBase members
 *      call: explicit cudaStream_t, each matrix's own stride (the reference
 *      kernel reads out_deriv with out_value's pitch, cnsl-cu-kernels.cu:293).
 *  (3) FUSED entry points -- implicit-GEMM convolution forward / input-gradient
 *      / weight-gradient, the affine (fully connected) GEMMs, the fused SGD
 *      update and the max-pool-with-index pair.  They replace SEQUENCES of
 *      reference launches; each comment names the sequence.
 *
 * Only float is built: the reference instantiates BaseFloat = float only
 * (cnslmat/conv2D.cc:81,96,138 hard-code CuMatrix<BaseFloat>), so its cudaD_*
 * launchers are unreachable.
 *
 * All pointers are DEVICE pointers.  Nothing here allocates, synchronises or
 * touches the host: every entry point is CUDA-graph capturable.
 */
#ifndef CNSL_CNSLMAT_CNSL_CU_KERNELS_H_
#define CNSL_CNSLMAT_CNSL_CU_KERNELS_H_

#include <cuda_runtime_api.h>
#include "cu-matrixdim.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library state ------------------------------------------------------ */

/* Stream used by the LEGACY launchers of group (1). */
void kcnn_set_stream(cudaStream_t stream);
cudaStream_t kcnn_get_stream(void);
/* Number of kernels this library has launched since load / since the last
 * reset (bench.py reports it as "gpu_launches"). */
unsigned long long kcnn_launch_count(void);
void kcnn_reset_launch_count(void);
/* Programmatic dependent launch for the library's launches from now on (default on; KCNN_PDL=0 in the
 * environment disables it for good).  Returns the previous setting. */
int kcnn_set_pdl(int on);
/* Per-launch timing for the benchmark's roofline table: between _start and _stop every kernel this
 * library launches (outside stream capture) is bracketed by CUDA events on its stream and tagged with the
 * label and algorithmic FLOPs / bytes announced last by kcnn_profile_label (the work is attributed to the
 * first launch under a label).  _stop synchronises the device and returns the number of records; _get
 * returns record i: demangled kernel name, label, duration in ms, work, grid. */
void kcnn_profile_start(void);
int kcnn_profile_stop(void);
/* 1 between _start and _stop.  Callers that fork work onto side streams keep everything on ONE stream while
 * this is set, so that a record is the duration of its kernel alone and not of a kernel sharing the SMs. */
int kcnn_profile_active(void);
void kcnn_profile_label(const char *label, double flops, double bytes);
int kcnn_profile_get(int i, char *kernel, int kernel_len, char *label, int label_len, float *ms,
                     double *flops, double *bytes, unsigned int *grid3);
/* "sm_100a" build tag and the ABI revision of this header. */
const char *kcnn_build_info(void);
int kcnn_abi_version(void);

/* Arithmetic used by the GEMM-shaped entry points of group (3). */
enum {
  KCNN_MATH_FP32_SIMT = 0,   /* FP32 FMA on the CUDA cores; 1e-5 class         */
  KCNN_MATH_TF32_TC = 1      /* tcgen05.mma kind::tf32, FP32 accumulate in TMEM */
};

/* Max-pool window modes (the two bools of Maxpool_prop, cu-matrix.h:479). */
enum { KCNN_POOL_PLAIN = 0, KCNN_POOL_OVERLAP = 1, KCNN_POOL_OVERLAP2D = 2 };

/* ---- (1) legacy launchers: cnsl-cu-kernels.h:25-45 ------------------------ */

void cudaF_span_row_to_convmat(dim3 Gr, dim3 Bl, const float *in, MatrixDim in_dim,
                               float *span, MatrixDim span_dim, int in_height,
                               int in_width, int in_channel, int kernel_height,
                               int kernel_width, int row_offset);           /* :25 */
void cudaF_convmat_to_out(dim3 Gr, dim3 Bl, const float *convMat, MatrixDim conv_dim,
                          float *out, MatrixDim out_dim, int out_height,
                          int out_width, int num_sample);                   /* :28 */
void cudaF_add_mat_rep_vec(dim3 Gr, dim3 Bl, const float *vec, int rep, float *out,
                           MatrixDim out_dim);                              /* :29 */
void cudaF_flip_mat(dim3 Gr, dim3 Bl, const float *orig, MatrixDim orig_dim,
                    int kernel_height, int kernel_width, int group, float *flip,
                    MatrixDim flip_dim);                                    /* :30 */
void cudaF_pad_zero(dim3 Gr, dim3 Bl, const float *orig, MatrixDim orig_dim,
                    int orig_height, int orig_width, int kernel_height,
                    int kernel_width, float *padmat, MatrixDim padmat_dim); /* :31 */
void cudaF_tp_block(dim3 Gr, dim3 Bl, const float *in, MatrixDim in_dim, float *out,
                    MatrixDim out_dim, int block_size);                     /* :32 */
void cudaF_tp_inside_block(dim3 Gr, dim3 Bl, const float *in, MatrixDim in_dim,
                           float *out, MatrixDim out_dim, int block_size);  /* :33 */
void cudaF_mod_permute_row(dim3 Gr, dim3 Bl, const float *in, MatrixDim in_dim,
                           float *out, MatrixDim out_dim, int block_size,
                           int in_channel);                                 /* :34 */
void cudaF_copy_rows_at(dim3 Gr, dim3 Bl, const float *src, MatrixDim src_dim,
                        float *dest, MatrixDim dest_dim, int row_offset);   /* :35 */
void cudaF_maxpool_prop(dim3 Gr, dim3 Bl, const float *src, MatrixDim src_dim,
                        float *pool, MatrixDim pool_dim, int in_height_,
                        int in_width_, int pool_height_dim_, int pool_width_dim_,
                        int pool_channel_dim_);                             /* :36 */
void cudaF_maxpool_backprop(dim3 Gr, dim3 Bl, const float *in_val, MatrixDim in_val_dim,
                            const float *out_val, MatrixDim out_val_dim,
                            const float *out_deriv, MatrixDim out_deriv_dim,
                            float *dest, MatrixDim dest_dim, int in_height_,
                            int in_width_, int pool_height_dim_, int pool_width_dim_,
                            int pool_channel_dim_);                         /* :37 */
void cudaF_maxpoolchannel_overlap_prop(dim3 Gr, dim3 Bl, const float *src,
                                       MatrixDim src_dim, float *pool, MatrixDim pool_dim,
                                       int in_height_, int in_width_, int pool_height_dim_,
                                       int pool_width_dim_, int pool_channel_dim_); /* :39 */
void cudaF_maxpoolchannel_overlap_backprop(dim3 Gr, dim3 Bl, const float *in_val,
                                           MatrixDim in_val_dim, const float *out_val,
                                           MatrixDim out_val_dim, const float *out_deriv,
                                           MatrixDim out_deriv_dim, float *dest,
                                           MatrixDim dest_dim, int in_height_, int in_width_,
                                           int pool_height_dim_, int pool_width_dim_,
                                           int pool_channel_dim_);          /* :40 */
void cudaF_maxpoolchannel_overlap2D_prop(dim3 Gr, dim3 Bl, const float *src,
                                         MatrixDim src_dim, float *pool, MatrixDim pool_dim,
                                         int in_height_, int in_width_, int pool_height_dim_,
                                         int pool_width_dim_, int pool_channel_dim_); /* :42 */
void cudaF_maxpoolchannel_overlap2D_backprop(dim3 Gr, dim3 Bl, const float *in_val,
                                             MatrixDim in_val_dim, const float *out_val,
                                             MatrixDim out_val_dim, const float *out_deriv,
                                             MatrixDim out_deriv_dim, float *dest,
                                             MatrixDim dest_dim, int in_height_,
                                             int in_width_, int pool_height_dim_,
                                             int pool_width_dim_, int pool_channel_dim_); /* :43 */

/* ---- (2) stream-ordered launchers --------------------------------------- */

/* this[i,j] += vec[j / rep].  Replaces _add_mat_rep_vec (cnsl-cu-kernels.cu:61-75). */
void cudaF_add_mat_rep_vec_s(cudaStream_t st, const float *vec, int rep, float *out,
                             MatrixDim out_dim);
/* flip[g*ks + r, c] = orig[c*ks + (ks-1-r), g].  _flip_mat (.cu:78-97).
 * in_channel = flip_dim.cols. */
void cudaF_flip_mat_s(cudaStream_t st, const float *orig, MatrixDim orig_dim,
                      int kernel_height, int kernel_width, int group, float *flip,
                      MatrixDim flip_dim);
/* Zero border of kernel_height-1 / kernel_width-1 per side.  _pad_zero (.cu:100-134). */
void cudaF_pad_zero_s(cudaStream_t st, const float *orig, MatrixDim orig_dim,
                      int orig_height, int orig_width, int kernel_height,
                      int kernel_width, float *padmat, MatrixDim padmat_dim);
/* out[c, n*bs + p] = in[n, c*bs + p].  _tp_block (.cu:138-161). */
void cudaF_tp_block_s(cudaStream_t st, const float *in, MatrixDim in_dim, float *out,
                      MatrixDim out_dim, int block_size);
/* out[n*bs + p, g] = in[n, g*bs + p].  _tp_inside_block (.cu:165-185). */
void cudaF_tp_inside_block_s(cudaStream_t st, const float *in, MatrixDim in_dim,
                             float *out, MatrixDim out_dim, int block_size);
/* out[(i % C)*bs + i / C, :] = in[i, :].  _mod_permute_row (.cu:189-210). */
void cudaF_mod_permute_row_s(cudaStream_t st, const float *in, MatrixDim in_dim,
                             float *out, MatrixDim out_dim, int block_size,
                             int in_channel);
/* dest[i + row_offset, :] = src[i, :].  _copy_rows_at (.cu:214-228). */
void cudaF_copy_rows_at_s(cudaStream_t st, const float *src, MatrixDim src_dim,
                          float *dest, MatrixDim dest_dim, int row_offset);
/* 3-D max pooling.  mode KCNN_POOL_PLAIN: _maxpool_prop (.cu:231-269);
 * OVERLAP: .cu:310-356; OVERLAP2D: .cu:405-452.  -1e20 sentinel, strict '<',
 * first maximum in c -> w -> h order wins (bit-exact incl. signed zero / NaN). */
void cudaF_maxpool_prop_s(cudaStream_t st, const float *src, MatrixDim src_dim,
                          float *pool, MatrixDim pool_dim, int in_height, int in_width,
                          int pool_height_dim, int pool_width_dim, int pool_channel_dim,
                          int mode);
/* Reference-exact routing: dest = err at EVERY window element equal to the
 * pooled value, other elements untouched.  _maxpool_backprop (.cu:271-308) and
 * the overlap variants (.cu:358-403, 454-503; those accumulate -- done with
 * atomics here, the reference kernels race).  zero_others != 0 additionally
 * writes 0 to the non-maximal elements of each window (PLAIN mode only), which
 * folds the caller's kSetZero pass (nnet0/nnet-component-nnet0.cc:889) into
 * the same kernel. */
void cudaF_maxpool_backprop_s(cudaStream_t st, const float *in_val, MatrixDim in_val_dim,
                              const float *out_val, MatrixDim out_val_dim,
                              const float *out_deriv, MatrixDim out_deriv_dim,
                              float *dest, MatrixDim dest_dim, int in_height,
                              int in_width, int pool_height_dim, int pool_width_dim,
                              int pool_channel_dim, int mode, int zero_others);

/* ---- (3) fused entry points ---------------------------------------------- */

/* Max pooling that also records, per output, the window position
 * (c*pw*ph + w*ph + h, one byte) of the first maximum; PLAIN mode, window
 * <= 256 elements.  Same values as cudaF_maxpool_prop_s. */
void cudaF_maxpool_prop_index(cudaStream_t st, const float *src, MatrixDim src_dim,
                              float *pool, MatrixDim pool_dim, unsigned char *index,
                              int index_stride, int in_height, int in_width,
                              int pool_height_dim, int pool_width_dim,
                              int pool_channel_dim);
/* Exact-routing backward from the recorded index: dest = err at the recorded
 * element, 0 at the other window elements (whole dest written; no input
 * values read).  Equals the reference whenever a window has a unique maximum. */
void cudaF_maxpool_backprop_index(cudaStream_t st, const unsigned char *index,
                                  int index_stride, const float *out_deriv,
                                  MatrixDim out_deriv_dim, float *dest, MatrixDim dest_dim,
                                  int in_height, int in_width, int pool_height_dim,
                                  int pool_width_dim, int pool_channel_dim);

/* Convolution forward as ONE implicit GEMM:
 *   out[n, g*OH*OW + ow*OH + oh] = bias[g] + sum_{c,kw,kh}
 *        Xpad[n, c, ow+kw, oh+kh] * kernel[(c*KW + kw)*KH + kh, g]
 * OH = H + 2*pad_h - KH + 1, OW likewise.  Replaces, for
 * ConvolutionComponent::Propagate (nnet0/nnet-component-nnet0.cc:423-446):
 * _pad_zero + [_span_row_to_convmat + SGEMM + _copy_rows_at]* + _convmat_to_out
 * + _add_mat_rep_vec.  bias may be NULL.  concat == 0 writes the raw
 * [(pos*N + n) x G] matrix of CuMatrixBase::Conv2D(..., concat=false)
 * (cnslmat/conv2D.cc:199). */
void cudaF_conv2d_fprop(cudaStream_t st, int math, const float *in, MatrixDim in_dim,
                        const float *kernel, MatrixDim kernel_dim, const float *bias,
                        float *out, MatrixDim out_dim, int in_height, int in_width,
                        int in_channel, int pad_height, int pad_width, int kernel_height,
                        int kernel_width, int group, int concat);
/* Input gradient:
 *   in_deriv[n, c, w, h] = sum_{g,kw,kh} out_deriv[n, g, w+pad_w-kw, h+pad_h-kh]
 *                          * kernel[(c*KW + kw)*KH + kh, g]
 * One kernel for both branches of ConvolutionComponent::Backprop (:499-540):
 * replaces TpInsideBlock + FlipMat + transpose + TpBlock + PaddingZero + Conv2D
 * + TpBlock, or PaddingZero + FlipMat + Conv2D. */
void cudaF_conv2d_dgrad(cudaStream_t st, int math, const float *out_deriv,
                        MatrixDim out_deriv_dim, const float *kernel, MatrixDim kernel_dim,
                        float *in_deriv, MatrixDim in_deriv_dim, int in_height,
                        int in_width, int in_channel, int pad_height, int pad_width,
                        int kernel_height, int kernel_width, int group);
/* Weight gradient (un-normalised), rows already in linear_params_ order:
 *   kernel_grad[(c*KW + kw)*KH + kh, g] = sum_{n,ow,oh} Xpad[n,c,ow+kw,oh+kh]
 *                                         * out_deriv[n, g, ow, oh]
 *   bias_grad[g] = sum_{n,ow,oh} out_deriv[n, g, ow, oh]        (may be NULL)
 * Replaces, in ConvolutionComponent::Update (:745-765, 775): PaddingZero +
 * TpBlock + TpInsideBlock + Conv2D(concat=false) + ModPermuteRow + AddRowSumMat.
 * workspace: device scratch of kcnn_conv2d_wgrad_workspace() bytes (split-K
 * partial sums of the generic kernels; the TMA path reduces its K-splits inside
 * the kernel, through distributed shared memory), may be NULL when that returns 0. */
void cudaF_conv2d_wgrad(cudaStream_t st, int math, const float *in_value,
                        MatrixDim in_value_dim, const float *out_deriv,
                        MatrixDim out_deriv_dim, float *kernel_grad,
                        MatrixDim kernel_grad_dim, float *bias_grad, void *workspace,
                        int in_height, int in_width, int in_channel, int pad_height,
                        int pad_width, int kernel_height, int kernel_width, int group);
size_t kcnn_conv2d_wgrad_workspace(int num_rows, int in_height, int in_width,
                                   int in_channel, int pad_height, int pad_width,
                                   int kernel_height, int kernel_width, int group);

/* Fully connected layer (AffineComponent, nnet2/nnet-component.cc:1216-1258).
 * W is [out_dim x in_dim].
 *   fprop : out = 1 * bias^T + in * W^T          (CopyRowsFromVec + SGEMM, :1224-1227)
 *   dgrad : in_deriv = out_deriv * W             (:1246-1247)
 *   wgrad : w_grad = out_deriv^T * in_value, bias_grad = column sums of out_deriv
 *           (un-normalised; the SGEMM of nnet0/nnet-component-nnet0.cc:1141 and
 *           the AddRowSumMat of :1137) */
void cudaF_affine_fprop(cudaStream_t st, int math, const float *in, MatrixDim in_dim,
                        const float *w, MatrixDim w_dim, const float *bias, float *out,
                        MatrixDim out_dim);
void cudaF_affine_dgrad(cudaStream_t st, int math, const float *out_deriv,
                        MatrixDim out_deriv_dim, const float *w, MatrixDim w_dim,
                        float *in_deriv, MatrixDim in_deriv_dim);
void cudaF_affine_wgrad(cudaStream_t st, int math, const float *in_value,
                        MatrixDim in_value_dim, const float *out_deriv,
                        MatrixDim out_deriv_dim, float *w_grad, MatrixDim w_grad_dim,
                        float *bias_grad);

/* Fused backward of one layer in the single-GPU (apply-immediately) case and for the
 * data-parallel (gradient-only) case.  Both return 1 when the fused TMA / tcgen05 path
 * ran and 0 when the shape or alignment is not eligible -- the caller then uses the
 * separate dgrad / wgrad / update entry points above (same results).
 *
 * cudaF_conv2d_backward: ConvolutionComponent::Backprop + Update
 * (nnet0/nnet-component-nnet0.cc:461-544, 738-777) from one channels-last staging copy
 * of out_deriv and one of in_value: in_deriv (skipped when NULL), then
 *   apply != 0:  prev = momentum*prev + a_decay*K + a_grad*dK ; K += prev ;
 *                bias += a_grad * db       (kernel_grad / bias_grad unused)
 *   apply == 0:  kernel_grad = dK ; bias_grad = db
 * staged_input: NULL, or the staging copy cudaF_conv2d_fprop_staged left for this in_value.
 * cudaF_affine_wgrad_sgd: FullyConnectedComponent::UpdateSimple (:1133-1143) with the
 * update applied in the weight-gradient GEMM's epilogue; the gradient is never stored. */
int cudaF_conv2d_backward(cudaStream_t st, int math, const float *in_value,
                          MatrixDim in_value_dim, const float *out_deriv,
                          MatrixDim out_deriv_dim, float *kernel, MatrixDim kernel_dim,
                          float *in_deriv, MatrixDim in_deriv_dim, float *kernel_grad,
                          MatrixDim kernel_grad_dim, float *bias_grad, float *prev_grad,
                          MatrixDim prev_grad_dim, float *bias, int apply, float momentum,
                          float decay_alpha, float grad_alpha, const float *staged_input,
                          int in_height, int in_width, int in_channel, int pad_height,
                          int pad_width, int kernel_height, int kernel_width, int group);
/* Forward pass that also LEAVES the channels-last staging copy of `in` in caller-owned memory
 * (kcnn_conv2d_staging_floats() floats, 16-byte aligned; 0 = this shape has no staging copy),
 * so that cudaF_conv2d_backward(..., staged_input = staging, ...) for the SAME in_value skips
 * its own pack.  Returns 1 when the copy WAS written (tensor-core TMA path taken), else 0. */
size_t kcnn_conv2d_staging_floats(int num_rows, int in_height, int in_width, int in_channel,
                                  int pad_height, int pad_width, int kernel_height,
                                  int kernel_width, int group);
int cudaF_conv2d_fprop_staged(cudaStream_t st, int math, const float *in, MatrixDim in_dim,
                               const float *kernel, MatrixDim kernel_dim, const float *bias,
                               float *out, MatrixDim out_dim, int in_height, int in_width,
                               int in_channel, int pad_height, int pad_width, int kernel_height,
                               int kernel_width, int group, int concat, float *staging);
/* The same forward pass with an activation fused into its epilogue (KCNN_ACT_RELU:
 * RectifiedLinearComponent::Propagate, upstream nnet2/nnet-component.cc:799-811, out = x > 0 ? x : 0
 * applied to the biased output before it is stored: `out` then holds what the ReLU component
 * would have produced and the pre-activation is never written).  staging as above. */
#define KCNN_ACT_NONE 0
#define KCNN_ACT_RELU 1
int cudaF_conv2d_fprop_act(cudaStream_t st, int math, const float *in, MatrixDim in_dim,
                           const float *kernel, MatrixDim kernel_dim, const float *bias, float *out,
                           MatrixDim out_dim, int in_height, int in_width, int in_channel,
                           int pad_height, int pad_width, int kernel_height, int kernel_width,
                           int group, int concat, float *staging, int act);
void cudaF_affine_fprop_act(cudaStream_t st, int math, const float *in, MatrixDim in_dim,
                            const float *w, MatrixDim w_dim, const float *bias, float *out,
                            MatrixDim out_dim, int act);
int cudaF_affine_wgrad_sgd(cudaStream_t st, int math, const float *in_value,
                           MatrixDim in_value_dim, const float *out_deriv,
                           MatrixDim out_deriv_dim, float *w, MatrixDim w_dim,
                           float *prev_grad, MatrixDim prev_grad_dim, float *bias,
                           float momentum, float decay_alpha, float grad_alpha);

/* Momentum / weight-decay SGD in ONE pass over the parameters:
 *   prev = momentum*prev; prev += decay_alpha*params; prev += grad_alpha*grad;
 *   params += prev
 * with the reference's four roundings (Scale, AddMat, AddMat, AddMat of
 * nnet0/nnet-component-nnet0.cc:769-772 and :1139-1142) kept in order. */
void cudaF_sgd_momentum_update(cudaStream_t st, float *params, MatrixDim params_dim,
                               float *prev_grad, MatrixDim prev_grad_dim,
                               const float *grad, MatrixDim grad_dim, float momentum,
                               float decay_alpha, float grad_alpha);
/* out[g] = sum over rows n and over the `inner` adjacent columns of map g of
 * m[n, g*inner + p]; m has out_count*inner columns.  inner = 1 is the column-sum
 * half of CuVectorBase::AddRowSumMat (:775, :1137); inner = OH*OW gives the
 * convolution's bias gradient straight from out_deriv (no TpInsideBlock copy). */
void cudaF_sum_rows_per_map(cudaStream_t st, const float *m, MatrixDim m_dim, int inner,
                            float *out);
/* vec[i] = alpha * grad[i] + vec[i]  (the bias AddRowSumMat tail, :775 / :1137). */
void cudaF_vec_axpy(cudaStream_t st, float *vec, const float *grad, int dim, float alpha);

/* ---- glue between hot-path layers --------------------------- */

/* RectifiedLinearComponent (nnet2/nnet-component.cc:799-827). */
void cudaF_relu_fprop(cudaStream_t st, const float *in, MatrixDim in_dim, float *out,
                      MatrixDim out_dim);
void cudaF_relu_bprop(cudaStream_t st, const float *out_value, MatrixDim out_value_dim,
                      const float *out_deriv, MatrixDim out_deriv_dim, float *in_deriv,
                      MatrixDim in_deriv_dim);
/* SoftmaxComponent::Propagate (:930-950): row softmax + floor 1e-20. */
void cudaF_softmax_fprop(cudaStream_t st, const float *in, MatrixDim in_dim, float *out,
                         MatrixDim out_dim);
/* SoftmaxComponent::Backprop (:952-1000). */
void cudaF_softmax_bprop(cudaStream_t st, const float *out_value, MatrixDim out_value_dim,
                         const float *out_deriv, MatrixDim out_deriv_dim, float *in_deriv,
                         MatrixDim in_deriv_dim);
/* NormalizeComponent (:576-639): out = in * max(2^-66, |row|^2 / D)^-0.5 per row, and its derivative
 * in_deriv = f out_deriv - (f == 2^33 ? 0 : f^3) / D (out_deriv . in) in. */
void cudaF_normalize_fprop(cudaStream_t st, const float *in, MatrixDim in_dim, float *out,
                           MatrixDim out_dim);
void cudaF_normalize_bprop(cudaStream_t st, const float *in_value, MatrixDim in_value_dim,
                           const float *out_deriv, MatrixDim out_deriv_dim, float *in_deriv,
                           MatrixDim in_deriv_dim);
/* Hard-label cross-entropy: deriv[i, label_i] = 1 / post[i, label_i], else 0;
 * objf_accum[0] += sum_i log post[i, label_i] (double, device). */
void cudaF_xent_deriv(cudaStream_t st, const float *post, MatrixDim post_dim,
                      const int *labels, float *deriv, MatrixDim deriv_dim,
                      double *objf_accum);

/* ---- (4) the fused training step: channels-last activations ----------------------------
 *
 * Entry points of NnetMinibatchUpdater's fused step (csrc/nnet2/nnet-fused.cc).  Between the
 * time-axis layers (in_height = kernel_height = 1) activations and derivatives are kept
 * channels-last -- dense [N][W][C] floats, C fastest, 16-byte aligned -- which is the layout the
 * convolution tensor maps read, so that what ConvolutionComponent::Propagate / Backprop / Update
 * (nnet0/nnet-component-nnet0.cc:423-446, 461-544, 738-777) do with 5-9 permute / pad copies per
 * layer, and round 1 did with one staging pack per operand, is tile addressing only.  All
 * tensor-core (TF32) only; every function returns 1 when it launched and 0 when the shape or an
 * alignment is not eligible (nothing launched: the caller falls back to the component path). */

/* 1 when the shape is one the channels-last entry points below accept (channel counts that are
 * multiples of 4, at least 8 input channels and 32 maps, widths up to 128; full-height: in_height a
 * multiple of 4 and kernel_width * in_height a multiple of 32). */
int kcnn_conv_time_shape_ok(int N, int W, int C, int pad_width, int kernel_width, int group);
int kcnn_conv_full_shape_ok(int N, int in_height, int in_width, int in_channel, int kernel_width, int group);
/* out[n][c*W + w] = in[n][w][c]: a channels-last activation copied back to the reference layout. */
void cudaF_cl_to_ref(cudaStream_t st, const float *in, int N, int W, int C, float *out, int ldo);

/* Forward of a time-axis layer.  x: [N][W][C].  out_cl != 0: out is [N][OW][G] (the next time-axis
 * layer's input); else out is the reference matrix, rows [G][OW] with pitch ldo (what a following
 * affine layer reads).  bias (may be NULL) and, with relu != 0, max(., 0) in the epilogue. */
int cudaF_conv_time_fprop_cl(cudaStream_t st, const float *x, int N, int W, int C, int pad_width,
                             int kernel_width, int group, const float *kernel, MatrixDim kernel_dim,
                             const float *bias, float *out, int out_cl, int ldo, int relu);
/* dx[N][W][C] = input gradient of dy[N][OW][G]; mask (may be NULL, [N][W][C]): dx = mask > 0 ? dx : 0,
 * the backward pass of the ReLU that produced this layer's input. */
int cudaF_conv_time_dgrad_cl(cudaStream_t st, const float *dy, int N, int W, int C, int pad_width,
                             int kernel_width, int group, const float *kernel, MatrixDim kernel_dim,
                             float *dx, const float *mask);
/* Weight gradient from x[N][W][C] and dy[N][OW][G], K-splits reduced inside the kernel.
 * apply != 0: the momentum / weight-decay step on (w = linear_params_, prev_grad) in the epilogue
 * (nnet0/nnet-component-nnet0.cc:767-773), the gradient is never stored; else w receives dK. */
int cudaF_conv_time_wgrad_cl(cudaStream_t st, const float *x, const float *dy, int N, int W, int C,
                             int pad_width, int kernel_width, int group, float *w, MatrixDim w_dim,
                             float *prev_grad, MatrixDim prev_grad_dim, int apply, float momentum,
                             float decay_alpha, float grad_alpha);
/* Full-height first layer (kernel_height = in_height, no padding): input in the reference layout,
 * output / out_deriv channels-last [N][OW][G]. */
int cudaF_conv_full_fprop_cl(cudaStream_t st, const float *in, MatrixDim in_dim, int in_height,
                             int in_width, int in_channel, int kernel_width, int group,
                             const float *kernel, MatrixDim kernel_dim, const float *bias, float *out,
                             int relu);
int cudaF_conv_full_dgrad_cl(cudaStream_t st, const float *dy, int num_rows, int in_height, int in_width,
                             int in_channel, int kernel_width, int group, const float *kernel,
                             MatrixDim kernel_dim, float *in_deriv, MatrixDim in_deriv_dim);
int cudaF_conv_full_wgrad_cl(cudaStream_t st, const float *in, MatrixDim in_dim, const float *dy,
                             int in_height, int in_width, int in_channel, int kernel_width, int group,
                             float *w, MatrixDim w_dim, float *prev_grad, MatrixDim prev_grad_dim,
                             int apply, float momentum, float decay_alpha, float grad_alpha);

/* Affine forward with the two element-wise components that follow it in nnet.config in its epilogue:
 * out = relu ? max(in W^T + b, 0) : in W^T + b; drop_out (may be NULL) = out .* scale, scale = low with
 * probability dp else high, drawn from *seed_dev exactly as DropoutComponent::Propagate draws it
 * (upstream nnet2/nnet-component.cc:3592-3620; the seed is advanced by cudaF_softmax_xent /
 * cudaF_bump_seeds once the forward pass is over). */
int cudaF_affine_fprop_fused(cudaStream_t st, const float *in, MatrixDim in_dim, const float *w,
                             MatrixDim w_dim, const float *bias, float *out, MatrixDim out_dim, int relu,
                             float *drop_out, MatrixDim drop_dim, float dp, float low, float high,
                             const unsigned long long *seed_dev);
/* Affine input gradient with the backward pass of the element-wise components that PRECEDE the layer:
 *   relu_out only:        in_deriv = relu_out > 0 ? dY W : 0                       (:813-827)
 *   relu_out + drop_out:  in_deriv = relu_out > 0 ? (dY W) * drop_out / relu_out : 0 (:3634-3636 then ReLU)
 * perm_r > 0: the input of this layer is a time-axis activation with perm_r positions per map and
 * in_deriv is written channels-last: logical column g*perm_r + pos goes to pos*(cols/perm_r) + g
 * (in_deriv is then dense, in_deriv_dim.stride = cols).  The masks are indexed like the logical matrix. */
int cudaF_affine_dgrad_fused(cudaStream_t st, const float *out_deriv, MatrixDim out_deriv_dim,
                             const float *w, MatrixDim w_dim, float *in_deriv, MatrixDim in_deriv_dim,
                             const float *relu_out, int relu_stride, const float *drop_out,
                             int drop_stride, int perm_r);

/* MaxpoolComponent on channels-last data (in_height = 1, plain mode): in [N][W][C] ->
 * out [N][W/pw][C/pc], comparison order, sentinel and tie behaviour of _maxpool_prop /
 * _maxpool_backprop (cnsl-cu-kernels.cu:231-308).  out_relu (may be NULL) additionally receives
 * max(out, 0): the ReLU that follows the pool.  ref_ld > 0: out / out_relu are written (backward: out
 * is read) in the reference layout, rows [c][w] with pitch ref_ld -- the pool feeds an affine layer;
 * out_deriv and in_deriv are always channels-last.  Backward writes every element of in_deriv (the
 * zero fill of nnet0/nnet-component-nnet0.cc:889 included); relu_gate != 0 multiplies by [in > 0],
 * the backward pass of the ReLU that precedes the pool. */
void cudaF_maxpool_prop_cl(cudaStream_t st, const float *in, int N, int W, int C, int pool_width_dim,
                           int pool_channel_dim, float *out, float *out_relu, int ref_ld);
void cudaF_maxpool_backprop_cl(cudaStream_t st, const float *in, const float *out, int ref_ld,
                               const float *out_deriv, int N, int W, int C, int pool_width_dim,
                               int pool_channel_dim, float *in_deriv, int relu_gate);

/* Column sums of several matrices in one launch (deterministic, two-stage).  Per job:
 *   KCNN_COLSUM_STORE        dst0[c]  = sum_r src[r][c]                     (float: a bias gradient)
 *   KCNN_COLSUM_AXPY         dst0[c] += alpha * sum                         (float: bias += lr * db, :775 / :1137)
 *   KCNN_COLSUM_STATS_RELU   dst0[c] += sum ; dst1[c] += #{src[r][c] > 0}   (double: NonlinearComponent::UpdateStats
 *   KCNN_COLSUM_STATS_VALUE  dst0[c] += sum                                  of a ReLU / softmax, :337-363)
 * perm_w > 0: src is a channels-last activation [rows][perm_w][perm_c] viewed as rows x (perm_w*perm_c)
 * and column w*perm_c + ch is accumulated into element ch*perm_w + w (the reference's [c][w] order).
 * scratch: kcnn_colsum_batch_scratch_bytes() bytes, zero-filled ONCE by the caller before first use. */
#define KCNN_COLSUM_STORE 0
#define KCNN_COLSUM_AXPY 1
#define KCNN_COLSUM_STATS_RELU 2
#define KCNN_COLSUM_STATS_VALUE 3
typedef struct {
  const float *src;
  int rows, cols, ld;
  int op;
  int perm_w, perm_c;
  void *dst0, *dst1;
  float alpha;
} KcnnColsumJob;
size_t kcnn_colsum_batch_scratch_bytes(const KcnnColsumJob *jobs, int njobs);
void cudaF_colsum_batch(cudaStream_t st, const KcnnColsumJob *jobs, int njobs, void *scratch);

/* Softmax + cross-entropy objective / derivative + softmax backward in one kernel, bit-identical to
 * cudaF_softmax_fprop -> cudaF_xent_deriv -> cudaF_softmax_bprop.  logits == NULL: post already holds
 * the posteriors.  Also advances the num_seeds (<= 4) dropout seeds.  Returns 0 (nothing launched)
 * for rows longer than 4096 columns. */
int cudaF_softmax_xent(cudaStream_t st, const float *logits, MatrixDim logits_dim, float *post,
                       MatrixDim post_dim, const int *labels, float *d_logits, MatrixDim d_dim,
                       double *objf_accum, unsigned long long *const *seeds, int num_seeds);
void cudaF_bump_seeds(cudaStream_t st, unsigned long long *const *seeds, int num_seeds);

#ifdef __cplusplus
}
#endif

#endif /* CNSL_CNSLMAT_CNSL_CU_KERNELS_H_ */
