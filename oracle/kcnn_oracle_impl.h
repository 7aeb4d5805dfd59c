/*
 * oracle/kcnn_oracle_impl.h -- body of the CPU oracle, included twice by
 * kcnn_oracle.c (REAL=float -> oraF_*, REAL=double -> oraD_*).
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the CPU ("else")
 * branches of the reference's CuMatrixBase extensions and of the nnet0
 * components that drive them.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call it.  The product
 * (kaldi-cnn_b200/) never links or loads it.
 *
 * Parity status.
 *   PINNED against outputs of the reference itself for the L0 rows of SURVEY 8(a)
 *   (a2 AddMatRepVec, a3 FlipMat, a4 PaddingZero, a5 TpBlock, a6 TpInsideBlock,
 *   a7 ModPermuteRow, a8 Maxpool_prop, a9 Maxpool_backprop): oracle/Makefile compiles
 *   the reference's own CUDA kernels (src/cnslmat/cnsl-cu-kernels.cu, unmodified, from
 *   /root/reference) into oracle/_ref/libcnsl_ref_kernels.so, and
 *   tests/test_gpu_reference_kernels.py runs them on the GPU box: reference kernel ==
 *   this oracle == the product, bit for bit.
 *   PINNED for the forward convolution (a1 Conv2D + a2 = a10 ConvolutionComponent::Propagate)
 *   against the reference's GPU path rebuilt from its own kernels (span_row_to_convmat,
 *   convmat_to_out, add_mat_rep_vec from oracle/_ref) and the cuBLAS SGEMM Kaldi's AddMatMat
 *   calls: tests/ref_conv_check.py, test_reference_conv2d_chain (<= 1e-5).
 *   PINNED for a11 / a12 (ConvolutionComponent Backprop, both input-gradient branches, and the
 *   gradient of Update) against the reference's OWN kernels chained in the host order of
 *   nnet0/nnet-component-nnet0.cc:461-544, 738-777 (PaddingZero, span_row_to_convmat, FlipMat,
 *   TpBlock, TpInsideBlock, ModPermuteRow from oracle/_ref + the cuBLAS SGEMM behind AddMatMat):
 *   tests/test_gpu_reference_chains.py, incl. C1a / C1b at N = 256 (<= 1e-5).  The host file that
 *   makes those calls is a patch on Kaldi r4510 and cannot itself be compiled in this image (no
 *   Kaldi tree, no BLAS headers), and it ships no golden vectors.
 *   RESTATEMENT ONLY for the momentum / weight-decay arithmetic of the updates (:767-775,
 *   :1136-1142: stock Kaldi Scale / AddMat / AddRowSumMat, whose sources are not in
 *   /root/reference) and for a14's FullyConnectedComponent: pinned by independent formulations --
 *   NumPy / einsum (oracle/oracle_np.py, tests/test_oracle_einsum.py, <= 1e-12 in FP64), the
 *   reference's finite-difference method in FP64, torch autograd of the whole training step in
 *   FP64 (tests/test_oracle_autograd.py, <= 1e-12 on objective, input derivative and every updated
 *   weight / bias / momentum matrix over two steps), and the committed fixtures (tests/golden/).
 *
 * Every function cites the reference lines it follows
 * (paths relative to /root/reference/src).
 *
 * Matrix convention (cudamatrix/cu-matrixdim.h upstream): element (r,c) of a
 * matrix lives at data[r*stride + c], stride >= cols.
 */

#ifndef REAL
#error "include from kcnn_oracle.c"
#endif

/* C[m x n] = alpha * op(A) * op(B) + beta * C ; the role Kaldi's AddMatMat
 * (cblas_sgemm) plays at cnslmat/conv2D.cc:139 and
 * nnet2/nnet-component.cc:1227,1247, nnet0/nnet-component-nnet0.cc:1141.
 * Plain loops, k-outer axpy form so gcc vectorises the inner loop; rows of C are
 * spread over the host threads (OpenMP) the way a threaded BLAS would. */
void ORA(gemm)(int transA, int transB, int m, int n, int k, REAL alpha,
               const REAL *A, int lda, const REAL *B, int ldb, REAL beta,
               REAL *C, int ldc) {
  if (ORA_BLAS_GEMM(transA, transB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)) return;
  for (int i = 0; i < m; i++) {
    REAL *c = C + (size_t)i * ldc;
    if (beta == (REAL)0) {
      for (int j = 0; j < n; j++) c[j] = 0;
    } else if (beta != (REAL)1) {
      for (int j = 0; j < n; j++) c[j] *= beta;
    }
  }
  if (!transB) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; i++) {
      REAL *c = C + (size_t)i * ldc;
      for (int p = 0; p < k; p++) {
        REAL a = alpha * (transA ? A[(size_t)p * lda + i] : A[(size_t)i * lda + p]);
        const REAL *b = B + (size_t)p * ldb;
        for (int j = 0; j < n; j++) c[j] += a * b[j];
      }
    }
  } else {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; i++) {
      REAL *c = C + (size_t)i * ldc;
      for (int j = 0; j < n; j++) {
        const REAL *b = B + (size_t)j * ldb;
        REAL s = 0;
        if (!transA) {
          const REAL *a = A + (size_t)i * lda;
          for (int p = 0; p < k; p++) s += a[p] * b[p];
        } else {
          for (int p = 0; p < k; p++) s += A[(size_t)p * lda + i] * b[p];
        }
        c[j] += alpha * s;
      }
    }
  }
}

/* CuMatrixBase::Conv2D, CPU branch.  cnslmat/conv2D.cc:44-201.
 *   x    [n_rows x H*W*C]   (this)
 *   kern [KH*KW*C x G]
 *   out  concat!=0: [n_rows x OH*OW*G]   (col2im, :187-196)
 *        concat==0: [OH*OW*n_rows x G]   (raw convMat, :199)
 * The reference splits the im2col rows by free memory (:69-93); the split does
 * not change any value, so the restatement runs one pass.
 * Returns 0, or -1 if the scratch allocation fails. */
int ORA(conv2d)(const REAL *x, int n_rows, int x_stride, const REAL *kern,
                int kern_stride, int H, int W, int C, int KH, int KW, int G,
                REAL *out, int out_stride, int concat) {
  int OH = H - KH + 1, OW = W - KW + 1;          /* :59-60 (stride ignored) */
  int span_h = OH * OW * n_rows;                 /* :65 */
  int span_w = KH * KW * C;                      /* :67 */
  int ks = KH * KW, q = OH;                      /* :117-118 */
  REAL *span = (REAL *)malloc((size_t)span_h * span_w * sizeof(REAL));
  REAL *conv = (REAL *)malloc((size_t)span_h * G * sizeof(REAL));
  if (!span || !conv) { free(span); free(conv); return -1; }
  /* 1. im2col, rows position-major / sample-minor.  :120-133 */
  for (int i = 0; i < span_h; i++) {
    int Ir = i % n_rows, I = i / n_rows;
    int Q = I % q + I / q * H;
    const REAL *xr = x + (size_t)Ir * x_stride;
    REAL *sr = span + (size_t)i * span_w;
    for (int j = 0; j < span_w; j++) {
      int Jr = j % ks, J = j / ks;
      int P = (Jr % KH) + (Jr / KH) * H;
      sr[j] = xr[Q + P + J * H * W];
    }
  }
  /* 2. convMat = span * kernel.  :138-139 (zero-initialised, beta = 1) */
  ORA(gemm)(0, 0, span_h, G, span_w, (REAL)1, span, span_w, kern, kern_stride,
            (REAL)0, conv, G);
  /* 3. col2im scatter or raw copy.  :172-200 */
  if (concat) {
    for (int i = 0; i < span_h; i++) {
      int Ir = i % n_rows, I = i / n_rows;
      for (int j = 0; j < G; j++)
        out[(size_t)Ir * out_stride + (I + j * OH * OW)] = conv[(size_t)i * G + j];
    }
  } else {
    for (int i = 0; i < span_h; i++)
      for (int j = 0; j < G; j++)
        out[(size_t)i * out_stride + j] = conv[(size_t)i * G + j];
  }
  free(span);
  free(conv);
  return 0;
}

/* CuMatrixBase::AddMatRepVec, CPU branch.  cnslmat/conv2D.cc:231-240. */
void ORA(add_mat_rep_vec)(REAL *m, int rows, int cols, int stride,
                          const REAL *vec, int rep) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) m[(size_t)i * stride + j] += vec[j / rep];
}

/* CuMatrixBase::FlipMat, CPU branch.  cnslmat/conv2D.cc:269-284.
 * in [KH*KW*C x G] -> flip [KH*KW*G x C]. */
void ORA(flip_mat)(const REAL *in, int in_stride, int KH, int KW, int C, int G,
                   REAL *flip, int flip_stride) {
  int ks = KH * KW;
  for (int i = 0; i < ks * G; i++) {
    int g = i / ks;
    int p = (g + 1) * ks - 1 - i;
    for (int j = 0; j < C; j++)
      flip[(size_t)i * flip_stride + j] = in[(size_t)(p + j * ks) * in_stride + g];
  }
}

/* CuMatrixBase::PaddingZero, CPU branch.  cnslmat/conv2D.cc:316-342.
 * in [rows x H*W*C] -> pad [rows x (H+2(KH-1))*(W+2(KW-1))*C]. */
void ORA(pad_zero)(const REAL *in, int rows, int in_stride, int H, int W, int C,
                   int KH, int KW, REAL *pad, int pad_stride) {
  int PH = H + 2 * (KH - 1), PW = W + 2 * (KW - 1), PS = PH * PW;
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < PS * C; j++) {
      int c = j / PS, p = j % PS, I = p % PH, J = p / PH;
      REAL v = 0;
      if (KH - 1 <= I && I < KH + H - 1 && KW - 1 <= J && J < KW + W - 1) {
        int mm = I - KH + 1, nn = J - KW + 1;
        v = in[(size_t)i * in_stride + (nn * H + mm) + c * (H * W)];
      }
      pad[(size_t)i * pad_stride + j] = v;
    }
}

/* CuMatrixBase::TpBlock, CPU branch.  cnslmat/conv2D.cc:375-385.
 * in [rows x C*bs] -> out [C x rows*bs]. */
void ORA(tp_block)(const REAL *in, int rows, int in_stride, int C, int bs,
                   REAL *out, int out_stride) {
  for (int i = 0; i < C; i++)
    for (int j = 0; j < rows * bs; j++)
      out[(size_t)i * out_stride + j] =
          in[(size_t)(j / bs) * in_stride + i * bs + j % bs];
}

/* CuMatrixBase::TpInsideBlock, CPU branch.  cnslmat/conv2D.cc:415-425.
 * in [rows x G*bs] -> out [rows*bs x G]. */
void ORA(tp_inside_block)(const REAL *in, int rows, int in_stride, int G, int bs,
                          REAL *out, int out_stride) {
  for (int i = 0; i < rows * bs; i++)
    for (int j = 0; j < G; j++)
      out[(size_t)i * out_stride + j] =
          in[(size_t)(i / bs) * in_stride + j * bs + i % bs];
}

/* CuMatrixBase::ModPermuteRow, CPU branch.  cnslmat/conv2D.cc:452-462. */
void ORA(mod_permute_row)(const REAL *in, int rows, int cols, int in_stride,
                          int C, int bs, REAL *out, int out_stride) {
  for (int i = 0; i < rows; i++) {
    int c = i % C, pos = i / C;
    for (int j = 0; j < cols; j++)
      out[(size_t)(c * bs + pos) * out_stride + j] = in[(size_t)i * in_stride + j];
  }
}

/* overlap2D helper: integer sqrt as the kernels compute it
 * (cnsl-cu-kernels.cu:421 "out_2d_map = sqrt(out_channel)" truncated to int). */
static int ORA(isqrt)(int v) {
  int r = (int)sqrt((double)v);
  return r;
}

/* CuMatrixBase::Maxpool_prop.  mode 0 = plain 3-D pooling
 * (cnsl-cu-kernels.cu:231-269, CPU twin conv2D.cc:531-557), 1 = overlap
 * (.cu:310-356), 2 = overlap2D (.cu:405-452; the CPU text at conv2D.cc:503-529
 * does not compile, so the kernel is the spec).  -1e20 sentinel, strict '<',
 * loop order c -> w -> h.  out_cols is trusted exactly as the reference trusts
 * out->NumCols() (conv2D.cc:469). */
void ORA(maxpool_prop)(const REAL *in, int rows, int in_stride, int H, int W,
                       int ph, int pw, int pc, int mode, REAL *out, int out_cols,
                       int out_stride) {
  if (mode == 1 || mode == 2) { ph = 1; pw = 1; }   /* .cu:316-317, 418-419 */
  int OH = H / ph, OW = W / pw;
  int o2 = 0, i2 = 0;
  if (mode == 2) { o2 = ORA(isqrt)(out_cols / (OH * OW)); i2 = o2 + pc - 1; }
  for (int i = 0; i < rows; i++) {
    const REAL *src = in + (size_t)i * in_stride;
    for (int j = 0; j < out_cols; j++) {
      int oc = j / (OH * OW), pos = j % (OH * OW), ow = pos / OH, oh = pos % OH;
      REAL val = (REAL)-1e20;
      if (mode == 2) {
        int cx0 = oc / o2, cy0 = oc % o2;
        for (int cx = 0; cx < pc; cx++)
          for (int cy = 0; cy < pc; cy++) {
            int ic = (cx0 + cx) * i2 + (cy0 + cy);
            REAL s = src[ic * H * W + pos];
            if (val < s) val = s;
          }
      } else {
        int start = (mode == 1 ? oc * H * W : oc * pc * H * W) + ow * pw * H + oh * ph;
        for (int c = 0; c < pc; c++)
          for (int w = 0; w < pw; w++)
            for (int h = 0; h < ph; h++) {
              REAL s = src[start + h + w * H + c * H * W];
              if (val < s) val = s;
            }
      }
      out[(size_t)i * out_stride + j] = val;
    }
  }
}

/* CuMatrixBase::Maxpool_backprop.  in_deriv must arrive zeroed (the component
 * does that, nnet0/nnet-component-nnet0.cc:889).
 * mode 0: plain -- GPU kernel semantics "dest = err" at every window element
 *   equal to the pooled value (cnsl-cu-kernels.cu:271-308); on a zeroed
 *   in_deriv and non-overlapping windows this equals the CPU AddMat text
 *   (conv2D.cc:660-679).
 * mode 1/2: overlapping windows accumulate (cnsl-cu-kernels.cu:358-403,
 *   454-503), done serially here (the kernels race). */
void ORA(maxpool_backprop)(const REAL *in, int rows, int in_stride,
                           const REAL *out_val, int ov_stride,
                           const REAL *out_deriv, int od_stride, int out_cols,
                           REAL *in_deriv, int id_stride, int H, int W, int ph,
                           int pw, int pc, int mode) {
  if (mode == 1 || mode == 2) { ph = 1; pw = 1; }
  int OH = H / ph, OW = W / pw;
  int o2 = 0, i2 = 0;
  if (mode == 2) { o2 = ORA(isqrt)(out_cols / (OH * OW)); i2 = o2 + pc - 1; }
  for (int i = 0; i < rows; i++) {
    const REAL *src = in + (size_t)i * in_stride;
    REAL *dst = in_deriv + (size_t)i * id_stride;
    for (int j = 0; j < out_cols; j++) {
      int oc = j / (OH * OW), pos = j % (OH * OW), ow = pos / OH, oh = pos % OH;
      REAL ov = out_val[(size_t)i * ov_stride + j];
      REAL err = out_deriv[(size_t)i * od_stride + j];
      if (mode == 2) {
        int cx0 = oc / o2, cy0 = oc % o2;
        for (int cx = 0; cx < pc; cx++)
          for (int cy = 0; cy < pc; cy++) {
            int idx = ((cx0 + cx) * i2 + (cy0 + cy)) * H * W + pos;
            if (ov == src[idx]) dst[idx] = dst[idx] + err;
          }
      } else {
        int start = (mode == 1 ? oc * H * W : oc * pc * H * W) + ow * pw * H + oh * ph;
        for (int c = 0; c < pc; c++)
          for (int w = 0; w < pw; w++)
            for (int h = 0; h < ph; h++) {
              int idx = start + h + w * H + c * H * W;
              if (ov == src[idx]) {
                if (mode == 0) dst[idx] = err;
                else dst[idx] = dst[idx] + err;
              }
            }
      }
    }
  }
}

/* ------------------------------------------------------------------------- */
/* nnet0 components, op for op (every pad / flip / Tp copy is materialised,   */
/* exactly as the reference does, so this is also the timed CPU "port").      */
/* ------------------------------------------------------------------------- */

static REAL *ORA(alloc)(size_t n) { return (REAL *)calloc(n ? n : 1, sizeof(REAL)); }

/* ConvolutionComponent::Propagate.  nnet0/nnet-component-nnet0.cc:423-446.
 * in [N x H*W*C] -> out [N x OH*OW*G], OH = H+2ph-KH+1. */
int ORA(conv_propagate)(const REAL *in, int N, int in_stride, const REAL *lin,
                        int lin_stride, const REAL *bias, int H, int W, int C,
                        int pad_h, int pad_w, int KH, int KW, int G, REAL *out,
                        int out_stride) {
  int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  int OH = Hp - KH + 1, OW = Wp - KW + 1, rc;
  if (pad_h > 0 || pad_w > 0) {                                  /* :430-435 */
    REAL *padded = ORA(alloc)((size_t)N * Hp * Wp * C);
    if (!padded) return -1;
    ORA(pad_zero)(in, N, in_stride, H, W, C, pad_h + 1, pad_w + 1, padded, Hp * Wp * C);
    rc = ORA(conv2d)(padded, N, Hp * Wp * C, lin, lin_stride, Hp, Wp, C, KH, KW, G,
                     out, out_stride, 1);
    free(padded);
  } else {                                                       /* :438 */
    rc = ORA(conv2d)(in, N, in_stride, lin, lin_stride, H, W, C, KH, KW, G, out,
                     out_stride, 1);
  }
  if (rc) return rc;
  ORA(add_mat_rep_vec)(out, N, OH * OW * G, out_stride, bias, OH * OW); /* :443 */
  return 0;
}

/* The branch rule of ConvolutionComponent::Backprop, :489-497.
 * Returns 1 when the reference takes the flip-kernel branch. */
int ORA(conv_backprop_uses_flip)(int pad_h, int pad_w, int KH, int KW, int OH, int OW) {
  int pkh = KH + 2 * (OH - pad_h - 1), pkw = KW + 2 * (OW - pad_w - 1);
  int poh = OH + 2 * (KH - pad_h - 1), pow_ = OW + 2 * (KW - pad_w - 1);
  return !(pkh * pkw < poh * pow_);
}

/* ConvolutionComponent::Backprop, input-derivative part only (the Update call
 * at :541-543 is ORA(conv_update)).  nnet0/nnet-component-nnet0.cc:461-540.
 * branch: -1 = the reference's own choice, 0 = force no-flip (:499-528),
 * 1 = force flip (:529-540).  in_deriv [N x H*W*C] must be pre-sized. */
int ORA(conv_backprop)(const REAL *out_deriv, int N, int od_stride,
                       const REAL *lin, int lin_stride, int H, int W, int C,
                       int pad_h, int pad_w, int KH, int KW, int G, int branch,
                       REAL *in_deriv, int id_stride) {
  int OH = H + 2 * pad_h - KH + 1, OW = W + 2 * pad_w - KW + 1;
  int ks = KH * KW, os = OH * OW, rc = 0;
  int pkh = KH + 2 * (OH - pad_h - 1), pkw = KW + 2 * (OW - pad_w - 1);
  int poh = OH + 2 * (KH - pad_h - 1), pow_ = OW + 2 * (KW - pad_w - 1);
  int flip = branch < 0 ? ORA(conv_backprop_uses_flip)(pad_h, pad_w, KH, KW, OH, OW) : branch;
  if (!flip) {
    /* :501-508  out_deriv -> TpInsideBlock -> FlipMat(OH,OW,N,G) */
    REAL *od_tp = ORA(alloc)((size_t)os * N * G);
    REAL *flip_od = ORA(alloc)((size_t)os * G * N);
    /* :510-520  linear^T -> TpBlock(C, ks) -> PaddingZero(KH,KW,G, OH-ph, OW-pw) */
    REAL *lin_tp = ORA(alloc)((size_t)G * ks * C);
    REAL *lin_tp2 = ORA(alloc)((size_t)C * ks * G);
    REAL *pad_k = ORA(alloc)((size_t)C * pkh * pkw * G);
    REAL *tmp = ORA(alloc)((size_t)C * H * W * N);
    if (!od_tp || !flip_od || !lin_tp || !lin_tp2 || !pad_k || !tmp) rc = -1;
    if (!rc) {
      ORA(tp_inside_block)(out_deriv, N, od_stride, G, os, od_tp, G);
      ORA(flip_mat)(od_tp, G, OH, OW, N, G, flip_od, N);
      for (int r = 0; r < ks * C; r++)                       /* AddMat kTrans :516 */
        for (int g = 0; g < G; g++)
          lin_tp[(size_t)g * ks * C + r] = lin[(size_t)r * lin_stride + g];
      ORA(tp_block)(lin_tp, G, ks * C, C, ks, lin_tp2, ks * G);
      ORA(pad_zero)(lin_tp2, C, ks * G, KH, KW, G, OH - pad_h, OW - pad_w, pad_k,
                    pkh * pkw * G);
      /* :524  pad_kernel.Conv2D(flip_out_deriv, pkh, pkw, G, OH, OW, N, &tmp, true) */
      rc = ORA(conv2d)(pad_k, C, pkh * pkw * G, flip_od, N, pkh, pkw, G, OH, OW, N,
                       tmp, H * W * N, 1);
      /* :525  in_deriv_tmp.TpBlock(N, H*W, in_deriv) */
      if (!rc) ORA(tp_block)(tmp, C, H * W * N, N, H * W, in_deriv, id_stride);
    }
    free(od_tp); free(flip_od); free(lin_tp); free(lin_tp2); free(pad_k); free(tmp);
  } else {
    /* :530-538 */
    REAL *pad_od = ORA(alloc)((size_t)N * poh * pow_ * G);
    REAL *flip_k = ORA(alloc)((size_t)ks * G * C);
    if (!pad_od || !flip_k) rc = -1;
    if (!rc) {
      ORA(pad_zero)(out_deriv, N, od_stride, OH, OW, G, KH - pad_h, KW - pad_w, pad_od,
                    poh * pow_ * G);
      ORA(flip_mat)(lin, lin_stride, KH, KW, C, G, flip_k, C);
      rc = ORA(conv2d)(pad_od, N, poh * pow_ * G, flip_k, C, poh, pow_, G, KH, KW, C,
                       in_deriv, id_stride, 1);
    }
    free(pad_od); free(flip_k);
  }
  return rc;
}

/* ConvolutionComponent::Update: weight gradient + momentum / weight-decay SGD
 * + bias.  nnet0/nnet-component-nnet0.cc:738-777.
 * When grad_out != NULL the un-normalised weight gradient [ks*C x G]
 * (linear_params_grad, :765) is copied there; when bias_grad_out != NULL the
 * column sums of out_deriv_tmp (:775) are copied there.  apply == 0 skips the
 * parameter update (used by the data-parallel tests: grad, reduce, apply). */
int ORA(conv_update)(const REAL *in_value, int N, int iv_stride,
                     const REAL *out_deriv, int od_stride, REAL *lin,
                     int lin_stride, REAL *bias, REAL *prev, int prev_stride,
                     int H, int W, int C, int pad_h, int pad_w, int KH, int KW,
                     int G, REAL learning_rate, REAL weight_decay, REAL momentum,
                     int apply, REAL *grad_out, REAL *bias_grad_out) {
  int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  int OH = Hp - KH + 1, OW = Wp - KW + 1, ks = KH * KW, os = OH * OW, rc = 0;
  REAL *iv_tmp = ORA(alloc)((size_t)C * N * Hp * Wp);            /* :745 */
  REAL *od_tmp = ORA(alloc)((size_t)os * N * G);                 /* :746 */
  REAL *lp_tmp = ORA(alloc)((size_t)ks * C * G);                 /* :748 */
  REAL *lp_grad = ORA(alloc)((size_t)ks * C * G);                /* :749 */
  if (!iv_tmp || !od_tmp || !lp_tmp || !lp_grad) rc = -1;
  if (!rc) {
    if (pad_h > 0 || pad_w > 0) {                                /* :751-754 */
      REAL *padded = ORA(alloc)((size_t)N * Hp * Wp * C);
      if (!padded) rc = -1;
      else {
        ORA(pad_zero)(in_value, N, iv_stride, H, W, C, pad_h + 1, pad_w + 1, padded,
                      Hp * Wp * C);
        ORA(tp_block)(padded, N, Hp * Wp * C, C, Hp * Wp, iv_tmp, N * Hp * Wp);
        free(padded);
      }
    } else {                                                     /* :757 */
      ORA(tp_block)(in_value, N, iv_stride, C, Hp * Wp, iv_tmp, N * Hp * Wp);
    }
  }
  if (!rc) {
    ORA(tp_inside_block)(out_deriv, N, od_stride, G, os, od_tmp, G);   /* :760 */
    /* :763  in_value_tmp.Conv2D(out_deriv_tmp, Hp, Wp, N, OH, OW, G, &tmp, false) */
    rc = ORA(conv2d)(iv_tmp, C, N * Hp * Wp, od_tmp, G, Hp, Wp, N, OH, OW, G, lp_tmp,
                     G, 0);
  }
  if (!rc) {
    ORA(mod_permute_row)(lp_tmp, ks * C, G, G, C, ks, lp_grad, G);     /* :765 */
    if (grad_out) memcpy(grad_out, lp_grad, (size_t)ks * C * G * sizeof(REAL));
    if (bias_grad_out)
      for (int g = 0; g < G; g++) {
        REAL s = 0;
        for (int i = 0; i < os * N; i++) s += od_tmp[(size_t)i * G + g];
        bias_grad_out[g] = s;
      }
    if (apply) {
      /* :767  double learning_rate = learning_rate_ / num_sample, narrowed to
       * BaseFloat when passed as the alpha of AddMat / AddRowSumMat. */
      REAL lr_f = learning_rate / (REAL)N;
      double lr = (double)lr_f;
      REAL a_decay = (REAL)(-1 * lr * (double)weight_decay);     /* :770 */
      REAL a_grad = (REAL)lr;                                    /* :771 */
      for (int r = 0; r < ks * C; r++)
        for (int g = 0; g < G; g++) {
          REAL *p = &prev[(size_t)r * prev_stride + g];
          REAL *w = &lin[(size_t)r * lin_stride + g];
          *p = *p * momentum;                                    /* :769 */
          *p = *p + a_decay * *w;                                /* :770 */
          *p = *p + a_grad * lp_grad[(size_t)r * G + g];         /* :771 */
          *w = *w + *p;                                          /* :772 */
        }
      for (int g = 0; g < G; g++) {                              /* :775 */
        REAL s = 0;
        for (int i = 0; i < os * N; i++) s += od_tmp[(size_t)i * G + g];
        bias[g] = a_grad * s + bias[g];
      }
    }
  }
  free(iv_tmp); free(od_tmp); free(lp_tmp); free(lp_grad);
  return rc;
}

/* AffineComponent::Propagate (inherited by FullyConnectedComponent).
 * nnet2/nnet-component.cc:1216-1228.  W is [out_dim x in_dim]. */
void ORA(fc_propagate)(const REAL *in, int N, int in_stride, const REAL *W,
                       int w_stride, const REAL *bias, int in_dim, int out_dim,
                       REAL *out, int out_stride) {
  for (int i = 0; i < N; i++)                                /* CopyRowsFromVec */
    for (int j = 0; j < out_dim; j++) out[(size_t)i * out_stride + j] = bias[j];
  ORA(gemm)(0, 1, N, out_dim, in_dim, (REAL)1, in, in_stride, W, w_stride, (REAL)1,
            out, out_stride);
}

/* AffineComponent::Backprop, derivative part.  nnet2/nnet-component.cc:1237-1247. */
void ORA(fc_backprop)(const REAL *out_deriv, int N, int od_stride, const REAL *W,
                      int w_stride, int in_dim, int out_dim, REAL *in_deriv,
                      int id_stride) {
  ORA(gemm)(0, 0, N, in_dim, out_dim, (REAL)1, out_deriv, od_stride, W, w_stride,
            (REAL)0, in_deriv, id_stride);
}

/* FullyConnectedComponent::UpdateSimple.  nnet0/nnet-component-nnet0.cc:1133-1143. */
void ORA(fc_update)(const REAL *in_value, int N, int iv_stride,
                    const REAL *out_deriv, int od_stride, REAL *W, int w_stride,
                    REAL *bias, REAL *prev, int prev_stride, int in_dim,
                    int out_dim, REAL learning_rate, REAL weight_decay,
                    REAL momentum) {
  REAL lr_f = learning_rate / (REAL)N;                           /* :1136 */
  double lr = (double)lr_f;
  REAL a_decay = (REAL)(-1 * lr * (double)weight_decay), a_grad = (REAL)lr;
  for (int j = 0; j < out_dim; j++) {                            /* :1137 */
    REAL s = 0;
    for (int i = 0; i < N; i++) s += out_deriv[(size_t)i * od_stride + j];
    bias[j] = a_grad * s + bias[j];
  }
  for (int r = 0; r < out_dim; r++)
    for (int c = 0; c < in_dim; c++) {
      REAL *p = &prev[(size_t)r * prev_stride + c];
      *p = *p * momentum;                                        /* :1139 */
      *p = *p + a_decay * W[(size_t)r * w_stride + c];           /* :1140 */
    }
  ORA(gemm)(1, 0, out_dim, in_dim, N, a_grad, out_deriv, od_stride, in_value,
            iv_stride, (REAL)1, prev, prev_stride);              /* :1141 */
  for (int r = 0; r < out_dim; r++)                              /* :1142 */
    for (int c = 0; c < in_dim; c++)
      W[(size_t)r * w_stride + c] += prev[(size_t)r * prev_stride + c];
}
