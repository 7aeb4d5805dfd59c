/*
 * oracle/kcnn_oracle.c -- CPU oracle for the kaldi-cnn CNN-layer hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see kcnn_oracle_impl.h).  Built by
 * __graft_entry__.build() / oracle/Makefile into oracle/liboracle.so and
 * loaded with ctypes by tests/, smoke() and bench.py's CPU-baseline legs.
 *
 * The same body is compiled for float (oraF_*, the reference's BaseFloat) and
 * for double (oraD_*, used to budget the FP32 / TF32 tolerances).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <dlfcn.h>

#define ORA_CAT(a, b) a##b

/* Optional BLAS back end for ORA(gemm): the reference's CPU build does its GEMMs in the BLAS
 * Kaldi links (AddMatMat -> cblas_sgemm, cnslmat/conv2D.cc:139, nnet2/nnet-component.cc:1227,
 * 1247, nnet0/nnet-component-nnet0.cc:1141), so the CPU BASELINE legs of bench.py route them
 * through the OpenBLAS that numpy bundles (ILP64 build, symbols scipy_cblas_?gemm64_), loaded
 * at run time with dlopen.  Parity tests keep the plain loops (deterministic summation order). */
typedef void (*ora_sgemm_fn)(int, int, int, long long, long long, long long, float, const float *, long long,
                             const float *, long long, float, float *, long long);
typedef void (*ora_dgemm_fn)(int, int, int, long long, long long, long long, double, const double *, long long,
                             const double *, long long, double, double *, long long);
static ora_sgemm_fn g_sgemm = NULL;
static ora_dgemm_fn g_dgemm = NULL;
static void (*g_blas_set_threads)(int) = NULL;

static int ora_blas_gemm_f(int tA, int tB, int m, int n, int k, float alpha, const float *A, int lda,
                           const float *B, int ldb, float beta, float *C, int ldc) {
  if (!g_sgemm) return 0;
  g_sgemm(101, tA ? 112 : 111, tB ? 112 : 111, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  return 1;
}
static int ora_blas_gemm_d(int tA, int tB, int m, int n, int k, double alpha, const double *A, int lda,
                           const double *B, int ldb, double beta, double *C, int ldc) {
  if (!g_dgemm) return 0;
  g_dgemm(101, tA ? 112 : 111, tB ? 112 : 111, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  return 1;
}

#define REAL float
#define ORA(name) ORA_CAT(oraF_, name)
#define ORA_BLAS_GEMM ora_blas_gemm_f
#include "kcnn_oracle_impl.h"
#undef REAL
#undef ORA
#undef ORA_BLAS_GEMM

#define REAL double
#define ORA(name) ORA_CAT(oraD_, name)
#define ORA_BLAS_GEMM ora_blas_gemm_d
#include "kcnn_oracle_impl.h"
#undef REAL
#undef ORA
#undef ORA_BLAS_GEMM

/* path: an ILP64 OpenBLAS (numpy.libs/libscipy_openblas64_*.so); NULL or "" switches back to
 * the plain loops.  Returns 1 when the BLAS is in use. */
int ora_use_blas(const char *path) {
  g_sgemm = NULL; g_dgemm = NULL; g_blas_set_threads = NULL;
  if (!path || !path[0]) return 0;
  void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 0;
  g_sgemm = (ora_sgemm_fn)dlsym(h, "scipy_cblas_sgemm64_");
  g_dgemm = (ora_dgemm_fn)dlsym(h, "scipy_cblas_dgemm64_");
  g_blas_set_threads = (void (*)(int))dlsym(h, "scipy_openblas_set_num_threads64_");
  if (!g_sgemm || !g_dgemm) { g_sgemm = NULL; g_dgemm = NULL; return 0; }
  return 1;
}

/* ---- glue between the hot-path layers, float only ---------- */

/* RectifiedLinearComponent::Propagate / Backprop.
 * nnet2/nnet-component.cc:799-827: out = max(in, 0); in_deriv = out_deriv * (out > 0). */
void oraF_relu_propagate(const float *in, int rows, int cols, int in_stride,
                         float *out, int out_stride) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      float v = in[(size_t)i * in_stride + j];
      out[(size_t)i * out_stride + j] = v > 0.0f ? v : 0.0f;
    }
}

void oraF_relu_backprop(const float *out_value, int rows, int cols, int ov_stride,
                        const float *out_deriv, int od_stride, float *in_deriv,
                        int id_stride) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++)
      in_deriv[(size_t)i * id_stride + j] =
          out_value[(size_t)i * ov_stride + j] > 0.0f
              ? out_deriv[(size_t)i * od_stride + j] : 0.0f;
}

/* SoftmaxComponent::Propagate.  nnet2/nnet-component.cc:930-950:
 * row-wise softmax (max-subtracted) then ApplyFloor(1e-20). */
void oraF_softmax_propagate(const float *in, int rows, int cols, int in_stride,
                            float *out, int out_stride) {
  for (int i = 0; i < rows; i++) {
    const float *x = in + (size_t)i * in_stride;
    float *y = out + (size_t)i * out_stride;
    float mx = x[0];
    for (int j = 1; j < cols; j++) if (x[j] > mx) mx = x[j];
    double sum = 0.0;
    for (int j = 0; j < cols; j++) { y[j] = expf(x[j] - mx); sum += y[j]; }
    float inv = (float)(1.0 / sum);
    for (int j = 0; j < cols; j++) { y[j] *= inv; if (y[j] < 1e-20f) y[j] = 1e-20f; }
  }
}

/* The objective nnet2's NnetUpdater::ComputeObjfAndDeriv forms for hard labels
 * (upstream nnet-update.cc, absent from the patch set; stated from the nnet2
 * contract): objf = sum_i log p[i, label_i]; deriv[i, label_i] = 1 / p[i, label_i],
 * zero elsewhere.  Returns the objective. */
double oraF_xent_objf_and_deriv(const float *post, int rows, int cols, int p_stride,
                                const int *labels, float *deriv, int d_stride) {
  double objf = 0.0;
  for (int i = 0; i < rows; i++) {
    for (int j = 0; j < cols; j++) deriv[(size_t)i * d_stride + j] = 0.0f;
    float p = post[(size_t)i * p_stride + labels[i]];
    objf += log((double)p);
    deriv[(size_t)i * d_stride + labels[i]] = 1.0f / p;
  }
  return objf;
}

/* SoftmaxComponent::Backprop.  nnet2/nnet-component.cc:952-1000:
 * in_deriv[i,:] = out[i,:] * (out_deriv[i,:] - dot(out[i,:], out_deriv[i,:])). */
void oraF_softmax_backprop(const float *out_value, int rows, int cols, int ov_stride,
                           const float *out_deriv, int od_stride, float *in_deriv,
                           int id_stride) {
  for (int i = 0; i < rows; i++) {
    const float *y = out_value + (size_t)i * ov_stride;
    const float *d = out_deriv + (size_t)i * od_stride;
    float *o = in_deriv + (size_t)i * id_stride;
    double dot = 0.0;
    for (int j = 0; j < cols; j++) dot += (double)y[j] * d[j];
    for (int j = 0; j < cols; j++) o[j] = y[j] * (d[j] - (float)dot);
  }
}

/* DropoutComponent::Propagate.  nnet2/nnet-component.cc:3592-3620, pass for pass on a matrix of
 * uniform [0, 1) draws (the reference fills `out` with CuRand::RandUniform first):
 *   out = u ; out += -dp ; out = heaviside(out) ; out *= (high - low) [if != 1] ;
 *   out += low [if != 0] ; out *= in.        high = (1 - dp*low) / (1 - dp). */
void oraF_dropout_propagate(const float *in, int rows, int cols, int in_stride, const float *uniform,
                            int u_stride, float dp, float low_scale, float *out, int out_stride) {
  float high_scale = (1.0 - (dp * low_scale)) / (1.0 - dp);
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      float v = uniform[(size_t)i * u_stride + j];
      v = v + (-dp);                                        /* Add(-dp) */
      v = v > 0.0f ? 1.0f : 0.0f;                           /* ApplyHeaviside */
      if ((high_scale - low_scale) != 1.0f) v = v * (high_scale - low_scale);
      if (low_scale != 0.0f) v = v + low_scale;
      out[(size_t)i * out_stride + j] = v * in[(size_t)i * in_stride + j];       /* MulElements(in) */
    }
}

/* DropoutComponent::Backprop.  :3622-3637: in_deriv->AddMatMatDivMat(out_deriv, out_value, in_value),
 * element-wise a * b / c, and a where c == 0 (Kaldi cu-kernels _add_mat_mat_div_mat). */
void oraF_dropout_backprop(const float *in_value, int rows, int cols, int iv_stride,
                           const float *out_value, int ov_stride, const float *out_deriv, int od_stride,
                           float *in_deriv, int id_stride) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      float a = out_deriv[(size_t)i * od_stride + j], b = out_value[(size_t)i * ov_stride + j],
            c = in_value[(size_t)i * iv_stride + j];
      in_deriv[(size_t)i * id_stride + j] = c == 0.0f ? a : a * b / c;
    }
}

/* NormalizeComponent::Propagate / Backprop.  nnet2/nnet-component.cc:576-639.
 *   f = (max(2^-66, row . row / D))^-0.5 ; out = f * in
 *   in_deriv = f * out_deriv - (f == 2^33 ? 0 : f^3) / D * (out_deriv . in) * in */
void oraF_normalize_propagate(const float *in, int rows, int cols, int in_stride, float *out,
                              int out_stride) {
  const float kNormFloor = (float)pow(2.0, -66);
  for (int i = 0; i < rows; i++) {
    const float *x = in + (size_t)i * in_stride;
    float p = 0.0f;
    for (int j = 0; j < cols; j++) p += (1.0f / cols) * x[j] * x[j];    /* AddDiagMat2(1/D, in, kNoTrans, 0) */
    if (p < kNormFloor) p = kNormFloor;                                  /* ApplyFloor */
    float f = powf(p, -0.5f);                                            /* ApplyPow(-0.5) */
    for (int j = 0; j < cols; j++) out[(size_t)i * out_stride + j] = x[j] * f;   /* MulRowsVec */
  }
}

void oraF_normalize_backprop(const float *in_value, int rows, int cols, int iv_stride,
                             const float *out_deriv, int od_stride, float *in_deriv, int id_stride) {
  const float kNormFloor = (float)pow(2.0, -66);
  for (int i = 0; i < rows; i++) {
    const float *x = in_value + (size_t)i * iv_stride;
    const float *d = out_deriv + (size_t)i * od_stride;
    float *o = in_deriv + (size_t)i * id_stride;
    float p = 0.0f;
    for (int j = 0; j < cols; j++) p += (1.0f / cols) * x[j] * x[j];
    if (p < kNormFloor) p = kNormFloor;
    float f = powf(p, -0.5f);
    for (int j = 0; j < cols; j++) o[j] = f * d[j];                      /* AddDiagVecMat(1, in_norm, out_deriv, 0) */
    float g = f;
    if (g == (float)(1.0 / sqrt((double)kNormFloor))) g = 0.0f;          /* ReplaceValue(1/sqrt(floor), 0) */
    g = g * g * g;                                                       /* ApplyPow(3) */
    float dot = 0.0f;
    for (int j = 0; j < cols; j++) dot += d[j] * x[j];                   /* AddDiagMatMat */
    dot *= g;                                                            /* MulElements */
    for (int j = 0; j < cols; j++) o[j] += (-1.0f / cols) * dot * x[j];  /* AddDiagVecMat(-1/D, ., in_value, 1) */
  }
}

/* NonlinearComponent::UpdateStats.  nnet2/nnet-component.cc:337-363: a FLOAT row-sum vector of
 * out_value (and of the derivative matrix, when given) is added to the DOUBLE accumulators;
 * count += rows.  deriv may be NULL (softmax). */
void oraF_nonlin_update_stats(const float *out_value, int rows, int cols, int ov_stride, const float *deriv,
                              int d_stride, double *value_sum, double *deriv_sum, double *count) {
  *count += rows;
  for (int j = 0; j < cols; j++) {
    float t = 0.0f;                                                      /* temp.AddRowSumMat(1, out_value, 0) */
    for (int i = 0; i < rows; i++) t += out_value[(size_t)i * ov_stride + j];
    value_sum[j] += (double)t;
    if (deriv) {
      float u = 0.0f;
      for (int i = 0; i < rows; i++) u += deriv[(size_t)i * d_stride + j];
      deriv_sum[j] += (double)u;
    }
  }
}

/* Host threads the GEMMs may use (bench.py's CPU legs set it to the box's core count;
 * tests leave it alone).  Returns the number in effect; 1 when built without OpenMP. */
int ora_set_num_threads(int n) {
  if (n > 0 && g_blas_set_threads) g_blas_set_threads(n);
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}
