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
