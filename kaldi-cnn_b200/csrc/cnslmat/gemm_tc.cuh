// kaldi-cnn_b200/csrc/cnslmat/gemm_tc.cuh
//
// Interface of the tcgen05 / TMEM TF32 implicit-GEMM kernels (kernels_gemm_tc.cu).
// Each function returns true when it launched the tensor-core kernel for the
// given shape, false when the shape is outside what the kernel supports (the
// caller then uses the FP32 CUDA-core kernel, which is at least as accurate).

#ifndef KCNN_GEMM_TC_CUH_
#define KCNN_GEMM_TC_CUH_

#include "kcnn_common.cuh"

namespace kcnn {

bool tc_conv_fprop(cudaStream_t st, const float *in, MatrixDim id, const float *kernel,
                   MatrixDim kd, const float *bias, float *out, MatrixDim od, int N, int H, int W,
                   int C, int ph, int pw, int KH, int KW, int G, int concat);
bool tc_conv_dgrad(cudaStream_t st, const float *out_deriv, MatrixDim odd, const float *kernel,
                   MatrixDim kd, float *in_deriv, MatrixDim idd, int N, int H, int W, int C,
                   int ph, int pw, int KH, int KW, int G);
bool tc_conv_wgrad(cudaStream_t st, const float *in_value, MatrixDim ivd, const float *out_deriv,
                   MatrixDim odd, float *kernel_grad, MatrixDim kgd, void *workspace, int N, int H,
                   int W, int C, int ph, int pw, int KH, int KW, int G);
size_t tc_conv_wgrad_workspace(int N, int H, int W, int C, int ph, int pw, int KH, int KW, int G);
bool tc_affine_fprop(cudaStream_t st, const float *in, MatrixDim id, const float *w, MatrixDim wd,
                     const float *bias, float *out, MatrixDim od);
bool tc_affine_dgrad(cudaStream_t st, const float *out_deriv, MatrixDim odd, const float *w,
                     MatrixDim wd, float *in_deriv, MatrixDim idd);
bool tc_affine_wgrad(cudaStream_t st, const float *in_value, MatrixDim ivd, const float *out_deriv,
                     MatrixDim odd, float *w_grad, MatrixDim wgd);

}  // namespace kcnn

#endif
