// placeholder until the tcgen05 kernels land (next commit)
#include "gemm_tc.cuh"
namespace kcnn {
bool tc_conv_fprop(cudaStream_t, const float *, MatrixDim, const float *, MatrixDim, const float *, float *, MatrixDim, int, int, int, int, int, int, int, int, int, int) { return false; }
bool tc_conv_dgrad(cudaStream_t, const float *, MatrixDim, const float *, MatrixDim, float *, MatrixDim, int, int, int, int, int, int, int, int, int) { return false; }
bool tc_conv_wgrad(cudaStream_t, const float *, MatrixDim, const float *, MatrixDim, float *, MatrixDim, void *, int, int, int, int, int, int, int, int, int) { return false; }
size_t tc_conv_wgrad_workspace(int, int, int, int, int, int, int, int, int) { return 0; }
bool tc_affine_fprop(cudaStream_t, const float *, MatrixDim, const float *, MatrixDim, const float *, float *, MatrixDim) { return false; }
bool tc_affine_dgrad(cudaStream_t, const float *, MatrixDim, const float *, MatrixDim, float *, MatrixDim) { return false; }
bool tc_affine_wgrad(cudaStream_t, const float *, MatrixDim, const float *, MatrixDim, float *, MatrixDim) { return false; }
}
