// cudamatrix/cu-common.h -- shim: CU_SAFE_CALL and n_blocks as the reference's
// host code uses them (cnslmat/conv2D.cc:103, 108).
#ifndef KALDI_CUDAMATRIX_CU_COMMON_H_
#define KALDI_CUDAMATRIX_CU_COMMON_H_
#include <cuda_runtime_api.h>
#include "base/kaldi-common.h"
#include "cudamatrix/cu-matrixdim.h"

#define CU_SAFE_CALL(fun)                                                              \
  {                                                                                    \
    int32 ret;                                                                         \
    if ((ret = (fun)) != 0) {                                                          \
      KALDI_ERR << "cudaError_t " << ret << " : \"" << cudaGetErrorString((cudaError_t)ret) \
                << "\" returned from '" << #fun << "'";                               \
    }                                                                                  \
  }

namespace kaldi {
inline int32 n_blocks(int32 size, int32 block_size) {
  return size / block_size + ((size % block_size == 0) ? 0 : 1);
}
}  // namespace kaldi
#endif
