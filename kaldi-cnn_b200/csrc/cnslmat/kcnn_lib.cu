// kaldi-cnn_b200/csrc/cnslmat/kcnn_lib.cu -- library state of libkaldicnn_b200.so.

#include "kcnn_common.cuh"

namespace kcnn {
unsigned long long g_launch_count = 0;
cudaStream_t g_legacy_stream = 0;
}  // namespace kcnn

extern "C" {

void kcnn_set_stream(cudaStream_t stream) { kcnn::g_legacy_stream = stream; }
cudaStream_t kcnn_get_stream(void) { return kcnn::g_legacy_stream; }
unsigned long long kcnn_launch_count(void) { return kcnn::g_launch_count; }
void kcnn_reset_launch_count(void) { kcnn::g_launch_count = 0; }
const char *kcnn_build_info(void) {
  return "kaldi-cnn_b200 sm_100a (compute_100a) nvcc " __DATE__;
}
int kcnn_abi_version(void) { return 1; }

}  // extern "C"
