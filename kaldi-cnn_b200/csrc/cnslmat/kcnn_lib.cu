// kaldi-cnn_b200/csrc/cnslmat/kcnn_lib.cu -- library state of libkaldicnn_b200.so.

#include "kcnn_common.cuh"

#include <stdlib.h>

namespace kcnn {
unsigned long long g_launch_count = 0;
cudaStream_t g_legacy_stream = 0;

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    // Opt-in: measured on the C2 step it LOSES 2.5 % (0.859 vs 0.837 ms) -- the early-scheduled
    // dependents take SM slots from the tail of the running kernel and gain little, because
    // a captured graph already keeps kernel-to-kernel gaps near 1 us.
    const char *e = getenv("KCNN_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
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
