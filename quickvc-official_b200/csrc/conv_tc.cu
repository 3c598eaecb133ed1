// conv_tc.cu -- placeholder until the tcgen05 kernel lands (next commit).
#include "common.cuh"
namespace qvc {
int launch_conv_tc(const qvc_conv_args&, cudaStream_t) {
  set_error("tcgen05 back end not built yet");
  return QVC_ERR_UNSUPPORTED;
}
}  // namespace qvc
