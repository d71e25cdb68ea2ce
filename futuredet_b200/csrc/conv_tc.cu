// tcgen05 / TMEM arm of fd_conv_forward (placeholder until the tensor-core kernel lands).
#include "conv_common.cuh"

namespace fd {

int conv_forward_tc(const ConvArgs&, int precision, cudaStream_t) {
  return set_error(-2, "fd_conv_forward: tensor-core precision %d not available in this build", precision);
}

}  // namespace fd
