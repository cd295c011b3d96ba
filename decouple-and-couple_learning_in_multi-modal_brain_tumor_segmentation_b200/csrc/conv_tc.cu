// Placeholder until the tcgen05 kernel lands: reports "unsupported" so the engine uses the fp32 kernels.
#include "conv_tc.cuh"

namespace dcl {
int tc_pack_weights(const float*, int cout, int cin, TcWeights* out) {
  out->dev = nullptr; out->cout = cout; out->cin = cin; out->bytes = 0;
  return 0;
}
bool tc_conv_supported(int, int, int, int) { return false; }
int launch_conv3d_k3_tc(const ConvSrc&, const ConvDst&, const TcWeights&, int, int, bool, cudaStream_t) {
  set_error("tensor-core convolution not built");
  return -2;
}
}  // namespace dcl
