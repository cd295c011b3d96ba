// tcgen05 / TMEM implicit-GEMM 3x3x3 convolution (precision modes DCL_BF16X3 and DCL_BF16).
#pragma once
#include "common.cuh"

namespace dcl {

// Weights of one convolution as bf16 UMMA operand tiles (hi and lo halves), device memory.
struct TcWeights {
  void* dev = nullptr;
  int cout = 0, cin = 0;
  int64_t bytes = 0;
};

// Packs a PyTorch (cout, cin, 3,3,3) fp32 weight (host) into the kernel's B-operand layout.
int tc_pack_weights(const float* w_host, int cout, int cin, TcWeights* out);
// True when launch_conv3d_k3_tc handles a cubic g^3 input with these channel counts.
bool tc_conv_supported(int cin, int cout, int g, int stride, bool split);
// Same contract as launch_conv3d_k3 (dense single-source input, fused input norm/activation, bias,
// residual).  split = true: bf16x3 (hi*hi + hi*lo + lo*hi), false: plain bf16 operands.
int launch_conv3d_k3_tc(const ConvSrc& src, const ConvDst& dst, const TcWeights& w, int cout, int g, bool split,
                        cudaStream_t st);

}  // namespace dcl
