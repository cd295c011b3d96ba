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

// Packs a PyTorch (cout, cin, taps) fp32 weight (host) into the kernel's B-operand layout (taps = 27 | 1).
int tc_pack_weights(const float* w_host, int cout, int cin, int taps, TcWeights* out);
// fp32 NCDHW (two-source concat, fused norm + activation) -> bf16 channel-blocked [cin_pad/8][D][H][W][8]
int launch_prep_blocked(const ConvSrc& src, int in_d, int in_h, int in_w, void* out, cudaStream_t st);
// general implicit-GEMM convolution on a blocked bf16 input (taps = 27: kernel 3 pad 1, taps = 1: pointwise)
int launch_conv_gemm(const void* a_blocked, const TcWeights& w, const ConvDst& dst, int in_d, int in_h, int in_w,
                     int stride, int taps, cudaStream_t st);
// True when launch_conv3d_k3_tc handles a cubic g^3 input with these channel counts.
bool tc_conv_supported(int cin, int cout, int g, int stride, bool split);
// Same contract as launch_conv3d_k3 (dense single-source input, fused input norm/activation, bias,
// residual).  split = true: bf16x3 (hi*hi + hi*lo + lo*hi), false: plain bf16 operands.
int launch_conv3d_k3_tc(const ConvSrc& src, const ConvDst& dst, const TcWeights& w, int cout, int g, bool split,
                        cudaStream_t st);

// fp32 [rows][512] (optionally LayerNorm'ed) -> bf16 blocked [64][rows][8]
int launch_prep_rows(const float* x, const float* gamma, const float* beta, int rows, void* out, cudaStream_t st);
// y[m][n] = a[m][:] . w[n][:] + bias (+GELU) (+residual), a blocked bf16, w packed by tc_pack_weights(taps = 1)
int launch_linear_tc(const void* a_blocked, const void* w_packed, const float* bias, const float* residual, float* y,
                     int m, int n, int k, bool gelu, cudaStream_t st);

}  // namespace dcl
