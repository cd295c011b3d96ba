// tcgen05 / TMEM implicit-GEMM 3x3x3 convolution (precision modes DCL_F16X3 and DCL_BF16).
#pragma once
#include "common.cuh"

namespace dcl {

// Weights of one convolution as bf16 UMMA operand tiles (hi and lo halves), device memory.
struct TcWeights {
  void* dev = nullptr;
  int cout = 0, cin = 0;
  int64_t bytes = 0;
  int64_t lo_off = 0;   // split mode: byte offset of the "lo" image (same layout as the "hi" image at dev); 0 = plain bf16
  float out_mul = 1.f;  // split mode: the weights are stored times 2^k (so that their fp16 lo halves stay normal numbers);
                        // every accumulator read is multiplied by out_mul = 2^-k (exact)
  int layout = 0;       // 0 canonical [tap][cin/8][cout][8] (+ lo image); 1-3: rolling-kernel orders (tc_pack_weights)
  // lazily built copy in the slab kernel's streaming order for one tile width (launch_slab_conv)
  mutable void* slab_dev = nullptr;
  mutable int slab_ntile = 0;
};
void tc_free_weights(TcWeights* w);   // frees dev and slab_dev

// Packs a PyTorch (cout, cin, taps) fp32 weight (host) into the kernel's B-operand layout (taps = 27 | 1).
// roll_layout: the conv will run on the rolling kernel (stacked-kh weight order for 16-channel outputs)
// x3: split every weight into fp16 hi + lo images (precision mode DCL_F16X3)
int tc_pack_weights(const float* w_host, int cout, int cin, int taps, bool roll_layout, TcWeights* out, bool x3 = false);
// fp32 NCDHW (two-source concat, fused norm + activation) -> bf16 channel-blocked [cin_pad/8][D][H][W][8]
// x3: out holds 2 * cin_pad / 8 chunks, the hi planes followed by the lo planes
int launch_prep_blocked(const ConvSrc& src, int in_d, int in_h, int in_w, void* out, cudaStream_t st, bool x3 = false);
// general implicit-GEMM convolution on a blocked bf16 input (taps = 27: kernel 3 pad 1, taps = 1: pointwise)
int launch_conv_gemm(const void* a_blocked, const TcWeights& w, const ConvDst& dst, int in_d, int in_h, int in_w,
                     int stride, int taps, cudaStream_t st, bool x3 = false);
// ---- "B-format" activations of the bf16 pipeline: bf16, channel-blocked [C/8][spatial][8] ------------------
// Split-fp16 mode (DCL_F16X3, `x3` below): every B-format tensor is [2][C/8][spatial][8] - the fp16 "hi" planes
// followed by the fp16 "lo" planes (value = hi + lo, 22 significant bits); every MMA kernel forms
// a_hi*w_hi + a_lo*w_hi + a_hi*w_lo in its fp32 accumulator.
// Fused transform of a conv input: InstanceNorm (from raw sums or explicit mean/rstd; all null = none) + activation.
struct BNorm {
  const stat_t* sums = nullptr;
  float inv_n = 0.f;
  const float* mean = nullptr;
  const float* rstd = nullptr;
  int act = ACT_NONE;
};
struct RollArgs {
  const void* xb = nullptr;          // B-format input (cin = 16 | 32) ...
  const float* x4 = nullptr;         // ... or the fp32 NCDHW 4-channel strided view InitConv reads
  int64_t s4c = 0, s4d = 0, s4h = 0;
  const PatchDesc* desc = nullptr;   // ... or the same view described in device memory (graph-replayable forward)
  BNorm norm;
  const float* bias = nullptr;
  const float* out_scale = nullptr;
  const void* resb = nullptr;        // B-format residual
  void* yb = nullptr;                // B-format output
  stat_t* stats = nullptr;           // 2*cout fixed-point sums += (sum, sum of squares) of the outputs
  bool x3 = false;                   // split-fp16 tensors and weights
};
// General implicit-GEMM convolution / linear layer on B-format input (conv_gemm.cu).
struct GemmArgs {
  const void* a0 = nullptr; int c0 = 0;      // B-format source with c0 channels (multiple of 8)
  const void* a1 = nullptr;                  // optional second source holding the remaining input channels
  int D = 1, H = 1, W = 1, stride = 1, taps = 27;   // input dims; taps = 27 (k3 pad 1) | 1 (pointwise)
  const float* bias = nullptr;
  const float* out_scale = nullptr;
  int out_mode = 2;                          // 0 fp32 NCDHW, 1 fp32 row-major [m][cout], 2 B-format
  void* y = nullptr;
  const void* residual = nullptr;            // same format as y
  stat_t* stats = nullptr;                   // 2*cout fixed-point sums += (sum, sum of squares) of the outputs
  int gelu = 0;
  bool x3 = false;                           // split-fp16 sources / weights / B-format output and residual
};
int launch_gemm_conv(const GemmArgs& g, const TcWeights& w, cudaStream_t st);
// stride-1 3x3x3 conv with the input tile staged once ("slab" kernel, fused input norm); W in {16,32,64,128}
bool slab_conv_supported(int cin, int cout, int d, int h, int w, int stride, int taps);
int launch_slab_conv(const GemmArgs& g, const BNorm* norm, const TcWeights& w, cudaStream_t st);

// True when the rolling kernel handles a cubic g^3 stride-1 conv with these channel counts.
bool tc_conv_supported(int cin, int cout, int g, int stride, bool split);
int launch_roll_conv(const RollArgs& a, const TcWeights& w, int cout, int g, cudaStream_t st);

// Rolling stride-2 kernel for EnDown1 (16 -> 32 channels, 128^3 -> 64^3), B-format in / out (conv_s2.cu)
bool s2_roll_supported(int cin, int cout, int g);
int launch_s2_roll_conv(const void* xb, const TcWeights& w, const float* bias, void* yb, stat_t* stats, cudaStream_t st,
                        bool x3 = false);
bool s2_roll_supported_x3(int cin, int cout, int g);

// ---- HBM-bound B-format kernels (bf16_ops.cu) ---------------------------------------------------------------
// y = act(norm(x)) (+ res), all B-format
int launch_norm_act_b(const void* x, const BNorm& n, const void* res, void* y, int channels, int64_t spatial,
                      cudaStream_t st, bool x3 = false);
// B-format -> fp32 NCDHW
int launch_unblock(const void* x, float* y, int channels, int64_t spatial, cudaStream_t st, bool x3 = false);
// tokens = convert_dim(act(norm(x[chunk0*8 : chunk0*8 + channels]))) (+ dense fp32 NCDHW copy)
// x3_chunks: 0, or the chunk count (C_total / 8) of the split-fp16 tensor x (its lo planes start there)
int launch_tokenise_b(const void* x, const BNorm& n, int chunk0, float* tokens, float* dense_or_null, int channels,
                      int grid, int p0, int p1, int p2, cudaStream_t st, int x3_chunks = 0);
// y (B-format) = split_dim(tokens * class_token)
int launch_untokenise_b(const float* tokens, const float* class_token, void* y, int channels, int grid, int p0, int p1,
                        int p2, cudaStream_t st, bool x3 = false);
// DeUp_Cat as one kernel: mt [8][cin/2][cin], w3a [cin/2][cin/2], bt [8][cin/2] (composed on the host)
// x3: mt / w3a hold the hi image followed by the lo image
int launch_deup_fused_b(const void* x, const void* skip, const float* mt, const float* w3a, const float* bt, void* y,
                        int cin, int gi, cudaStream_t st, bool x3 = false, float out_mul = 1.f);
// probs (fp32 NCDHW, 4 classes) = softmax(endconv(x)), x B-format 16 channels
// norm != nullptr: the input is act(norm(x)) + res, i.e. the DeBlock tail is applied while loading (bit-identical to
// running launch_norm_act_b first, including its bf16 rounding)
int launch_endconv_softmax_b(const void* x, const float* w, const float* b, float* probs, int64_t spatial,
                             cudaStream_t st, const BNorm* norm = nullptr, const void* res = nullptr,
                             const PatchDesc* desc = nullptr, bool x3 = false);

// fp32 [rows][512] (optionally LayerNorm'ed) -> bf16 blocked [64][rows][8]
// x3: out = [64 hi chunks][64 lo chunks] x rows
int launch_prep_rows(const float* x, const float* gamma, const float* beta, int rows, void* out, cudaStream_t st, bool x3 = false);
int launch_prep_rows2(const float* x0, const float* g0, const float* b0, int rows0, void* out0, const float* x1, const float* g1,
                      const float* b1, int rows1, void* out1, cudaStream_t st, bool x3 = false);
// y[m][n] = a[m][:] . w[n][:] + bias (+GELU) (+residual), a blocked bf16, w packed by tc_pack_weights(taps = 1)
// y_blocked != nullptr: the result goes out as bf16 [n/8][m][8] (the next GEMM's A operand) instead of fp32 `y`
// x3: a_blocked / y_blocked are split-fp16, w_packed holds the hi image followed by the lo image (n * k * 2 bytes each)
int launch_linear_tc(const void* a_blocked, const void* w_packed, const float* bias, const float* residual, float* y,
                     int m, int n, int k, bool gelu, cudaStream_t st, void* y_blocked = nullptr, bool x3 = false,
                     float acc_mul = 1.f);

}  // namespace dcl
