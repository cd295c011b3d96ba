// Shared declarations of the dcl_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>

namespace dcl {

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const std::string& msg);
#define DCL_CUDA_OK(expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::dcl::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
      return -2;                                                                               \
    }                                                                                          \
  } while (0)

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2 };

// Fused InstanceNorm statistics: per channel (sum, sum of squares) accumulated by conv epilogues as 64-bit
// FIXED-POINT integers (integer atomics are associative, so the result does not depend on CTA order and the
// forward stays bit-reproducible).  sum uses 24 fractional bits, sum of squares 16.
typedef long long stat_t;
constexpr double STAT_SCALE_S = 16777216.0, STAT_SCALE_Q = 65536.0;
__device__ __forceinline__ void stat_add(stat_t* slot, int c, float sum, float sumsq) {
  atomicAdd(reinterpret_cast<unsigned long long*>(slot + 2 * c), (unsigned long long)__double2ll_rn((double)sum * STAT_SCALE_S));
  atomicAdd(reinterpret_cast<unsigned long long*>(slot + 2 * c + 1), (unsigned long long)__double2ll_rn((double)sumsq * STAT_SCALE_Q));
}
// InstanceNorm3d(affine=False, eps=1e-5, biased variance) from the accumulated sums
__device__ __forceinline__ void stat_mean_rstd(const stat_t* slot, int c, float inv_n, float* mean, float* rstd) {
  const double mu = (double)slot[2 * c] * (1.0 / STAT_SCALE_S) * (double)inv_n;
  const double var = (double)slot[2 * c + 1] * (1.0 / STAT_SCALE_Q) * (double)inv_n - mu * mu;
  *mean = (float)mu;
  *rstd = (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + 1e-5));
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LRELU) return v > 0.f ? v : 0.01f * v;
  return v;
}

// ---- convolution family (conv_simt.cu) -------------------------------------------------------
// Input descriptor of a 3-D convolution: channels [0,c0) come from x0, [c0,c0+c1) from x1 (the
// skip concat of DeUp_Cat / the edge branch is never materialised).  x0 may be a strided view
// (patch of a volume); x1 is dense.  Optional per-channel instance-norm + activation is applied
// while the tile is staged ("prologue fusion"); zero padding is applied after it.
struct ConvSrc {
  const float* x0;
  const float* x1;
  int c0, c1;
  int64_t s0c, s0d, s0h;   // element strides of x0 (w stride is 1)
  const float* mean;       // per input channel (c0+c1) or nullptr
  const float* rstd;
  int act;
  // alternative to mean/rstd: raw per-channel (sum, sum of squares) pairs accumulated by the producing
  // kernel's epilogue (ConvDst::stats); the consumer derives mean / rstd itself (no finalize launch)
  const stat_t* sums;
  float inv_n;             // 1 / (elements per channel)
};

struct ConvDst {
  float* y;                // dense (cout, od, oh, ow)
  const float* bias;       // cout or nullptr
  const float* out_scale;  // per output channel multiplier applied after bias (dropout3d) or nullptr
  const float* residual;   // dense, same shape as y, added last; or nullptr
  stat_t* stats;           // nullptr, or 2*cout fixed-point sums: += (sum, sum of squares) of the stored values
};

// Packed k3 weights: [cin][27][cout_pad], cout_pad = round_up(cout, 16).
int launch_conv3d_k3(const ConvSrc& src, const ConvDst& dst, const float* w_packed, int cout, int cout_pad,
                     int in_d, int in_h, int in_w, int stride, cudaStream_t st);
// Pointwise conv: packed weights [cin][cout]; optional softmax over cout (<= 4) in the epilogue.
int launch_conv1x1(const ConvSrc& src, const ConvDst& dst, const float* w_packed, int cout, int64_t spatial,
                   bool softmax, cudaStream_t st);
// ConvTranspose3d k2 s2: packed weights [cin][8][cout] (tap = (kd*2+kh)*2+kw).
int launch_convt_k2s2(const float* x, float* y, const float* w_packed, const float* bias, int cin, int cout,
                      int in_d, int in_h, int in_w, cudaStream_t st);

// ---- normalisation (norm.cu) ---------------------------------------------------------------
// accum: 2*channels doubles, zeroed by the launcher; mean/rstd: floats (biased variance, eps 1e-5).
int launch_instnorm_stats(const float* x, int channels, int64_t spatial, double* accum, float* mean, float* rstd,
                          cudaStream_t st);
// y = act((x - mean) * rstd) + residual
int launch_norm_act_res(const float* x, const float* mean, const float* rstd, int act, const float* residual,
                        float* y, int channels, int64_t spatial, cudaStream_t st);
// tokens[(d/p0,h/p1,w/p2)][(c,p0,p1,p2)] = act(norm(x)); optional dense copy of act(norm(x)) too.
int launch_norm_act_tokenise(const float* x, const float* mean, const float* rstd, int act, float* tokens,
                             float* dense_or_null, int channels, int grid, int p0, int p1, int p2, cudaStream_t st);
// y[c][d][h][w] = tokens[...] * class_token[(c,p0,p1,p2)]   (split_dim of a scaled token matrix)
int launch_scale_untokenise(const float* tokens, const float* class_token, float* y, int channels, int grid, int p0,
                            int p1, int p2, cudaStream_t st);

// ---- token path (token.cu) -----------------------------------------------------------------
constexpr int TOKEN_DIM = 512;
constexpr int TOP_NUM = 128;
constexpr int SEQ = TOP_NUM + 1;
// idx_out[0..128) = indices of the 128 largest <token, feats[i]> in descending score order.
int launch_select_topk(const float* token, const float* feats, int n_tokens, float* score_scratch, int* idx_out,
                       cudaStream_t st);
// seq[0] = class_token; seq[1+i] = feats[idx[i]] + pe_row
int launch_build_sequence(const float* class_token, const float* feats, const int* idx, const float* pe_row,
                          float* seq, cudaStream_t st);
int launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, int rows, cudaStream_t st);
// y[m][n] = sum_k x[m][k] * w[n][k] + bias[n]  (+gelu) (+residual[m][n]);  w is the PyTorch (N,K) matrix.
int launch_linear(const float* x, const float* w, const float* bias, const float* residual, float* y, int m, int n,
                  int k, bool gelu, cudaStream_t st);
// 8 heads x 64; q: (mq,512), kv: (mk,1024) = [K | V]; out (mq,512)
// out_blocked != nullptr: the result is written as bf16 [64][mq][8] (GEMM A operand) instead of fp32 `out`
// x3: out_blocked is split-fp16 ([64 hi chunks][64 lo chunks] x mq rows)
int launch_attention(const float* q, const float* kv, float* out, int mq, int mk, cudaStream_t st, void* out_blocked = nullptr,
                     bool x3 = false);
// feats[idx[i]] = rows[i]   (i < 128)
int launch_scatter_rows(float* feats, const int* idx, const float* rows, int row_stride, cudaStream_t st);
// y = a + b + c  (n elements)
int launch_add3(const float* a, const float* b, const float* c, float* y, int64_t n, cudaStream_t st);

struct Floats16 { float v[16]; };
// Per-patch arguments of a forward, kept in device memory so that the captured CUDA graph of the forward is the same
// for every patch: the first convolution reads its source view from here, the dropout scale sits in keep[16].
struct PatchDesc {
  const float* x;           // first element of the (4,128,128,128) view
  long long sc, sd, sh;     // element strides of (C, X, Y); the Z stride is 1
  float keep[16];
  float* probs;             // nullptr, or where this patch's probabilities go instead of the handle's buffer
  float* aux[12];           // destinations of the auxiliary heads of this patch (nullptr = head not requested)
};
// debug: dst[idx] = %globaltimer (ns); DCL_STAMPS=1 places these between the stages of the forward
int launch_stamp(unsigned long long* dst, int idx, cudaStream_t st);
// *dst = v, passed by value (one tiny launch per patch, outside the graph)
int launch_patch_desc(PatchDesc* dst, const PatchDesc& v, cudaStream_t st);
// dst[0..16) = v, passed by value (keeps the per-patch dropout scale off the memcpy path)
int launch_fill16(float* dst, const Floats16& v, cudaStream_t st);

// ---- aux heads (aux.cu) --------------------------------------------------------------------
// x: (2, g,g,g) logits -> y: (2, g*s, g*s, g*s) = softmax_c(trilinear_upsample(x, align_corners=False))
// desc != nullptr: the destination is desc->aux[slot] (read on the device: the launch is graph-replayable)
int launch_upsample_softmax2(const float* x, float* y, int g, int scale, cudaStream_t st, const PatchDesc* desc = nullptr,
                             int slot = 0);

// ---- stitch / label tail (stitch.cu) -------------------------------------------------------
struct StitchBox {      // one rectangular copy of the crop-and-overwrite plan
  int dst[3];           // destination origin (global)
  int ext[3];           // extent
  int src[3];           // source origin (patch local)
};
int launch_stitch_copy(const float* probs, float* out, const StitchBox& box, int X, int Y, int Zout, cudaStream_t st);
int launch_accumulate(const float* probs, const int start[3], int gaussian, float* acc, float* wsum, int X, int Y,
                      int Z, cudaStream_t st);
// gather form of the weighted stitch: patch i keeps its probabilities at patch_probs + i * 4 * 128^3
struct GatherPlan {
  static constexpr int MAX = 128;
  int n;
  int slot_planes;          // 128^3 planes per patch slot: 4, or 16 when the six final auxiliary heads travel along
  int start[MAX][3];
};
int launch_gather_finalize(const float* patch_probs, const GatherPlan& plan, int gaussian, int X, int Y, int Z,
                           float* probs_out, uint8_t* labels, const uint8_t* target, unsigned long long* counts,
                           cudaStream_t st);
// The same blend from one slot POINTER per patch (a slot may live in a peer GPU's memory, mapped through CUDA IPC and
// read over NVLink: the multi-GPU owner-computes exchange) restricted to the rows x in [x0, x1) of the volume.
// slots[i] = first float of patch i's (slot_planes x 128^3) slot; outputs are indexed like the whole volume.
struct GatherSlots { const float* ptr[GatherPlan::MAX]; };
int launch_gather_finalize_slots(const GatherSlots& slots, const GatherPlan& plan, int gaussian, int X, int Y, int Z, int x0,
                                 int x1, float* probs_out, uint8_t* labels, const uint8_t* target, unsigned long long* counts,
                                 cudaStream_t st);
// probs (4 x total) [+ wsum] -> labels / normalised probs / 13 counters over voxels [v0, v0+nvox)
int launch_finalize_labels(const float* acc, const float* wsum, int64_t total, int64_t v0, int64_t nvox,
                           float* probs_out, uint8_t* labels, const uint8_t* target, unsigned long long* counts,
                           cudaStream_t st);

// 8-flip TTA (predict_cls.py:180-203): dst = flip(src[..., :Zd]) along the selected axes; out (+)= softmax_c(flip(yf))
int launch_flip_volume(const float* src, float* dst, int X, int Y, int Zs, int Zd, int fx, int fy, int fz, cudaStream_t st);
int launch_tta_accumulate(const float* yf, float* out, int X, int Y, int Z, int fx, int fy, int fz, int first, int last,
                          cudaStream_t st);

extern thread_local int64_t g_launches;   // incremented by every launcher

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The forward is a chain of ~200 short kernels; the stream-ordered gap between two of them is ~2.7 us on B200
// (tools/ubench/cta_launch.cu).  Kernels launched through launch_pdl may be scheduled while their predecessor is
// still running; they call pdl_wait() first thing, which returns once the predecessor has completed and its
// writes are visible, so nothing is read or written early - only the launch latency is hidden.
// Measured: inside the replayed CUDA graph of the forward the gain is within noise (2.303 vs 2.314 ms per patch),
// so the attribute is OFF by default (DCL_PDL=1 switches it on); the kernels are identical either way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // DCL_PDL=1 in the environment switches the launch attribute on
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace dcl
