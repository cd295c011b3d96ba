// Host side of the C ABI (include/dcl_b200.h): handle, weight repacking, the per-patch forward
// schedule and the sliding-window driver.  Every arithmetic step is a kernel from this library;
// there is no CPU compute path.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_fp16.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/dcl_b200.h"
#include "common.cuh"
#include "conv_tc.cuh"

namespace dcl {

static long long* g_trace_host_ptr = nullptr;
constexpr int TRACE_CAP_HOST = 4096;
cudaError_t trace_set_conv_tc(long long* p, int cta);
cudaError_t trace_set_conv_gemm(long long* p, int cta);

thread_local std::string g_error;
thread_local int64_t g_launches = 0;
void set_error(const std::string& msg) { g_error = msg; }
bool pdl_enabled() {
  static const bool on = getenv("DCL_PDL") != nullptr;
  return on;
}

#define DCL_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Weight catalogue: the 222 state_dict tensors of the reference.
// ---------------------------------------------------------------------------------------------
enum WKind { W_RAW, W_CONV3, W_CONV1, W_CONVT };
struct WSpec {
  std::string name;
  WKind kind;
  int cout, cin;       // conv kinds
  int64_t numel;
  bool aux;            // only needed when want_aux
};

static void add_conv(std::vector<WSpec>& v, const std::string& n, WKind k, int cout, int cin, bool aux = false) {
  int taps = k == W_CONV3 ? 27 : (k == W_CONVT ? 8 : 1);
  v.push_back({n + ".weight", k, cout, cin, (int64_t)cout * cin * taps, aux});
  v.push_back({n + ".bias", W_RAW, 0, 0, cout, aux});
}

static const char* const REGION_KEY[3] = {"01", "02", "04"};
static const char* const REGION_NUM[3] = {"1", "2", "4"};

static std::vector<WSpec> build_catalogue() {
  std::vector<WSpec> v;
  const int D = TOKEN_DIM;
  for (int r = 0; r < 3; ++r)
    v.push_back({std::string("label_") + REGION_KEY[r] + "_position_encoding.pe", W_RAW, 0, 0, 1024 * D, false});
  auto transformer = [&](const std::string& t) {
    std::string a = t + ".cross_attention_list.0.fn.";
    v.push_back({a + "norm.weight", W_RAW, 0, 0, D, false});
    v.push_back({a + "norm.bias", W_RAW, 0, 0, D, false});
    v.push_back({a + "norm2.weight", W_RAW, 0, 0, D, false});
    v.push_back({a + "norm2.bias", W_RAW, 0, 0, D, false});
    v.push_back({a + "fn.out_proj.weight", W_RAW, 0, 0, (int64_t)D * D, false});
    v.push_back({a + "fn.out_proj.bias", W_RAW, 0, 0, D, false});
    v.push_back({a + "fn.qkv.weight", W_RAW, 0, 0, (int64_t)3 * D * D, false});
    std::string f = t + ".cross_ffn_list.0.fn.";
    v.push_back({f + "norm.weight", W_RAW, 0, 0, D, false});
    v.push_back({f + "norm.bias", W_RAW, 0, 0, D, false});
    v.push_back({f + "fn.net.0.weight", W_RAW, 0, 0, (int64_t)D * D, false});
    v.push_back({f + "fn.net.0.bias", W_RAW, 0, 0, D, false});
    v.push_back({f + "fn.net.3.weight", W_RAW, 0, 0, (int64_t)D * D, false});
    v.push_back({f + "fn.net.3.bias", W_RAW, 0, 0, D, false});
  };
  for (int r = 0; r < 3; ++r) transformer(std::string("transformer_") + REGION_KEY[r]);
  v.push_back({"fusion_label_pos.pe", W_RAW, 0, 0, 1024 * D, false});
  transformer("fusion_transformer_1_2_4");
  for (int r = 0; r < 3; ++r) add_conv(v, std::string("conv_semantic_") + REGION_NUM[r], W_CONV3, 128, 256);
  for (int r = 0; r < 3; ++r) add_conv(v, std::string("conv_mid_fea_") + REGION_NUM[r], W_CONV3, 32, 96);
  const std::string u = "Unet_list.";
  add_conv(v, u + "InitConv.conv", W_CONV3, 16, 4);
  auto block = [&](const std::string& n, int c, bool aux = false) {
    add_conv(v, n + ".conv1", W_CONV3, c, c, aux);
    add_conv(v, n + ".conv2", W_CONV3, c, c, aux);
  };
  block(u + "EnBlock1", 16); block(u + "EnBlock1_1", 16);
  add_conv(v, u + "EnDown1.conv", W_CONV3, 32, 16);
  block(u + "EnBlock2_1", 32); block(u + "EnBlock2_2", 32);
  add_conv(v, u + "EnDown2.conv", W_CONV3, 64, 32);
  block(u + "EnBlock3_1", 64); block(u + "EnBlock3_2", 64);
  add_conv(v, u + "EnDown3.conv", W_CONV3, 128, 64);
  block(u + "EnBlock4_1", 128); block(u + "EnBlock4_2", 128);
  add_conv(v, u + "EnDown_4.conv", W_CONV3, 256, 128);
  const std::string d = "decoder.";
  add_conv(v, d + "down_channel", W_CONV1, 128, 256);
  block(d + "Enblock8_1", 128); block(d + "Enblock8_2", 128);
  auto up = [&](const std::string& n, int c) {   // c = in_channels, out = c/2
    add_conv(v, n + ".conv1", W_CONV1, c / 2, c);
    add_conv(v, n + ".conv2", W_CONVT, c / 2, c / 2);
    add_conv(v, n + ".conv3", W_CONV1, c / 2, c);
  };
  up(d + "DeUp4", 128); block(d + "DeBlock4", 64); block(d + "DeBlock4_1", 64);
  up(d + "DeUp3", 64);  block(d + "DeBlock3", 32); block(d + "DeBlock3_1", 32);
  up(d + "DeUp2", 32);  block(d + "DeBlock2", 16); block(d + "DeBlock2_1", 16);
  add_conv(v, d + "endconv", W_CONV1, 4, 16);
  auto sem_head = [&](const std::string& n) {
    for (int r = 0; r < 3; ++r) {
      add_conv(v, n + ".supervise_label_" + REGION_NUM[r], W_CONV3, 32, 128, true);
      add_conv(v, n + ".down_label_" + REGION_NUM[r], W_CONV3, 2, 32, true);
    }
  };
  auto edge_head = [&](const std::string& n) {
    for (int r = 0; r < 3; ++r) {
      add_conv(v, n + ".edge_supervise_label_" + REGION_NUM[r], W_CONV3, 8, 32, true);
      add_conv(v, n + ".edge_down_label_" + REGION_NUM[r], W_CONV3, 2, 8, true);
    }
  };
  sem_head("supervise_label"); edge_head("edge_supervise_label");
  sem_head("mid_supervise_label"); edge_head("mid_edge_supervise_label");
  for (int r = 0; r < 3; ++r) {
    v.push_back({std::string("e_token_") + REGION_KEY[r], W_RAW, 0, 0, D, false});
    v.push_back({std::string("s_token_") + REGION_KEY[r], W_RAW, 0, 0, D, false});
  }
  add_conv(v, "sum_fusion", W_CONV3, 256, 128);
  add_conv(v, "conv_64_to_32", W_CONV3, 32, 32);
  return v;
}

static const std::vector<WSpec>& catalogue() {
  static const std::vector<WSpec> c = build_catalogue();
  return c;
}

// ---------------------------------------------------------------------------------------------
struct ConvW {          // packed conv weights on the device
  float* w = nullptr;   // k3: [cin][27][cout_pad]   k1: [cin][cout]   convT: [cin][8][cout]
  float* b = nullptr;
  float* raw = nullptr; // PyTorch layout (used by the tensor-core path, which packs its own operand tiles)
  TcWeights tc;         // 16-bit operand tiles: bf16 (DCL_BF16) or fp16 hi + lo images (DCL_F16X3); built when precision != FP32
  int cout = 0, cin = 0, cout_pad = 0;
};

struct Transformer {
  float *n1w, *n1b, *n2w, *n2b, *wqkv, *wout, *bout, *fnw, *fnb, *w0, *b0, *w3, *b3;
  void *pq = nullptr, *pkv = nullptr, *pout = nullptr, *p0 = nullptr, *p3 = nullptr;   // 16-bit UMMA operand tiles (tensor-core modes)
  float mq = 1.f, mkv = 1.f, mout = 1.f, m0 = 1.f, m3 = 1.f;                           // their accumulator multipliers (split mode)
};

struct StatSlot { float* mean; float* rstd; };

// scratch of one coupler (token path); three copies so the three regions can run on concurrent streams
struct TokScratch {
  float* score;
  float* seq[4];
  float *ln_a, *ln_b, *qbuf, *kvbuf, *obuf, *eqs, *sqe, *cross, *ffn_ln, *ffn_h;
  void *tok_a, *tok_b;
};

}  // namespace dcl

using namespace dcl;

struct dcl_handle {
  int device = -1;
  dcl_config cfg{};
  bool ready = false;
  int64_t launches = 0;
  std::vector<void*> allocs;
  int64_t alloc_bytes = 0;
  bool dry_run = false;    // size the workspace without touching the device

  std::map<std::string, std::vector<float>> host_w;   // raw state_dict tensors as set
  std::map<std::string, float*> dev_raw;               // device copies (PyTorch layout)
  std::map<std::string, ConvW> conv;                   // by module name (without .weight)
  ConvW edge_merged, sem_merged;
  Transformer tr[4];
  float *e_tok[3], *s_tok[3], *pe[4];

  // ---- activations (all fp32, dense NCDHW unless noted) ----
  float *l_t0[4], *l_a[4], *l_t1[4], *l_x[4];   // encoder levels 0..3: input, scratch, block-1 out, block-2 out (skip)
  float* x4;                                     // (256,16^3)
  float *e_down, *e_raw, *s_raw;                 // (32,32^3), (96,32^3), (384,16^3)
  float *E[3], *S[3];                            // token matrices (2048,512) / (1024,512)
  float *edge_dense[3], *sem_dense[3];           // dense IN+LeakyReLU outputs (stages / aux)
  float *sup_edge[3], *sup_sem[3];               // aux-head inputs after the couplers
  float *aux_t1, *aux_t2;                        // aux conv scratch
  float* score;                                  // (2048)
  int* topk;                                     // (13,128)
  float* seq[4];                                 // 4 x (129,512)
  float *ln_a, *ln_b, *qbuf, *kvbuf, *obuf;      // attention scratch
  float *eqs, *sqe, *cross, *ffn_ln, *ffn_h;
  float* coupler_out[4];                         // (258,512) x3, (129,512)
  float *f_tok, *f_fea, *fused_dense, *enc;
  float *d8_0, *d8_a, *d8_b, *d8_1, *d8_2;
  float *up_u1[3], *up_u2[3], *dl_in[3], *dl_a[3], *dl_b[3], *dl_1[3], *dl_2[3];   // decoder levels (32^3, 64^3, 128^3)
  float* probs;                                  // (4,128^3)
  float* keep_dev;                               // (16)  (bf16 mode: aliases patch_desc->keep)
  PatchDesc* patch_desc = nullptr;               // per-patch arguments of the graph-replayed bf16 forward
  unsigned long long* stamps = nullptr;          // debug (DCL_STAMPS=1): %globaltimer between the stages of the forward
  struct FwdGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0; int eager_runs = 0; };
  std::map<int, FwdGraph> fwd_graphs;            // captured tensor-core forwards, keyed by the mask of requested auxiliary heads
  int fwd_eager_runs = 0;
  bool fwd_graph_off = false;
  double* stat_accum;                            // (2*512)
  void* blk = nullptr;                           // bf16 channel-blocked conv input (DCL_BF16 only)
  void *tok_a = nullptr, *tok_b = nullptr;       // bf16 blocked token matrices feeding the linear GEMMs
  TokScratch ts[6];                              // [0] aliases the buffers above; region r works in ts[2r], ts[2r+1]
  cudaStream_t aux_stream[3] = {nullptr, nullptr, nullptr};   // [2] = capture stream of the forward graph
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  cudaStream_t tok_stream[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // coupler lanes 1..5 (lane 0 = caller's stream)
  cudaEvent_t ev_tok[3][3] = {}, ev_tok_join[5] = {};                           // per region: A1 done, A2 done, A4 done
  // ---- bf16 pipeline (DCL_BF16): B-format activations = bf16 [C/8][spatial][8] ----
  void *b_t0[4], *b_a[4], *b_t1[4], *b_x[4];     // encoder levels (16@128, 32@64, 64@32, 128@16)
  void *b_x4, *b_edown, *b_eraw, *b_sraw, *b_fused, *b_enc, *b_nrm;
  void *b_d8[5];                                 // d8_0, a, b, 1, 2
  void *b_dl[3][5];                              // decoder levels: in, a, b, 1, 2
  stat_t* stat_arena = nullptr;                  // STAT_SLOTS x 1024 fixed-point sums, zeroed once per forward
  int stat_used = 0;
  struct DeUpW { float *mt, *w3a, *bt; float out_mul = 1.f; } deup[3];     // mt / w3a: bf16 rows padded by 8 elements (the kernel's smem image)
  float *end_w = nullptr, *end_b = nullptr;
  struct BStage { const void* p; int c; int64_t spatial; };
  std::map<std::string, BStage> bstages;
  std::vector<StatSlot> stat_slots;
  int stat_next = 0;

  // ---- volume level ----
  float* vol_probs = nullptr;  int64_t vol_probs_cap = 0;
  float* vol_wsum = nullptr;   int64_t vol_wsum_cap = 0;
  float* gather_buf = nullptr; int64_t gather_cap = 0;      // gather-form stitch: one probability slot per patch
  float* shard_buf = nullptr;  int64_t shard_cap = 0;       // multi-GPU: this rank's slots, exported to the peers (CUDA IPC)
  float* tta_vol = nullptr;    int64_t tta_vol_cap = 0;      // flipped copy of the volume (TTA)
  float* tta_sum = nullptr;    int64_t tta_sum_cap = 0;      // running sum of the un-flipped softmaxes (TTA)
  float* stage_vol = nullptr;  int64_t stage_vol_cap = 0;
  float* stage_probs = nullptr; int64_t stage_probs_cap = 0;
  uint8_t* stage_labels = nullptr; uint8_t* stage_target = nullptr; int64_t stage_lab_cap = 0;
  unsigned long long* counts_dev = nullptr;
  // host entry point: the volume is uploaded on a copy stream in three x-slabs, patches wait only for the slab they need
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_up[3] = {nullptr, nullptr, nullptr}, ev_st = nullptr;
  int up_bound[3] = {0, 0, 0};      // slab k has landed when x < up_bound[k] is on the device; 0 = no staged upload
  bool up_active = false;
  // two patches in flight (bf16 sliding window): a twin handle (own workspace, graph and weights copy) runs every other
  // patch on a second stream; the overlap accumulates stay in patch order through an event chain
  static constexpr int MAX_LANES = 4;
  dcl_handle* twin[MAX_LANES - 1] = {nullptr, nullptr, nullptr};
  int lanes = 0;                                   // lanes built so far (0 = none)
  cudaStream_t lane_stream[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_lane_fork = nullptr, ev_lane_acc[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};

  std::map<std::string, std::pair<const float*, int64_t>> stages;

  // ---- per-class event timing (dcl_profile_*) ----
  struct ProfRec { cudaEvent_t a, b; int cls, cls2; double work; };
  bool profiling = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t prof_begin(cudaStream_t st) {
    cudaEvent_t e;
    if (!event_pool.empty()) { e = event_pool.back(); event_pool.pop_back(); } else cudaEventCreate(&e);
    cudaEventRecord(e, st);
    return e;
  }
  void prof_end(cudaEvent_t a, int cls, double work, cudaStream_t st, int cls2 = -1) {
    cudaEvent_t b;
    if (!event_pool.empty()) { b = event_pool.back(); event_pool.pop_back(); } else cudaEventCreate(&b);
    cudaEventRecord(b, st);
    prof.push_back({a, b, cls, cls2, work});
  }
  // scoped helper for the kernel classes that are only timed (no work figure)
  struct ProfScope {
    dcl_handle* h; cudaStream_t st; int cls; cudaEvent_t ev;
    ProfScope(dcl_handle* h_, cudaStream_t st_, int cls_) : h(h_), st(st_), cls(cls_), ev(h_->profiling ? h_->prof_begin(st_) : nullptr) {}
    ~ProfScope() { if (ev) h->prof_end(ev, cls, 0.0, st); }
  };
};

namespace dcl {

// tensor-core pipeline (B-format activations, tcgen05 kernels): DCL_BF16 and the split-fp16 mode DCL_F16X3
static inline bool is_tc(const dcl_handle* h) { return h->cfg.precision == DCL_BF16 || h->cfg.precision == DCL_F16X3; }
static inline bool is_x3(const dcl_handle* h) { return h->cfg.precision == DCL_F16X3; }

static int dev_alloc(dcl_handle* h, void** p, int64_t bytes) {
  h->alloc_bytes += (bytes + 255) / 256 * 256;
  if (h->dry_run) { *p = nullptr; return 0; }
  DCL_CUDA_OK(cudaMalloc(p, (size_t)bytes));
  h->allocs.push_back(*p);
  return 0;
}
static int falloc(dcl_handle* h, float** p, int64_t numel) { return dev_alloc(h, (void**)p, numel * 4); }

static const int64_t P3 = 128LL * 128 * 128;
static const int LVL_C[4] = {16, 32, 64, 128};
static const int LVL_G[4] = {128, 64, 32, 16};
static const int STAT_SLOTS = 64;

static int allocate_workspace(dcl_handle* h) {
  // the dense fp32 NCDHW activations of the convolutional path exist only in the FFMA mode: the tensor-core modes keep
  // those tensors in B-format (below) - 2.0 GB less per handle, and a volume call runs three handles (lanes)
  const bool dense = !is_tc(h);
  for (int l = 0; l < 4; ++l) {
    h->l_t0[l] = h->l_a[l] = h->l_t1[l] = h->l_x[l] = nullptr;
    if (!dense) continue;
    int64_t n = (int64_t)LVL_C[l] * LVL_G[l] * LVL_G[l] * LVL_G[l];
    DCL_TRY(falloc(h, &h->l_t0[l], n)); DCL_TRY(falloc(h, &h->l_a[l], n));
    DCL_TRY(falloc(h, &h->l_t1[l], n)); DCL_TRY(falloc(h, &h->l_x[l], n));
  }
  const int64_t g16 = 16 * 16 * 16, g32 = 32 * 32 * 32;
  h->x4 = h->e_down = h->e_raw = h->s_raw = nullptr;
  if (dense) {
    DCL_TRY(falloc(h, &h->x4, 256 * g16));
    DCL_TRY(falloc(h, &h->e_down, 32 * g32)); DCL_TRY(falloc(h, &h->e_raw, 96 * g32));
    DCL_TRY(falloc(h, &h->s_raw, 384 * g16));
  }
  for (int r = 0; r < 3; ++r) {
    DCL_TRY(falloc(h, &h->E[r], 2048 * 512)); DCL_TRY(falloc(h, &h->S[r], 1024 * 512));
    DCL_TRY(falloc(h, &h->edge_dense[r], 32 * g32)); DCL_TRY(falloc(h, &h->sem_dense[r], 128 * g16));
    DCL_TRY(falloc(h, &h->sup_edge[r], 32 * g32)); DCL_TRY(falloc(h, &h->sup_sem[r], 128 * g16));
  }
  DCL_TRY(falloc(h, &h->aux_t1, 32 * g32)); DCL_TRY(falloc(h, &h->aux_t2, 2 * g32));
  DCL_TRY(falloc(h, &h->score, 2048));
  DCL_TRY(dev_alloc(h, (void**)&h->topk, 13 * 128 * sizeof(int)));
  for (int i = 0; i < 4; ++i) DCL_TRY(falloc(h, &h->seq[i], SEQ * 512));
  DCL_TRY(falloc(h, &h->ln_a, 258 * 512)); DCL_TRY(falloc(h, &h->ln_b, 258 * 512));
  DCL_TRY(falloc(h, &h->qbuf, 258 * 512)); DCL_TRY(falloc(h, &h->kvbuf, 258 * 1024));
  DCL_TRY(falloc(h, &h->obuf, 258 * 512));
  DCL_TRY(falloc(h, &h->eqs, SEQ * 512)); DCL_TRY(falloc(h, &h->sqe, SEQ * 512));
  DCL_TRY(falloc(h, &h->cross, 258 * 512)); DCL_TRY(falloc(h, &h->ffn_ln, 258 * 512));
  DCL_TRY(falloc(h, &h->ffn_h, 258 * 512));
  for (int i = 0; i < 4; ++i) DCL_TRY(falloc(h, &h->coupler_out[i], 258 * 512));
  DCL_TRY(falloc(h, &h->f_tok, 512)); DCL_TRY(falloc(h, &h->f_fea, 1024 * 512));
  h->fused_dense = h->enc = h->d8_0 = h->d8_a = h->d8_b = h->d8_1 = h->d8_2 = nullptr;
  if (dense) {
    DCL_TRY(falloc(h, &h->fused_dense, 128 * g16)); DCL_TRY(falloc(h, &h->enc, 256 * g16));
    DCL_TRY(falloc(h, &h->d8_0, 128 * g16)); DCL_TRY(falloc(h, &h->d8_a, 128 * g16));
    DCL_TRY(falloc(h, &h->d8_b, 128 * g16)); DCL_TRY(falloc(h, &h->d8_1, 128 * g16));
    DCL_TRY(falloc(h, &h->d8_2, 128 * g16));
  }
  for (int l = 0; l < 3; ++l) {   // decoder level l: channels 64/32/16 at 32/64/128
    h->up_u1[l] = h->up_u2[l] = h->dl_in[l] = h->dl_a[l] = h->dl_b[l] = h->dl_1[l] = h->dl_2[l] = nullptr;
    if (!dense) continue;
    int c = 64 >> l, g = 32 << l;
    int64_t n = (int64_t)c * g * g * g;
    DCL_TRY(falloc(h, &h->up_u1[l], n / 8)); DCL_TRY(falloc(h, &h->up_u2[l], n));
    DCL_TRY(falloc(h, &h->dl_in[l], n)); DCL_TRY(falloc(h, &h->dl_a[l], n)); DCL_TRY(falloc(h, &h->dl_b[l], n));
    DCL_TRY(falloc(h, &h->dl_1[l], n)); DCL_TRY(falloc(h, &h->dl_2[l], n));
  }
  DCL_TRY(falloc(h, &h->probs, 4 * P3));
  if (is_tc(h)) {
    const int64_t E = is_x3(h) ? 4 : 2;      // bytes per element of a B-format tensor (split-fp16: hi + lo planes)
    DCL_TRY(dev_alloc(h, &h->blk, 2 * 1024 * 1024 * E));      // blocked input of an auxiliary-head convolution (<= 128 ch @ 16^3, 32 ch @ 32^3)
    DCL_TRY(dev_alloc(h, &h->tok_a, 258 * 512 * E));
    DCL_TRY(dev_alloc(h, &h->tok_b, 258 * 512 * E));
    for (int l = 0; l < 4; ++l) {
      int64_t bytes = (int64_t)LVL_C[l] * LVL_G[l] * LVL_G[l] * LVL_G[l] * E;
      DCL_TRY(dev_alloc(h, &h->b_t0[l], bytes)); DCL_TRY(dev_alloc(h, &h->b_a[l], bytes));
      DCL_TRY(dev_alloc(h, &h->b_t1[l], bytes)); DCL_TRY(dev_alloc(h, &h->b_x[l], bytes));
    }
    const int64_t v16 = 16 * 16 * 16, v32 = 32 * 32 * 32;
    DCL_TRY(dev_alloc(h, &h->b_x4, 256 * v16 * E)); DCL_TRY(dev_alloc(h, &h->b_edown, 32 * v32 * E));
    DCL_TRY(dev_alloc(h, &h->b_eraw, 96 * v32 * E)); DCL_TRY(dev_alloc(h, &h->b_sraw, 384 * v16 * E));
    DCL_TRY(dev_alloc(h, &h->b_fused, 128 * v16 * E)); DCL_TRY(dev_alloc(h, &h->b_enc, 256 * v16 * E));
    DCL_TRY(dev_alloc(h, &h->b_nrm, 32 * (P3 / 8) * E));
    for (int i = 0; i < 5; ++i) DCL_TRY(dev_alloc(h, &h->b_d8[i], 128 * v16 * E));
    for (int l = 0; l < 3; ++l) {
      int c = 64 >> l, g = 32 << l;
      for (int i = 0; i < 5; ++i) DCL_TRY(dev_alloc(h, &h->b_dl[l][i], (int64_t)c * g * g * g * E));
    }
    DCL_TRY(dev_alloc(h, (void**)&h->stat_arena, (int64_t)STAT_SLOTS * 1024 * sizeof(stat_t)));
  }
  if (is_tc(h)) {
    DCL_TRY(dev_alloc(h, (void**)&h->stamps, 16 * sizeof(unsigned long long)));
    DCL_TRY(dev_alloc(h, (void**)&h->patch_desc, sizeof(PatchDesc)));
    h->keep_dev = h->dry_run ? nullptr : h->patch_desc->keep;
  } else {
    DCL_TRY(falloc(h, &h->keep_dev, 16));
  }
  DCL_TRY(dev_alloc(h, (void**)&h->stat_accum, 2 * 512 * sizeof(double)));
  if (!h->dry_run) DCL_CUDA_OK(cudaMemset(h->stat_accum, 0, 2 * 512 * sizeof(double)));
  h->stat_slots.resize(48);
  for (auto& s : h->stat_slots) { DCL_TRY(falloc(h, &s.mean, 512)); DCL_TRY(falloc(h, &s.rstd, 512)); }
  DCL_TRY(dev_alloc(h, (void**)&h->counts_dev, 16 * sizeof(unsigned long long)));
  h->ts[0] = TokScratch{h->score, {h->seq[0], h->seq[1], h->seq[2], h->seq[3]}, h->ln_a, h->ln_b, h->qbuf, h->kvbuf, h->obuf,
                        h->eqs, h->sqe, h->cross, h->ffn_ln, h->ffn_h, h->tok_a, h->tok_b};
  for (int i = 1; i < 6; ++i) {
    TokScratch& t = h->ts[i];
    DCL_TRY(falloc(h, &t.score, 2048));
    for (int k = 0; k < 4; ++k) DCL_TRY(falloc(h, &t.seq[k], SEQ * 512));
    DCL_TRY(falloc(h, &t.ln_a, 258 * 512)); DCL_TRY(falloc(h, &t.ln_b, 258 * 512));
    DCL_TRY(falloc(h, &t.qbuf, 258 * 512)); DCL_TRY(falloc(h, &t.kvbuf, 258 * 1024));
    DCL_TRY(falloc(h, &t.obuf, 258 * 512));
    DCL_TRY(falloc(h, &t.eqs, SEQ * 512)); DCL_TRY(falloc(h, &t.sqe, SEQ * 512));
    DCL_TRY(falloc(h, &t.cross, 258 * 512)); DCL_TRY(falloc(h, &t.ffn_ln, 258 * 512));
    DCL_TRY(falloc(h, &t.ffn_h, 258 * 512));
    t.tok_a = t.tok_b = nullptr;
    if (is_tc(h)) {
      DCL_TRY(dev_alloc(h, &t.tok_a, 258 * 512 * (is_x3(h) ? 4 : 2)));
      DCL_TRY(dev_alloc(h, &t.tok_b, 258 * 512 * (is_x3(h) ? 4 : 2)));
    }
  }
  return 0;
}

static int64_t workspace_estimate(const dcl_config* cfg) {
  dcl_handle tmp;
  if (cfg) tmp.cfg = *cfg;
  tmp.dry_run = true;
  if (allocate_workspace(&tmp) != 0) return -1;
  return tmp.alloc_bytes;
}

// ---- weight upload / packing ----------------------------------------------------------------
static int upload(dcl_handle* h, const std::vector<float>& v, float** out) {
  DCL_TRY(falloc(h, out, (int64_t)v.size()));
  DCL_CUDA_OK(cudaMemcpy(*out, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

// srcs: list of (weight, bias) host tensors concatenated along cout.
static int pack_conv(dcl_handle* h, WKind kind, const std::vector<const std::vector<float>*>& ws,
                     const std::vector<const std::vector<float>*>& bs, int cout_each, int cin, ConvW* out,
                     bool stride1 = true) {
  const int parts = (int)ws.size();
  const int cout = cout_each * parts;
  out->cout = cout; out->cin = cin;
  if (out->tc.slab_dev) { cudaFree(out->tc.slab_dev); out->tc.slab_dev = nullptr; out->tc.slab_ntile = 0; }   // stale repack
  std::vector<float> bias((size_t)cout);
  for (int p = 0; p < parts; ++p)
    for (int c = 0; c < cout_each; ++c) bias[p * cout_each + c] = (*bs[p])[c];
  std::vector<float> packed, raw;
  if (kind == W_CONV3) {
    out->cout_pad = (cout + 15) / 16 * 16;
    packed.assign((size_t)cin * 27 * out->cout_pad, 0.f);
    raw.resize((size_t)cout * cin * 27);
    for (int p = 0; p < parts; ++p)
      for (int co = 0; co < cout_each; ++co)
        for (int ci = 0; ci < cin; ++ci)
          for (int t = 0; t < 27; ++t) {
            float v = (*ws[p])[((size_t)co * cin + ci) * 27 + t];
            packed[((size_t)ci * 27 + t) * out->cout_pad + p * cout_each + co] = v;
            raw[((size_t)(p * cout_each + co) * cin + ci) * 27 + t] = v;
          }
  } else if (kind == W_CONV1) {
    out->cout_pad = cout;
    packed.resize((size_t)cin * cout);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci) packed[(size_t)ci * cout + co] = (*ws[0])[(size_t)co * cin + ci];
  } else {   // ConvTranspose3d weight is (Cin, Cout, 2,2,2)
    out->cout_pad = cout;
    packed.resize((size_t)cin * 8 * cout);
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co)
        for (int t = 0; t < 8; ++t) packed[((size_t)ci * 8 + t) * cout + co] = (*ws[0])[((size_t)ci * cout + co) * 8 + t];
  }
  DCL_TRY(upload(h, packed, &out->w));
  DCL_TRY(upload(h, bias, &out->b));
  if (is_tc(h) && (kind == W_CONV3 || kind == W_CONV1)) {
    if (kind == W_CONV1) raw = *ws[0];   // (cout, cin)
    DCL_TRY(tc_pack_weights(raw.data(), cout, cin, kind == W_CONV3 ? 27 : 1,
                            kind == W_CONV3 && stride1 && tc_conv_supported(cin, cout, cin == 32 ? 64 : 128, 1, is_x3(h)), &out->tc,
                            is_x3(h)));
    h->allocs.push_back(out->tc.dev);
  }
  return 0;
}

static bool needed(const dcl_handle* h, const WSpec& s) { return !s.aux || h->cfg.want_aux; }

static int prepare(dcl_handle* h) {
  for (auto& kv : h->fwd_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);      // weights are about to move
  h->fwd_graphs.clear();
  h->fwd_eager_runs = 0;
  for (const auto& s : catalogue())
    if (needed(h, s) && !h->host_w.count(s.name)) {
      set_error("weight not set: " + s.name);
      return DCL_ERR_WEIGHTS;
    }
  auto W = [&](const std::string& n) -> const std::vector<float>* { return &h->host_w.at(n); };
  std::map<std::string, int> conv_modules;
  for (const auto& s : catalogue())
    if (s.kind != W_RAW) conv_modules[s.name.substr(0, s.name.size() - 7)] = 1;
  for (const auto& s : catalogue()) {
    if (!needed(h, s)) continue;
    if (s.kind == W_RAW) {
      if (s.name.size() > 5 && s.name.compare(s.name.size() - 5, 5, ".bias") == 0 &&
          conv_modules.count(s.name.substr(0, s.name.size() - 5)))
        continue;   // conv bias: packed together with its weight
      if (s.numel == 1024 * TOKEN_DIM) {   // positional-encoding buffer: only row 0 is ever added
        std::vector<float> row(h->host_w.at(s.name).begin(), h->host_w.at(s.name).begin() + TOKEN_DIM);
        DCL_TRY(upload(h, row, &h->dev_raw[s.name]));
      } else {
        DCL_TRY(upload(h, h->host_w.at(s.name), &h->dev_raw[s.name]));
      }
    } else {
      std::string mod = s.name.substr(0, s.name.size() - 7);
      // conv_64_to_32 is the one 32->32 convolution with stride 2: it runs on the GEMM kernel, not the rolling one
      DCL_TRY(pack_conv(h, s.kind, {W(s.name)}, {W(mod + ".bias")}, s.cout, s.cin, &h->conv[mod], mod != "conv_64_to_32"));
    }
  }
  DCL_TRY(pack_conv(h, W_CONV3, {W("conv_mid_fea_1.weight"), W("conv_mid_fea_2.weight"), W("conv_mid_fea_4.weight")},
                    {W("conv_mid_fea_1.bias"), W("conv_mid_fea_2.bias"), W("conv_mid_fea_4.bias")}, 32, 96,
                    &h->edge_merged));
  DCL_TRY(pack_conv(h, W_CONV3,
                    {W("conv_semantic_1.weight"), W("conv_semantic_2.weight"), W("conv_semantic_4.weight")},
                    {W("conv_semantic_1.bias"), W("conv_semantic_2.bias"), W("conv_semantic_4.bias")}, 128, 256,
                    &h->sem_merged));
  auto R = [&](const std::string& n) { return h->dev_raw.at(n); };
  const char* tnames[4] = {"transformer_01", "transformer_02", "transformer_04", "fusion_transformer_1_2_4"};
  for (int i = 0; i < 4; ++i) {
    std::string a = std::string(tnames[i]) + ".cross_attention_list.0.fn.";
    std::string f = std::string(tnames[i]) + ".cross_ffn_list.0.fn.";
    Transformer& t = h->tr[i];
    t.n1w = R(a + "norm.weight"); t.n1b = R(a + "norm.bias"); t.n2w = R(a + "norm2.weight"); t.n2b = R(a + "norm2.bias");
    t.wqkv = R(a + "fn.qkv.weight"); t.wout = R(a + "fn.out_proj.weight"); t.bout = R(a + "fn.out_proj.bias");
    t.fnw = R(f + "norm.weight"); t.fnb = R(f + "norm.bias");
    t.w0 = R(f + "fn.net.0.weight"); t.b0 = R(f + "fn.net.0.bias");
    t.w3 = R(f + "fn.net.3.weight"); t.b3 = R(f + "fn.net.3.bias");
    if (is_tc(h)) {
      const std::vector<float>& qkv = h->host_w.at(a + "fn.qkv.weight");
      auto pack = [&](const float* w, int n, void** out, float* mul) -> int {
        TcWeights tw;
        DCL_TRY(tc_pack_weights(w, n, 512, 1, false, &tw, is_x3(h)));
        h->allocs.push_back(tw.dev);
        *out = tw.dev;
        *mul = tw.out_mul;
        return 0;
      };
      DCL_TRY(pack(qkv.data(), 512, &t.pq, &t.mq));
      DCL_TRY(pack(qkv.data() + 512 * 512, 1024, &t.pkv, &t.mkv));
      DCL_TRY(pack(h->host_w.at(a + "fn.out_proj.weight").data(), 512, &t.pout, &t.mout));
      DCL_TRY(pack(h->host_w.at(f + "fn.net.0.weight").data(), 512, &t.p0, &t.m0));
      DCL_TRY(pack(h->host_w.at(f + "fn.net.3.weight").data(), 512, &t.p3, &t.m3));
    }
  }
  for (int r = 0; r < 3; ++r) {
    h->e_tok[r] = R(std::string("e_token_") + REGION_KEY[r]);
    h->s_tok[r] = R(std::string("s_token_") + REGION_KEY[r]);
    h->pe[r] = R(std::string("label_") + REGION_KEY[r] + "_position_encoding.pe");
  }
  h->pe[3] = R("fusion_label_pos.pe");
  if (is_tc(h)) {
    // DeUp_Cat composed into one linear map per transposed-conv tap (see bf16_ops.cu)
    const char* upn[3] = {"decoder.DeUp4", "decoder.DeUp3", "decoder.DeUp2"};
    for (int l = 0; l < 3; ++l) {
      const int C = 128 >> l, CH = C / 2;
      const std::string n = upn[l];
      const std::vector<float>&w1 = h->host_w.at(n + ".conv1.weight"), &b1 = h->host_w.at(n + ".conv1.bias"),
                               &wt = h->host_w.at(n + ".conv2.weight"), &bt = h->host_w.at(n + ".conv2.bias"),
                               &w3 = h->host_w.at(n + ".conv3.weight"), &b3 = h->host_w.at(n + ".conv3.bias");
      std::vector<float> mt((size_t)8 * CH * C), w3a((size_t)CH * CH), btc((size_t)8 * CH);
      std::vector<double> tmp((size_t)CH * C), tb(CH);
      for (int o = 0; o < CH; ++o)
        for (int sidx = 0; sidx < CH; ++sidx) w3a[(size_t)o * CH + sidx] = w3[(size_t)o * C + sidx];   // skip comes first in the cat
      for (int t = 0; t < 8; ++t) {
        // tmp[j][c] = sum_i wt[i][j][t] * w1[i][c];  tb[j] = sum_i wt[i][j][t] * b1[i] + bt[j]
        for (int j = 0; j < CH; ++j) {
          double bj = bt[j];
          for (int c = 0; c < C; ++c) tmp[(size_t)j * C + c] = 0.0;
          for (int i = 0; i < CH; ++i) {
            const double wv = wt[((size_t)i * CH + j) * 8 + t];
            bj += wv * b1[i];
            for (int c = 0; c < C; ++c) tmp[(size_t)j * C + c] += wv * w1[(size_t)i * C + c];
          }
          tb[j] = bj;
        }
        for (int o = 0; o < CH; ++o) {
          double bo = b3[o];
          for (int j = 0; j < CH; ++j) bo += (double)w3[(size_t)o * C + CH + j] * tb[j];
          btc[(size_t)t * CH + o] = (float)bo;
          for (int c = 0; c < C; ++c) {
            double a = 0.0;
            for (int j = 0; j < CH; ++j) a += (double)w3[(size_t)o * C + CH + j] * tmp[(size_t)j * C + c];
            mt[((size_t)t * CH + o) * C + c] = (float)a;
          }
        }
      }
      // the kernel keeps these as bf16 with rows padded by 8 elements (conflict-free fragment reads): store exactly that
      // image, so its prologue is a straight 16-byte copy instead of thousands of scalar loads + conversions per CTA
      // (split-fp16: the hi image is followed by the lo image = bf16(w - hi))
      const bool x3 = is_x3(h);
      auto to_bf16_padded = [x3](const std::vector<float>& src, int rows, int cols) {
        const size_t image = (size_t)rows * (cols + 8);                  // bf16 elements
        std::vector<float> out((x3 ? 2 : 1) * image / 2, 0.f);           // 2 bf16 per float slot
        uint16_t* o = reinterpret_cast<uint16_t*>(out.data());
        auto rn = [](float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); };
        for (int r = 0; r < rows; ++r)
          for (int c = 0; c < cols; ++c) {
            const float f = src[(size_t)r * cols + c];
            if (x3) {        // split mode: fp16 hi image, fp16 lo image
              const __half hh = __float2half_rn(f);
              const __half hl = __float2half_rn(f - __half2float(hh));
              memcpy(&o[(size_t)r * (cols + 8) + c], &hh, 2);
              memcpy(&o[image + (size_t)r * (cols + 8) + c], &hl, 2);
            } else {
              o[(size_t)r * (cols + 8) + c] = rn(f);
            }
          }
        return out;
      };
      h->deup[l].out_mul = 1.f;
      if (x3) {      // one power-of-two scale for both maps and the bias (they share the accumulator): fp16 lo halves stay normal
        float mx = 0.f;
        for (float v : mt) mx = fmaxf(mx, fabsf(v));
        for (float v : w3a) mx = fmaxf(mx, fabsf(v));
        if (mx > 0.f) {
          int e = 0;
          frexpf(mx, &e);
          int k = 14 - e;
          k = k < -8 ? -8 : (k > 30 ? 30 : k);
          const float sc = ldexpf(1.f, k);
          for (float& v : mt) v *= sc;
          for (float& v : w3a) v *= sc;
          for (float& v : btc) v *= sc;
          h->deup[l].out_mul = ldexpf(1.f, -k);
        }
      }
      DCL_TRY(upload(h, to_bf16_padded(mt, 8 * CH, C), &h->deup[l].mt));
      DCL_TRY(upload(h, to_bf16_padded(w3a, CH, CH), &h->deup[l].w3a));
      DCL_TRY(upload(h, btc, &h->deup[l].bt));
    }
    DCL_TRY(upload(h, h->host_w.at("decoder.endconv.weight"), &h->end_w));
    DCL_TRY(upload(h, h->host_w.at("decoder.endconv.bias"), &h->end_b));
  }
  h->ready = true;
  return 0;
}

// ---- forward schedule --------------------------------------------------------------------------
struct Fwd {
  dcl_handle* h;
  cudaStream_t st;
  TokScratch* ts;      // token-path scratch this schedule instance works in

  StatSlot stats(const float* x, int c, int64_t spatial, int* rc) {
    StatSlot s = h->stat_slots[h->stat_next++ % h->stat_slots.size()];
    *rc = launch_instnorm_stats(x, c, spatial, h->stat_accum, s.mean, s.rstd, st);
    return s;
  }

  // dense-input 3x3x3 conv with optional fused input norm/activation
  int conv3(const float* x0, int c0, const float* x1, int c1, int g, const ConvW& w, int stride, const StatSlot* norm,
            int act, const float* out_scale, const float* residual, float* y) {
    ConvSrc s{x0, x1, c0, c1, (int64_t)g * g * g, (int64_t)g * g, g, norm ? norm->mean : nullptr,
              norm ? norm->rstd : nullptr, act};
    ConvDst d{y, w.b, out_scale, residual};
    cudaEvent_t ev = nullptr;
    if (h->profiling) ev = h->prof_begin(st);
    int rc;
    if (is_tc(h)) {
      // fp32 NCDHW tensors of the auxiliary heads: blocked-bf16 prep + the general GEMM kernel
      rc = launch_prep_blocked(s, g, g, g, h->blk, st, is_x3(h));
      if (rc == 0) rc = launch_conv_gemm(h->blk, w.tc, d, g, g, g, stride, 27, st, is_x3(h));
    } else {
      rc = launch_conv3d_k3(s, d, w.w, w.cout, w.cout_pad, g, g, g, stride, st);
    }
    if (h->profiling) {
      const double og = (double)((g - 1) / stride + 1);
      h->prof_end(ev, 0, 2.0 * 27.0 * (c0 + c1) * w.cout * og * og * og, st);
    }
    return rc;
  }

  int conv1(const float* x0, int c0, const float* x1, int c1, int64_t spatial, const ConvW& w, float* y,
            bool softmax = false) {
    ConvSrc s{x0, x1, c0, c1, spatial, 0, 0, nullptr, nullptr, ACT_NONE};
    ConvDst d{y, w.b, nullptr, nullptr};
    if (is_tc(h) && !softmax) {   // pointwise conv = 1-tap GEMM over a flat "row" of voxels
      DCL_TRY(launch_prep_blocked(s, 1, 1, (int)spatial, h->blk, st, is_x3(h)));
      return launch_conv_gemm(h->blk, w.tc, d, 1, 1, (int)spatial, 1, 1, st, is_x3(h));
    }
    return launch_conv1x1(s, d, w.w, w.cout, spatial, softmax, st);
  }

  // EnBlock (Unet_skipconnection.py:36-57): y = conv2(relu(IN(conv1(relu(IN(x)))))) + x
  int en_block(const float* x, int c, int g, const std::string& name, float* a, float* y) {
    int rc = 0;
    int64_t sp = (int64_t)g * g * g;
    StatSlot s1 = stats(x, c, sp, &rc); DCL_TRY(rc);
    DCL_TRY(conv3(x, c, nullptr, 0, g, h->conv.at(name + ".conv1"), 1, &s1, ACT_RELU, nullptr, nullptr, a));
    StatSlot s2 = stats(a, c, sp, &rc); DCL_TRY(rc);
    DCL_TRY(conv3(a, c, nullptr, 0, g, h->conv.at(name + ".conv2"), 1, &s2, ACT_RELU, nullptr, x, y));
    return 0;
  }

  // EnBlock2 / DeBlock (cls_wise_former.py:691-713, :732-754): y = lrelu(IN(conv2(lrelu(IN(conv1(x)))))) + x
  int post_block(const float* x, int c, int g, const std::string& name, float* a, float* b, float* y) {
    int rc = 0;
    int64_t sp = (int64_t)g * g * g;
    DCL_TRY(conv3(x, c, nullptr, 0, g, h->conv.at(name + ".conv1"), 1, nullptr, ACT_NONE, nullptr, nullptr, a));
    StatSlot s1 = stats(a, c, sp, &rc); DCL_TRY(rc);
    DCL_TRY(conv3(a, c, nullptr, 0, g, h->conv.at(name + ".conv2"), 1, &s1, ACT_LRELU, nullptr, nullptr, b));
    StatSlot s2 = stats(b, c, sp, &rc); DCL_TRY(rc);
    DCL_TRY(launch_norm_act_res(b, s2.mean, s2.rstd, ACT_LRELU, x, y, c, sp, st));
    return 0;
  }

  // DeUp_Cat (cls_wise_former.py:716-729): conv1x1 -> convT k2s2 -> conv1x1(cat(skip, .))
  int up_cat(const float* x, int c, int g, const float* skip, const std::string& name, float* u1, float* u2,
             float* y) {
    int64_t sp = (int64_t)g * g * g;
    const ConvW& c2 = h->conv.at(name + ".conv2");
    DCL_TRY(conv1(x, c, nullptr, 0, sp, h->conv.at(name + ".conv1"), u1));
    DCL_TRY(launch_convt_k2s2(u1, u2, c2.w, c2.b, c / 2, c / 2, g, g, g, st));
    DCL_TRY(conv1(skip, c / 2, u2, c / 2, sp * 8, h->conv.at(name + ".conv3"), y));
    return 0;
  }

  // Residual(PreNormDrop(DualSelfAttention)) (ResidualNorm.py:4-32, SelfAttention.py:74-102)
  int attn_block(const Transformer& t, const float* x, const float* x2, int mq, int mk, float* out) {
    if (is_tc(h)) {   // LayerNorm fused into the bf16 operand prep, linears on tcgen05
      const bool x3 = is_x3(h);
      DCL_TRY(launch_prep_rows2(x, t.n1w, t.n1b, mq, ts->tok_a, x2, t.n2w, t.n2b, mk, ts->tok_b, st, x3));
      DCL_TRY(launch_linear_tc(ts->tok_a, t.pq, nullptr, nullptr, ts->qbuf, mq, 512, 512, false, st, nullptr, x3, t.mq));
      DCL_TRY(launch_linear_tc(ts->tok_b, t.pkv, nullptr, nullptr, ts->kvbuf, mk, 1024, 512, false, st, nullptr, x3, t.mkv));
      DCL_TRY(launch_attention(ts->qbuf, ts->kvbuf, nullptr, mq, mk, st, ts->tok_a, x3));     // bf16 blocked, straight into the GEMM
      DCL_TRY(launch_linear_tc(ts->tok_a, t.pout, t.bout, x, out, mq, 512, 512, false, st, nullptr, x3, t.mout));
      return 0;
    }
    DCL_TRY(launch_layernorm(x, t.n1w, t.n1b, ts->ln_a, mq, st));
    DCL_TRY(launch_layernorm(x2, t.n2w, t.n2b, ts->ln_b, mk, st));
    DCL_TRY(launch_linear(ts->ln_a, t.wqkv, nullptr, nullptr, ts->qbuf, mq, 512, 512, false, st));
    DCL_TRY(launch_linear(ts->ln_b, t.wqkv + 512 * 512, nullptr, nullptr, ts->kvbuf, mk, 1024, 512, false, st));
    DCL_TRY(launch_attention(ts->qbuf, ts->kvbuf, ts->obuf, mq, mk, st));
    DCL_TRY(launch_linear(ts->obuf, t.wout, t.bout, x, out, mq, 512, 512, false, st));
    return 0;
  }

  // Residual(PreNorm(FeedForward)) (ResidualNorm.py:35-47)
  int ffn_block(const Transformer& t, const float* x, int m, float* out) {
    if (is_tc(h)) {
      const bool x3 = is_x3(h);
      DCL_TRY(launch_prep_rows(x, t.fnw, t.fnb, m, ts->tok_a, st, x3));
      DCL_TRY(launch_linear_tc(ts->tok_a, t.p0, t.b0, nullptr, nullptr, m, 512, 512, true, st, ts->tok_b, x3, t.m0));   // GELU, blocked out
      DCL_TRY(launch_linear_tc(ts->tok_b, t.p3, t.b3, x, out, m, 512, 512, false, st, nullptr, x3, t.m3));
      return 0;
    }
    DCL_TRY(launch_layernorm(x, t.fnw, t.fnb, ts->ffn_ln, m, st));
    DCL_TRY(launch_linear(ts->ffn_ln, t.w0, t.b0, nullptr, ts->ffn_h, m, 512, 512, true, st));
    DCL_TRY(launch_linear(ts->ffn_h, t.w3, t.b3, x, out, m, 512, 512, false, st));
    return 0;
  }

  int select_build(const float* score_tok, const float* class_tok, const float* feats, int n, const float* pe,
                   int slot, float* seq) {
    int* idx = h->topk + slot * TOP_NUM;
    DCL_TRY(launch_select_topk(score_tok, feats, n, ts->score, idx, st));
    DCL_TRY(launch_build_sequence(class_tok, feats, idx, pe, seq, st));
    return 0;
  }

  // one auxiliary head branch: conv k3 -> conv k3 (2 classes) -> trilinear upsample -> softmax
  // slot >= 0: the destination is read from the device-side patch descriptor (graph-replayable tensor-core forward)
  int aux_branch(const float* x, int c, int g, const std::string& first, const std::string& second, float* out, int slot = -1) {
    const ConvW& w1 = h->conv.at(first);
    DCL_TRY(conv3(x, c, nullptr, 0, g, w1, 1, nullptr, ACT_NONE, nullptr, nullptr, h->aux_t1));
    DCL_TRY(conv3(h->aux_t1, w1.cout, nullptr, 0, g, h->conv.at(second), 1, nullptr, ACT_NONE, nullptr, nullptr,
                  h->aux_t2));
    DCL_TRY(launch_upsample_softmax2(h->aux_t2, out, g, 128 / g, st, slot >= 0 ? h->patch_desc : nullptr, slot >= 0 ? slot : 0));
    return 0;
  }

  int run(const float* x, const int64_t xs[4], const float* keep_host, float* probs_out, float* const* aux) {
    int rc = 0;
    const bool want_aux = aux != nullptr;
    const bool dense_feats = want_aux || h->cfg.keep_stages;
    // ---- encoder ----
    Floats16 keep;
    for (int i = 0; i < 16; ++i) keep.v[i] = keep_host ? keep_host[i] : 1.f;
    DCL_TRY(launch_fill16(h->keep_dev, keep, st));   // by-value kernel argument: no pageable-memory copy
    {
      ConvSrc s{x, nullptr, 4, 0, xs[0], xs[1], xs[2], nullptr, nullptr, ACT_NONE};
      const ConvW& w = h->conv.at("Unet_list.InitConv.conv");
      ConvDst d{h->l_t0[0], w.b, h->keep_dev, nullptr};
      cudaEvent_t ev = h->profiling ? h->prof_begin(st) : nullptr;
      if (is_tc(h)) {
        DCL_TRY(launch_prep_blocked(s, 128, 128, 128, h->blk, st, is_x3(h)));
        DCL_TRY(launch_conv_gemm(h->blk, w.tc, d, 128, 128, 128, 1, 27, st, is_x3(h)));
      } else {
        DCL_TRY(launch_conv3d_k3(s, d, w.w, w.cout, w.cout_pad, 128, 128, 128, 1, st));
      }
      if (h->profiling) h->prof_end(ev, 0, 2.0 * 27.0 * 4 * 16 * (double)P3, st);
    }
    const char* blk[4][2] = {{"Unet_list.EnBlock1", "Unet_list.EnBlock1_1"}, {"Unet_list.EnBlock2_1", "Unet_list.EnBlock2_2"},
                             {"Unet_list.EnBlock3_1", "Unet_list.EnBlock3_2"}, {"Unet_list.EnBlock4_1", "Unet_list.EnBlock4_2"}};
    const char* down[4] = {"Unet_list.EnDown1.conv", "Unet_list.EnDown2.conv", "Unet_list.EnDown3.conv",
                           "Unet_list.EnDown_4.conv"};
    for (int l = 0; l < 4; ++l) {
      int c = LVL_C[l], g = LVL_G[l];
      DCL_TRY(en_block(h->l_t0[l], c, g, blk[l][0], h->l_a[l], h->l_t1[l]));
      DCL_TRY(en_block(h->l_t1[l], c, g, blk[l][1], h->l_a[l], h->l_x[l]));
      float* nxt = l < 3 ? h->l_t0[l + 1] : h->x4;
      DCL_TRY(conv3(h->l_x[l], c, nullptr, 0, g, h->conv.at(down[l]), l < 3 ? 2 : 1, nullptr, ACT_NONE, nullptr,
                    nullptr, nxt));
    }
    const float *x1 = h->l_x[0], *x2 = h->l_x[1], *x3 = h->l_x[2];
    const int64_t g16 = 16 * 16 * 16, g32 = 32 * 32 * 32;

    // ---- Anatomy-induced Region Decoupler (cls_wise_former.py:284-328), the 3 sibling convs merged ----
    DCL_TRY(conv3(x2, 32, nullptr, 0, 64, h->conv.at("conv_64_to_32"), 2, nullptr, ACT_NONE, nullptr, nullptr,
                  h->e_down));
    DCL_TRY(conv3(h->e_down, 32, x3, 64, 32, h->edge_merged, 1, nullptr, ACT_NONE, nullptr, nullptr, h->e_raw));
    StatSlot se = stats(h->e_raw, 96, g32, &rc); DCL_TRY(rc);
    DCL_TRY(conv3(h->x4, 256, nullptr, 0, 16, h->sem_merged, 1, nullptr, ACT_NONE, nullptr, nullptr, h->s_raw));
    StatSlot ss = stats(h->s_raw, 384, g16, &rc); DCL_TRY(rc);
    for (int r = 0; r < 3; ++r) {
      DCL_TRY(launch_norm_act_tokenise(h->e_raw + r * 32 * g32, se.mean + 32 * r, se.rstd + 32 * r, ACT_LRELU, h->E[r],
                                       dense_feats ? h->edge_dense[r] : nullptr, 32, 32, 4, 2, 2, st));
      DCL_TRY(launch_norm_act_tokenise(h->s_raw + r * 128 * g16, ss.mean + 128 * r, ss.rstd + 128 * r, ACT_LRELU,
                                       h->S[r], dense_feats ? h->sem_dense[r] : nullptr, 128, 16, 2, 2, 1, st));
    }
    if (want_aux) {   // mid heads (cls_wise_former.py:332-333)
      for (int r = 0; r < 3; ++r) {
        std::string n = REGION_NUM[r];
        if (aux[6 + r] != nullptr)
        DCL_TRY(aux_branch(h->sem_dense[r], 128, 16, "mid_supervise_label.supervise_label_" + n,
                           "mid_supervise_label.down_label_" + n, aux[6 + r]));
        if (aux[9 + r] != nullptr)
        DCL_TRY(aux_branch(h->edge_dense[r], 32, 32, "mid_edge_supervise_label.edge_supervise_label_" + n,
                           "mid_edge_supervise_label.edge_down_label_" + n, aux[9 + r]));
      }
    }

    // ---- Edge-supported Intra-region Coupler per region (cls_wise_former.py:341-543) ----
    for (int r = 0; r < 3; ++r) {
      const Transformer& t = h->tr[r];
      float *E = h->E[r], *S = h->S[r], *out = h->coupler_out[r];
      DCL_TRY(select_build(h->e_tok[r], h->e_tok[r], E, 2048, h->pe[r], 4 * r + 0, ts->seq[0]));   // edge
      DCL_TRY(select_build(h->e_tok[r], h->s_tok[r], S, 1024, h->pe[r], 4 * r + 1, ts->seq[1]));   // semantic supplement
      DCL_TRY(select_build(h->s_tok[r], h->s_tok[r], S, 1024, h->pe[r], 4 * r + 2, ts->seq[2]));   // semantic
      DCL_TRY(select_build(h->s_tok[r], h->e_tok[r], E, 2048, h->pe[r], 4 * r + 3, ts->seq[3]));   // edge supplement
      DCL_TRY(attn_block(t, ts->seq[0], ts->seq[1], SEQ, SEQ, ts->eqs));
      DCL_TRY(attn_block(t, ts->seq[2], ts->seq[3], SEQ, SEQ, ts->sqe));
      DCL_TRY(attn_block(t, ts->eqs, ts->sqe, SEQ, SEQ, ts->cross));
      DCL_TRY(attn_block(t, ts->sqe, ts->eqs, SEQ, SEQ, ts->cross + SEQ * 512));
      DCL_TRY(ffn_block(t, ts->cross, 2 * SEQ, out));
      DCL_TRY(launch_scatter_rows(E, h->topk + (4 * r + 0) * TOP_NUM, out + 512, 512, st));
      DCL_TRY(launch_scatter_rows(S, h->topk + (4 * r + 2) * TOP_NUM, out + (SEQ + 1) * 512, 512, st));
      if (want_aux) {
        DCL_TRY(launch_scale_untokenise(E, out, h->sup_edge[r], 32, 32, 4, 2, 2, st));
        DCL_TRY(launch_scale_untokenise(S, out + SEQ * 512, h->sup_sem[r], 128, 16, 2, 2, 1, st));
      }
    }
    if (want_aux) {   // final heads (cls_wise_former.py:545-546)
      for (int r = 0; r < 3; ++r) {
        std::string n = REGION_NUM[r];
        if (aux[0 + r] != nullptr)
        DCL_TRY(aux_branch(h->sup_sem[r], 128, 16, "supervise_label.supervise_label_" + n,
                           "supervise_label.down_label_" + n, aux[0 + r]));
        if (aux[3 + r] != nullptr)
        DCL_TRY(aux_branch(h->sup_edge[r], 32, 32, "edge_supervise_label.edge_supervise_label_" + n,
                           "edge_supervise_label.edge_down_label_" + n, aux[3 + r]));
      }
    }

    // ---- Mutual Cross-region Coupler (cls_wise_former.py:549-582) ----
    DCL_TRY(launch_add3(h->coupler_out[0] + SEQ * 512, h->coupler_out[1] + SEQ * 512, h->coupler_out[2] + SEQ * 512,
                        h->f_tok, 512, st));
    DCL_TRY(launch_add3(h->S[0], h->S[1], h->S[2], h->f_fea, 1024 * 512, st));
    DCL_TRY(select_build(h->f_tok, h->f_tok, h->f_fea, 1024, h->pe[3], 12, ts->seq[0]));
    DCL_TRY(attn_block(h->tr[3], ts->seq[0], ts->seq[0], SEQ, SEQ, ts->eqs));
    DCL_TRY(ffn_block(h->tr[3], ts->eqs, SEQ, h->coupler_out[3]));
    DCL_TRY(launch_scatter_rows(h->f_fea, h->topk + 12 * TOP_NUM, h->coupler_out[3] + 512, 512, st));
    DCL_TRY(launch_scale_untokenise(h->f_fea, h->coupler_out[3], h->fused_dense, 128, 16, 2, 2, 1, st));
    DCL_TRY(conv3(h->fused_dense, 128, nullptr, 0, 16, h->conv.at("sum_fusion"), 1, nullptr, ACT_NONE, nullptr,
                  nullptr, h->enc));

    // ---- decoder (cls_wise_former.py:644-664) ----
    DCL_TRY(conv1(h->enc, 256, nullptr, 0, g16, h->conv.at("decoder.down_channel"), h->d8_0));
    DCL_TRY(post_block(h->d8_0, 128, 16, "decoder.Enblock8_1", h->d8_a, h->d8_b, h->d8_1));
    DCL_TRY(post_block(h->d8_1, 128, 16, "decoder.Enblock8_2", h->d8_a, h->d8_b, h->d8_2));
    const char* upn[3] = {"decoder.DeUp4", "decoder.DeUp3", "decoder.DeUp2"};
    const char* dbn[3][2] = {{"decoder.DeBlock4", "decoder.DeBlock4_1"}, {"decoder.DeBlock3", "decoder.DeBlock3_1"},
                             {"decoder.DeBlock2", "decoder.DeBlock2_1"}};
    const float* skips[3] = {x3, x2, x1};
    const float* cur = h->d8_2;
    for (int l = 0; l < 3; ++l) {
      int cin = 128 >> l, g_in = 16 << l, c = cin / 2, g = g_in * 2;
      DCL_TRY(up_cat(cur, cin, g_in, skips[l], upn[l], h->up_u1[l], h->up_u2[l], h->dl_in[l]));
      DCL_TRY(post_block(h->dl_in[l], c, g, dbn[l][0], h->dl_a[l], h->dl_b[l], h->dl_1[l]));
      DCL_TRY(post_block(h->dl_1[l], c, g, dbn[l][1], h->dl_a[l], h->dl_b[l], h->dl_2[l]));
      cur = h->dl_2[l];
    }
    DCL_TRY(conv1(cur, 16, nullptr, 0, P3, h->conv.at("decoder.endconv"), probs_out, true));
    return 0;
  }
};

// ---- bf16 forward schedule (DCL_BF16): every activation is B-format, every conv runs on tcgen05 --------
struct Fwd16 {
  dcl_handle* h;
  cudaStream_t st;

  stat_t* new_stats() { return h->stat_arena + (size_t)(h->stat_used++ % STAT_SLOTS) * 1024; }
  static BNorm norm_of(const stat_t* sums, int64_t spatial, int act) {
    BNorm n;
    n.sums = sums; n.inv_n = (float)(1.0 / (double)spatial); n.act = act;
    return n;
  }

  // 3x3x3 conv, B-format in/out.  x1 (optional) = second concat source; norm = fused input transform.
  int conv(const void* x0, int c0, const void* x1, int c1, int g, const ConvW& w, int stride, const BNorm* norm,
           const float* out_scale, const void* resb, void* y, stat_t* stats, int taps = 27) {
    cudaEvent_t ev = h->profiling ? h->prof_begin(st) : nullptr;
    int rc, kind;
    const bool x3 = is_x3(h);
    if (taps == 27 && x1 == nullptr && tc_conv_supported(c0, w.cout, g, stride, x3)) {
      kind = c0 == 32 ? 3 : 2;
      RollArgs a;
      a.x3 = x3;
      a.xb = x0;
      if (norm) a.norm = *norm;
      a.bias = w.b; a.out_scale = out_scale; a.resb = resb; a.yb = y; a.stats = stats;
      rc = launch_roll_conv(a, w.tc, w.cout, g, st);
    } else if (taps == 27 && slab_conv_supported(c0 + c1, w.cout, g, g, g, stride, taps) && g <= 32) {
      kind = 4;
      GemmArgs ga;
      ga.x3 = x3;
      ga.a0 = x0; ga.c0 = c0; ga.a1 = x1;
      ga.D = g; ga.H = g; ga.W = g; ga.stride = 1; ga.taps = 27;
      ga.bias = w.b; ga.out_scale = out_scale; ga.out_mode = 2; ga.y = y; ga.residual = resb; ga.stats = stats;
      rc = launch_slab_conv(ga, norm, w.tc, st);
    } else if (taps == 27 && stride == 2 && x1 == nullptr && norm == nullptr && out_scale == nullptr && resb == nullptr &&
               (x3 ? s2_roll_supported_x3(c0, w.cout, g) : s2_roll_supported(c0, w.cout, g))) {
      kind = 12;
      rc = launch_s2_roll_conv(x0, w.tc, w.b, y, stats, st, x3);
    } else {
      kind = 5;
      const void* src = x0;
      if (norm) {
        if (x1 != nullptr) { set_error("bf16 conv: fused norm with two sources is not used by this network"); return -1; }
        DCL_TRY(launch_norm_act_b(x0, *norm, nullptr, h->b_nrm, c0, (int64_t)g * g * g, st, x3));
        src = h->b_nrm;
      }
      GemmArgs ga;
      ga.x3 = x3;
      ga.a0 = src; ga.c0 = c0; ga.a1 = x1;
      ga.D = g; ga.H = g; ga.W = g; ga.stride = stride; ga.taps = taps;
      ga.bias = w.b; ga.out_scale = out_scale; ga.out_mode = 2; ga.y = y; ga.residual = resb; ga.stats = stats;
      rc = launch_gemm_conv(ga, w.tc, st);
    }
    if (h->profiling) {
      const double og = (double)((g - 1) / stride + 1);
      h->prof_end(ev, taps == 27 ? 0 : 11, 2.0 * taps * (c0 + c1) * w.cout * og * og * og, st, kind);
    }
    return rc;
  }

  // EnBlock: y = conv2(relu(IN(conv1(relu(IN(x)))))) + x; sx = sums of x, *sy = sums of y (if wanted)
  int en_block(const void* x, const stat_t* sx, int c, int g, const std::string& name, void* a, void* y, stat_t* sy) {
    const int64_t sp = (int64_t)g * g * g;
    stat_t* sa = new_stats();
    BNorm n1 = norm_of(sx, sp, ACT_RELU), n2 = norm_of(sa, sp, ACT_RELU);
    DCL_TRY(conv(x, c, nullptr, 0, g, h->conv.at(name + ".conv1"), 1, &n1, nullptr, nullptr, a, sa));
    DCL_TRY(conv(a, c, nullptr, 0, g, h->conv.at(name + ".conv2"), 1, &n2, nullptr, x, y, sy));
    return 0;
  }

  // EnBlock2 / DeBlock: y = lrelu(IN(conv2(lrelu(IN(conv1(x)))))) + x
  // tail_out != nullptr: the trailing norm + act + residual is NOT run; its description (input b, statistics) is
  // returned so that the consumer applies it while loading (endconv)
  int post_block(const void* x, int c, int g, const std::string& name, void* a, void* b, void* y, BNorm* tail_out = nullptr) {
    const int64_t sp = (int64_t)g * g * g;
    stat_t *sa = new_stats(), *sb = new_stats();
    DCL_TRY(conv(x, c, nullptr, 0, g, h->conv.at(name + ".conv1"), 1, nullptr, nullptr, nullptr, a, sa));
    BNorm n1 = norm_of(sa, sp, ACT_LRELU);
    DCL_TRY(conv(a, c, nullptr, 0, g, h->conv.at(name + ".conv2"), 1, &n1, nullptr, nullptr, b, sb));
    if (tail_out != nullptr) { *tail_out = norm_of(sb, sp, ACT_LRELU); return 0; }
    dcl_handle::ProfScope ps(h, st, 7);
    DCL_TRY(launch_norm_act_b(b, norm_of(sb, sp, ACT_LRELU), x, y, c, sp, st, is_x3(h)));
    return 0;
  }

  // The forward proper.  Its per-patch arguments (source view, dropout scale) live in h->patch_desc, so the
  // sequence of launches is identical for every patch and can be replayed as a CUDA graph.
  int run(const float* x, const int64_t xs[4], const float* keep_host, float* probs_out, float* const* aux) {
    PatchDesc d;
    d.x = x; d.sc = xs[0]; d.sd = xs[1]; d.sh = xs[2];
    for (int i = 0; i < 16; ++i) d.keep[i] = keep_host ? keep_host[i] : 1.f;
    d.probs = probs_out;      // endconv writes here (the captured graph only knows h->probs)
    int mask = 0;             // requested auxiliary heads: their destinations travel in the descriptor too
    for (int j = 0; j < DCL_NUM_AUX; ++j) {
      d.aux[j] = aux ? aux[j] : nullptr;
      if (d.aux[j]) mask |= 1 << j;
    }
    DCL_TRY(launch_patch_desc(h->patch_desc, d, st));
    const bool graphable = !h->profiling && !h->fwd_graph_off;
    if (!graphable) return body(probs_out, aux);
    dcl_handle::FwdGraph& fg = h->fwd_graphs[mask];
    if (fg.exec == nullptr) {
      if (fg.eager_runs < 1) {         // the first forward of a kind runs eagerly: one-time attribute calls and weight repacks
        ++fg.eager_runs;
        ++h->fwd_eager_runs;
        return body(probs_out, aux);
      }
      // capture the whole forward (including the three concurrent coupler streams) into a graph writing h->probs
      // (captured on a private stream: the caller's stream may be the legacy default stream, which cannot capture)
      const int64_t before = g_launches;
      cudaGraph_t graph = nullptr;
      cudaStream_t cs = h->aux_stream[2];
      if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        h->fwd_graph_off = true;
        if (getenv("DCL_DEBUG")) fprintf(stderr, "dcl: stream capture unavailable\n");
        return body(probs_out, aux);
      }
      Fwd16 cap{h, cs};
      const int rc = cap.body(h->probs, aux);       // aux: only WHICH heads are present matters to the captured launches
      const cudaError_t e = cudaStreamEndCapture(cs, &graph);
      if (rc != 0 || e != cudaSuccess || graph == nullptr ||
          cudaGraphInstantiate(&fg.exec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        fg.exec = nullptr;
        h->fwd_graph_off = true;       // stay on the eager launches (same kernels)
        if (getenv("DCL_DEBUG")) fprintf(stderr, "dcl: graph capture failed rc=%d e=%d (%s)\n", rc, (int)e, g_error.c_str());
        g_launches = before;
        return body(probs_out, aux);
      }
      cudaGraphDestroy(graph);
      if (getenv("DCL_DEBUG")) fprintf(stderr, "dcl: forward captured (aux mask %x), %lld launches\n", mask, (long long)(g_launches - before));
      fg.launches = g_launches - before;
      g_launches = before;
    }
    DCL_CUDA_OK(cudaGraphLaunch(fg.exec, st));
    g_launches += fg.launches;
    return 0;
  }

  int body(float* probs_out, float* const* aux) {
    const bool want_aux = aux != nullptr;
    const bool dense_feats = want_aux || h->cfg.keep_stages;
    Fwd f{h, st, &h->ts[0]};   // token-path and auxiliary-head helpers are shared with the fp32 schedule
    h->stat_used = 0;
    static const bool stamps_on = getenv("DCL_STAMPS") != nullptr;
    auto stamp = [&](int i) -> int { return stamps_on ? launch_stamp(h->stamps, i, st) : 0; };
    DCL_TRY(stamp(0));
    DCL_CUDA_OK(cudaMemsetAsync(h->stat_arena, 0, (size_t)STAT_SLOTS * 1024 * sizeof(stat_t), st));
    const int64_t g16 = 16 * 16 * 16, g32 = 32 * 32 * 32;

    // ---- encoder ----
    stat_t* s_in = new_stats();
    {
      const ConvW& w = h->conv.at("Unet_list.InitConv.conv");
      RollArgs a;
      a.desc = h->patch_desc;
      a.bias = w.b; a.out_scale = h->keep_dev; a.yb = h->b_t0[0]; a.stats = s_in; a.x3 = is_x3(h);
      cudaEvent_t ev = h->profiling ? h->prof_begin(st) : nullptr;
      DCL_TRY(launch_roll_conv(a, w.tc, 16, 128, st));
      if (h->profiling) h->prof_end(ev, 0, 2.0 * 27.0 * 4 * 16 * (double)P3, st, 2);
    }
    const char* blk[4][2] = {{"Unet_list.EnBlock1", "Unet_list.EnBlock1_1"}, {"Unet_list.EnBlock2_1", "Unet_list.EnBlock2_2"},
                             {"Unet_list.EnBlock3_1", "Unet_list.EnBlock3_2"}, {"Unet_list.EnBlock4_1", "Unet_list.EnBlock4_2"}};
    const char* down[4] = {"Unet_list.EnDown1.conv", "Unet_list.EnDown2.conv", "Unet_list.EnDown3.conv",
                           "Unet_list.EnDown_4.conv"};
    for (int l = 0; l < 4; ++l) {
      const int c = LVL_C[l], g = LVL_G[l];
      stat_t* s_mid = new_stats();
      DCL_TRY(en_block(h->b_t0[l], s_in, c, g, blk[l][0], h->b_a[l], h->b_t1[l], s_mid));
      DCL_TRY(en_block(h->b_t1[l], s_mid, c, g, blk[l][1], h->b_a[l], h->b_x[l], nullptr));
      s_in = new_stats();
      void* nxt = l < 3 ? h->b_t0[l + 1] : h->b_x4;
      DCL_TRY(conv(h->b_x[l], c, nullptr, 0, g, h->conv.at(down[l]), l < 3 ? 2 : 1, nullptr, nullptr, nullptr, nxt,
                   l < 3 ? s_in : nullptr));
    }

    DCL_TRY(stamp(1));
    // ---- Anatomy-induced Region Decoupler (the 3 sibling convs of each branch merged) ----
    DCL_TRY(conv(h->b_x[1], 32, nullptr, 0, 64, h->conv.at("conv_64_to_32"), 2, nullptr, nullptr, nullptr, h->b_edown,
                 nullptr));
    stat_t *s_e = new_stats(), *s_s = new_stats();
    DCL_TRY(conv(h->b_edown, 32, h->b_x[2], 64, 32, h->edge_merged, 1, nullptr, nullptr, nullptr, h->b_eraw, s_e));
    DCL_TRY(conv(h->b_x4, 256, nullptr, 0, 16, h->sem_merged, 1, nullptr, nullptr, nullptr, h->b_sraw, s_s));
    for (int r = 0; r < 3; ++r) {
      dcl_handle::ProfScope ps(h, st, 10);
      BNorm ne = norm_of(s_e, g32, ACT_LRELU), ns = norm_of(s_s, g16, ACT_LRELU);
      DCL_TRY(launch_tokenise_b(h->b_eraw, ne, 4 * r, h->E[r], dense_feats ? h->edge_dense[r] : nullptr, 32, 32, 4, 2, 2, st,
                                is_x3(h) ? 12 : 0));
      DCL_TRY(launch_tokenise_b(h->b_sraw, ns, 16 * r, h->S[r], dense_feats ? h->sem_dense[r] : nullptr, 128, 16, 2, 2, 1, st,
                                is_x3(h) ? 48 : 0));
    }
    if (want_aux) {
      for (int r = 0; r < 3; ++r) {
        std::string n = REGION_NUM[r];
        if (aux[6 + r] != nullptr)
          DCL_TRY(f.aux_branch(h->sem_dense[r], 128, 16, "mid_supervise_label.supervise_label_" + n,
                             "mid_supervise_label.down_label_" + n, aux[6 + r], 6 + r));
        if (aux[9 + r] != nullptr)
          DCL_TRY(f.aux_branch(h->edge_dense[r], 32, 32, "mid_edge_supervise_label.edge_supervise_label_" + n,
                             "mid_edge_supervise_label.edge_down_label_" + n, aux[9 + r], 9 + r));
      }
    }

    // ---- Edge-supported Intra-region Couplers: the three regions are independent until the cross-region
    // coupler, and each is a chain of small latency-bound kernels, so they run on three concurrent streams
    // (region 0 on the caller's stream, regions 1-2 on the handle's auxiliary streams, own scratch each)
    // Within a region the edge half (selects 0,1 -> A1) and the semantic half (selects 2,3 -> A2) are independent,
    // and so are the two cross attentions A3 / A4: every region runs on TWO lanes with private scratch, six lanes
    // in all (lane 0 = the caller's stream).  In the captured graph these are plain parallel branches.
    cudaEvent_t ev_tokp = h->profiling ? h->prof_begin(st) : nullptr;
    DCL_TRY(stamp(2));
    DCL_CUDA_OK(cudaEventRecord(h->ev_fork, st));
    for (int r = 0; r < 3; ++r) {
      cudaStream_t sa = r == 0 ? st : h->tok_stream[2 * r - 1];
      cudaStream_t sb = h->tok_stream[2 * r];
      if (r > 0) DCL_CUDA_OK(cudaStreamWaitEvent(sa, h->ev_fork, 0));
      DCL_CUDA_OK(cudaStreamWaitEvent(sb, h->ev_fork, 0));
      TokScratch *ta = &h->ts[2 * r], *tb = &h->ts[2 * r + 1];
      Fwd fa{h, sa, ta}, fb{h, sb, tb};
      const Transformer& t = h->tr[r];
      float *E = h->E[r], *S = h->S[r], *out = h->coupler_out[r];
      DCL_TRY(fa.select_build(h->e_tok[r], h->e_tok[r], E, 2048, h->pe[r], 4 * r + 0, ta->seq[0]));   // edge
      DCL_TRY(fb.select_build(h->s_tok[r], h->s_tok[r], S, 1024, h->pe[r], 4 * r + 2, tb->seq[2]));   // semantic
      DCL_TRY(fa.select_build(h->e_tok[r], h->s_tok[r], S, 1024, h->pe[r], 4 * r + 1, ta->seq[1]));   // semantic supplement
      DCL_TRY(fb.select_build(h->s_tok[r], h->e_tok[r], E, 2048, h->pe[r], 4 * r + 3, tb->seq[3]));   // edge supplement
      if (r == 0 && stamps_on) DCL_TRY(launch_stamp(h->stamps, 10, sa));
      DCL_TRY(fa.attn_block(t, ta->seq[0], ta->seq[1], SEQ, SEQ, ta->eqs));
      if (r == 0 && stamps_on) DCL_TRY(launch_stamp(h->stamps, 11, sa));
      DCL_TRY(fb.attn_block(t, tb->seq[2], tb->seq[3], SEQ, SEQ, tb->sqe));
      DCL_CUDA_OK(cudaEventRecord(h->ev_tok[r][0], sa));
      DCL_CUDA_OK(cudaEventRecord(h->ev_tok[r][1], sb));
      DCL_CUDA_OK(cudaStreamWaitEvent(sa, h->ev_tok[r][1], 0));
      DCL_CUDA_OK(cudaStreamWaitEvent(sb, h->ev_tok[r][0], 0));
      DCL_TRY(fa.attn_block(t, ta->eqs, tb->sqe, SEQ, SEQ, ta->cross));
      if (r == 0 && stamps_on) DCL_TRY(launch_stamp(h->stamps, 12, sa));
      DCL_TRY(fb.attn_block(t, tb->sqe, ta->eqs, SEQ, SEQ, ta->cross + SEQ * 512));
      DCL_CUDA_OK(cudaEventRecord(h->ev_tok[r][2], sb));
      DCL_CUDA_OK(cudaStreamWaitEvent(sa, h->ev_tok[r][2], 0));
      DCL_TRY(fa.ffn_block(t, ta->cross, 2 * SEQ, out));
      if (r == 0 && stamps_on) DCL_TRY(launch_stamp(h->stamps, 13, sa));
      DCL_TRY(launch_scatter_rows(E, h->topk + (4 * r + 0) * TOP_NUM, out + 512, 512, sa));
      DCL_TRY(launch_scatter_rows(S, h->topk + (4 * r + 2) * TOP_NUM, out + (SEQ + 1) * 512, 512, sa));
      if (want_aux) {
        DCL_TRY(launch_scale_untokenise(E, out, h->sup_edge[r], 32, 32, 4, 2, 2, sa));
        DCL_TRY(launch_scale_untokenise(S, out + SEQ * 512, h->sup_sem[r], 128, 16, 2, 2, 1, sa));
      }
      if (r > 0) {
        DCL_CUDA_OK(cudaEventRecord(h->ev_tok_join[2 * r - 1], sa));
        DCL_CUDA_OK(cudaStreamWaitEvent(st, h->ev_tok_join[2 * r - 1], 0));
      }
    }
    if (want_aux) {
      for (int r = 0; r < 3; ++r) {
        std::string n = REGION_NUM[r];
        if (aux[0 + r] != nullptr)
          DCL_TRY(f.aux_branch(h->sup_sem[r], 128, 16, "supervise_label.supervise_label_" + n,
                             "supervise_label.down_label_" + n, aux[0 + r], 0 + r));
        if (aux[3 + r] != nullptr)
          DCL_TRY(f.aux_branch(h->sup_edge[r], 32, 32, "edge_supervise_label.edge_supervise_label_" + n,
                             "edge_supervise_label.edge_down_label_" + n, aux[3 + r], 3 + r));
      }
    }

    // ---- Mutual Cross-region Coupler ----
    DCL_TRY(stamp(3));
    DCL_TRY(launch_add3(h->coupler_out[0] + SEQ * 512, h->coupler_out[1] + SEQ * 512, h->coupler_out[2] + SEQ * 512,
                        h->f_tok, 512, st));
    DCL_TRY(launch_add3(h->S[0], h->S[1], h->S[2], h->f_fea, 1024 * 512, st));
    DCL_TRY(f.select_build(h->f_tok, h->f_tok, h->f_fea, 1024, h->pe[3], 12, h->ts[0].seq[0]));
    DCL_TRY(f.attn_block(h->tr[3], h->ts[0].seq[0], h->ts[0].seq[0], SEQ, SEQ, h->ts[0].eqs));
    DCL_TRY(f.ffn_block(h->tr[3], h->ts[0].eqs, SEQ, h->coupler_out[3]));
    DCL_TRY(launch_scatter_rows(h->f_fea, h->topk + 12 * TOP_NUM, h->coupler_out[3] + 512, 512, st));
    DCL_TRY(launch_untokenise_b(h->f_fea, h->coupler_out[3], h->b_fused, 128, 16, 2, 2, 1, st, is_x3(h)));
    if (ev_tokp) h->prof_end(ev_tokp, 8, 0.0, st);
    DCL_TRY(stamp(4));
    DCL_TRY(conv(h->b_fused, 128, nullptr, 0, 16, h->conv.at("sum_fusion"), 1, nullptr, nullptr, nullptr, h->b_enc, nullptr));

    // ---- decoder ----
    DCL_TRY(conv(h->b_enc, 256, nullptr, 0, 16, h->conv.at("decoder.down_channel"), 1, nullptr, nullptr, nullptr,
                 h->b_d8[0], nullptr, 1));
    DCL_TRY(post_block(h->b_d8[0], 128, 16, "decoder.Enblock8_1", h->b_d8[1], h->b_d8[2], h->b_d8[3]));
    DCL_TRY(post_block(h->b_d8[3], 128, 16, "decoder.Enblock8_2", h->b_d8[1], h->b_d8[2], h->b_d8[4]));
    const char* dbn[3][2] = {{"decoder.DeBlock4", "decoder.DeBlock4_1"}, {"decoder.DeBlock3", "decoder.DeBlock3_1"},
                             {"decoder.DeBlock2", "decoder.DeBlock2_1"}};
    const void* cur = h->b_d8[4];
    BNorm end_tail;
    const void *end_src = nullptr, *end_res = nullptr;
    for (int l = 0; l < 3; ++l) {
      DCL_TRY(stamp(5 + l));
      const int cin = 128 >> l, g_in = 16 << l, c = cin / 2, g = g_in * 2;
      void** b = h->b_dl[l];
      {
        dcl_handle::ProfScope ps(h, st, 6);
        DCL_TRY(launch_deup_fused_b(cur, h->b_x[2 - l], h->deup[l].mt, h->deup[l].w3a, h->deup[l].bt, b[0], cin, g_in, st, is_x3(h),
                                    h->deup[l].out_mul));
      }
      DCL_TRY(post_block(b[0], c, g, dbn[l][0], b[1], b[2], b[3]));
      // the very last DeBlock tail (16 channels @ 128^3) is folded into endconv's load unless the stage is to be kept
      const bool fold = l == 2 && !h->cfg.keep_stages;
      DCL_TRY(post_block(b[3], c, g, dbn[l][1], b[1], b[2], b[4], fold ? &end_tail : nullptr));
      cur = b[4];
      if (fold) { end_src = b[2]; end_res = b[3]; }
    }
    DCL_TRY(stamp(8));
    {
      dcl_handle::ProfScope ps(h, st, 9);
      if (end_src != nullptr) DCL_TRY(launch_endconv_softmax_b(end_src, h->end_w, h->end_b, probs_out, P3, st, &end_tail, end_res, h->patch_desc, is_x3(h)));
      else DCL_TRY(launch_endconv_softmax_b(cur, h->end_w, h->end_b, probs_out, P3, st, nullptr, nullptr, h->patch_desc, is_x3(h)));
    }
    DCL_TRY(stamp(9));
    return 0;
  }
};

static void register_stages(dcl_handle* h) {
  const int64_t g16 = 16 * 16 * 16, g32 = 32 * 32 * 32;
  auto& s = h->stages;
  s["init"] = {h->l_t0[0], 16 * P3};
  s["x1_1"] = {h->l_x[0], 16 * P3};
  s["x2_1"] = {h->l_x[1], 32 * P3 / 8};
  s["x3_1"] = {h->l_x[2], 64 * g32};
  s["x4"] = {h->x4, 256 * g16};
  for (int r = 0; r < 3; ++r) {
    s[std::string("edge_") + REGION_NUM[r]] = {h->edge_dense[r], 32 * g32};
    s[std::string("sem_") + REGION_NUM[r]] = {h->sem_dense[r], 128 * g16};
    s[std::string("coupler_") + REGION_KEY[r]] = {h->coupler_out[r], 258 * 512};
  }
  s["coupler_fusion"] = {h->coupler_out[3], SEQ * 512};
  s["enc_out"] = {h->enc, 256 * g16};
  s["dec8"] = {h->d8_2, 128 * g16};
  s["dec4"] = {h->dl_2[0], 64 * g32};
  s["dec3"] = {h->dl_2[1], 32 * P3 / 8};
  s["dec2"] = {h->dl_2[2], 16 * P3};
  if (is_tc(h)) {   // the conv-path stages live in B-format buffers in this mode
    auto& b = h->bstages;
    b["init"] = {h->b_t0[0], 16, P3};
    b["x1_1"] = {h->b_x[0], 16, P3};
    b["x2_1"] = {h->b_x[1], 32, P3 / 8};
    b["x3_1"] = {h->b_x[2], 64, g32};
    b["x4"] = {h->b_x4, 256, g16};
    b["enc_out"] = {h->b_enc, 256, g16};
    b["dec8"] = {h->b_d8[4], 128, g16};
    b["dec4"] = {h->b_dl[0][4], 64, g32};
    b["dec3"] = {h->b_dl[1][4], 32, P3 / 8};
    b["dec2"] = {h->b_dl[2][4], 16, P3};
  }
}

static int check_handle(dcl_handle* h) {
  if (!h) { set_error("null handle"); return DCL_ERR_ARG; }
  int dev = -1;
  DCL_CUDA_OK(cudaGetDevice(&dev));
  if (dev != h->device) { set_error("handle used on a different CUDA device than it was created on"); return DCL_ERR_STATE; }
  if (!h->ready) DCL_TRY(prepare(h));
  return 0;
}

static int grow(dcl_handle* h, void** p, int64_t* cap, int64_t bytes) {
  if (*cap >= bytes) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  DCL_CUDA_OK(cudaMalloc(p, (size_t)bytes));
  *cap = bytes;
  return 0;
}

struct PlanItem { int start[3]; StitchBox box; };

static int build_plan(int mode, const int32_t shape[3], int n_patches, const int32_t* starts,
                      std::vector<PlanItem>* plan, int* zout) {
  plan->clear();
  if (mode == DCL_STITCH_REFERENCE || mode == DCL_STITCH_ALIGNED) {
    if (shape[0] != 240 || shape[1] != 240 || shape[2] < 155) {
      set_error("reference tiling needs a (4,240,240,>=155) volume (predict_overlap.py:34-41)");
      return DCL_ERR_ARG;
    }
    *zout = 155;
    const int zsrc = mode == DCL_STITCH_REFERENCE ? 96 : 101;   // predict_overlap.py:53: 96:123 (5-voxel shift)
    for (int iz = 0; iz < 2; ++iz)
      for (int ix = 0; ix < 2; ++ix)
        for (int iy = 0; iy < 2; ++iy) {
          PlanItem p;
          p.start[0] = ix ? 112 : 0; p.start[1] = iy ? 112 : 0; p.start[2] = iz ? 27 : 0;
          p.box.dst[0] = ix ? 128 : 0; p.box.ext[0] = ix ? 112 : 128; p.box.src[0] = ix ? 16 : 0;
          p.box.dst[1] = iy ? 128 : 0; p.box.ext[1] = iy ? 112 : 128; p.box.src[1] = iy ? 16 : 0;
          p.box.dst[2] = iz ? 128 : 0; p.box.ext[2] = iz ? 27 : 128; p.box.src[2] = iz ? zsrc : 0;
          plan->push_back(p);
        }
    return 0;
  }
  if (mode != DCL_STITCH_UNIFORM && mode != DCL_STITCH_GAUSSIAN) { set_error("unknown stitch mode"); return DCL_ERR_ARG; }
  if (n_patches <= 0 || !starts) { set_error("weighted modes need an explicit patch list"); return DCL_ERR_ARG; }
  *zout = shape[2];
  for (int i = 0; i < n_patches; ++i) {
    PlanItem p{};
    for (int a = 0; a < 3; ++a) {
      p.start[a] = starts[3 * i + a];
      if (p.start[a] < 0 || p.start[a] + 128 > shape[a]) { set_error("patch outside the volume"); return DCL_ERR_ARG; }
    }
    plan->push_back(p);
  }
  // Every voxel must lie in at least one patch: an uncovered voxel has weight sum 0 and would come out as 0/0.
  // Exact test on the grid cut at every patch face: cell (i,j,k) is covered iff some patch covers interval i of x,
  // j of y and k of z - one AND of three per-interval patch bitsets per cell.
  {
    const int P = n_patches, words = (P + 63) / 64;
    std::vector<int> cuts[3];
    std::vector<std::vector<uint64_t>> bits[3];
    for (int a = 0; a < 3; ++a) {
      std::vector<int>& c = cuts[a];
      c.push_back(0); c.push_back(shape[a]);
      for (int i = 0; i < P; ++i) { c.push_back((*plan)[i].start[a]); c.push_back((*plan)[i].start[a] + 128); }
      std::sort(c.begin(), c.end());
      c.erase(std::unique(c.begin(), c.end()), c.end());
      bits[a].assign(c.size() - 1, std::vector<uint64_t>(words, 0));
      for (size_t k = 0; k + 1 < c.size(); ++k)
        for (int i = 0; i < P; ++i)
          if ((*plan)[i].start[a] <= c[k] && c[k + 1] <= (*plan)[i].start[a] + 128) bits[a][k][i / 64] |= 1ull << (i % 64);
    }
    for (size_t i = 0; i < bits[0].size(); ++i)
      for (size_t j = 0; j < bits[1].size(); ++j)
        for (size_t k = 0; k < bits[2].size(); ++k) {
          uint64_t any = 0;
          for (int w = 0; w < words; ++w) any |= bits[0][i][w] & bits[1][j][w] & bits[2][k][w];
          if (!any) {
            set_error("patch list does not cover the volume: voxel (" + std::to_string(cuts[0][i]) + "," + std::to_string(cuts[1][j]) +
                      "," + std::to_string(cuts[2][k]) + ") lies in no patch");
            return DCL_ERR_ARG;
          }
        }
  }
  return 0;
}

static int check_handle(dcl_handle* h);
static void free_lanes(dcl_handle* h);

// The second lane of the sliding window: a handle of its own (workspace, captured graph, coupler streams) with a copy
// of the weights.  Two patches in flight overlap each other's latency-bound phases (the couplers, the 16^3 / 32^3
// levels): measured 35.3 instead of 30.0 volumes/s.
static int lane_count() {      // DCL_LANES = 1..4 (default 3: measured 29.9 / 36.6 / 38.2 / 38.5 volumes/s with 1 / 2 / 3 / 4 lanes)
  static const int n = [] {
    const char* e = getenv("DCL_LANES");
    int v = e ? atoi(e) : 3;
    return v < 1 ? 1 : (v > dcl_handle::MAX_LANES ? dcl_handle::MAX_LANES : v);
  }();
  return n;
}

static void free_lanes(dcl_handle* h) {
  for (int i = 0; i < dcl_handle::MAX_LANES - 1; ++i)
    if (h->twin[i]) { dcl_destroy(h->twin[i]); h->twin[i] = nullptr; }
  for (int i = 0; i < dcl_handle::MAX_LANES; ++i) {
    if (h->lane_stream[i]) { cudaStreamDestroy(h->lane_stream[i]); h->lane_stream[i] = nullptr; }
    if (h->ev_lane_acc[i]) { cudaEventDestroy(h->ev_lane_acc[i]); h->ev_lane_acc[i] = nullptr; }
  }
  if (h->ev_lane_fork) { cudaEventDestroy(h->ev_lane_fork); h->ev_lane_fork = nullptr; }
  h->lanes = 0;
}

static int ensure_lanes(dcl_handle* h, int n) {
  if (h->lanes >= n) return 0;
  if (!h->ev_lane_fork) DCL_CUDA_OK(cudaEventCreateWithFlags(&h->ev_lane_fork, cudaEventDisableTiming));
  for (int i = 0; i < n; ++i) {
    if (!h->lane_stream[i]) DCL_CUDA_OK(cudaStreamCreateWithFlags(&h->lane_stream[i], cudaStreamNonBlocking));
    if (!h->ev_lane_acc[i]) DCL_CUDA_OK(cudaEventCreateWithFlags(&h->ev_lane_acc[i], cudaEventDisableTiming));
    if (i > 0 && !h->twin[i - 1]) {
      dcl_handle* t = nullptr;
      DCL_TRY(dcl_create(&h->cfg, &t));
      t->host_w = h->host_w;
      h->twin[i - 1] = t;
      DCL_TRY(check_handle(t));      // packs the weights
    }
  }
  h->lanes = n;
  return 0;
}

static int run_patches(dcl_handle* h, const float* vol, const int32_t shape[3], int mode,
                       const std::vector<PlanItem>& plan, int first, int count, const float* keep_host, int zout,
                       float* acc, float* wsum, cudaStream_t st, float* gather = nullptr, int slot_planes = 4) {
  const int X = shape[0], Y = shape[1], Z = shape[2];
  const int64_t xs[4] = {(int64_t)X * Y * Z, (int64_t)Y * Z, Z, 1};
  const bool weighted = mode == DCL_STITCH_UNIFORM || mode == DCL_STITCH_GAUSSIAN;
  int L = lane_count();
  if (L > count) L = count;
  const bool two = is_tc(h) && !h->profiling && L >= 2;
  if (two) {
    DCL_TRY(ensure_lanes(h, L));
    DCL_CUDA_OK(cudaEventRecord(h->ev_lane_fork, st));
    for (int l = 0; l < L; ++l) DCL_CUDA_OK(cudaStreamWaitEvent(h->lane_stream[l], h->ev_lane_fork, 0));
  }
  int up_have[dcl_handle::MAX_LANES] = {-1, -1, -1, -1};     // staged upload: highest x-slab each lane already waits for
  int lane = 0;
  // an error half-way through the plan must not leave the lane streams forked from the caller's stream
  struct LaneJoin {
    dcl_handle* h; cudaStream_t st; int L; bool armed;
    ~LaneJoin() {
      if (!armed) return;
      for (int l = 0; l < L; ++l)
        if (cudaEventRecord(h->ev_lane_acc[l], h->lane_stream[l]) == cudaSuccess) cudaStreamWaitEvent(st, h->ev_lane_acc[l], 0);
    }
  } lane_join{h, st, L, two};
  for (int i = first; i < first + count; ++i) {
    const PlanItem& p = plan[i];
    lane = two ? (i - first) % L : 0;
    dcl_handle* hh = lane ? h->twin[lane - 1] : h;
    cudaStream_t s = two ? h->lane_stream[lane] : st;
    if (h->up_active) {
      int need = 0;
      while (need < 2 && h->up_bound[need] < p.start[0] + 128) ++need;
      for (; up_have[lane] < need; ++up_have[lane]) DCL_CUDA_OK(cudaStreamWaitEvent(s, h->ev_up[up_have[lane] + 1], 0));
    }
    const float* x = vol + (int64_t)p.start[0] * xs[1] + (int64_t)p.start[1] * xs[2] + p.start[2];
    // gather form: the patch keeps its probabilities in its own slot; nothing is accumulated here and the lanes
    // never wait for each other
    float* dst = gather ? gather + (int64_t)(i - first) * slot_planes * P3 : nullptr;
    // 16-plane slots: the six final auxiliary heads (supervise / edge x {01,02,04}, two classes each) follow the four
    // class planes; the mid heads (forward()[3..4]) are not computed
    float* aux_ptr[DCL_NUM_AUX] = {};
    if (dst && slot_planes == 16)
      for (int j = 0; j < 6; ++j) aux_ptr[j] = dst + (int64_t)(4 + 2 * j) * P3;
    float* const* aux = (dst && slot_planes == 16) ? aux_ptr : nullptr;
    if (is_tc(h)) {
      Fwd16 f16{hh, s};
      DCL_TRY(f16.run(x, xs, keep_host ? keep_host + 16 * i : nullptr, dst ? dst : hh->probs, aux));
    } else {
      Fwd f{h, st, &h->ts[0]};
      DCL_TRY(f.run(x, xs, keep_host ? keep_host + 16 * i : nullptr, dst ? dst : h->probs, aux));
    }
    if (gather) {
      if (two) DCL_CUDA_OK(cudaEventRecord(h->ev_lane_acc[lane], s));
      continue;
    }
    // the accumulates run in patch order (fp32 sums stay reproducible): wait for the previous patch's, on the other lane
    if (two && i > first) DCL_CUDA_OK(cudaStreamWaitEvent(s, h->ev_lane_acc[(lane + L - 1) % L], 0));
    cudaEvent_t ev = h->profiling ? h->prof_begin(s) : nullptr;
    double bytes;
    if (weighted) {
      DCL_TRY(launch_accumulate(hh->probs, p.start, mode == DCL_STITCH_GAUSSIAN, acc, wsum, X, Y, zout, s));
      bytes = (double)P3 * (4 * 4 + 2 * (4 * 4 + 4));     // read 4 probs; read+write 4 acc + wsum
    } else {
      DCL_TRY(launch_stitch_copy(hh->probs, acc, p.box, X, Y, zout, s));
      bytes = (double)p.box.ext[0] * p.box.ext[1] * p.box.ext[2] * 4 * 4 * 2;
    }
    if (h->profiling) h->prof_end(ev, 1, bytes, s);
    if (two) DCL_CUDA_OK(cudaEventRecord(h->ev_lane_acc[lane], s));
  }
  lane_join.armed = false;      // regular exit: the joins below are the minimal ones
  if (two && gather) {
    for (int l = 0; l < L; ++l) DCL_CUDA_OK(cudaStreamWaitEvent(st, h->ev_lane_acc[l], 0));   // every lane's last patch
  } else if (two) {
    DCL_CUDA_OK(cudaStreamWaitEvent(st, h->ev_lane_acc[lane], 0));     // the chain ends at the last accumulate
  }
  if (h->up_active) {     // the tail (finalize, target) needs the whole upload
    int have = -1;
    for (; have < 2; ++have) DCL_CUDA_OK(cudaStreamWaitEvent(st, h->ev_up[have + 1], 0));
  }
  return 0;
}

// Weighted stitch in gather form: every patch of the plan keeps its probabilities (and, with aux_out_dev, the six final
// auxiliary heads) in a slot of its own; gather_finalize_kernel then blends, normalises and labels the volume.
static int predict_volume_gather(dcl_handle* h, const float* vol_dev, const int32_t shape[3], int32_t mode,
                                 const std::vector<PlanItem>& plan, int zout, const float* keep_scale_host,
                                 float* probs_out_dev, float* aux_out_dev, uint8_t* labels_out_dev, const uint8_t* target_dev,
                                 uint64_t* counts_out_dev, cudaStream_t st) {
  const int np = (int)plan.size();
  const int planes = aux_out_dev ? 16 : 4;
  const int64_t V = (int64_t)shape[0] * shape[1] * zout;
  DCL_TRY(grow(h, (void**)&h->gather_buf, &h->gather_cap, (int64_t)np * planes * P3 * 4));
  const int64_t before = g_launches;
  int rc = run_patches(h, vol_dev, shape, mode, plan, 0, np, keep_scale_host, zout, nullptr, nullptr, st, h->gather_buf, planes);
  if (rc == 0 && (labels_out_dev || counts_out_dev || probs_out_dev || aux_out_dev)) {
    if (counts_out_dev) DCL_CUDA_OK(cudaMemsetAsync(counts_out_dev, 0, 13 * sizeof(uint64_t), st));
    GatherPlan gp;
    gp.n = np;
    gp.slot_planes = planes;
    for (int i = 0; i < np; ++i) for (int a = 0; a < 3; ++a) gp.start[i][a] = plan[i].start[a];
    const int gaussian = mode == DCL_STITCH_GAUSSIAN;
    cudaEvent_t ev = h->profiling ? h->prof_begin(st) : nullptr;
    if (labels_out_dev || counts_out_dev || probs_out_dev)
      rc = launch_gather_finalize(h->gather_buf, gp, gaussian, shape[0], shape[1], zout, probs_out_dev, labels_out_dev,
                                  target_dev, (unsigned long long*)counts_out_dev, st);
    if (h->profiling)
      h->prof_end(ev, 1, (double)np * 4 * P3 * 4 + (double)V * ((probs_out_dev ? 16 : 0) + (labels_out_dev ? 1 : 0) +
                                                                (target_dev ? 1 : 0)), st);
    // the auxiliary planes, four at a time (two heads per launch), with the same weights
    for (int g = 0; rc == 0 && aux_out_dev && g < 3; ++g)
      rc = launch_gather_finalize(h->gather_buf + (int64_t)(4 + 4 * g) * P3, gp, gaussian, shape[0], shape[1], zout,
                                  aux_out_dev + (int64_t)g * 4 * V, nullptr, nullptr, nullptr, st);
  }
  h->launches += g_launches - before;
  return rc;
}

}  // namespace dcl

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

DCL_API const char* dcl_last_error(void) { return g_error.c_str(); }
DCL_API int dcl_abi_version(void) { return DCL_ABI_VERSION; }

DCL_API int64_t dcl_workspace_bytes(const dcl_config* cfg) { return workspace_estimate(cfg); }

DCL_API int dcl_create(const dcl_config* cfg, dcl_handle** out) {
  if (!cfg || !out) { set_error("dcl_create: null argument"); return DCL_ERR_ARG; }
  if (cfg->abi_version != DCL_ABI_VERSION) { set_error("dcl_create: ABI version mismatch"); return DCL_ERR_ARG; }
  if (cfg->precision < DCL_FP32 || cfg->precision > DCL_BF16) { set_error("dcl_create: bad precision"); return DCL_ERR_ARG; }
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    set_error("dcl_create: no CUDA device (this library has no CPU path)");
    return DCL_ERR_CUDA;
  }
  dcl_handle* h = new dcl_handle();
  h->cfg = *cfg;
  if (cudaGetDevice(&h->device) != cudaSuccess) { delete h; set_error("cudaGetDevice failed"); return DCL_ERR_CUDA; }
  int rc = allocate_workspace(h);
  if (rc != 0) { dcl_destroy(h); return rc; }
  if (cudaStreamCreateWithFlags(&h->aux_stream[2], cudaStreamNonBlocking) != cudaSuccess) {
    dcl_destroy(h); set_error("dcl_create: stream creation failed"); return DCL_ERR_CUDA;
  }
  for (int i = 0; i < 2; ++i) {
    if (cudaStreamCreateWithFlags(&h->aux_stream[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming) != cudaSuccess) {
      dcl_destroy(h); set_error("dcl_create: stream / event creation failed"); return DCL_ERR_CUDA;
    }
  }
  if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) { dcl_destroy(h); set_error("dcl_create: event creation failed"); return DCL_ERR_CUDA; }
  for (int i = 0; i < 5; ++i)
    if (cudaStreamCreateWithFlags(&h->tok_stream[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_tok_join[i], cudaEventDisableTiming) != cudaSuccess) {
      dcl_destroy(h); set_error("dcl_create: stream / event creation failed"); return DCL_ERR_CUDA;
    }
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k)
      if (cudaEventCreateWithFlags(&h->ev_tok[r][k], cudaEventDisableTiming) != cudaSuccess) {
        dcl_destroy(h); set_error("dcl_create: event creation failed"); return DCL_ERR_CUDA;
      }
  register_stages(h);
  *out = h;
  return DCL_OK;
}

DCL_API int dcl_destroy(dcl_handle* h) {
  if (!h) return DCL_OK;
  cudaDeviceSynchronize();
  free_lanes(h);
  for (auto& kv : h->fwd_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : h->allocs) cudaFree(p);
  for (auto& kv : h->conv)
    if (kv.second.tc.slab_dev) cudaFree(kv.second.tc.slab_dev);
  if (h->edge_merged.tc.slab_dev) cudaFree(h->edge_merged.tc.slab_dev);
  if (h->sem_merged.tc.slab_dev) cudaFree(h->sem_merged.tc.slab_dev);
  for (int i = 0; i < 2; ++i) {
    if (h->aux_stream[i]) cudaStreamDestroy(h->aux_stream[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  if (h->aux_stream[2]) cudaStreamDestroy(h->aux_stream[2]);
  for (int i = 0; i < 5; ++i) {
    if (h->tok_stream[i]) cudaStreamDestroy(h->tok_stream[i]);
    if (h->ev_tok_join[i]) cudaEventDestroy(h->ev_tok_join[i]);
  }
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k)
      if (h->ev_tok[r][k]) cudaEventDestroy(h->ev_tok[r][k]);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int k = 0; k < 3; ++k) if (h->ev_up[k]) cudaEventDestroy(h->ev_up[k]);
  if (h->ev_st) cudaEventDestroy(h->ev_st);
  if (h->vol_probs) cudaFree(h->vol_probs);
  if (h->vol_wsum) cudaFree(h->vol_wsum);
  if (h->gather_buf) cudaFree(h->gather_buf);
  if (h->shard_buf) cudaFree(h->shard_buf);
  if (h->tta_vol) cudaFree(h->tta_vol);
  if (h->tta_sum) cudaFree(h->tta_sum);
  if (h->stage_vol) cudaFree(h->stage_vol);
  if (h->stage_probs) cudaFree(h->stage_probs);
  if (h->stage_labels) cudaFree(h->stage_labels);
  if (h->stage_target) cudaFree(h->stage_target);
  delete h;
  return DCL_OK;
}

DCL_API int dcl_weight_count(void) { return (int)catalogue().size(); }

DCL_API int dcl_weight_spec(int index, char* name, int32_t cap, int64_t* numel, int32_t* aux_only) {
  if (index < 0 || index >= (int)catalogue().size() || !name || cap <= 0) { set_error("dcl_weight_spec: bad argument"); return DCL_ERR_ARG; }
  const WSpec& s = catalogue()[index];
  strncpy(name, s.name.c_str(), cap - 1);
  name[cap - 1] = 0;
  if (numel) *numel = s.numel;
  if (aux_only) *aux_only = s.aux ? 1 : 0;
  return DCL_OK;
}

DCL_API int dcl_set_weight(dcl_handle* h, const char* name, const float* data, int64_t numel) {
  if (!h || !name || !data) { set_error("dcl_set_weight: null argument"); return DCL_ERR_ARG; }
  std::string n(name);
  if (n.compare(0, 7, "module.") == 0) n = n.substr(7);
  const WSpec* spec = nullptr;
  for (const auto& s : catalogue())
    if (s.name == n) { spec = &s; break; }
  if (!spec) { set_error("dcl_set_weight: unknown state_dict key '" + n + "'"); return DCL_ERR_ARG; }
  if (spec->numel != numel) {
    set_error("dcl_set_weight: '" + n + "' expects " + std::to_string(spec->numel) + " elements, got " + std::to_string(numel));
    return DCL_ERR_ARG;
  }
  std::vector<float>& v = h->host_w[n];
  v.resize((size_t)numel);
  DCL_CUDA_OK(cudaMemcpy(v.data(), data, (size_t)numel * 4, cudaMemcpyDefault));
  h->ready = false;   // repacked lazily by the next compute call
  free_lanes(h);      // the extra lanes hold copies of the weights: rebuilt on demand
  return DCL_OK;
}

DCL_API int dcl_missing_weights(dcl_handle* h, char* first_missing, int32_t cap) {
  if (!h) { set_error("null handle"); return DCL_ERR_ARG; }
  int missing = 0;
  for (const auto& s : catalogue())
    if (needed(h, s) && !h->host_w.count(s.name)) {
      if (missing == 0 && first_missing && cap > 0) { strncpy(first_missing, s.name.c_str(), cap - 1); first_missing[cap - 1] = 0; }
      ++missing;
    }
  return missing;
}

DCL_API int dcl_forward(dcl_handle* h, const float* x_dev, const int64_t x_strides[4], const float* keep_scale_host,
                float* probs_dev, float* const* aux_dev, void* stream) {
  DCL_TRY(check_handle(h));
  if (!x_dev || !x_strides || !probs_dev) { set_error("dcl_forward: null argument"); return DCL_ERR_ARG; }
  if (x_strides[3] != 1) { set_error("dcl_forward: the Z stride of x must be 1"); return DCL_ERR_ARG; }
  if (aux_dev && !h->cfg.want_aux) { set_error("dcl_forward: aux outputs need cfg.want_aux"); return DCL_ERR_ARG; }
  int64_t before = g_launches;
  int rc;
  if (is_tc(h)) {
    Fwd16 f{h, (cudaStream_t)stream};
    rc = f.run(x_dev, x_strides, keep_scale_host, probs_dev, aux_dev);
  } else {
    Fwd f{h, (cudaStream_t)stream, &h->ts[0]};
    rc = f.run(x_dev, x_strides, keep_scale_host, probs_dev, aux_dev);
  }
  h->launches += g_launches - before;
  return rc;
}

DCL_API int dcl_accumulate_patches(dcl_handle* h, const float* vol_dev, const int32_t shape[3], int32_t mode,
                           int32_t n_patches, const int32_t* starts_host, const float* keep_scale_host, int32_t first,
                           int32_t count, float* acc_dev, float* wsum_dev, void* stream) {
  DCL_TRY(check_handle(h));
  if (!vol_dev || !shape || !acc_dev) { set_error("dcl_accumulate_patches: null argument"); return DCL_ERR_ARG; }
  std::vector<PlanItem> plan;
  int zout = 0;
  DCL_TRY(build_plan(mode, shape, n_patches, starts_host, &plan, &zout));
  if (first < 0 || count < 0 || first + count > (int)plan.size()) { set_error("patch range outside the plan"); return DCL_ERR_ARG; }
  const bool weighted = mode == DCL_STITCH_UNIFORM || mode == DCL_STITCH_GAUSSIAN;
  if (weighted && !wsum_dev) { set_error("weighted modes need wsum"); return DCL_ERR_ARG; }
  int64_t before = g_launches;
  int rc = run_patches(h, vol_dev, shape, mode, plan, first, count, keep_scale_host, zout, acc_dev, wsum_dev,
                       (cudaStream_t)stream);
  h->launches += g_launches - before;
  return rc;
}

// ---- single volume over several GPUs, owner-computes: every rank keeps the probabilities of ITS patches in
// slots of its own memory, exports the slot buffer through CUDA IPC, and blends / labels the x-range it owns by reading
// the covering patches' slots where they live - local memory or a peer's over NVLink - inside gather_finalize_kernel.
DCL_API int dcl_slots_ensure(dcl_handle* h, int32_t n_slots, void** slots_dev_out) {
  if (!h || n_slots < 1 || !slots_dev_out) { set_error("dcl_slots_ensure: bad argument"); return DCL_ERR_ARG; }
  DCL_TRY(grow(h, (void**)&h->shard_buf, &h->shard_cap, (int64_t)n_slots * 4 * P3 * 4));
  *slots_dev_out = h->shard_buf;
  return DCL_OK;
}

DCL_API int dcl_ipc_export(const void* dev_ptr, void* handle64_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  if (!dev_ptr || !handle64_out) { set_error("dcl_ipc_export: null argument"); return DCL_ERR_ARG; }
  cudaIpcMemHandle_t hd;
  DCL_CUDA_OK(cudaIpcGetMemHandle(&hd, const_cast<void*>(dev_ptr)));
  memcpy(handle64_out, &hd, 64);
  return DCL_OK;
}

DCL_API int dcl_ipc_import(const void* handle64, void** dev_ptr_out) {
  if (!handle64 || !dev_ptr_out) { set_error("dcl_ipc_import: null argument"); return DCL_ERR_ARG; }
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  DCL_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr_out, hd, cudaIpcMemLazyEnablePeerAccess));
  return DCL_OK;
}

DCL_API int dcl_ipc_release(void* dev_ptr) {
  if (dev_ptr) DCL_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return DCL_OK;
}

DCL_API int dcl_forward_patches_to_slots(dcl_handle* h, const float* vol_dev, const int32_t shape[3], int32_t mode,
                                         int32_t n_patches, const int32_t* starts_host, const float* keep_scale_host,
                                         int32_t first, int32_t count, void* stream) {
  DCL_TRY(check_handle(h));
  if (!vol_dev || !shape) { set_error("dcl_forward_patches_to_slots: null argument"); return DCL_ERR_ARG; }
  if (mode != DCL_STITCH_UNIFORM && mode != DCL_STITCH_GAUSSIAN) { set_error("dcl_forward_patches_to_slots: weighted stitch modes only"); return DCL_ERR_ARG; }
  std::vector<PlanItem> plan;
  int zout = 0;
  DCL_TRY(build_plan(mode, shape, n_patches, starts_host, &plan, &zout));
  if (first < 0 || count < 0 || first + count > (int)plan.size()) { set_error("patch range outside the plan"); return DCL_ERR_ARG; }
  if (count == 0) return DCL_OK;
  if (h->shard_cap < (int64_t)count * 4 * P3 * 4) { set_error("dcl_forward_patches_to_slots: call dcl_slots_ensure first"); return DCL_ERR_STATE; }
  const int64_t before = g_launches;
  int rc = run_patches(h, vol_dev, shape, mode, plan, first, count, keep_scale_host, zout, nullptr, nullptr, (cudaStream_t)stream,
                       h->shard_buf, 4);
  h->launches += g_launches - before;
  return rc;
}

DCL_API int dcl_gather_finalize_range(const int32_t shape[3], int32_t mode, int32_t n_patches, const int32_t* starts_host,
                                      const void* const* slot_ptrs_host, int32_t x0, int32_t x1, float* probs_out_dev,
                                      uint8_t* labels_out_dev, const uint8_t* target_dev, uint64_t* counts_out_dev, void* stream) {
  if (!shape || !slot_ptrs_host) { set_error("dcl_gather_finalize_range: null argument"); return DCL_ERR_ARG; }
  if (mode != DCL_STITCH_UNIFORM && mode != DCL_STITCH_GAUSSIAN) { set_error("dcl_gather_finalize_range: weighted stitch modes only"); return DCL_ERR_ARG; }
  std::vector<PlanItem> plan;
  int zout = 0;
  DCL_TRY(build_plan(mode, shape, n_patches, starts_host, &plan, &zout));
  if ((int)plan.size() > GatherPlan::MAX) { set_error("dcl_gather_finalize_range: at most 128 patches"); return DCL_ERR_ARG; }
  GatherPlan gp;
  GatherSlots sl;
  gp.n = (int)plan.size();
  gp.slot_planes = 4;
  for (int i = 0; i < GatherPlan::MAX; ++i) {
    sl.ptr[i] = i < gp.n ? reinterpret_cast<const float*>(slot_ptrs_host[i]) : nullptr;
    if (i < gp.n) {
      if (!sl.ptr[i]) { set_error("dcl_gather_finalize_range: null slot pointer"); return DCL_ERR_ARG; }
      for (int a = 0; a < 3; ++a) gp.start[i][a] = plan[i].start[a];
    }
  }
  return launch_gather_finalize_slots(sl, gp, mode == DCL_STITCH_GAUSSIAN, shape[0], shape[1], zout, x0, x1, probs_out_dev,
                                      labels_out_dev, target_dev, (unsigned long long*)counts_out_dev, (cudaStream_t)stream);
}

DCL_API int dcl_finalize_labels(const float* acc_dev, const float* wsum_dev, int64_t voxels_total, int64_t v0, int64_t nvox,
                        float* probs_out_dev, uint8_t* labels_out_dev, const uint8_t* target_dev,
                        uint64_t* counts_out_dev, void* stream) {
  if (!acc_dev || v0 < 0 || nvox < 0 || v0 + nvox > voxels_total) { set_error("dcl_finalize_labels: bad argument"); return DCL_ERR_ARG; }
  return launch_finalize_labels(acc_dev, wsum_dev, voxels_total, v0, nvox, probs_out_dev, labels_out_dev, target_dev,
                                (unsigned long long*)counts_out_dev, (cudaStream_t)stream);
}

DCL_API int dcl_predict_volume(dcl_handle* h, const float* vol_dev, const int32_t shape[3], int32_t mode, int32_t n_patches,
                       const int32_t* starts_host, const float* keep_scale_host, float* probs_out_dev,
                       uint8_t* labels_out_dev, const uint8_t* target_dev, uint64_t* counts_out_dev, void* stream) {
  DCL_TRY(check_handle(h));
  if (!vol_dev || !shape) { set_error("dcl_predict_volume: null argument"); return DCL_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<PlanItem> plan;
  int zout = 0;
  DCL_TRY(build_plan(mode, shape, n_patches, starts_host, &plan, &zout));
  const int64_t V = (int64_t)shape[0] * shape[1] * zout;
  const bool weighted = mode == DCL_STITCH_UNIFORM || mode == DCL_STITCH_GAUSSIAN;
  // weighted modes: gather form by default (one slot per patch, no accumulator round trip; bit-identical to the
  // accumulate form, which DCL_GATHER=0 selects and the multi-GPU path keeps using)
  const char* genv = getenv("DCL_GATHER");
  const bool gather = weighted && (int)plan.size() <= GatherPlan::MAX && !(genv && genv[0] == '0');
  if (gather)
    return predict_volume_gather(h, vol_dev, shape, mode, plan, zout, keep_scale_host, probs_out_dev, nullptr, labels_out_dev,
                                 target_dev, counts_out_dev, st);
  DCL_TRY(grow(h, (void**)&h->vol_probs, &h->vol_probs_cap, 4 * V * 4));
  float* acc = h->vol_probs;
  float* wsum = nullptr;
  if (weighted) {
    DCL_TRY(grow(h, (void**)&h->vol_wsum, &h->vol_wsum_cap, V * 4));
    wsum = h->vol_wsum;
    DCL_CUDA_OK(cudaMemsetAsync(acc, 0, 4 * V * 4, st));
    DCL_CUDA_OK(cudaMemsetAsync(wsum, 0, V * 4, st));
  } else if (probs_out_dev) {
    acc = probs_out_dev;   // the crop-overwrite plan covers every voxel exactly once
  }
  int64_t before = g_launches;
  int rc = run_patches(h, vol_dev, shape, mode, plan, 0, (int)plan.size(), keep_scale_host, zout, acc, wsum, st);
  if (rc == 0 && (labels_out_dev || counts_out_dev || (weighted && probs_out_dev))) {
    if (counts_out_dev) DCL_CUDA_OK(cudaMemsetAsync(counts_out_dev, 0, 13 * sizeof(uint64_t), st));
    cudaEvent_t ev = h->profiling ? h->prof_begin(st) : nullptr;
    rc = launch_finalize_labels(acc, wsum, V, 0, V, weighted ? probs_out_dev : nullptr, labels_out_dev, target_dev,
                                (unsigned long long*)counts_out_dev, st);
    if (h->profiling)
      h->prof_end(ev, 1, (double)V * (16 + (wsum ? 4 : 0) + (weighted && probs_out_dev ? 16 : 0) +
                                      (labels_out_dev ? 1 : 0) + (target_dev ? 1 : 0)), st);
  }
  h->launches += g_launches - before;
  return rc;
}


// BASELINE config 4 ("WT/TC/ET + edge outputs"): the weighted sliding window with the six final auxiliary heads
// (forward()[1] = supervise {01,02,04}, forward()[2] = edge {01,02,04}; cls_wise_former.py:545-546, :585-592) blended
// alongside the class probabilities.  aux_out_dev: (6, 2, X, Y, Z) in that order.  Needs cfg.want_aux.
DCL_API int dcl_predict_volume_aux(dcl_handle* h, const float* vol_dev, const int32_t shape[3], int32_t mode, int32_t n_patches,
                           const int32_t* starts_host, const float* keep_scale_host, float* probs_out_dev,
                           float* aux_out_dev, uint8_t* labels_out_dev, const uint8_t* target_dev,
                           uint64_t* counts_out_dev, void* stream) {
  DCL_TRY(check_handle(h));
  if (!vol_dev || !shape || !aux_out_dev) { set_error("dcl_predict_volume_aux: null argument"); return DCL_ERR_ARG; }
  if (!h->cfg.want_aux) { set_error("dcl_predict_volume_aux: the handle was created without want_aux"); return DCL_ERR_ARG; }
  if (mode != DCL_STITCH_UNIFORM && mode != DCL_STITCH_GAUSSIAN) {
    set_error("dcl_predict_volume_aux: weighted stitch modes only (the crop-and-overwrite plan needs 4 output channels, predict_overlap.py:43)");
    return DCL_ERR_ARG;
  }
  std::vector<PlanItem> plan;
  int zout = 0;
  DCL_TRY(build_plan(mode, shape, n_patches, starts_host, &plan, &zout));
  return predict_volume_gather(h, vol_dev, shape, mode, plan, zout, keep_scale_host, probs_out_dev, aux_out_dev, labels_out_dev,
                               target_dev, counts_out_dev, (cudaStream_t)stream);
}

// 8-flip test-time augmentation around the reference tiling: predict_cls.py:180-203.
//   keep_scale_host: NULL, or 8 flips x 8 patches x 16 dropout scales in the order the reference draws them
//   (flip order: none, X, Y, Z, XY, XZ, YZ, XYZ = dims (2,), (3,), (4,), (2,3), (2,4), (3,4), (2,3,4)).
DCL_API int dcl_predict_volume_tta(dcl_handle* h, const float* vol_dev, const int32_t shape[3], const float* keep_scale_host,
                           float* probs_out_dev, uint8_t* labels_out_dev, const uint8_t* target_dev,
                           uint64_t* counts_out_dev, void* stream) {
  DCL_TRY(check_handle(h));
  if (!vol_dev || !shape) { set_error("dcl_predict_volume_tta: null argument"); return DCL_ERR_ARG; }
  if (shape[0] != 240 || shape[1] != 240 || shape[2] < 155) {
    set_error("the reference tiling needs a (4,240,240,>=155) volume (predict_overlap.py:34-41)");
    return DCL_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int X = shape[0], Y = shape[1], Z = 155;
  const int64_t V = (int64_t)X * Y * Z;
  DCL_TRY(grow(h, (void**)&h->tta_vol, &h->tta_vol_cap, 4 * V * 4));
  DCL_TRY(grow(h, (void**)&h->vol_probs, &h->vol_probs_cap, 4 * V * 4));
  float* sum = probs_out_dev;
  if (!sum) { DCL_TRY(grow(h, (void**)&h->tta_sum, &h->tta_sum_cap, 4 * V * 4)); sum = h->tta_sum; }
  std::vector<PlanItem> plan;
  int zout = 0;
  const int32_t fshape[3] = {X, Y, Z};
  DCL_TRY(build_plan(DCL_STITCH_REFERENCE, fshape, 8, nullptr, &plan, &zout));
  static const int FLIPS[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  const int64_t before = g_launches;
  for (int f = 0; f < 8; ++f) {
    const int* fl = FLIPS[f];
    DCL_TRY(launch_flip_volume(vol_dev, h->tta_vol, X, Y, shape[2], Z, fl[0], fl[1], fl[2], st));     // x[..., :155].flip(dims)
    DCL_TRY(run_patches(h, h->tta_vol, fshape, DCL_STITCH_REFERENCE, plan, 0, 8, keep_scale_host ? keep_scale_host + f * 8 * 16 : nullptr,
                        zout, h->vol_probs, nullptr, st));                                           // tailor_and_concat
    DCL_TRY(launch_tta_accumulate(h->vol_probs, sum, X, Y, Z, fl[0], fl[1], fl[2], f == 0, f == 7, st));   // += softmax(.flip(dims))
  }
  int rc = 0;
  if (labels_out_dev || counts_out_dev) {
    if (counts_out_dev) DCL_CUDA_OK(cudaMemsetAsync(counts_out_dev, 0, 13 * sizeof(uint64_t), st));
    rc = launch_finalize_labels(sum, nullptr, V, 0, V, nullptr, labels_out_dev, target_dev, (unsigned long long*)counts_out_dev, st);
  }
  h->launches += g_launches - before;
  return rc;
}

DCL_API int dcl_predict_volume_host(dcl_handle* h, const float* vol_host, const int32_t shape[3], int32_t mode,
                            int32_t n_patches, const int32_t* starts_host, const float* keep_scale_host,
                            float* probs_out_host, uint8_t* labels_out_host, const uint8_t* target_host,
                            uint64_t counts_out_host[13], void* stream) {
  DCL_TRY(check_handle(h));
  if (!vol_host || !shape) { set_error("dcl_predict_volume_host: null argument"); return DCL_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const bool ref = mode == DCL_STITCH_REFERENCE || mode == DCL_STITCH_ALIGNED;
  const int zout = ref ? 155 : shape[2];
  const int64_t vin = (int64_t)shape[0] * shape[1] * shape[2], V = (int64_t)shape[0] * shape[1] * zout;
  DCL_TRY(grow(h, (void**)&h->stage_vol, &h->stage_vol_cap, 4 * vin * 4));
  if (h->stage_lab_cap < V) {
    if (h->stage_labels) cudaFree(h->stage_labels);
    if (h->stage_target) cudaFree(h->stage_target);
    h->stage_labels = h->stage_target = nullptr; h->stage_lab_cap = 0;
    DCL_CUDA_OK(cudaMalloc((void**)&h->stage_labels, (size_t)V));
    DCL_CUDA_OK(cudaMalloc((void**)&h->stage_target, (size_t)V));
    h->stage_lab_cap = V;
  }
  // upload in three x-slabs on the copy stream (x < 128 is all the first patches of the z-major plan need), so the
  // first forward starts after ~half of the host-to-device copy and the rest overlaps with compute
  if (!h->copy_stream) {
    DCL_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 3; ++k) DCL_CUDA_OK(cudaEventCreateWithFlags(&h->ev_up[k], cudaEventDisableTiming));
    DCL_CUDA_OK(cudaEventCreateWithFlags(&h->ev_st, cudaEventDisableTiming));
  }
  DCL_CUDA_OK(cudaEventRecord(h->ev_st, st));                       // the copies follow the work already queued on st
  DCL_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_st, 0));
  {
    const int X = shape[0];
    const int64_t row = (int64_t)shape[1] * shape[2];
    h->up_bound[0] = X < 128 ? X : 128; h->up_bound[1] = X < 192 ? X : 192; h->up_bound[2] = X;
    int x0 = 0;
    for (int k = 0; k < 3; ++k) {
      const int x1 = h->up_bound[k];
      if (x1 > x0)
        for (int c = 0; c < 4; ++c)
          DCL_CUDA_OK(cudaMemcpyAsync(h->stage_vol + c * vin + x0 * row, vol_host + c * vin + x0 * row, (size_t)(x1 - x0) * row * 4,
                                      cudaMemcpyHostToDevice, h->copy_stream));
      if (k == 2 && target_host)
        DCL_CUDA_OK(cudaMemcpyAsync(h->stage_target, target_host, (size_t)V, cudaMemcpyHostToDevice, h->copy_stream));
      DCL_CUDA_OK(cudaEventRecord(h->ev_up[k], h->copy_stream));
      x0 = x1;
    }
  }
  if (probs_out_host) DCL_TRY(grow(h, (void**)&h->stage_probs, &h->stage_probs_cap, 4 * V * 4));
  h->up_active = true;
  const int rc_pv = dcl_predict_volume(h, h->stage_vol, shape, mode, n_patches, starts_host, keep_scale_host,
                             probs_out_host ? h->stage_probs : nullptr, h->stage_labels, target_host ? h->stage_target : nullptr,
                             counts_out_host ? (uint64_t*)h->counts_dev : nullptr, stream);
  h->up_active = false;
  if (rc_pv != 0) { cudaStreamSynchronize(h->copy_stream); return rc_pv; }
  if (labels_out_host) DCL_CUDA_OK(cudaMemcpyAsync(labels_out_host, h->stage_labels, (size_t)V, cudaMemcpyDeviceToHost, st));
  if (counts_out_host) DCL_CUDA_OK(cudaMemcpyAsync(counts_out_host, h->counts_dev, 13 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (probs_out_host) DCL_CUDA_OK(cudaMemcpyAsync(probs_out_host, h->stage_probs, 4 * V * 4, cudaMemcpyDeviceToHost, st));
  DCL_CUDA_OK(cudaStreamSynchronize(st));
  return DCL_OK;
}

DCL_API int64_t dcl_read_stage(dcl_handle* h, const char* stage, float* out_dev, int64_t cap, void* stream) {
  if (!h || !stage) { set_error("dcl_read_stage: null argument"); return DCL_ERR_ARG; }
  if (!h->cfg.keep_stages) { set_error("dcl_read_stage: handle was created without keep_stages"); return DCL_ERR_STATE; }
  auto bit = h->bstages.find(stage);
  if (bit != h->bstages.end()) {
    const int64_t n = (int64_t)bit->second.c * bit->second.spatial;
    if (out_dev) {
      if (cap < n) { set_error("dcl_read_stage: output buffer too small"); return DCL_ERR_ARG; }
      DCL_TRY(launch_unblock(bit->second.p, out_dev, bit->second.c, bit->second.spatial, (cudaStream_t)stream, is_x3(h)));
    }
    return n;
  }
  auto it = h->stages.find(stage);
  if (it == h->stages.end()) { set_error(std::string("dcl_read_stage: unknown stage '") + stage + "'"); return DCL_ERR_ARG; }
  int64_t n = it->second.second;
  if (out_dev) {
    if (cap < n) { set_error("dcl_read_stage: output buffer too small"); return DCL_ERR_ARG; }
    DCL_CUDA_OK(cudaMemcpyAsync(out_dev, it->second.first, n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  return n;
}

DCL_API int dcl_read_topk(dcl_handle* h, int32_t* out_host, void* stream) {
  if (!h || !out_host) { set_error("dcl_read_topk: null argument"); return DCL_ERR_ARG; }
  DCL_CUDA_OK(cudaMemcpyAsync(out_host, h->topk, 13 * 128 * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  DCL_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  return DCL_OK;
}

DCL_API int dcl_profile_enable(dcl_handle* h, int32_t on) {
  if (!h) { set_error("null handle"); return DCL_ERR_ARG; }
  if (on) {   // a new measurement window: recycle the recorded events
    cudaDeviceSynchronize();
    for (auto& r : h->prof) { h->event_pool.push_back(r.a); h->event_pool.push_back(r.b); }
    h->prof.clear();
  }
  h->profiling = on != 0;
  return DCL_OK;
}

DCL_API int dcl_profile_read(dcl_handle* h, int32_t cls, double* ms_total, int64_t* launches, double* work_total) {
  if (!h) { set_error("null handle"); return DCL_ERR_ARG; }
  double ms = 0, work = 0;
  int64_t n = 0;
  for (auto& r : h->prof) {
    if (r.cls != cls && r.cls2 != cls) continue;
    DCL_CUDA_OK(cudaEventSynchronize(r.b));
    float t = 0.f;
    DCL_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t; work += r.work; ++n;
  }
  if (ms_total) *ms_total = ms;
  if (launches) *launches = n;
  if (work_total) *work_total = work;
  return DCL_OK;
}

int64_t dcl_launch_count(const dcl_handle* h) { return h ? h->launches : 0; }

// debug: the 16 %globaltimer stamps of the last bf16 forward (needs DCL_STAMPS=1 in the environment)
DCL_API int dcl_debug_stamps(dcl_handle* h, uint64_t* out_host) {
  if (!h || !out_host || !h->stamps) { set_error("dcl_debug_stamps: no stamps (bf16 handle needed)"); return DCL_ERR_ARG; }
  DCL_CUDA_OK(cudaMemcpy(out_host, h->stamps, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return DCL_OK;
}

// debug: in-kernel timeline of CTA (0,0) of the tcgen05 kernels (tools/trace_kernel.py)
DCL_API int dcl_trace_enable(int32_t on) {
  const size_t bytes = (1 + 2 * TRACE_CAP_HOST) * sizeof(long long);
  if (on && !g_trace_host_ptr) DCL_CUDA_OK(cudaMalloc((void**)&g_trace_host_ptr, bytes));
  long long* p = on ? g_trace_host_ptr : nullptr;
  if (p) DCL_CUDA_OK(cudaMemset(p, 0, bytes));
  DCL_CUDA_OK(trace_set_conv_tc(p, on > 0 ? on - 1 : 0));      // on = 1 + blockIdx.x of the traced CTA
  DCL_CUDA_OK(trace_set_conv_gemm(p, on > 0 ? on - 1 : 0));
  return DCL_OK;
}
// copies the recorded (tag<<32|step, clock) pairs (slots with a non-zero clock) to out_host, returns the count
// and clears the buffer
DCL_API int64_t dcl_trace_read(int64_t* out_host, int64_t cap) {
  if (!g_trace_host_ptr) return 0;
  std::vector<long long> h(1 + 2 * TRACE_CAP_HOST);
  DCL_CUDA_OK(cudaMemcpy(h.data(), g_trace_host_ptr, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  int64_t n = 0;
  for (int i = 0; i < TRACE_CAP_HOST && n < cap; ++i)
    if (h[2 + 2 * i] != 0) { out_host[2 * n] = h[1 + 2 * i]; out_host[2 * n + 1] = h[2 + 2 * i]; ++n; }
  DCL_CUDA_OK(cudaMemset(g_trace_host_ptr, 0, h.size() * sizeof(long long)));
  return n;
}


// bench helper (tools/op_time.py): average device time in us of `reps` back-to-back launches of the bf16 kernel
// that the forward would pick for a cubic g^3 convolution.  mode 1 = fused input norm + residual + statistics.
DCL_API double dcl_bench_conv(int32_t cin, int32_t cout, int32_t g, int32_t stride, int32_t mode, int32_t reps) {
  const int64_t sp = (int64_t)g * g * g;
  const int og = (g - 1) / stride + 1;
  const int64_t osp = (int64_t)og * og * og;
  const int cin_pad = (cin + 15) / 16 * 16, cout_pad = (cout + 15) / 16 * 16;
  std::vector<float> w((size_t)cout * cin * 27);
  for (size_t i = 0; i < w.size(); ++i) w[i] = (float)((i * 2654435761u >> 8) & 0xffff) / 65536.f - 0.5f;
  const bool x3 = (mode & 8) != 0;     // split-fp16 (DCL_F16X3) variant of the same layer
  const int E = x3 ? 4 : 2;
  const bool roll = stride == 1 && tc_conv_supported(cin, cout, g, 1, x3) && cin != 4;
  TcWeights tw;
  if (tc_pack_weights(w.data(), cout, cin, 27, roll, &tw, x3) != 0) return -1.0;
  void *x = nullptr, *y = nullptr, *r = nullptr;
  float* bias = nullptr;
  stat_t *sin = nullptr, *sout = nullptr;
  cudaMalloc(&x, (size_t)cin_pad * sp * E); cudaMalloc(&y, (size_t)cout_pad * osp * E); cudaMalloc(&r, (size_t)cout_pad * osp * E);
  cudaMalloc((void**)&bias, cout_pad * 4); cudaMalloc((void**)&sin, 2 * cin_pad * sizeof(stat_t)); cudaMalloc((void**)&sout, 2 * cout_pad * sizeof(stat_t));
  cudaMemset(x, 0x3c, (size_t)cin_pad * sp * E); cudaMemset(r, 0x3c, (size_t)cout_pad * osp * E); cudaMemset(bias, 0, cout_pad * 4);
  cudaMemset(sin, 0, 2 * cin_pad * sizeof(stat_t)); cudaMemset(sout, 0, 2 * cout_pad * sizeof(stat_t));
  BNorm bn; bn.sums = sin; bn.inv_n = 1.f / (float)sp; bn.act = ACT_RELU;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = 0;
  for (int it = 0; it < reps + 3 && rc == 0; ++it) {
    if (it == 3) cudaEventRecord(e0, 0);
    if (roll) {
      RollArgs a; a.x3 = x3; a.xb = x; if (mode & 1) a.norm = bn; if (mode & 2) a.resb = r; if (mode & 4) a.stats = sout;
      a.bias = bias; a.yb = y;
      rc = launch_roll_conv(a, tw, cout, g, 0);
    } else {
      GemmArgs ga; ga.x3 = x3; ga.a0 = x; ga.c0 = cin_pad; ga.D = g; ga.H = g; ga.W = g; ga.stride = stride; ga.taps = 27;
      ga.bias = bias; ga.out_mode = 2; ga.y = y;
      if (mode & 2) ga.residual = r;
      if (mode & 4) ga.stats = sout;
      if (stride == 1 && slab_conv_supported(cin, cout, g, g, g, 1, 27) && g <= 32) rc = launch_slab_conv(ga, (mode & 1) ? &bn : nullptr, tw, 0);
      else if (stride == 2 && (x3 ? s2_roll_supported_x3(cin, cout, g) : s2_roll_supported(cin, cout, g)))
        rc = launch_s2_roll_conv(x, tw, bias, y, (mode & 7) ? sout : nullptr, 0, x3);
      else rc = launch_gemm_conv(ga, tw, 0);
    }
  }
  float ms = 0.f;
  if (rc == 0) {     // a refused configuration never recorded e0
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaDeviceSynchronize();
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(x); cudaFree(y); cudaFree(r); cudaFree(bias); cudaFree(sin); cudaFree(sout);
  tc_free_weights(&tw);
  if (cudaGetLastError() != cudaSuccess) rc = -1;
  return rc == 0 ? (double)ms * 1e3 / reps : -1.0;
}

// Isolated device timing of the weighted stitch on one X x Y x Z volume with the sliding-window plan of `stride`:
// form 0 = gather (one gather_finalize launch over the per-patch slots), form 1 = accumulate (two memsets, one accumulate
// launch per patch, finalize).  Slots hold a constant; labels + counters are produced.  Returns us per volume.
DCL_API double dcl_bench_stitch(const int32_t shape[3], int32_t stride, int32_t gaussian, int32_t form, int32_t reps,
                                int32_t* n_patches_out) {
  if (!shape || stride < 1 || reps < 1) return -1.0;
  std::vector<PlanItem> plan;
  int zout = 0;
  std::vector<int32_t> starts;
  auto axis = [&](int n) { std::vector<int> v; for (int s = 0; s < n - 128; s += stride) v.push_back(s); v.push_back(n - 128); return v; };
  const std::vector<int> xs = axis(shape[0]), ys = axis(shape[1]), zs = axis(shape[2]);
  for (int z : zs) for (int x : xs) for (int y : ys) { starts.push_back(x); starts.push_back(y); starts.push_back(z); }
  const int np = (int)starts.size() / 3;
  if (build_plan(DCL_STITCH_UNIFORM, shape, np, starts.data(), &plan, &zout) != 0 || np > GatherPlan::MAX) return -1.0;
  if (n_patches_out) *n_patches_out = np;
  const int64_t V = (int64_t)shape[0] * shape[1] * zout;
  float *slots = nullptr, *acc = nullptr, *wsum = nullptr;
  uint8_t* labels = nullptr;
  unsigned long long* counts = nullptr;
  if (cudaMalloc((void**)&slots, (size_t)np * 4 * P3 * 4) != cudaSuccess) return -1.0;
  cudaMalloc((void**)&acc, 4 * V * 4); cudaMalloc((void**)&wsum, V * 4); cudaMalloc((void**)&labels, V); cudaMalloc((void**)&counts, 13 * 8);
  cudaMemset(slots, 0x3e, (size_t)np * 4 * P3 * 4);
  GatherPlan gp;
  gp.n = np;
  gp.slot_planes = 4;
  for (int i = 0; i < np; ++i) for (int a = 0; a < 3; ++a) gp.start[i][a] = plan[i].start[a];
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = 0;
  for (int it = 0; it < reps + 3 && rc == 0; ++it) {
    if (it == 3) cudaEventRecord(e0, 0);
    cudaMemsetAsync(counts, 0, 13 * 8, 0);
    if (form == 0) {
      rc = launch_gather_finalize(slots, gp, gaussian, shape[0], shape[1], zout, nullptr, labels, nullptr, counts, 0);
    } else {
      cudaMemsetAsync(acc, 0, 4 * V * 4, 0); cudaMemsetAsync(wsum, 0, V * 4, 0);
      for (int i = 0; i < np && rc == 0; ++i)
        rc = launch_accumulate(slots + (int64_t)i * 4 * P3, plan[i].start, gaussian, acc, wsum, shape[0], shape[1], zout, 0);
      if (rc == 0) rc = launch_finalize_labels(acc, wsum, V, 0, V, nullptr, labels, nullptr, counts, 0);
    }
  }
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(slots); cudaFree(acc); cudaFree(wsum); cudaFree(labels); cudaFree(counts);
  return rc == 0 ? (double)ms * 1e3 / reps : -1.0;
}

DCL_API int dcl_op_instnorm_stats(const float* x, int32_t channels, int64_t spatial, float* mean, float* rstd, void* stream) {
  if (!x || !mean || !rstd || channels <= 0 || channels > 512) { set_error("dcl_op_instnorm_stats: bad argument"); return DCL_ERR_ARG; }
  double* accum = nullptr;
  DCL_CUDA_OK(cudaMalloc((void**)&accum, 2 * 512 * sizeof(double)));
  cudaMemsetAsync(accum, 0, 2 * 512 * sizeof(double), (cudaStream_t)stream);
  int rc = launch_instnorm_stats(x, channels, spatial, accum, mean, rstd, (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(accum);
  return rc;
}

DCL_API int dcl_op_conv3d_k3(const float* x0, int32_t c0, const float* x1, int32_t c1, const int32_t in_dhw[3], const float* w,
                     const float* bias, int32_t cout, int32_t stride, const float* norm_mean, const float* norm_rstd,
                     int32_t act, const float* residual, float* y, int32_t impl, double* stats_out, void* stream) {
  if (!x0 || !w || !y || !in_dhw || cout <= 0 || c0 <= 0) { set_error("dcl_op_conv3d_k3: bad argument"); return DCL_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int cin = c0 + c1;
  std::vector<float> wh((size_t)cout * cin * 27);
  DCL_CUDA_OK(cudaMemcpy(wh.data(), w, wh.size() * 4, cudaMemcpyDefault));
  const int64_t sp = (int64_t)in_dhw[0] * in_dhw[1] * in_dhw[2];
  ConvSrc s{x0, x1, c0, c1, sp, (int64_t)in_dhw[1] * in_dhw[2], in_dhw[2], norm_mean, norm_rstd, act};
  ConvDst d{y, bias, nullptr, residual, nullptr};
  if (stats_out && impl == 0) { set_error("dcl_op_conv3d_k3: fused statistics need a tensor-core impl"); return DCL_ERR_ARG; }
  int rc;
  if (impl == 0) {
    const int cout_pad = (cout + 15) / 16 * 16;
    std::vector<float> packed((size_t)cin * 27 * cout_pad, 0.f);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < 27; ++t) packed[((size_t)ci * 27 + t) * cout_pad + co] = wh[((size_t)co * cin + ci) * 27 + t];
    float* wp = nullptr;
    DCL_CUDA_OK(cudaMalloc((void**)&wp, packed.size() * 4));
    DCL_CUDA_OK(cudaMemcpy(wp, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
    rc = launch_conv3d_k3(s, d, wp, cout, cout_pad, in_dhw[0], in_dhw[1], in_dhw[2], stride, st);
    cudaStreamSynchronize(st);
    cudaFree(wp);
  } else {
    const bool x3 = impl == 1;           // split-fp16 operands (DCL_F16X3)
    const int E = x3 ? 4 : 2;            // bytes per B-format element
    TcWeights tw;
    const bool cubic = in_dhw[0] == in_dhw[1] && in_dhw[1] == in_dhw[2];
    const bool roll = cubic && x1 == nullptr && tc_conv_supported(cin, cout, in_dhw[0], stride, x3) && cin != 4;
    DCL_TRY(tc_pack_weights(wh.data(), cout, cin, 27, roll, &tw, x3));
    if (roll) {
      // rolling kernel: B-format in / out, so convert around it (raw input; the norm is fused in the kernel)
      void *xb = nullptr, *yb = nullptr, *rb = nullptr;
      DCL_CUDA_OK(cudaMalloc(&xb, (size_t)cin * sp * E));
      DCL_CUDA_OK(cudaMalloc(&yb, (size_t)cout * sp * E));
      ConvSrc raw{x0, nullptr, c0, 0, sp, (int64_t)in_dhw[1] * in_dhw[2], in_dhw[2], nullptr, nullptr, ACT_NONE};
      rc = launch_prep_blocked(raw, in_dhw[0], in_dhw[1], in_dhw[2], xb, st, x3);
      if (rc == 0 && residual) {
        DCL_CUDA_OK(cudaMalloc(&rb, (size_t)cout * sp * E));
        ConvSrc rs{residual, nullptr, cout, 0, sp, (int64_t)in_dhw[1] * in_dhw[2], in_dhw[2], nullptr, nullptr, ACT_NONE};
        rc = launch_prep_blocked(rs, in_dhw[0], in_dhw[1], in_dhw[2], rb, st, x3);
      }
      RollArgs a;
      a.x3 = x3;
      a.xb = xb; a.norm.mean = norm_mean; a.norm.rstd = norm_rstd; a.norm.act = act;
      stat_t* sfix = nullptr;
      if (stats_out) {
        DCL_CUDA_OK(cudaMalloc((void**)&sfix, 2 * cout * sizeof(stat_t)));
        DCL_CUDA_OK(cudaMemsetAsync(sfix, 0, 2 * cout * sizeof(stat_t), st));
      }
      a.bias = bias; a.resb = rb; a.yb = yb; a.stats = sfix;
      if (rc == 0) rc = launch_roll_conv(a, tw, cout, in_dhw[0], st);
      if (rc == 0) rc = launch_unblock(yb, y, cout, sp, st, x3);
      cudaStreamSynchronize(st);
      if (stats_out && rc == 0) {   // decode the fixed-point sums into the caller's doubles
        std::vector<stat_t> hs(2 * cout);
        std::vector<double> hd(2 * cout);
        DCL_CUDA_OK(cudaMemcpy(hs.data(), sfix, hs.size() * sizeof(stat_t), cudaMemcpyDeviceToHost));
        for (int c = 0; c < cout; ++c) { hd[2 * c] = (double)hs[2 * c] / STAT_SCALE_S; hd[2 * c + 1] = (double)hs[2 * c + 1] / STAT_SCALE_Q; }
        DCL_CUDA_OK(cudaMemcpy(stats_out, hd.data(), hd.size() * sizeof(double), cudaMemcpyDefault));
      }
      if (sfix) cudaFree(sfix);
      cudaFree(xb); cudaFree(yb); if (rb) cudaFree(rb);
    } else {
      if (stats_out) { cudaFree(tw.dev); set_error("dcl_op_conv3d_k3: fused statistics need the rolling kernel"); return DCL_ERR_ARG; }
      void* blk = nullptr;
      DCL_CUDA_OK(cudaMalloc(&blk, (size_t)((cin + 15) / 16 * 16) * sp * E));
      if (stride == 2 && cubic && x1 == nullptr && norm_mean == nullptr && act == ACT_NONE && residual == nullptr &&
          (x3 ? s2_roll_supported_x3(cin, cout, in_dhw[0]) : s2_roll_supported(cin, cout, in_dhw[0]))) {
        // rolling stride-2 kernel (EnDown1): B-format in / out
        const int64_t osp = sp / 8;
        void* yb = nullptr;
        DCL_CUDA_OK(cudaMalloc(&yb, (size_t)cout * osp * E));
        rc = launch_prep_blocked(s, in_dhw[0], in_dhw[1], in_dhw[2], blk, st, x3);
        if (rc == 0) rc = launch_s2_roll_conv(blk, tw, bias, yb, nullptr, st, x3);
        if (rc == 0) rc = launch_unblock(yb, y, cout, osp, st, x3);
        cudaStreamSynchronize(st);
        cudaFree(yb);
      } else if (slab_conv_supported(cin, cout, in_dhw[0], in_dhw[1], in_dhw[2], stride, 27)) {
        // slab kernel: raw blocked input, norm fused in the kernel, fp32 NCDHW output
        ConvSrc raw = s;
        raw.mean = nullptr; raw.rstd = nullptr; raw.act = ACT_NONE;
        rc = launch_prep_blocked(raw, in_dhw[0], in_dhw[1], in_dhw[2], blk, st, x3);
        GemmArgs ga;
        ga.x3 = x3;
        ga.a0 = blk; ga.c0 = (cin + 15) / 16 * 16;
        ga.D = in_dhw[0]; ga.H = in_dhw[1]; ga.W = in_dhw[2]; ga.stride = 1; ga.taps = 27;
        ga.bias = bias; ga.out_mode = 0; ga.y = y; ga.residual = residual;
        BNorm bn;
        bn.mean = norm_mean; bn.rstd = norm_rstd; bn.act = act;
        if (rc == 0) rc = launch_slab_conv(ga, &bn, tw, st);
      } else {
        rc = launch_prep_blocked(s, in_dhw[0], in_dhw[1], in_dhw[2], blk, st, x3);
        if (rc == 0) rc = launch_conv_gemm(blk, tw, d, in_dhw[0], in_dhw[1], in_dhw[2], stride, 27, st, x3);
      }
      cudaStreamSynchronize(st);
      cudaFree(blk);
    }
    cudaStreamSynchronize(st);
    tc_free_weights(&tw);
  }
  return rc;
}

}  // extern "C"
