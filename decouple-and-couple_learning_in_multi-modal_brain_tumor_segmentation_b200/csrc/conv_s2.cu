// Rolling tcgen05 kernels for the stride-2 convolutions nn.Conv3d(C, C', k3, s2, p1) of the encoder and the edge
// branch: EnDown1 16->32 on the 128^3 level (Unet_skipconnection.py:60-68, :122), EnDown2 32->64 and conv_64_to_32
// 32->32 on the 64^3 level (:126, cls_wise_former.py:284-286).  The general im2col GEMM (conv_gemm.cu) gathers
// 27 x 16-byte vectors per output voxel through cp.async and is L2-gather bound there (EnDown1: ~180 us; EnDown2 33 us,
// conv_64_to_32 32 us); this kernel stages every input row once: 37 / 18.6 / 15.0 us.  DCL_S2GEN=0 keeps the 64^3
// layers on the GEMM.
//
//   GEMM view   M = 128 outputs = RPT output rows x GO ow (RPT = 128 / GO), N = CO, K = 27 taps x CI.
//   CTA         TH = RPT x NT output rows x all GO ow, walking along od; a ring of NSLOT staged input planes (an output
//               plane needs input planes 2od-1, 2od, 2od+1 and the walk advances by two).
//   staging     a stride-2 tap reads every other voxel, which no UMMA descriptor can express (a core matrix is 8
//               rows 16 bytes apart).  The producers therefore DE-INTERLEAVE while staging: a staged plane holds
//               four blocks (odd|even w) x (odd|even input row), each [rows][GO positions] of 16 bytes.  Output
//               (oh, ow) tap (kh, kw) reads input row 2oh+kh-1 and column 2ow+kw-1, i.e. row list odd/even/odd at
//               index oh-1|oh|oh and column list odd/even/odd at index ow-1|ow|ow: every tap of an RPT-row tile is
//               one contiguous 128-position run again.  kw = 0 starts one position early; the lanes with ow = 0
//               are switched off with the MMA's disable-output-lane mask (zero padding).
//   warps       0-3 epilogue (bias, fused statistics, B-format stores), 4 MMA issuer, 5-12 producers.
#include <stdlib.h>

#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace dcl {

using namespace tc;

// X3_: split-fp16 (DCL_F16X3): staged planes and weights carry their lo halves behind the hi halves (as in global
// memory), every (tap, K step) issues a_hi*w_hi + a_lo*w_hi + a_hi*w_lo.
template <int CI_, int CO_, int GI_, int NT_, int NSLOT_, bool X3_ = false>
struct S2Cfg {
  static constexpr int CI = CI_, CO = CO_, GI = GI_, GO = GI_ / 2, NT = NT_, NSLOT = NSLOT_;
  static constexpr bool X3 = X3_;
  static constexpr int RPT = 128 / GO;                      // output rows per 128-row MMA tile
  static constexpr int TH = RPT * NT;                       // output rows per CTA
  static constexpr int ROWS = 2 * TH + 1;                   // staged input rows 2*oh0-1 .. 2*oh0+2*TH-1
  static constexpr int N_ODD = TH + 1, N_EVEN = TH;         // odd / even input rows among them
  static constexpr int B_OO = 8;                            // positions 0..7 = pad (keeps every block 128-byte aligned)
  static constexpr int B_OE = B_OO + N_ODD * GO;
  static constexpr int B_EO = B_OE + N_EVEN * GO;
  static constexpr int B_EE = B_EO + N_ODD * GO;
  static constexpr int NPOS = B_EE + N_EVEN * GO;           // 16-byte positions per channel chunk
  static constexpr int KC = CI / 8;
  static constexpr int KCS = X3 ? 2 * KC : KC;              // staged chunks
  static constexpr int KSTEPS = CI / 16;                    // K = 16 MMAs per tap
  static constexpr int SLOT_BYTES = KCS * NPOS * 16;
  static constexpr int ACC_COLS = NT * CO;
  static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;     // two accumulator buffers
  static constexpr int W_HALF = 27 * CI * CO * 2;
  static constexpr int W_BYTES = X3 ? 2 * W_HALF : W_HALF;
  static constexpr int OFF_W = NSLOT * SLOT_BYTES;
  static constexpr int OFF_BIAS = OFF_W + W_BYTES;
  static constexpr int OFF_RED = OFF_BIAS + CO * 4;         // [4 epilogue warps][CO][2] partial statistics
  static constexpr int OFF_BAR = OFF_RED + 4 * CO * 2 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + (2 * NSLOT + 4) * 8 + 16;
  static constexpr int EPI_WARPS = 4, PROD_WARPS = 8;
  static constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;
  static constexpr int NPROD = PROD_WARPS * 32;
  // disable-output-lane mask of the kw = 0 taps: M rows with ow = 0, i.e. the multiples of GO
  static constexpr uint32_t mask_word(int k) { return GO >= 32 ? ((32 * k) % GO == 0 ? 1u : 0u) : 0x00010001u; }
  static_assert(GO == 16 || GO == 32 || GO == 64, "one MMA tile = whole output rows");
  static_assert(CI % 16 == 0 && CO % 16 == 0 && CO <= 256, "channel counts");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM allocation is a power of two");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(OFF_W % 128 == 0 && SLOT_BYTES % 128 == 0, "operand alignment");
};

struct S2Params {
  const uint4* xb;      // B-format input, CI channels @ GI^3
  const uint4* w;       // packed weights, layout 0: [tap][cin/8][CO][8]
  const float* bias;
  uint4* yb;            // B-format output, CO channels @ GO^3
  stat_t* stats;        // 2*CO fixed-point sums or nullptr
  float acc_mul;        // accumulators are multiplied by this (split mode: 2^-k of the weight scale; else 1)
  int dsplit;
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
conv3d_k3s2_roll_kernel(S2Params prm) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  constexpr int CI = C::CI, CO = C::CO, GI = C::GI, GO = C::GO, NT = C::NT, NSLOT = C::NSLOT, RPT = C::RPT, TH = C::TH;
  constexpr int ROWS = C::ROWS, NPOS = C::NPOS, KC = C::KC, KCS = C::KCS, SLOT_BYTES = C::SLOT_BYTES, ACC_COLS = C::ACC_COLS;
  constexpr bool X3 = C::X3;
  constexpr int EPI_WARPS = C::EPI_WARPS, THREADS = C::THREADS, NPROD = C::NPROD;
  extern __shared__ __align__(128) uint8_t smem[];
  float* s_bias = reinterpret_cast<float*>(smem + C::OFF_BIAS);
  float* s_red = reinterpret_cast<float*>(smem + C::OFF_RED);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_empty = bar_full + NSLOT;
  uint64_t* bar_acc_full = bar_empty + NSLOT;
  uint64_t* bar_acc_empty = bar_acc_full + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform role index
  const int dsplit = prm.dsplit;
  const int ht = blockIdx.x / dsplit;
  const int ds = blockIdx.x - ht * dsplit;
  const int oh0 = ht * TH;
  const int od0 = (ds * GO) / dsplit, od1 = ((ds + 1) * GO) / dsplit;
  const int n_out = od1 - od0;
  const int n_in = 2 * n_out + 1;                  // staged plane j holds input plane 2*od0 - 1 + j
  constexpr int64_t SPI = (int64_t)GI * GI * GI, SPO = (int64_t)GO * GO * GO;

  for (int i = tid; i < C::W_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem + C::OFF_W)[i] = __ldg(prm.w + i);
  for (int i = tid; i < NSLOT * KCS * 8; i += THREADS) {   // the pad positions of every slot chunk (read by masked lanes only)
    const int sc = i >> 3;
    *reinterpret_cast<uint4*>(smem + (size_t)(sc / KCS) * SLOT_BYTES + (size_t)((sc % KCS) * NPOS + (i & 7)) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = tid; i < CO; i += THREADS) s_bias[i] = prm.bias ? prm.bias[i] : 0.f;
  for (int i = tid; i < 4 * CO * 2; i += THREADS) s_red[i] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&bar_full[s], NPROD); mbar_init(&bar_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_empty[b], EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) tmem_alloc(s_tmem, C::TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);

  if (warp >= EPI_WARPS + 1) {
    // =============================== producers: de-interleaving stage ==============================
    const int pt = tid - (EPI_WARPS + 1) * 32;
    for (int j = 0; j < n_in; ++j) {
      const int s = j % NSLOT;
      mbar_wait(&bar_empty[s], ((uint32_t)(j / NSLOT) & 1u) ^ 1u);
      const int d_in = 2 * od0 - 1 + j;
      const bool d_ok = d_in >= 0;                  // the high side never leaves the volume (2*(GO-1)+1 = GI-1)
      uint8_t* slot = smem + (size_t)s * SLOT_BYTES;
      // item = (chunk, staged row, w): consecutive threads read consecutive 16-byte vectors of one row
      constexpr int ITEMS = KCS * ROWS * GI;      // (split-fp16: global chunk KC + k is the lo half of chunk k, as in the slot)
      constexpr int U = 8;                        // 16-byte loads in flight per producer thread (the stage is load-latency bound)
      for (int e0 = pt; e0 < ITEMS; e0 += U * NPROD) {
        uint4 v[U];
        int pos[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int e = e0 + u * NPROD;
          v[u] = make_uint4(0u, 0u, 0u, 0u);
          pos[u] = -1;
          if (e < ITEMS) {
            const int kc = e / (ROWS * GI);
            const int rem = e - kc * (ROWS * GI);
            const int r = rem / GI, w = rem - r * GI;
            const int h_in = 2 * oh0 - 1 + r;
            // r even <-> odd input row (index r/2), r odd <-> even input row (index (r-1)/2)
            const int base = (w & 1) ? ((r & 1) ? C::B_OE : C::B_OO) : ((r & 1) ? C::B_EE : C::B_EO);
            pos[u] = kc * NPOS + base + (r >> 1) * GO + (w >> 1);
            if (d_ok && h_in >= 0) v[u] = __ldg(prm.xb + (int64_t)kc * SPI + ((int64_t)d_in * GI + h_in) * GI + w);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (pos[u] >= 0) *reinterpret_cast<uint4*>(slot + (size_t)pos[u] * 16) = v[u];
      }
      fence_proxy_async();
      mbar_arrive(&bar_full[s]);
    }
  } else if (warp == EPI_WARPS) {
    // =============================== MMA issuer ====================================================
    constexpr uint32_t idesc = umma_idesc_16(128, CO, X3);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t b_base = umma_desc(smem_base + C::OFF_W, CO * 16, 128);
    for (int i = 0; i < n_out; ++i) {
      const int b = i & 1;
      if (i == 0) {
        mbar_wait(&bar_full[0], 0);
        mbar_wait(&bar_full[1 % NSLOT], 0);
      } else {
        mbar_wait(&bar_full[(2 * i + 1) % NSLOT], (uint32_t)((2 * i + 1) / NSLOT) & 1u);
      }
      mbar_wait(&bar_full[(2 * i + 2) % NSLOT], (uint32_t)((2 * i + 2) / NSLOT) & 1u);
      mbar_wait(&bar_acc_empty[b], ((uint32_t)(i >> 1) & 1u) ^ 1u);
      tc_fence_after();
      uint64_t a_kd[3];
#pragma unroll
      for (int kd = 0; kd < 3; ++kd)
        a_kd[kd] = umma_desc(smem_base + (uint32_t)(((2 * i + kd) % NSLOT) * SLOT_BYTES), NPOS * 16, 128);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * ACC_COLS + t * CO);
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            // kh = 0: odd rows at index RPT*t, kh = 1: even rows at RPT*t, kh = 2: odd rows at RPT*t+1
            const int ridx = RPT * t + (kh == 2 ? 1 : 0);
#pragma unroll
            for (int kwi = 0; kwi < 3; ++kwi) {
              const int kw = kwi == 0 ? 1 : (kwi == 1 ? 0 : 2);     // the first MMA of a tile must be unmasked
              const int base = (kw == 1) ? (kh == 1 ? C::B_EE : C::B_EO) : (kh == 1 ? C::B_OE : C::B_OO);
              const int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
              for (int ks = 0; ks < C::KSTEPS; ++ks) {
#pragma unroll
                for (int v = 0; v < (X3 ? 3 : 1); ++v) {
                  const uint64_t ad = a_kd[kd] + (uint64_t)(uint32_t)(base + ridx * GO - (kw == 0 ? 1 : 0) + ks * 2 * NPOS + (v == 1 ? KC * NPOS : 0));
                  const uint64_t bd = b_base + (uint64_t)((tap * CI * CO * 2 + ks * 2 * CO * 16 + (v == 2 ? C::W_HALF : 0)) >> 4);
                  const uint32_t accum = (kd | kh | kwi | ks | v) != 0 ? 1u : 0u;
                  if (kw == 0)      // ow = 0: the M rows that are multiples of GO
                    umma_bf16_masked_ws(d_tmem, ad, bd, idesc, accum, C::mask_word(0), C::mask_word(1), C::mask_word(2), C::mask_word(3));
                  else umma_bf16_ws(d_tmem, ad, bd, idesc, accum);
                }
              }
            }
          }
        }
      }
      umma_commit_ws(&bar_acc_full[b]);
      umma_commit_ws(&bar_empty[(2 * i) % NSLOT]);          // planes 2i and 2i+1 are not needed again
      umma_commit_ws(&bar_empty[(2 * i + 1) % NSLOT]);
    }
    __syncwarp();
  } else {
    // =============================== epilogue =======================================================
    // 16 output channels at a time.  Statistics: up to 32 channels are summed per thread over the CTA's whole walk and
    // reduced once at the end; wider layers reduce over the warp per tile (fixed order) and one lane adds into this
    // warp's own shared-memory partials.  Either way the sums do not depend on any scheduling.
    constexpr bool PER_THREAD = CO <= 32;
    float st_s[PER_THREAD ? CO : 1], st_q[PER_THREAD ? CO : 1];
#pragma unroll
    for (int c = 0; c < (PER_THREAD ? CO : 1); ++c) { st_s[c] = 0.f; st_q[c] = 0.f; }
    const int m = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int i = 0; i < n_out; ++i) {
      const int b = i & 1;
      const int od = od0 + i;
      mbar_wait(&bar_acc_full[b], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oh = oh0 + RPT * t + m / GO, ow = m % GO;
        const int64_t off = ((int64_t)od * GO + oh) * GO + ow;
#pragma unroll
        for (int g16 = 0; g16 < CO / 16; ++g16) {
          uint32_t acc[16];
          tmem_ld16(lane_addr + (uint32_t)(b * ACC_COLS + t * CO + g16 * 16), acc);
          tmem_ld_wait();
          if (t == NT - 1 && g16 == CO / 16 - 1) {
            tc_fence_before();
            mbar_arrive(&bar_acc_empty[b]);
          }
          float val[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) val[k] = __uint_as_float(acc[k]) * prm.acc_mul + s_bias[g16 * 16 + k];
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            uint4 o;
            if constexpr (X3) {
              uint4 l;
              split_x2(val[h8 * 8 + 0], val[h8 * 8 + 1], o.x, l.x);
              split_x2(val[h8 * 8 + 2], val[h8 * 8 + 3], o.y, l.y);
              split_x2(val[h8 * 8 + 4], val[h8 * 8 + 5], o.z, l.z);
              split_x2(val[h8 * 8 + 6], val[h8 * 8 + 7], o.w, l.w);
              prm.yb[(int64_t)(CO / 8 + g16 * 2 + h8) * SPO + off] = l;
            } else {
              o.x = pack_bf16x2(val[h8 * 8 + 0], val[h8 * 8 + 1]);
              o.y = pack_bf16x2(val[h8 * 8 + 2], val[h8 * 8 + 3]);
              o.z = pack_bf16x2(val[h8 * 8 + 4], val[h8 * 8 + 5]);
              o.w = pack_bf16x2(val[h8 * 8 + 6], val[h8 * 8 + 7]);
            }
            prm.yb[(int64_t)(g16 * 2 + h8) * SPO + off] = o;
          }
          if constexpr (PER_THREAD) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              st_s[g16 * 16 + k] += val[k];
              st_q[g16 * 16 + k] += val[k] * val[k];
            }
          } else {
            if (prm.stats != nullptr) {
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                float a = val[k], q = val[k] * val[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  a += __shfl_xor_sync(0xffffffffu, a, o);
                  q += __shfl_xor_sync(0xffffffffu, q, o);
                }
                if (lane == 0) {
                  const int c = g16 * 16 + k;
                  s_red[(warp * CO + c) * 2] += a;
                  s_red[(warp * CO + c) * 2 + 1] += q;
                }
              }
            }
          }
        }
      }
    }
    if (prm.stats != nullptr) {
      // CTA-level reduction (fixed order), then one pair of atomics per channel and CTA
      if constexpr (PER_THREAD) {
#pragma unroll
        for (int c = 0; c < CO; ++c) {
          float a = st_s[c], q = st_q[c];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
          }
          if (lane == 0) { s_red[(warp * CO + c) * 2] = a; s_red[(warp * CO + c) * 2 + 1] = q; }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      for (int c = m; c < CO; c += EPI_WARPS * 32) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < EPI_WARPS; ++w4) { a += s_red[(w4 * CO + c) * 2]; q += s_red[(w4 * CO + c) * 2 + 1]; }
        stat_add(prm.stats, c, a, q);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

using S2EnDown1 = S2Cfg<16, 32, 128, 2, 5>;     // 16 -> 32 @ 128^3
using S2EnDown2 = S2Cfg<32, 64, 64, 1, 3>;      // 32 -> 64 @ 64^3: 110 KB of weights leave room for three staged planes
using S2Edge = S2Cfg<32, 32, 64, 1, 4>;         // conv_64_to_32: 32 -> 32 @ 64^3
using S2EnDown1X3 = S2Cfg<16, 32, 128, 1, 4, true>;   // split-fp16: two output rows per CTA, 221 KB (the 64^3 layers' split
                                                      // weights alone are 221 KB: they stay on the GEMM kernel)

static bool s2_general_enabled() {      // DCL_S2GEN=0 sends the 64^3 layers back to the im2col GEMM
  static const bool on = [] { const char* e = getenv("DCL_S2GEN"); return e == nullptr || e[0] != '0'; }();
  return on;
}

bool s2_roll_supported_x3(int cin, int cout, int g) { return cin == 16 && cout == 32 && g == 128; }

bool s2_roll_supported(int cin, int cout, int g) {
  if (cin == 16 && cout == 32 && g == 128) return true;
  if (!s2_general_enabled()) return false;
  return g == 64 && cin == 32 && (cout == 64 || cout == 32);
}

template <class C>
static int launch_s2(const void* xb, const TcWeights& w, const float* bias, void* yb, stat_t* stats, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    DCL_CUDA_OK(cudaFuncSetAttribute(conv3d_k3s2_roll_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  S2Params p;
  p.xb = reinterpret_cast<const uint4*>(xb);
  p.w = reinterpret_cast<const uint4*>(w.dev);
  p.bias = bias;
  p.yb = reinterpret_cast<uint4*>(yb);
  p.stats = stats;
  p.acc_mul = C::X3 ? w.out_mul : 1.f;
  const int htiles = C::GO / C::TH;
  p.dsplit = 148 / htiles;
  if (p.dsplit > C::GO) p.dsplit = C::GO;
  DCL_CUDA_OK(launch_pdl(conv3d_k3s2_roll_kernel<C>, dim3(htiles * p.dsplit), dim3(C::THREADS), (size_t)(C::SMEM_BYTES), st, p));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// xb: B-format CI channels @ GI^3, yb: B-format CO channels @ (GI/2)^3; w packed by tc_pack_weights(taps = 27,
// roll_layout = false).  The geometry is recovered from the weights (cin, cout): 16 -> 32 is the 128^3 level, the
// 32-input layers are the 64^3 level.
int launch_s2_roll_conv(const void* xb, const TcWeights& w, const float* bias, void* yb, stat_t* stats, cudaStream_t st, bool x3) {
  if (w.dev == nullptr) { set_error("s2_roll_conv: weights not packed"); return -1; }
  if (x3) {
    if (w.lo_off == 0 || !(w.cin == 16 && w.cout == 32)) { set_error("s2_roll_conv: unsupported split-fp16 layer"); return -1; }
    return launch_s2<S2EnDown1X3>(xb, w, bias, yb, stats, st);
  }
  if (w.cin == 16 && w.cout == 32) return launch_s2<S2EnDown1>(xb, w, bias, yb, stats, st);
  if (w.cin == 32 && w.cout == 64) return launch_s2<S2EnDown2>(xb, w, bias, yb, stats, st);
  if (w.cin == 32 && w.cout == 32) return launch_s2<S2Edge>(xb, w, bias, yb, stats, st);
  set_error("s2_roll_conv: unsupported channel counts");
  return -1;
}

}  // namespace dcl
