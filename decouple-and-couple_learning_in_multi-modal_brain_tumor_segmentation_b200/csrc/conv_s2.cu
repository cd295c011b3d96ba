// Rolling tcgen05 kernel for the first stride-2 convolution of the encoder: EnDown1, nn.Conv3d(16, 32, k3, s2, p1)
// on the 128^3 level (Unet_skipconnection.py:60-68, :122).  The general im2col GEMM (conv_gemm.cu) gathers
// 27 x 16-byte vectors per output voxel through cp.async and is L2-gather bound there (~180 us); this kernel
// stages every input row once.
//
//   GEMM view   M = 128 outputs = 2 output rows x 64 ow, N = 32, K = 27 taps x 16.
//   CTA         TH = 4 output rows x all 64 ow, walking along od; a ring of 5 staged input planes (an output plane
//               needs input planes 2od-1, 2od, 2od+1 and the walk advances by two).
//   staging     a stride-2 tap reads every other voxel, which no UMMA descriptor can express (a core matrix is 8
//               rows 16 bytes apart).  The producers therefore DE-INTERLEAVE while staging: a staged plane holds
//               four blocks (odd|even w) x (odd|even input row), each [rows][64 positions] of 16 bytes.  Output
//               (oh, ow) tap (kh, kw) reads input row 2oh+kh-1 and column 2ow+kw-1, i.e. row list odd/even/odd at
//               index oh-1|oh|oh and column list odd/even/odd at index ow-1|ow|ow: every tap of a 2-row tile is
//               one contiguous 128-position run again.  kw = 0 starts one position early; the lanes with ow = 0
//               are switched off with the MMA's disable-output-lane mask (zero padding).
//   warps       0-3 epilogue (bias, fused statistics, B-format stores), 4 MMA issuer, 5-12 producers.
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace dcl {

using namespace tc;

namespace s2 {
constexpr int CI = 16, CO = 32, GI = 128, GO = 64, TH = 4;
constexpr int ROWS = 2 * TH + 1;                   // staged input rows 2*oh0-1 .. 2*oh0+2*TH-1
constexpr int N_ODD = TH + 1, N_EVEN = TH;         // odd / even input rows among them
constexpr int B_OO = 8;                            // positions 0..7 = pad (keeps every block 128-byte aligned)
constexpr int B_OE = B_OO + N_ODD * 64;
constexpr int B_EO = B_OE + N_EVEN * 64;
constexpr int B_EE = B_EO + N_ODD * 64;
constexpr int NPOS = B_EE + N_EVEN * 64;           // 16-byte positions per channel chunk
constexpr int KC = CI / 8;
constexpr int SLOT_BYTES = KC * NPOS * 16;
constexpr int NSLOT = 5;
constexpr int NT = TH / 2;                         // 128-output tiles per output plane
constexpr int ACC_COLS = NT * CO;
constexpr int TMEM_COLS = 128;                     // 2 buffers x 64 columns
constexpr int W_BYTES = 27 * CI * CO * 2;
constexpr int OFF_W = NSLOT * SLOT_BYTES;
constexpr int OFF_BIAS = OFF_W + W_BYTES;
constexpr int OFF_BAR = OFF_BIAS + CO * 4;
constexpr int SMEM_BYTES = OFF_BAR + (2 * NSLOT + 4) * 8 + 16;
constexpr int EPI_WARPS = 4, PROD_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;
constexpr int NPROD = PROD_WARPS * 32;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(2 * ACC_COLS <= TMEM_COLS, "accumulators exceed the allocation");
}  // namespace s2

struct S2Params {
  const uint4* xb;      // B-format input, 16 channels @ 128^3
  const uint4* w;       // packed weights, layout 0: [tap][cin/8][32][8]
  const float* bias;
  uint4* yb;            // B-format output, 32 channels @ 64^3
  stat_t* stats;        // 2*32 fixed-point sums or nullptr
  int dsplit;
};

__global__ void __launch_bounds__(s2::THREADS, 1)
conv3d_k3s2_roll_kernel(S2Params prm) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  using namespace s2;
  extern __shared__ __align__(128) uint8_t smem[];
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_empty = bar_full + NSLOT;
  uint64_t* bar_acc_full = bar_empty + NSLOT;
  uint64_t* bar_acc_empty = bar_acc_full + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dsplit = prm.dsplit;
  const int ht = blockIdx.x / dsplit;
  const int ds = blockIdx.x - ht * dsplit;
  const int oh0 = ht * TH;
  const int od0 = (ds * GO) / dsplit, od1 = ((ds + 1) * GO) / dsplit;
  const int n_out = od1 - od0;
  const int n_in = 2 * n_out + 1;                  // staged plane j holds input plane 2*od0 - 1 + j
  constexpr int64_t SPI = (int64_t)GI * GI * GI, SPO = (int64_t)GO * GO * GO;

  for (int i = tid; i < W_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem + OFF_W)[i] = __ldg(prm.w + i);
  for (int i = tid; i < NSLOT * KC * 8; i += THREADS) {   // the pad positions of every slot chunk (read by masked lanes only)
    const int sc = i >> 3;
    *reinterpret_cast<uint4*>(smem + (size_t)(sc / KC) * SLOT_BYTES + (size_t)((sc % KC) * NPOS + (i & 7)) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < CO) s_bias[tid] = prm.bias ? prm.bias[tid] : 0.f;
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&bar_full[s], NPROD); mbar_init(&bar_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_empty[b], EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) tmem_alloc(s_tmem, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp >= EPI_WARPS + 1) {
    // =============================== producers: de-interleaving stage ==============================
    const int pt = tid - (EPI_WARPS + 1) * 32;
    for (int j = 0; j < n_in; ++j) {
      const int s = j % NSLOT;
      mbar_wait(&bar_empty[s], ((uint32_t)(j / NSLOT) & 1u) ^ 1u);
      const int d_in = 2 * od0 - 1 + j;
      const bool d_ok = d_in >= 0;                  // the high side never leaves the volume (2*63+1 = 127)
      uint8_t* slot = smem + (size_t)s * SLOT_BYTES;
      // item = (chunk, staged row, w): consecutive threads read consecutive 16-byte vectors of one row
      constexpr int ITEMS = KC * ROWS * GI;
      for (int e0 = pt; e0 < ITEMS; e0 += 4 * NPROD) {
        uint4 v[4];
        int pos[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * NPROD;
          v[u] = make_uint4(0u, 0u, 0u, 0u);
          pos[u] = -1;
          if (e < ITEMS) {
            const int kc = e / (ROWS * GI);
            const int rem = e - kc * (ROWS * GI);
            const int r = rem / GI, w = rem - r * GI;
            const int h_in = 2 * oh0 - 1 + r;
            // r even <-> odd input row (index r/2), r odd <-> even input row (index (r-1)/2)
            const int base = (w & 1) ? ((r & 1) ? B_OE : B_OO) : ((r & 1) ? B_EE : B_EO);
            pos[u] = kc * NPOS + base + (r >> 1) * 64 + (w >> 1);
            if (d_ok && h_in >= 0) v[u] = __ldg(prm.xb + (int64_t)kc * SPI + ((int64_t)d_in * GI + h_in) * GI + w);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (pos[u] >= 0) *reinterpret_cast<uint4*>(slot + (size_t)pos[u] * 16) = v[u];
      }
      fence_proxy_async();
      mbar_arrive(&bar_full[s]);
    }
  } else if (warp == EPI_WARPS) {
    // =============================== MMA issuer ====================================================
    constexpr uint32_t idesc = umma_idesc_bf16(128, CO);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t b_base = umma_desc(smem_base + OFF_W, CO * 16, 128);
    for (int i = 0; i < n_out; ++i) {
      const int b = i & 1;
      if (i == 0) {
        mbar_wait(&bar_full[0], 0);
        mbar_wait(&bar_full[1], 0);
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
            // kh = 0: odd rows at index 2t, kh = 1: even rows at 2t, kh = 2: odd rows at 2t+1
            const int ridx = 2 * t + (kh == 2 ? 1 : 0);
#pragma unroll
            for (int kwi = 0; kwi < 3; ++kwi) {
              const int kw = kwi == 0 ? 1 : (kwi == 1 ? 0 : 2);     // the first MMA of a tile must be unmasked
              const int base = (kw == 1) ? (kh == 1 ? B_EE : B_EO) : (kh == 1 ? B_OE : B_OO);
              const int tap = (kd * 3 + kh) * 3 + kw;
              const uint64_t ad = a_kd[kd] + (uint64_t)(uint32_t)(base + ridx * 64 - (kw == 0 ? 1 : 0));
              const uint64_t bd = b_base + (uint64_t)((tap * CI * CO * 2) >> 4);
              const uint32_t accum = (kd | kh | kwi) != 0 ? 1u : 0u;
              if (kw == 0) umma_bf16_masked_ws(d_tmem, ad, bd, idesc, accum, 1u, 0u, 1u, 0u);   // ow = 0: M rows 0, 64
              else umma_bf16_ws(d_tmem, ad, bd, idesc, accum);
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
    float st_s[CO], st_q[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) { st_s[c] = 0.f; st_q[c] = 0.f; }
    const int m = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int i = 0; i < n_out; ++i) {
      const int b = i & 1;
      const int od = od0 + i;
      mbar_wait(&bar_acc_full[b], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        uint32_t acc[2][16];
        tmem_ld16(lane_addr + (uint32_t)(b * ACC_COLS + t * CO), acc[0]);
        tmem_ld16(lane_addr + (uint32_t)(b * ACC_COLS + t * CO + 16), acc[1]);
        tmem_ld_wait();
        if (t == NT - 1) {
          tc_fence_before();
          mbar_arrive(&bar_acc_empty[b]);
        }
        const int oh = oh0 + 2 * t + (m >> 6), ow = m & 63;
        const int64_t off = ((int64_t)od * GO + oh) * GO + ow;
#pragma unroll
        for (int kc = 0; kc < CO / 8; ++kc) {
          float val[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = kc * 8 + k;
            val[k] = __uint_as_float(acc[c / 16][c % 16]) + s_bias[c];
            st_s[c] += val[k];
            st_q[c] += val[k] * val[k];
          }
          uint4 o;
          o.x = pack_bf16x2(val[0], val[1]);
          o.y = pack_bf16x2(val[2], val[3]);
          o.z = pack_bf16x2(val[4], val[5]);
          o.w = pack_bf16x2(val[6], val[7]);
          prm.yb[(int64_t)kc * SPO + off] = o;
        }
      }
    }
    if (prm.stats != nullptr) {
      // CTA-level reduction (fixed order), then one pair of atomics per channel and CTA
      float* s_red = reinterpret_cast<float*>(smem);      // the staged planes are dead by now: [4 warps][32][2]
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
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (warp == 0) {
        const int c = lane;      // CO == 32
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
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

bool s2_roll_supported(int cin, int cout, int g) { return cin == s2::CI && cout == s2::CO && g == s2::GI; }

// xb: B-format 16ch @ 128^3, yb: B-format 32ch @ 64^3; w packed by tc_pack_weights(taps = 27, roll_layout = false)
int launch_s2_roll_conv(const void* xb, const TcWeights& w, const float* bias, void* yb, stat_t* stats, cudaStream_t st) {
  if (w.dev == nullptr || w.cin != s2::CI || w.cout != s2::CO) { set_error("s2_roll_conv: expects the 16->32 weights"); return -1; }
  static bool configured = false;
  if (!configured) {
    DCL_CUDA_OK(cudaFuncSetAttribute(conv3d_k3s2_roll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s2::SMEM_BYTES));
    configured = true;
  }
  S2Params p;
  p.xb = reinterpret_cast<const uint4*>(xb);
  p.w = reinterpret_cast<const uint4*>(w.dev);
  p.bias = bias;
  p.yb = reinterpret_cast<uint4*>(yb);
  p.stats = stats;
  const int htiles = s2::GO / s2::TH;
  p.dsplit = 148 / htiles;
  DCL_CUDA_OK(launch_pdl(conv3d_k3s2_roll_kernel, dim3(htiles * p.dsplit), dim3(s2::THREADS), (size_t)(s2::SMEM_BYTES), st, p));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dcl
