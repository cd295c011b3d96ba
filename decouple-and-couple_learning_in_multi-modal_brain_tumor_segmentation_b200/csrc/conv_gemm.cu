// General tcgen05 implicit-GEMM convolution (precision mode DCL_BF16) for every conv the rolling kernel
// (conv_tc.cu) does not take: any Cin / Cout, kernel 3 (27 taps) or 1, stride 1 or 2, any input size.
// Replaces nn.Conv3d of Unet_skipconnection.py:60-78 (EnDown), cls_wise_former.py:284-328 (region
// decoupler), :257-263 (sum_fusion), :691-713 / :732-754 at 16^3 and 32^3.
//
//   prep kernel   fp32 NCDHW (optionally two concatenated sources) -> InstanceNorm + activation ->
//                 bf16, channel-blocked [Cin/8][D][H][W][8]: every (voxel, 8 channels) is one 16-byte
//                 vector, so a shifted / strided tap of 128 output voxels is 128 x 16-byte gathers.
//   GEMM kernel   M = 128 output voxels, N = n_tile output channels, K = taps x Cin.
//     producers (4 warps)  cp.async 16 B with zero fill (padding + tail rows) straight into the K-major
//                          no-swizzle UMMA operand layout, one (tap, <=64 channel) stage at a time;
//                          weights come pre-packed in the same layout
//     MMA (1 thread)       tcgen05.mma kind::f16 into one TMEM accumulator, tcgen05.commit frees stages
//     epilogue (4 warps)   tcgen05.ld -> +bias, x channel scale, +residual -> fp32 NCDHW
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <math.h>
#include <stdlib.h>

namespace dcl {

using namespace tc;

cudaError_t trace_set_conv_gemm(long long* p) { return trace_set_local(p); }

// ---------------------------------------------------------------------------------------------
// prep: norm + act + bf16 + channel blocking (zero-pads channels up to a multiple of 16)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_blocked_kernel(ConvSrc src, int64_t spatial, int in_h, int in_w, uint4* __restrict__ out) {
  const int kc = blockIdx.y;
  const int cin = src.c0 + src.c1;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= spatial) return;
  // x0 may be a strided view (s0c/s0d/s0h); x1 is dense
  const int w = (int)(p % in_w);
  const int64_t t = p / in_w;
  const int h = (int)(t % in_h);
  const int64_t d = t / in_h;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = kc * 8 + k;
    float x = 0.f;
    if (c < cin) {
      if (c < src.c0)
        x = __ldg(src.x0 + (int64_t)c * src.s0c + d * src.s0d + (int64_t)h * src.s0h + w);
      else
        x = __ldg(src.x1 + (int64_t)(c - src.c0) * spatial + p);
      if (src.sums != nullptr) {
        float mu, rs;
        stat_mean_rstd(src.sums, c, src.inv_n, &mu, &rs);
        x = (x - mu) * rs;
      } else if (src.mean != nullptr) {
        x = (x - __ldg(src.mean + c)) * __ldg(src.rstd + c);
      }
      x = apply_act(x, src.act);
    }
    v[k] = x;
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]);
  o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]);
  o.w = pack_bf16x2(v[6], v[7]);
  out[(int64_t)kc * spatial + p] = o;
}

int launch_prep_blocked(const ConvSrc& src, int in_d, int in_h, int in_w, void* out, cudaStream_t st) {
  const int cin_pad = (src.c0 + src.c1 + 15) / 16 * 16;
  const int64_t spatial = (int64_t)in_d * in_h * in_w;
  dim3 grid((unsigned)((spatial + 255) / 256), cin_pad / 8);
  prep_blocked_kernel<<<grid, 256, 0, st>>>(src, spatial, in_h, in_w, reinterpret_cast<uint4*>(out));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// GEMM kernel
// ---------------------------------------------------------------------------------------------
struct GemmConvParams {
  const uint4* a;          // B-format input, channels [0, c0_chunks*8)
  const uint4* a1;         // optional second source for the remaining channels (concat never materialised)
  int c0_chunks;
  const uint4* w;          // packed weights [taps][cin_pad/8][cout_pad][8 bf16]
  const float* bias;       // cout or nullptr
  const float* out_scale;  // cout or nullptr
  const void* residual;    // same format as y, or nullptr
  void* y;
  stat_t* stats;           // 2*cout fixed-point sums += per-channel (sum, sum of squares) of the outputs, or nullptr
  int cin_pad, cout, cout_pad, n_tile;
  int D, H, W, OD, OH, OW, stride, taps, kstage;
  int out_mode;            // 0: fp32 NCDHW [cout][m]   1: fp32 row-major [m][cout]   2: B-format bf16
  int gelu;                // exact (erf) GELU after the bias
};

constexpr int G_EPI_WARPS = 4;
constexpr int G_PROD_WARPS = 4;
constexpr int G_THREADS = (G_EPI_WARPS + 1 + G_PROD_WARPS) * 32;
constexpr int G_PROD_T0 = (G_EPI_WARPS + 1) * 32;
constexpr int G_NPROD = G_PROD_WARPS * 32;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Epilogue of one 128-row accumulator tile (TMEM columns [0, n_tile) at lane_addr): bias, channel scale,
// GELU, residual, store in the requested format, per-channel statistics into s_stat[warp][2][n_tile].
__device__ __forceinline__ void gemm_epilogue_tile(const GemmConvParams& p, uint32_t lane_addr, int64_t m, int64_t m_total,
                                                   int n0, float* s_stat, int warp, int lane) {
  const bool row_ok = m < m_total;
  float* yf = reinterpret_cast<float*>(p.y);
  const float* rf = reinterpret_cast<const float*>(p.residual);
  for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
    uint32_t acc[16];
    tmem_ld16(lane_addr + (uint32_t)c0, acc);
    tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int co = n0 + c0 + k;
      float val = 0.f;
      if (row_ok && co < p.cout) {
        val = __uint_as_float(acc[k]) + (p.bias ? __ldg(p.bias + co) : 0.f);
        if (p.out_scale) val *= __ldg(p.out_scale + co);
        if (p.gelu) val = 0.5f * val * (1.f + erff(val * 0.70710678118654752440f));
      }
      v[k] = val;
    }
    if (row_ok) {
      if (p.out_mode == 1) {          // 16 consecutive outputs of one row: 4 x 16-byte stores
        const int64_t off = m * p.cout + n0 + c0;
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
          if (rf) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(rf + off + k));
            v[k] += r.x; v[k + 1] += r.y; v[k + 2] += r.z; v[k + 3] += r.w;
          }
          *reinterpret_cast<float4*>(yf + off + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
        }
      } else if (p.out_mode == 2) {   // B-format: two 16-byte vectors (8 channels each) per voxel
        const uint4* rb = reinterpret_cast<const uint4*>(p.residual);
        uint4* yb = reinterpret_cast<uint4*>(p.y);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int kc = (n0 + c0) / 8 + hf;
          if (kc * 8 < p.cout_pad) {
            if (rb) {
              const uint4 rv = __ldg(rb + (int64_t)kc * m_total + m);
              const uint32_t* pr = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                v[8 * hf + 2 * k] += __uint_as_float(pr[k] << 16);
                v[8 * hf + 2 * k + 1] += __uint_as_float(pr[k] & 0xffff0000u);
              }
            }
            uint4 o;
            o.x = pack_bf16x2(v[8 * hf], v[8 * hf + 1]);
            o.y = pack_bf16x2(v[8 * hf + 2], v[8 * hf + 3]);
            o.z = pack_bf16x2(v[8 * hf + 4], v[8 * hf + 5]);
            o.w = pack_bf16x2(v[8 * hf + 6], v[8 * hf + 7]);
            yb[(int64_t)kc * m_total + m] = o;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int co = n0 + c0 + k;
          if (co < p.cout) {
            const int64_t off = (int64_t)co * m_total + m;
            if (rf) v[k] += __ldg(rf + off);
            yf[off] = v[k];
          }
        }
      }
    }
    if (p.stats != nullptr) {
      // per-channel sums over the 32 rows of this warp: transposed butterfly (16 shuffles per quantity),
      // lane l ends up with channel ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1)
      float q[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) { if (!row_ok) v[k] = 0.f; q[k] = v[k] * v[k]; }
#pragma unroll
      for (int width = 8, bit = 16; width >= 1; width >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int k = 0; k < width; ++k) {
          const float sv = up ? v[k] : v[k + width], kv = up ? v[k + width] : v[k];
          const float sq = up ? q[k] : q[k + width], kq = up ? q[k + width] : q[k];
          v[k] = kv + __shfl_xor_sync(0xffffffffu, sv, bit);
          q[k] = kq + __shfl_xor_sync(0xffffffffu, sq, bit);
        }
      }
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
      q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
      if ((lane & 1) == 0) {
        const int ch = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        s_stat[(warp * 2) * p.n_tile + ch] += v[0];       // slot owned by this lane: no race
        s_stat[(warp * 2 + 1) * p.n_tile + ch] += q[0];
      }
    }
  }
}

// wait for the cp.async groups of the last K+1 stages one by one (oldest first) and publish them
template <int K>
__device__ __forceinline__ void drain_stages(uint64_t* bar_full, int total, int ns) {
  if (K < total) {
    cp_async_wait<K>();
    fence_proxy_async();
    mbar_arrive(&bar_full[(total - 1 - K) % ns]);
  }
  if constexpr (K > 0) drain_stages<K - 1>(bar_full, total, ns);
}

// G_NS shared-memory stages; producers keep LAG = G_NS - 2 cp.async groups in flight
template <int G_NS>
__global__ void __launch_bounds__(G_THREADS, 2)
conv_gemm_kernel(GemmConvParams p) {
  constexpr int LAG = G_NS - 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const int a_bytes = 2 * p.kstage * 2048;               // [chunk][128 rows][16 B]
  const int b_bytes = 2 * p.kstage * p.n_tile * 16;      // [chunk][n_tile rows][16 B]
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + G_NS * stage_bytes);
  uint64_t* bar_empty = bar_full + G_NS;
  uint64_t* bar_acc = bar_empty + G_NS;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc + 1);
  float* s_stat = reinterpret_cast<float*>(s_tmem + 2);     // [4 warps][2][n_tile] partial sums

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) trace_event(0, 0);     // kernel entry
  const int64_t m_total = (int64_t)p.OD * p.OH * p.OW;
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  const int n0 = blockIdx.y * p.n_tile;
  const int cpt = p.cin_pad / (16 * p.kstage);           // stages per tap
  const int total = p.taps * cpt;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.n_tile) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < G_NS; ++s) { mbar_init(&bar_full[s], G_NPROD); mbar_init(&bar_empty[s], 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == G_EPI_WARPS) tmem_alloc(s_tmem, tmem_cols);
  for (int i = tid; i < 8 * p.n_tile; i += G_THREADS) s_stat[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (tid == 0) trace_event(1, 0);   // setup done

  if (warp >= G_EPI_WARPS + 1) {
    // =============================== producers ===================================================
    const int pt = tid - G_PROD_T0;                       // 0..127 = A row of this thread
    const int64_t m = m0 + pt;
    const bool row_ok = m < m_total;
    int ow = 0, oh = 0, od = 0;
    if (row_ok) {
      ow = (int)(m % p.OW);
      const int64_t t = m / p.OW;
      oh = (int)(t % p.OH);
      od = (int)(t / p.OH);
    }
    const int64_t sp_in = (int64_t)p.D * p.H * p.W;
    const int pad = p.taps == 27 ? 1 : 0;
    const uint32_t smem_base = smem_u32(smem);
    const int n_chunks = 2 * p.kstage;
    // chunk kc of the (virtually concatenated) input lives at a + kc*sp_in (+ delta1 once past source 0)
    const int64_t delta1 = p.a1 ? (p.a1 - p.a) - (int64_t)p.c0_chunks * sp_in : 0;
    const uint32_t b_seg = (uint32_t)p.n_tile * 16;        // bytes of one chunk row block of B
    int it = 0;
    for (int tap = 0; tap < p.taps; ++tap) {
      int kd = 0, kh = 0, kw = 0;
      if (p.taps == 27) { kd = tap / 9; kh = (tap - kd * 9) / 3; kw = tap - kd * 9 - kh * 3; }
      const int id = od * p.stride + kd - pad, ih = oh * p.stride + kh - pad, iw = ow * p.stride + kw - pad;
      const bool ok = row_ok && (unsigned)id < (unsigned)p.D && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
      const uint4* a_row = p.a + (ok ? ((int64_t)id * p.H + ih) * p.W + iw : 0);
      const uint32_t a_bytes_ok = ok ? 16u : 0u;
      const uint4* b_tap = p.w + (int64_t)tap * (p.cin_pad / 8) * p.cout_pad + n0;
      for (int kg = 0; kg < cpt; ++kg, ++it) {
        const int s = it % G_NS;
        mbar_wait(&bar_empty[s], ((uint32_t)(it / G_NS) & 1u) ^ 1u);
        const int kc0 = kg * n_chunks;
        const uint32_t st_base = smem_base + (uint32_t)(s * stage_bytes);
        if (pt == 0) {
          // weights: one bulk async copy (UBLKCP) per 8-channel chunk, completion counted in bytes on full[s]
          const uint32_t bar = smem_u32(&bar_full[s]);
          asm volatile("mbarrier.expect_tx.shared::cta.b64 [%1], %0;" ::"r"((uint32_t)n_chunks * b_seg), "r"(bar) : "memory");
          const uint4* b_src = b_tap + (int64_t)kc0 * p.cout_pad;
          for (int c = 0; c < n_chunks; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             st_base + (uint32_t)a_bytes + (uint32_t)c * b_seg),
                         "l"(b_src + (int64_t)c * p.cout_pad), "r"(b_seg), "r"(bar)
                         : "memory");
        }
        // activations: this thread's row, one 16-byte zero-filling cp.async per chunk
        const uint32_t a_dst = st_base + (uint32_t)(pt * 16);
        const int64_t step = ok ? sp_in : 0, d1 = ok ? delta1 : 0;   // masked rows never form an out-of-range address
        const uint4* a_src = a_row + (int64_t)kc0 * step;
        for (int c = 0; c < n_chunks; ++c) {
          cp_async16(a_dst + (uint32_t)(c * 2048), a_src + (kc0 + c >= p.c0_chunks ? d1 : 0), a_bytes_ok);
          a_src += step;
        }
        cp_async_commit();
        if (pt == 0) trace_event(2, it);   // stage issued
        if (it >= LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async();
          mbar_arrive(&bar_full[(it - LAG) % G_NS]);
        }
      }
    }
    drain_stages<LAG - 1>(bar_full, total, G_NS);
  } else if (warp == G_EPI_WARPS) {
    // =============================== MMA issuer ==================================================
    {   // all 32 lanes run the loop; the elected lane issues
      const uint32_t idesc = umma_idesc_bf16(128, p.n_tile);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t b_lbo = (uint32_t)p.n_tile * 16;
      const uint64_t a_desc0 = umma_desc(smem_base, 2048, 128);
      const uint64_t b_desc0 = umma_desc(smem_base + (uint32_t)a_bytes, b_lbo, 128);
      for (int it = 0; it < total; ++it) {
        const int s = it % G_NS;
        mbar_wait(&bar_full[s], (uint32_t)(it / G_NS) & 1u);
        tc_fence_after();
        if (lane == 0) trace_event(3, it);              // stage landed, MMAs issued next
        uint64_t ad = a_desc0 + (uint64_t)((uint32_t)(s * stage_bytes) >> 4);
        uint64_t bd = b_desc0 + (uint64_t)((uint32_t)(s * stage_bytes) >> 4);
        for (int ks = 0; ks < p.kstage; ++ks) {
          umma_bf16_ws(tmem_base, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
          ad += 256u;                       // 2 chunks x 2048 B
          bd += (uint64_t)(b_lbo >> 3);     // 2 chunks x n_tile*16 B, in 16-byte units
        }
        umma_commit_ws(&bar_empty[s]);
      }
      umma_commit_ws(bar_acc);
    }
    __syncwarp();
  } else {
    // =============================== epilogue ====================================================
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    if (tid == 0) trace_event(4, 0);   // accumulator complete
    gemm_epilogue_tile(p, tmem_base + ((uint32_t)(warp * 32) << 16), m0 + warp * 32 + lane, m_total, n0, s_stat, warp, lane);
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");   // the 4 epilogue warps only
      for (int c = tid; c < p.n_tile; c += G_EPI_WARPS * 32) {
        if (n0 + c < p.cout) {   // fixed summation order over the 4 epilogue warps
          const float a = (s_stat[c] + s_stat[2 * p.n_tile + c]) + (s_stat[4 * p.n_tile + c] + s_stat[6 * p.n_tile + c]);
          const float q = (s_stat[p.n_tile + c] + s_stat[3 * p.n_tile + c]) + (s_stat[5 * p.n_tile + c] + s_stat[7 * p.n_tile + c]);
          stat_add(p.stats, n0 + c, a, q);
        }
      }
    }
  }

  if (tid == 0) trace_event(5, 0);     // epilogue done
  tc_fence_before();
  __syncthreads();
  if (warp == G_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// "Slab" kernel: stride-1 3x3x3 convolution whose input tile is staged ONCE.
//   CTA        MT consecutive 128-voxel output tiles of one plane (R = MT*128/W full-width rows) x n_tile
//              output channels.  The input slab = 3 planes x (R+2) rows x W voxels x Cin is fetched once with
//              zero-filling cp.async (B-format is already the UMMA operand layout), InstanceNorm + activation
//              applied in place; the 27 taps are 27 start addresses into it.  Rows have NO halo columns: a
//              kw = 0 / 2 tap simply reads the neighbouring position and the output lanes whose neighbour would
//              wrap to another row (w = 0 / w = W-1) are switched off with the MMA's disable-output-lane mask,
//              which is exactly zero padding.  Versus the im2col GEMM above this cuts L2->SM traffic for the
//              activations 27x; weights stream through a small ring with bulk async copies (one tap per stage).
//   warps      0-3 epilogue, 4 MMA issuer, 5-12 slab producers, 13 weight streamer.
// ---------------------------------------------------------------------------------------------
struct SlabParams {
  GemmConvParams g;         // a / a1 / c0_chunks / w / bias / residual / y / stats / cout / n_tile / D,H,W / out_mode
  const stat_t* sums;       // fused input InstanceNorm (+ activation), as in the rolling kernel
  float inv_n;
  const float* mean;
  const float* rstd;
  int act;
  int mt;                   // output tiles per CTA
  int rows;                 // R = mt * 128 / W
  int kc_pass;              // 8-channel chunks resident per pass (Cin is processed in npass passes)
  int npass;
  int nb;                   // weight ring depth (taps)
  uint32_t mk0[4], mk2[4];  // output lanes switched off for kw = 0 (w == 0) and kw = 2 (w == W-1)
  uint32_t idesc;
};

constexpr int S_PROD_WARPS = 8;
constexpr int S_THREADS = (G_EPI_WARPS + 1 + S_PROD_WARPS + 1) * 32;
constexpr int S_WEIGHT_WARP = G_EPI_WARPS + 1 + S_PROD_WARPS;
// taps are processed with kw in the order 1,0,2: the first MMA of a tile must write every lane (accumulate = 0),
// so it has to be an unmasked (kw = 1) tap
__device__ __forceinline__ int slab_tap(int j, int* kd, int* kh, int* kw) {
  const int kdh = j / 3, i = j - kdh * 3;
  *kd = kdh / 3; *kh = kdh - *kd * 3; *kw = i == 0 ? 1 : (i == 1 ? 0 : 2);
  return kdh * 3 + *kw;
}
constexpr int S_PROD_T0 = (G_EPI_WARPS + 1) * 32;
constexpr int S_NPROD = S_PROD_WARPS * 32;

__device__ __forceinline__ void umma_bf16_masked(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
      : "memory");
}

__global__ void __launch_bounds__(S_THREADS, 1)
conv_slab_kernel(SlabParams sp) {
  const GemmConvParams& p = sp.g;
  extern __shared__ __align__(128) uint8_t smem[];
  const int W = p.W, H = p.H, D = p.D;
  const int R = sp.rows;
  const int npos = 3 * (R + 2) * W + 2;                  // +1 pad position in front and behind
  const int slab_bytes = sp.kc_pass * npos * 16;
  const int b_tap_bytes = sp.kc_pass * p.n_tile * 16;    // one tap, all resident channels
  const int b_stage = 3 * b_tap_bytes;                   // ring stage = the 3 kw taps of one (kd,kh)
  uint64_t* bar_bfull = reinterpret_cast<uint64_t*>(smem + slab_bytes + sp.nb * b_stage);
  uint64_t* bar_bempty = bar_bfull + sp.nb;
  uint64_t* bar_slab_full = bar_bempty + sp.nb;
  uint64_t* bar_slab_empty = bar_slab_full + 1;
  uint64_t* bar_slab_land = bar_slab_empty + 1;          // bulk copies of the slab have landed (byte count)
  uint64_t* bar_acc = bar_slab_land + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc + 1);
  float* s_stat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_tmem + 2) + 15) & ~(uintptr_t)15);  // [4 warps][2][n_tile]
  float* s_scale = s_stat + 8 * p.n_tile;                // [cin_pad]  rstd
  float* s_shift = s_scale + p.cin_pad;                  // [cin_pad]  -mean * rstd

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) trace_event(0, 0);
  const int64_t m_total = (int64_t)D * H * W;
  const int64_t m0 = (int64_t)blockIdx.x * (sp.mt * 128);
  const int n0 = blockIdx.y * p.n_tile;
  const int d_out = (int)(m0 / ((int64_t)H * W));
  const int h0 = (int)((m0 - (int64_t)d_out * H * W) / W);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < sp.mt * p.n_tile) tmem_cols <<= 1;
  const bool has_norm = sp.sums != nullptr || sp.mean != nullptr;

  if (tid == 0) {
    for (int s = 0; s < sp.nb; ++s) { mbar_init(&bar_bfull[s], 1); mbar_init(&bar_bempty[s], 1); }
    mbar_init(bar_slab_full, S_NPROD);
    mbar_init(bar_slab_empty, 1);
    mbar_init(bar_slab_land, 1);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == G_EPI_WARPS) tmem_alloc(s_tmem, tmem_cols);
  for (int i = tid; i < 8 * p.n_tile; i += S_THREADS) s_stat[i] = 0.f;
  if (has_norm)
    for (int c = tid; c < p.cin_pad; c += S_THREADS) {
      float m = 0.f, r = 1.f;
      if (sp.sums != nullptr) stat_mean_rstd(sp.sums, c, sp.inv_n, &m, &r);
      else { m = sp.mean[c]; r = sp.rstd[c]; }
      s_scale[c] = r;
      s_shift[c] = -m * r;
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t smem_base = smem_u32(smem);
  const int kcs_total = p.cin_pad / 8;
  if (tid == 0) trace_event(1, 0);

  if (warp == S_WEIGHT_WARP) {
    // =============================== weight streamer =============================================
    // one ring stage = one tap (all resident channels); lane c issues the bulk async copy of chunk c
    {
      const uint32_t b_seg = (uint32_t)p.n_tile * 16;
      int bit = 0;
      for (int pass = 0; pass < sp.npass; ++pass) {
        const int kc_base = pass * sp.kc_pass;
        for (int kdh = 0; kdh < 9; ++kdh, ++bit) {
          const int s = bit % sp.nb;
          const uint32_t bar = smem_u32(&bar_bfull[s]);
          if (lane == 0) {
            mbar_wait(&bar_bempty[s], ((uint32_t)(bit / sp.nb) & 1u) ^ 1u);
            trace_event(10, bit);   // weight stage free, issuing 3 taps
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"((uint32_t)(3 * sp.kc_pass) * b_seg), "r"(bar)
                         : "memory");
          }
          __syncwarp();
          const uint32_t b_dst = smem_base + (uint32_t)(slab_bytes + s * b_stage);
          for (int e = lane; e < 3 * sp.kc_pass; e += 32) {
            const int i = e / sp.kc_pass, c = e - i * sp.kc_pass;
            const int kw = i == 0 ? 1 : (i == 1 ? 0 : 2);            // kw order 1,0,2 (see slab_tap)
            const uint4* b_src = p.w + ((int64_t)(kdh * 3 + kw) * kcs_total + kc_base + c) * p.cout_pad + n0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             b_dst + (uint32_t)(i * b_tap_bytes) + (uint32_t)c * b_seg),
                         "l"(b_src), "r"(b_seg), "r"(bar)
                         : "memory");
          }
        }
      }
    }
  } else if (warp >= G_EPI_WARPS + 1) {
    // =============================== slab producers ==============================================
    const int pt = tid - S_PROD_T0;
    const int64_t sp_in = m_total;
    const int64_t delta1 = p.a1 ? (p.a1 - p.a) - (int64_t)p.c0_chunks * sp_in : 0;
    const bool identity = !has_norm && sp.act == ACT_NONE;
    const int items_per_chunk = 3 * (R + 2) * W;
    for (int pass = 0; pass < sp.npass; ++pass) {
      const int kc_base = pass * sp.kc_pass;
      if (pass > 0) mbar_wait(bar_slab_empty, (uint32_t)(pass - 1) & 1u);   // MMAs of the previous pass are done
      // ---- the slab: every (chunk, plane, row) is W contiguous 16-byte vectors in global memory AND in the
      // slab, so it is one bulk async copy; rows outside the volume are zero-filled by the thread instead
      const int n_rows = 3 * (R + 2);
      const uint32_t row_bytes = (uint32_t)W * 16u;
      if (pt == 0) {
        int vd = 0, vh = 0;
        for (int pl = 0; pl < 3; ++pl) vd += (unsigned)(d_out - 1 + pl) < (unsigned)D ? 1 : 0;
        for (int r = 0; r < R + 2; ++r) vh += (unsigned)(h0 - 1 + r) < (unsigned)H ? 1 : 0;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"((uint32_t)(vd * vh * sp.kc_pass) * row_bytes),
                     "r"(smem_u32(bar_slab_land))
                     : "memory");
      }
      asm volatile("bar.sync 2, %0;" ::"n"(S_NPROD) : "memory");   // expect_tx is registered before any copy lands
      for (int e = pt; e < sp.kc_pass * n_rows; e += S_NPROD) {
        const int c = e / n_rows;
        const int rr = e - c * n_rows;
        const int pl = rr / (R + 2);
        const int r = rr - pl * (R + 2);
        const int d_in = d_out - 1 + pl, h_in = h0 - 1 + r;
        const int kc = kc_base + c;
        const uint32_t dst = smem_base + (uint32_t)((c * npos + 1 + rr * W) * 16);
        if ((unsigned)d_in < (unsigned)D && (unsigned)h_in < (unsigned)H) {
          const uint4* src = p.a + (int64_t)kc * sp_in + (kc >= p.c0_chunks ? delta1 : 0) + ((int64_t)d_in * H + h_in) * W;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                       "l"(src), "r"(row_bytes), "r"(smem_u32(bar_slab_land))
                       : "memory");
        } else {
          uint4* z = reinterpret_cast<uint4*>(smem + (size_t)(c * npos + 1 + rr * W) * 16);
          for (int i = 0; i < W; ++i) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      if (pt < 2 * sp.kc_pass)   // the two pad positions of each chunk (only read by masked-off lanes; keep them finite)
        *reinterpret_cast<uint4*>(smem + (size_t)((pt >> 1) * npos + ((pt & 1) ? npos - 1 : 0)) * 16) = make_uint4(0u, 0u, 0u, 0u);
      if (pt == 0) trace_event(11, pass);   // slab copies issued
      mbar_wait(bar_slab_land, (uint32_t)pass & 1u);
      if (pt == 0) trace_event(12, pass);   // slab landed
      if (!identity) {
        // in-place InstanceNorm + activation.  A warp walks whole rows (all index math once per row, lanes on
        // consecutive 16-byte vectors = conflict free); W = 16 packs two rows per warp pass.
        const int lanes_per_row = W < 32 ? W : 32;
        const int rows_per_iter = 32 / lanes_per_row;
        const int sub = lane / lanes_per_row, li = lane - sub * lanes_per_row;
        const int total_rows = sp.kc_pass * n_rows;
        constexpr int UR = 4;                               // independent rows in flight per lane (latency hiding)
        const int pw = warp - (G_EPI_WARPS + 1);
        for (int row0 = pw * rows_per_iter * UR; row0 < total_rows; row0 += S_PROD_WARPS * rows_per_iter * UR) {
          uint4* qp[UR];
          uint4 v[UR];
          int kcs[UR];
          bool ok[UR];
#pragma unroll
          for (int u = 0; u < UR; ++u) {
            const int row = row0 + u * rows_per_iter + sub;
            const int c = row / n_rows;
            const int rr = row - c * n_rows;
            const int pl = rr / (R + 2);
            const int r = rr - pl * (R + 2);
            const int d_in = d_out - 1 + pl, h_in = h0 - 1 + r;
            ok[u] = row < total_rows && li < W && (unsigned)d_in < (unsigned)D && (unsigned)h_in < (unsigned)H;
            kcs[u] = kc_base + c;
            qp[u] = reinterpret_cast<uint4*>(smem + (size_t)(c * npos + 1 + rr * W + li) * 16);
            if (ok[u]) v[u] = *qp[u];
          }
          for (int i = 0; i < W; i += 32) {                 // W = 64 / 128: further 32-vector segments of the rows
#pragma unroll
            for (int u = 0; u < UR; ++u) {
              if (!ok[u]) continue;
              if (i > 0) v[u] = qp[u][i];
              const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + kcs[u] * 8), sc1 = *reinterpret_cast<const float4*>(s_scale + kcs[u] * 8 + 4);
              const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + kcs[u] * 8), sh1 = *reinterpret_cast<const float4*>(s_shift + kcs[u] * 8 + 4);
              const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
              const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
              uint32_t* pv = reinterpret_cast<uint32_t*>(&v[u]);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float fx = fmaf(__uint_as_float(pv[k] << 16), sc[2 * k], sh[2 * k]);
                float fy = fmaf(__uint_as_float(pv[k] & 0xffff0000u), sc[2 * k + 1], sh[2 * k + 1]);
                if (sp.act == ACT_RELU) { fx = fmaxf(fx, 0.f); fy = fmaxf(fy, 0.f); }
                else if (sp.act == ACT_LRELU) { fx = fmaxf(fx, 0.01f * fx); fy = fmaxf(fy, 0.01f * fy); }
                pv[k] = pack_bf16x2(fx, fy);
              }
              qp[u][i] = v[u];
            }
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_slab_full);
      if (pt == 0) trace_event(13, pass);   // slab transformed + published
    }
  } else if (warp == G_EPI_WARPS) {
    // =============================== MMA issuer ==================================================
    {   // all 32 lanes run the loop; the elected lane issues (see umma_bf16_ws)
      const uint32_t idesc = sp.idesc;
      const uint32_t b_lbo = (uint32_t)p.n_tile * 16;
      const uint32_t a_lbo = (uint32_t)npos * 16;
      const int rows_per_tile = 128 / W;
      // descriptors = base + (byte offset >> 4): the single issuing thread only adds small integers per MMA
      const uint64_t a_desc0 = umma_desc(smem_base, a_lbo, 128);
      const uint64_t b_desc0 = umma_desc(smem_base + (uint32_t)slab_bytes, b_lbo, 128);
      const uint32_t a_ks = 2u * (uint32_t)npos, b_ks = 2u * (uint32_t)p.n_tile;    // K-step strides in 16-byte units
      const uint32_t a_tile = (uint32_t)(rows_per_tile * W);
      const int nks = sp.kc_pass / 2;
      int bit = 0;
      for (int pass = 0; pass < sp.npass; ++pass) {
        mbar_wait(bar_slab_full, (uint32_t)pass & 1u);
        tc_fence_after();
        for (int kdh = 0; kdh < 9; ++kdh, ++bit) {
          const int kd = kdh / 3, kh = kdh - kd * 3;
          const uint32_t pos_row = (uint32_t)(1 + (kd * (R + 2) + kh) * W);
          const int s = bit % sp.nb;
          mbar_wait(&bar_bfull[s], (uint32_t)(bit / sp.nb) & 1u);
          tc_fence_after();
          if (lane == 0) trace_event(14, bit);   // weights of this (kd,kh) landed, issuing MMAs
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int kw = i == 0 ? 1 : (i == 1 ? 0 : 2);             // same order as the weight streamer
            const uint64_t b_tap = b_desc0 + (uint64_t)((uint32_t)(s * b_stage + i * b_tap_bytes) >> 4);
            const uint64_t a_tap = a_desc0 + (uint64_t)(pos_row + (uint32_t)(kw - 1));
            const uint32_t q0 = kw == 0 ? sp.mk0[0] : sp.mk2[0], q1 = kw == 0 ? sp.mk0[1] : sp.mk2[1];
            const uint32_t q2 = kw == 0 ? sp.mk0[2] : sp.mk2[2], q3 = kw == 0 ? sp.mk0[3] : sp.mk2[3];
            const uint32_t accum = (pass | kdh | i) != 0 ? 1u : 0u;
            for (int t = 0; t < sp.mt; ++t) {
              const uint32_t d_tmem = tmem_base + (uint32_t)(t * p.n_tile);
              uint64_t ad = a_tap + (uint64_t)((uint32_t)t * a_tile), bd = b_tap;
              uint32_t acc_t = accum;
              for (int ks = 0; ks < nks; ++ks) {
                if (kw == 1) umma_bf16_ws(d_tmem, ad, bd, idesc, acc_t);
                else umma_bf16_masked_ws(d_tmem, ad, bd, idesc, acc_t, q0, q1, q2, q3);
                ad += a_ks; bd += b_ks; acc_t = 1u;
              }
            }
          }
          umma_commit_ws(&bar_bempty[s]);
        }
        umma_commit_ws(bar_slab_empty);
      }
      umma_commit_ws(bar_acc);
    }
    __syncwarp();
  } else {
    // =============================== epilogue ====================================================
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    if (tid == 0) trace_event(15, 0);     // accumulators complete
    for (int t = 0; t < sp.mt; ++t)
      gemm_epilogue_tile(p, tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * p.n_tile),
                         m0 + t * 128 + warp * 32 + lane, m_total, n0, s_stat, warp, lane);
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
      for (int c = tid; c < p.n_tile; c += G_EPI_WARPS * 32) {
        if (n0 + c < p.cout) {
          const float a = (s_stat[c] + s_stat[2 * p.n_tile + c]) + (s_stat[4 * p.n_tile + c] + s_stat[6 * p.n_tile + c]);
          const float q = (s_stat[p.n_tile + c] + s_stat[3 * p.n_tile + c]) + (s_stat[5 * p.n_tile + c] + s_stat[7 * p.n_tile + c]);
          stat_add(p.stats, n0 + c, a, q);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == G_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// n_tile: a multiple of 16 that divides cout_pad, <= 256, chosen so the grid has enough CTAs
static int pick_n_tile(int cout_pad, int64_t m_tiles) {
  int best = 16;
  for (int n = 16; n <= 256 && n <= cout_pad; n += 16) {
    if (cout_pad % n != 0) continue;
    const int64_t ctas = m_tiles * (cout_pad / n);
    if (ctas >= 120 || n == 16) best = n;     // the largest tile that still fills the machine
    else break;
  }
  return best;
}

static int launch_gemm_params(GemmConvParams& p, cudaStream_t st);

int launch_gemm_conv(const GemmArgs& g, const TcWeights& w, cudaStream_t st) {
  GemmConvParams p;
  const int cin_pad = (w.cin + 15) / 16 * 16;
  if (w.dev == nullptr || g.a0 == nullptr || (g.c0 % 8) != 0 || (g.a1 == nullptr && g.c0 < w.cin) ||
      (g.a1 != nullptr && g.c0 >= w.cin)) {
    set_error("gemm_conv: bad source description");
    return -1;
  }
  p.a = reinterpret_cast<const uint4*>(g.a0);
  p.a1 = reinterpret_cast<const uint4*>(g.a1);
  p.c0_chunks = g.a1 ? g.c0 / 8 : cin_pad / 8;
  p.w = reinterpret_cast<const uint4*>(w.dev);
  p.bias = g.bias; p.out_scale = g.out_scale; p.residual = g.residual; p.y = g.y; p.stats = g.stats;
  p.cin_pad = cin_pad;
  p.cout = w.cout;
  p.cout_pad = (w.cout + 15) / 16 * 16;
  p.D = g.D; p.H = g.H; p.W = g.W;
  p.stride = g.stride; p.taps = g.taps;
  p.out_mode = g.out_mode; p.gelu = g.gelu;
  p.OD = (g.D - 1) / g.stride + 1; p.OH = (g.H - 1) / g.stride + 1; p.OW = (g.W - 1) / g.stride + 1;
  if (p.out_mode == 1 && (p.cout % 16) != 0) { set_error("gemm_conv: row-major output needs cout % 16 == 0"); return -1; }
  return launch_gemm_params(p, st);
}

// ---- slab kernel launcher ---------------------------------------------------------------------
bool slab_conv_supported(int cin, int cout, int d, int h, int w, int stride, int taps) {
  (void)cout;
  if (stride != 1 || taps != 27) return false;
  if (!(w == 16 || w == 32 || w == 64 || w == 128)) return false;
  if (((int64_t)h * w) % 128 != 0 || d < 1) return false;
  const int cin_pad = (cin + 15) / 16 * 16;
  return cin_pad <= 256;
}

int launch_slab_conv(const GemmArgs& g, const BNorm* norm, const TcWeights& w, cudaStream_t st) {
  if (!slab_conv_supported(w.cin, w.cout, g.D, g.H, g.W, g.stride, g.taps) || w.dev == nullptr || g.a0 == nullptr) {
    set_error("slab_conv: unsupported shape");
    return -1;
  }
  SlabParams sp;
  GemmConvParams& p = sp.g;
  const int cin_pad = (w.cin + 15) / 16 * 16;
  p.a = reinterpret_cast<const uint4*>(g.a0);
  p.a1 = reinterpret_cast<const uint4*>(g.a1);
  p.c0_chunks = g.a1 ? g.c0 / 8 : cin_pad / 8;
  p.w = reinterpret_cast<const uint4*>(w.dev);
  p.bias = g.bias; p.out_scale = g.out_scale; p.residual = g.residual; p.y = g.y; p.stats = g.stats;
  p.cin_pad = cin_pad; p.cout = w.cout; p.cout_pad = (w.cout + 15) / 16 * 16;
  p.D = g.D; p.H = g.H; p.W = g.W; p.OD = g.D; p.OH = g.H; p.OW = g.W;
  p.stride = 1; p.taps = 27; p.kstage = 1;
  p.out_mode = g.out_mode; p.gelu = g.gelu;
  sp.sums = norm ? norm->sums : nullptr;
  sp.inv_n = norm ? norm->inv_n : 0.f;
  sp.mean = norm ? norm->mean : nullptr;
  sp.rstd = norm ? norm->rstd : nullptr;
  sp.act = norm ? norm->act : ACT_NONE;
  const int64_t m_total = (int64_t)g.D * g.H * g.W;
  const int64_t plane = (int64_t)g.H * g.W;
  const int smem_cap = 227 * 1024;
  // configuration search: tiles per CTA (mt), output channels per CTA (nt), input-channel passes; cost model =
  // waves x (fixed + max(MMA issue time, L2->SM fill time)) in SM clocks
  int best_mt = 0, best_nt = 0, best_pass = 0, best_nb = 0;
  double best_cost = 1e30;
  for (int mt = 1; mt <= 4; mt <<= 1) {
    if (plane % (mt * 128) != 0) continue;
    const int64_t m_ctas = m_total / (mt * 128);
    const int rows = mt * 128 / g.W;
    const int npos = 3 * (rows + 2) * g.W + 2;
    for (int nt = 16; nt <= 256 && nt <= p.cout_pad; nt += 16) {
      if (p.cout_pad % nt != 0 || mt * nt > 512) continue;
      for (int npass = 1; npass <= 2; ++npass) {
        if ((cin_pad / 16) % npass != 0) continue;
        const int kc_pass = cin_pad / 8 / npass;
        const int slab = kc_pass * npos * 16;
        const int b_stage = 3 * kc_pass * nt * 16;
        const int fixed = (2 * 8 + 4) * 8 + 16 + 8 * nt * 4 + 2 * cin_pad * 4 + 64;
        int nb = (smem_cap - slab - fixed) / b_stage;
        if (nb > 6) nb = 6;
        if (nb < 2) continue;
        const double ctas = (double)m_ctas * (p.cout_pad / nt);
        const double waves = ceil(ctas / 148.0);
        const double per_mma = nt <= 64 ? 60.0 : nt / 2.0 + 12.0;
        const double mma = (double)npass * 27 * mt * (kc_pass / 2) * per_mma;
        const double fill = ((double)slab * npass + 27.0 * cin_pad * nt * 2) / 28.0;
        const double cost = waves * (8000.0 + (mma > fill ? mma : fill) + (nb < 3 ? 4000.0 : 0.0) + 1500.0 * (npass - 1));
        if (cost < best_cost) { best_cost = cost; best_mt = mt; best_nt = nt; best_pass = npass; best_nb = nb; }
      }
    }
  }
  if (best_mt == 0) { set_error("slab_conv: tile does not fit shared memory"); return -1; }
  sp.mt = best_mt; p.n_tile = best_nt; sp.npass = best_pass; sp.nb = best_nb;
  sp.idesc = umma_idesc_bf16(128, best_nt);
  for (int j = 0; j < 4; ++j) {
    uint32_t a = 0, b = 0;
    for (int i = 0; i < 32; ++i) {
      const int wpos = (32 * j + i) % g.W;
      if (wpos == 0) a |= 1u << i;
      if (wpos == g.W - 1) b |= 1u << i;
    }
    sp.mk0[j] = a; sp.mk2[j] = b;
  }
  sp.rows = best_mt * 128 / g.W;
  sp.kc_pass = cin_pad / 8 / best_pass;
  const int npos = 3 * (sp.rows + 2) * g.W + 2;
  const int smem_bytes = sp.kc_pass * npos * 16 + sp.nb * 3 * sp.kc_pass * p.n_tile * 16 + (2 * sp.nb + 4) * 8 + 16 +
                         8 * p.n_tile * 4 + 2 * cin_pad * 4 + 64;
  static bool configured = false;
  if (!configured) {
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap));
    configured = true;
  }
  dim3 grid((unsigned)(m_total / (sp.mt * 128)), p.cout_pad / p.n_tile);
  conv_slab_kernel<<<grid, S_THREADS, smem_bytes, st>>>(sp);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// fp32-NCDHW-out convenience wrapper used by the mixed pipeline (prep kernel output as the only source)
int launch_conv_gemm(const void* a_blocked, const TcWeights& w, const ConvDst& dst, int in_d, int in_h, int in_w,
                     int stride, int taps, cudaStream_t st) {
  GemmArgs g;
  g.a0 = a_blocked; g.c0 = (w.cin + 15) / 16 * 16;
  g.D = in_d; g.H = in_h; g.W = in_w; g.stride = stride; g.taps = taps;
  g.bias = dst.bias; g.out_scale = dst.out_scale; g.residual = dst.residual; g.y = dst.y; g.stats = dst.stats;
  g.out_mode = 0;
  return launch_gemm_conv(g, w, st);
}

static int launch_gemm_params(GemmConvParams& p, cudaStream_t st) {
  const int k16 = p.cin_pad / 16;
  p.kstage = k16 % 4 == 0 ? 4 : (k16 % 3 == 0 ? 3 : (k16 % 2 == 0 ? 2 : 1));
  const int64_t m_total = (int64_t)p.OD * p.OH * p.OW;
  const int64_t m_tiles = (m_total + 127) / 128;
  p.n_tile = pick_n_tile(p.cout_pad, m_tiles);
  const int stage_bytes = 2 * p.kstage * 2048 + 2 * p.kstage * p.n_tile * 16;
  // deep enough to cover the L2 round trip with small stages; short K loops (linears) take all stages at once
  const int total_stages = p.taps * (p.cin_pad / (16 * p.kstage));
  int ns = stage_bytes <= 12 * 1024 ? 8 : stage_bytes <= 16 * 1024 ? 6 : stage_bytes <= 26 * 1024 ? 4 : 3;
  if (total_stages <= 8 && 8 * stage_bytes <= 150 * 1024) ns = 8;
  const int smem_bytes = ns * stage_bytes + (2 * ns + 1) * 8 + 16 + 8 * p.n_tile * 4;
  if (smem_bytes > 160 * 1024) { set_error("conv_gemm: stage does not fit shared memory"); return -1; }
  static bool configured = false;
  if (!configured) {
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)m_tiles, p.cout_pad / p.n_tile);
  if (ns == 8) conv_gemm_kernel<8><<<grid, G_THREADS, smem_bytes, st>>>(p);
  else if (ns == 6) conv_gemm_kernel<6><<<grid, G_THREADS, smem_bytes, st>>>(p);
  else if (ns == 4) conv_gemm_kernel<4><<<grid, G_THREADS, smem_bytes, st>>>(p);
  else conv_gemm_kernel<3><<<grid, G_THREADS, smem_bytes, st>>>(p);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Linear layers of the couplers (nn.Linear in SelfAttention.py:62-66, ResidualNorm.py:38-44) as the same
// GEMM: rows = tokens.  prep_rows fuses the preceding nn.LayerNorm(512) (ResidualNorm.py:14-32).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 int rows, uint4* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * TOKEN_DIM) + lane * 4;   // 16 elements / lane
  float v[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = __ldg(xr + j);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
  if (gamma != nullptr) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / TOKEN_DIM);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) { const float a = v[k] - mean; q += a * a; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / TOKEN_DIM) + 1e-5f);
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (v[k] - mean) * rstd * __ldg(gamma + lane * 16 + k) + __ldg(beta + lane * 16 + k);
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint4 o;
    o.x = pack_bf16x2(v[8 * c], v[8 * c + 1]);
    o.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
    o.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
    o.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
    out[(int64_t)(lane * 2 + c) * rows + row] = o;
  }
}

int launch_prep_rows(const float* x, const float* gamma, const float* beta, int rows, void* out, cudaStream_t st) {
  prep_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, gamma, beta, rows, reinterpret_cast<uint4*>(out));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_linear_tc(const void* a_blocked, const void* w_packed, const float* bias, const float* residual, float* y,
                     int m, int n, int k, bool gelu, cudaStream_t st) {
  if (n % 16 != 0 || k % 16 != 0) { set_error("linear_tc: n and k must be multiples of 16"); return -1; }
  TcWeights w;
  w.dev = const_cast<void*>(w_packed); w.cin = k; w.cout = n;
  GemmArgs g;
  g.a0 = a_blocked; g.c0 = k;
  g.D = 1; g.H = 1; g.W = m; g.stride = 1; g.taps = 1;
  g.bias = bias; g.residual = residual; g.y = y;
  g.out_mode = 1; g.gelu = gelu ? 1 : 0;
  return launch_gemm_conv(g, w, st);
}

}  // namespace dcl
