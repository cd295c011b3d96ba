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

cudaError_t trace_set_conv_gemm(long long* p, int cta) { return trace_set_local(p, cta); }

// ---------------------------------------------------------------------------------------------
// prep: norm + act + bf16 + channel blocking (zero-pads channels up to a multiple of 16)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_blocked_kernel(ConvSrc src, int64_t spatial, int in_h, int in_w, uint4* __restrict__ out, int64_t lo_off) {
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
  if (lo_off != 0) {      // split-fp16: hi planes, then lo planes
    uint4 l;
    split8(v, o, l);
    out[lo_off + (int64_t)kc * spatial + p] = l;
  } else {
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
  }
  out[(int64_t)kc * spatial + p] = o;
}

int launch_prep_blocked(const ConvSrc& src, int in_d, int in_h, int in_w, void* out, cudaStream_t st, bool x3) {
  const int cin_pad = (src.c0 + src.c1 + 15) / 16 * 16;
  const int64_t spatial = (int64_t)in_d * in_h * in_w;
  dim3 grid((unsigned)((spatial + 255) / 256), cin_pad / 8);
  prep_blocked_kernel<<<grid, 256, 0, st>>>(src, spatial, in_h, in_w, reinterpret_cast<uint4*>(out),
                                            x3 ? (int64_t)(cin_pad / 8) * spatial : 0);
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
  // split-fp16 (DCL_F16X3): the lo planes of a source follow its hi planes (chunk kc of source 0 at + c0_chunks *
  // spatial, of source 1 at + c1 chunks * spatial), the lo weight image sits w_lo 16-byte units behind the hi image,
  // B-format outputs / residuals carry their lo planes cout_pad / 8 chunks behind the hi planes
  int x3;
  int64_t w_lo;
  float acc_mul;           // accumulators are multiplied by this (split mode: 2^-k of the power-of-two weight scale; else 1)
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
// STACK: the accumulator holds the three kw partial products P_0 | P_1 | P_2 (n_tile columns each, see the slab
// kernel); out[m] = P_0[m-1] + P_1[m] + P_2[m+1] without the terms that cross a row end (rows are `roww` voxels,
// a tile is whole rows).  Neighbouring lanes come from shuffles, neighbouring warps through `xchg`.
template <bool STACK = false>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmConvParams& p, uint32_t lane_addr, int64_t m, int64_t m_total,
                                                   int n0, float* s_stat, const float* s_bo, int warp, int lane,
                                                   int roww = 0, float* xchg = nullptr, int grp0 = 0, int c0_begin = 0,
                                                   int c0_step = 16, int stat_slot = -1, int bar_id = 1) {
  // warp = TMEM lane quarter (0..3); stat_slot = row pair of s_stat this warp owns (default: warp)
  if (stat_slot < 0) stat_slot = warp;
  const bool row_ok = m < m_total;
  float* yf = reinterpret_cast<float*>(p.y);
  const float* rf = reinterpret_cast<const float*>(p.residual);
  int grp = grp0;   // running 16-column group index: the exchange buffer alternates with it
  for (int c0 = c0_begin; c0 < p.n_tile; c0 += c0_step, ++grp) {
    uint32_t acc[16];
    if (!STACK) {
      tmem_ld16(lane_addr + (uint32_t)c0, acc);
      tmem_ld_wait();
    } else {
      uint32_t a0[16], a2[16];
      tmem_ld16(lane_addr + (uint32_t)c0, a0);
      tmem_ld16(lane_addr + (uint32_t)(p.n_tile + c0), acc);
      tmem_ld16(lane_addr + (uint32_t)(2 * p.n_tile + c0), a2);
      tmem_ld_wait();
      const int wpos = (warp * 32 + lane) % roww;
      const bool wide = roww > 32;                       // a row spans several warps: exchange the edge lanes
      float* xs = xchg + (grp & 1) * (G_EPI_WARPS * 32);
      if (wide) {
        if (lane == 31) {
#pragma unroll
          for (int k = 0; k < 16; k += 4)
            *reinterpret_cast<float4*>(xs + (warp * 2) * 16 + k) =
                make_float4(__uint_as_float(a0[k]), __uint_as_float(a0[k + 1]), __uint_as_float(a0[k + 2]), __uint_as_float(a0[k + 3]));
        }
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 16; k += 4)
            *reinterpret_cast<float4*>(xs + (warp * 2 + 1) * 16 + k) =
                make_float4(__uint_as_float(a2[k]), __uint_as_float(a2[k + 1]), __uint_as_float(a2[k + 2]), __uint_as_float(a2[k + 3]));
        }
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(G_EPI_WARPS * 32) : "memory");
      }
      const bool has_l = wpos > 0, has_r = wpos < roww - 1;
      // branch-free edge lanes: every lane reads the (valid) neighbour-warp slots as broadcasts and selects
      float el[16], er[16];
      if (wide) {
        const float* xl = xs + ((warp > 0 ? warp - 1 : 0) * 2) * 16;
        const float* xr = xs + ((warp < G_EPI_WARPS - 1 ? warp + 1 : warp) * 2 + 1) * 16;
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(xl + k), r4 = *reinterpret_cast<const float4*>(xr + k);
          el[k] = l4.x; el[k + 1] = l4.y; el[k + 2] = l4.z; el[k + 3] = l4.w;
          er[k] = r4.x; er[k + 1] = r4.y; er[k + 2] = r4.z; er[k + 3] = r4.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) { el[k] = 0.f; er[k] = 0.f; }   // rows end at warp edges: the edge terms are masked
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float up = __shfl_up_sync(0xffffffffu, __uint_as_float(a0[k]), 1);
        float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(a2[k]), 1);
        up = lane == 0 ? el[k] : up;
        dn = lane == 31 ? er[k] : dn;
        acc[k] = __float_as_uint((__uint_as_float(acc[k]) + (has_l ? up : 0.f)) + (has_r ? dn : 0.f));
      }
    }
    // bias / channel scale of these 16 columns come from shared memory (staged once per CTA, zero / one beyond cout):
    // vector loads up front instead of one dependent global load per element
    float bs[16], os[16];
#pragma unroll
    for (int k = 0; k < 16; k += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(s_bo + c0 + k);
      const float4 o4 = *reinterpret_cast<const float4*>(s_bo + p.n_tile + c0 + k);
      bs[k] = b4.x; bs[k + 1] = b4.y; bs[k + 2] = b4.z; bs[k + 3] = b4.w;
      os[k] = o4.x; os[k + 1] = o4.y; os[k + 2] = o4.z; os[k + 3] = o4.w;
    }
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float val = (__uint_as_float(acc[k]) * p.acc_mul + bs[k]) * os[k];
      v[k] = (row_ok && n0 + c0 + k < p.cout) ? val : 0.f;
    }
    if (p.gelu) {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = 0.5f * v[k] * (1.f + erff(v[k] * 0.70710678118654752440f));   // gelu(0) = 0
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
            const int64_t lo_off = p.x3 ? (int64_t)(p.cout_pad / 8) * m_total : 0;
            if (rb) {
#pragma unroll
              for (int part = 0; part < 2; ++part) {
                if (part == 1 && !p.x3) break;
                const uint4 rv = __ldg(rb + (part ? lo_off : 0) + (int64_t)kc * m_total + m);
                const uint32_t* pr = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (p.x3) {       // split mode: fp16 halves
                    const float2 t = unpack_f16x2(pr[k]);
                    v[8 * hf + 2 * k] += t.x;
                    v[8 * hf + 2 * k + 1] += t.y;
                  } else {
                    v[8 * hf + 2 * k] += __uint_as_float(pr[k] << 16);
                    v[8 * hf + 2 * k + 1] += __uint_as_float(pr[k] & 0xffff0000u);
                  }
                }
              }
            }
            uint4 o;
            if (p.x3) {
              uint4 l;
              split_x2(v[8 * hf], v[8 * hf + 1], o.x, l.x);
              split_x2(v[8 * hf + 2], v[8 * hf + 3], o.y, l.y);
              split_x2(v[8 * hf + 4], v[8 * hf + 5], o.z, l.z);
              split_x2(v[8 * hf + 6], v[8 * hf + 7], o.w, l.w);
              yb[lo_off + (int64_t)kc * m_total + m] = l;
            } else {
              o.x = pack_bf16x2(v[8 * hf], v[8 * hf + 1]);
              o.y = pack_bf16x2(v[8 * hf + 2], v[8 * hf + 3]);
              o.z = pack_bf16x2(v[8 * hf + 4], v[8 * hf + 5]);
              o.w = pack_bf16x2(v[8 * hf + 6], v[8 * hf + 7]);
            }
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
        s_stat[(stat_slot * 2) * p.n_tile + ch] += v[0];       // slot owned by this lane: no race
        s_stat[(stat_slot * 2 + 1) * p.n_tile + ch] += q[0];
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
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  constexpr int LAG = G_NS - 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const int a_half = 2 * p.kstage * 2048;                // [chunk][128 rows][16 B]
  const int b_half = 2 * p.kstage * p.n_tile * 16;       // [chunk][n_tile rows][16 B]
  const int a_bytes = a_half << p.x3;                    // split-fp16: the lo chunks follow the hi chunks
  const int b_bytes = b_half << p.x3;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + G_NS * stage_bytes);
  uint64_t* bar_empty = bar_full + G_NS;
  uint64_t* bar_acc = bar_empty + G_NS;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc + 1);
  float* s_stat = reinterpret_cast<float*>(s_tmem + 2);     // [4 warps][2][n_tile] partial sums
  float* s_bo = s_stat + 8 * p.n_tile;                      // [n_tile] bias, [n_tile] channel scale

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform role index
  long long* const tbuf = trace_begin();
  if (tid == 0) trace_event(tbuf, 0, 0);     // kernel entry
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
  for (int i = tid; i < p.n_tile; i += G_THREADS) {
    const bool in = n0 + i < p.cout;
    s_bo[i] = in && p.bias ? __ldg(p.bias + n0 + i) : 0.f;
    s_bo[p.n_tile + i] = in && p.out_scale ? __ldg(p.out_scale + n0 + i) : 1.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);
  if (tid == 0) trace_event(tbuf, 1, 0);   // setup done

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
    const int64_t lo0 = (int64_t)p.c0_chunks * sp_in, lo1 = (int64_t)(p.cin_pad / 8 - p.c0_chunks) * sp_in;   // x3 only
    const uint32_t b_seg = (uint32_t)p.n_tile * 16;        // bytes of one chunk row block of B
    int it = 0;
    if (p.taps == 1 && p.stride == 1 && p.a1 == nullptr) {
      // pointwise conv / linear layer: the 128 rows of a chunk are 2 KB of contiguous global memory, so a stage is
      // 2 x n_chunks bulk async copies issued by one thread (rows past m_total of a tail tile stay unwritten: GEMM
      // rows are independent and those rows are never stored)
      const uint32_t a_chunk = (uint32_t)(m_total - m0 < 128 ? m_total - m0 : 128) * 16u;
      for (int kg = 0; kg < cpt; ++kg, ++it) {
        const int s = it % G_NS;
        mbar_wait(&bar_empty[s], ((uint32_t)(it / G_NS) & 1u) ^ 1u);
        if (pt == 0) {
          const int kc0 = kg * n_chunks;
          const uint32_t st_base = smem_base + (uint32_t)(s * stage_bytes);
          const uint32_t bar = smem_u32(&bar_full[s]);
          asm volatile("mbarrier.expect_tx.shared::cta.b64 [%1], %0;" ::"r"(((uint32_t)n_chunks * (b_seg + a_chunk)) << p.x3), "r"(bar) : "memory");
          const uint4* b_src = p.w + (int64_t)kc0 * p.cout_pad + n0;
          const uint4* a_src = p.a + (int64_t)kc0 * sp_in + m0;
          for (int part = 0; part <= p.x3; ++part) {
            const uint4* bs = b_src + (part ? p.w_lo : 0);
            const uint4* as = a_src + (part ? lo0 : 0);
            for (int c = 0; c < n_chunks; ++c) {
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                               st_base + (uint32_t)a_bytes + (uint32_t)(part * b_half) + (uint32_t)c * b_seg),
                           "l"(bs + (int64_t)c * p.cout_pad), "r"(b_seg), "r"(bar)
                           : "memory");
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                               st_base + (uint32_t)(part * a_half) + (uint32_t)(c * 2048)),
                           "l"(as + (int64_t)c * sp_in), "r"(a_chunk), "r"(bar)
                           : "memory");
            }
          }
        }
        mbar_arrive(&bar_full[s]);
      }
    } else {
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
          asm volatile("mbarrier.expect_tx.shared::cta.b64 [%1], %0;" ::"r"(((uint32_t)n_chunks * b_seg) << p.x3), "r"(bar) : "memory");
          const uint4* b_src = b_tap + (int64_t)kc0 * p.cout_pad;
          for (int part = 0; part <= p.x3; ++part)
            for (int c = 0; c < n_chunks; ++c)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                               st_base + (uint32_t)a_bytes + (uint32_t)(part * b_half) + (uint32_t)c * b_seg),
                           "l"(b_src + (part ? p.w_lo : 0) + (int64_t)c * p.cout_pad), "r"(b_seg), "r"(bar)
                           : "memory");
        }
        // activations: this thread's row, one 16-byte zero-filling cp.async per chunk
        const uint32_t a_dst = st_base + (uint32_t)(pt * 16);
        const int64_t step = ok ? sp_in : 0, d1 = ok ? delta1 : 0;   // masked rows never form an out-of-range address
        const uint4* a_src = a_row + (int64_t)kc0 * step;
        for (int c = 0; c < n_chunks; ++c) {
          const bool second = kc0 + c >= p.c0_chunks;
          const uint4* srcp = a_src + (second ? d1 : 0);
          cp_async16(a_dst + (uint32_t)(c * 2048), srcp, a_bytes_ok);
          if (p.x3) cp_async16(a_dst + (uint32_t)(a_half + c * 2048), srcp + (ok ? (second ? lo1 : lo0) : 0), a_bytes_ok);
          a_src += step;
        }
        cp_async_commit();
        if (pt == 0) trace_event(tbuf, 2, it);   // stage issued
        if (it >= LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async();
          mbar_arrive(&bar_full[(it - LAG) % G_NS]);
        }
      }
    }
    drain_stages<LAG - 1>(bar_full, total, G_NS);
    }
  } else if (warp == G_EPI_WARPS) {
    // =============================== MMA issuer ==================================================
    {   // all 32 lanes run the loop; the elected lane issues
      const uint32_t idesc = umma_idesc_16(128, p.n_tile, p.x3 != 0);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t b_lbo = (uint32_t)p.n_tile * 16;
      const uint64_t a_desc0 = umma_desc(smem_base, 2048, 128);
      const uint64_t b_desc0 = umma_desc(smem_base + (uint32_t)a_bytes, b_lbo, 128);
      for (int it = 0; it < total; ++it) {
        const int s = it % G_NS;
        mbar_wait(&bar_full[s], (uint32_t)(it / G_NS) & 1u);
        tc_fence_after();
        if (lane == 0) trace_event(tbuf, 3, it);              // stage landed, MMAs issued next
        uint64_t ad = a_desc0 + (uint64_t)((uint32_t)(s * stage_bytes) >> 4);
        uint64_t bd = b_desc0 + (uint64_t)((uint32_t)(s * stage_bytes) >> 4);
        for (int ks = 0; ks < p.kstage; ++ks) {
          umma_bf16_ws(tmem_base, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
          if (p.x3) {                       // + a_lo * w_hi + a_hi * w_lo
            umma_bf16_ws(tmem_base, ad + (uint64_t)((uint32_t)a_half >> 4), bd, idesc, 1u);
            umma_bf16_ws(tmem_base, ad, bd + (uint64_t)((uint32_t)b_half >> 4), idesc, 1u);
          }
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
    if (tid == 0) trace_event(tbuf, 4, 0);   // accumulator complete
    gemm_epilogue_tile(p, tmem_base + ((uint32_t)(warp * 32) << 16), m0 + warp * 32 + lane, m_total, n0, s_stat, s_bo, warp, lane);
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

  if (tid == 0) trace_event(tbuf, 5, 0);     // epilogue done
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
//              output channels.  The input slab = 3 planes x (R+2) rows x W voxels x Cin is fetched with one bulk
//              async copy per (8-channel chunk, plane) - B-format rows are contiguous and already the UMMA operand
//              layout - and InstanceNorm + activation are applied in place by 12 warps.
//   kw taps    are NOT shifted operand reads (a 16-byte shift breaks the 128-byte alignment of the core matrices
//              and costs ~80 instead of ~48 clk per MMA, tools/ubench).  Instead the three kw weight matrices are
//              stacked along N: one MMA with N = 3*n_tile reads the aligned A tile once and fills three
//              accumulators P_kw[m] = sum_{kd,kh,cin} X[d+kd-1][h+kh-1][w(m)] W[kd][kh][kw]; the epilogue forms
//              out[m] = P_0[m-1] + P_1[m] + P_2[m+1] with warp shuffles (and a 2-float-per-channel exchange
//              between neighbouring warps when a row spans more than one warp), dropping the terms that would
//              cross a row end - which is exactly the zero padding along w.
//   weights    repacked once per layer as [n tile][kd,kh][cin/8][kw][n_tile][8]: a ring stage (one (kd,kh),
//              all resident channels, 3 kw) is ONE bulk async copy.
//   warps      0-3 epilogue (+ transform), 4 MMA issuer, 5-12 transform, 13 loader (slab + weight ring).
// ---------------------------------------------------------------------------------------------
struct SlabParams {
  GemmConvParams g;         // a / a1 / c0_chunks / bias / residual / y / stats / cout / n_tile / D,H,W / out_mode
  const uint4* wslab;       // repacked weights
  int64_t wslab_lo;         // split-fp16: 16-byte units from the hi image to the lo image
  const stat_t* sums;       // fused input InstanceNorm (+ activation), as in the rolling kernel
  float inv_n;
  const float* mean;
  const float* rstd;
  int act;
  int mt;                   // output tiles per CTA
  int rows;                 // R = mt * 128 / W
  int kc_pass;              // 8-channel chunks resident per pass (Cin is processed in npass passes)
  int npass;
  int nb;                   // weight ring depth (stages)
  uint32_t idesc;           // M = 128, N = 3 * n_tile
};

constexpr int S_XF_WARPS = 12;                       // warps 0-3 and 5-12 transform the slab
constexpr int S_LOAD_WARP = G_EPI_WARPS + 1 + 8;     // 13
constexpr int S_THREADS = (S_LOAD_WARP + 1) * 32;

template <int ACT>
__device__ __forceinline__ uint4 slab_xf(uint4 v, const float (&sc)[8], const float (&sh)[8]) {
  uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float fx = fmaf(__uint_as_float(pv[k] << 16), sc[2 * k], sh[2 * k]);
    float fy = fmaf(__uint_as_float(pv[k] & 0xffff0000u), sc[2 * k + 1], sh[2 * k + 1]);
    if (ACT == ACT_RELU) { fx = fmaxf(fx, 0.f); fy = fmaxf(fy, 0.f); }
    else if (ACT == ACT_LRELU) { fx = fmaxf(fx, 0.01f * fx); fy = fmaxf(fy, 0.01f * fy); }
    pv[k] = pack_bf16x2(fx, fy);
  }
  return v;
}

// split-fp16 variant: value = hi + lo -> norm + act -> split again
template <int ACT>
__device__ __forceinline__ void slab_xf_run_x3(uint4* bh, uint4* bl, int n_vec, int lane, const float (&sc)[8], const float (&sh)[8]) {
  for (int i = lane; i < n_vec; i += 32) {
    float f[8];
    unpack8_x3<false>(bh[i], f);
    unpack8_x3<true>(bl[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = fmaf(f[k], sc[k], sh[k]);
      if (ACT == ACT_RELU) v = fmaxf(v, 0.f);
      else if (ACT == ACT_LRELU) v = fmaxf(v, 0.01f * v);
      f[k] = v;
    }
    uint4 h, l;
    split8(f, h, l);
    bh[i] = h;
    bl[i] = l;
  }
}

template <int ACT>
__device__ __forceinline__ void slab_xf_run(uint4* base, int n_vec, int lane, const float (&sc)[8], const float (&sh)[8]) {
  int i = lane;
  for (; i + 96 < n_vec; i += 128) {     // 4 independent vectors in flight per lane
    uint4 v0 = base[i], v1 = base[i + 32], v2 = base[i + 64], v3 = base[i + 96];
    base[i] = slab_xf<ACT>(v0, sc, sh);
    base[i + 32] = slab_xf<ACT>(v1, sc, sh);
    base[i + 64] = slab_xf<ACT>(v2, sc, sh);
    base[i + 96] = slab_xf<ACT>(v3, sc, sh);
  }
  for (; i < n_vec; i += 32) base[i] = slab_xf<ACT>(base[i], sc, sh);
}

__global__ void __launch_bounds__(S_THREADS, 1)
conv_slab_kernel(SlabParams sp) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  const GemmConvParams& p = sp.g;
  extern __shared__ __align__(128) uint8_t smem[];
  const int W = p.W, H = p.H, D = p.D;
  const int R = sp.rows;
  const int npos = 3 * (R + 2) * W;                      // 16-byte positions per channel chunk
  const int kcp = sp.kc_pass;                            // channel chunks per pass
  const int kcs = kcp << p.x3;                           // staged chunks per pass: split-fp16 keeps the lo chunks behind the hi chunks
  const int slab_bytes = kcs * npos * 16;
  const int nst = 3 * p.n_tile;                          // stacked N
  const int b_stage = kcs * nst * 16;                    // ring stage = one (kd,kh): [hi|lo][chunk][kw][n_tile] x 16 B
  uint64_t* bar_bfull = reinterpret_cast<uint64_t*>(smem + slab_bytes + sp.nb * b_stage);
  uint64_t* bar_bempty = bar_bfull + sp.nb;
  uint64_t* bar_slab_full = bar_bempty + sp.nb;          // slab transformed (one arrive per transform warp)
  uint64_t* bar_slab_empty = bar_slab_full + 1;          // MMAs of the pass are done with the slab
  uint64_t* bar_slab_land = bar_slab_empty + 1;          // bulk copies of the slab have landed (byte count)
  uint64_t* bar_acc = bar_slab_land + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc + 1);
  float* s_stat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_tmem + 2) + 15) & ~(uintptr_t)15);  // [4 warps][2][n_tile]
  float* s_scale = s_stat + 16 * p.n_tile;               // [cin_pad]  rstd      (s_stat: [2 groups][4 warps][2][n_tile])
  float* s_shift = s_scale + p.cin_pad;                  // [cin_pad]  -mean * rstd
  float* s_xchg = s_shift + p.cin_pad;                   // [2 groups][2 buffers][4 warps][2][16]
  float* s_bo = s_xchg + 4 * G_EPI_WARPS * 32;           // [n_tile] bias, [n_tile] channel scale

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform role index
  long long* const tbuf = trace_begin();
  if (tid == 0) trace_event(tbuf, 0, 0);
  const int64_t m_total = (int64_t)D * H * W;
  const int64_t m0 = (int64_t)blockIdx.x * (sp.mt * 128);
  const int n0 = blockIdx.y * p.n_tile;
  const int d_out = (int)(m0 / ((int64_t)H * W));
  const int h0 = (int)((m0 - (int64_t)d_out * H * W) / W);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < sp.mt * nst) tmem_cols <<= 1;
  const bool has_norm = sp.sums != nullptr || sp.mean != nullptr;
  // staged rows r = 0..R+1 hold input rows h0-1+r; [r_lo, r_hi] are inside the volume
  const int r_lo = h0 == 0 ? 1 : 0;
  const int r_hi = h0 + R >= H ? R : R + 1;

  if (tid == 0) {
    for (int s = 0; s < sp.nb; ++s) { mbar_init(&bar_bfull[s], 1); mbar_init(&bar_bempty[s], 1); }
    mbar_init(bar_slab_full, S_XF_WARPS);
    mbar_init(bar_slab_empty, 1);
    mbar_init(bar_slab_land, 1);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == G_EPI_WARPS) tmem_alloc(s_tmem, tmem_cols);
  for (int i = tid; i < 16 * p.n_tile; i += S_THREADS) s_stat[i] = 0.f;
  for (int i = tid; i < p.n_tile; i += S_THREADS) {
    const bool in = n0 + i < p.cout;
    s_bo[i] = in && p.bias ? __ldg(p.bias + n0 + i) : 0.f;
    s_bo[p.n_tile + i] = in && p.out_scale ? __ldg(p.out_scale + n0 + i) : 1.f;
  }
  for (int c = tid; c < p.cin_pad; c += S_THREADS) {
    float m = 0.f, r = 1.f;
    if (sp.sums != nullptr) stat_mean_rstd(sp.sums, c, sp.inv_n, &m, &r);
    else if (sp.mean != nullptr) { m = sp.mean[c]; r = sp.rstd[c]; }
    s_scale[c] = r;
    s_shift[c] = -m * r;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);
  const uint32_t smem_base = smem_u32(smem);
  const int kcs_total = p.cin_pad / 8;
  if (tid == 0) trace_event(tbuf, 1, 0);

  if (warp == S_LOAD_WARP) {
    // =============================== loader: slab + weight ring ==================================
    const int64_t sp_in = m_total;
    const int64_t delta1 = p.a1 ? (p.a1 - p.a) - (int64_t)p.c0_chunks * sp_in : 0;
    const int64_t lo0 = (int64_t)p.c0_chunks * sp_in, lo1 = (int64_t)(kcs_total - p.c0_chunks) * sp_in;   // x3 only
    const uint32_t run_bytes = (uint32_t)((r_hi - r_lo + 1) * W) * 16u;
    int vd = 0;
    for (int pl = 0; pl < 3; ++pl) vd += (unsigned)(d_out - 1 + pl) < (unsigned)D ? 1 : 0;
    const uint4* wsrc = sp.wslab + (int64_t)blockIdx.y * 9 * kcs_total * nst;
    int bit = 0;
    for (int pass = 0; pass < sp.npass; ++pass) {
      const int kc_base = pass * sp.kc_pass;
      if (pass > 0) mbar_wait(bar_slab_empty, (uint32_t)(pass - 1) & 1u);   // MMAs of the previous pass are done
      if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"((uint32_t)(vd * kcs) * run_bytes),
                     "r"(smem_u32(bar_slab_land))
                     : "memory");
      __syncwarp();
      for (int e = lane; e < kcs * 3; e += 32) {
        const int cc = e / 3, pl = e - cc * 3;             // staged chunk (hi chunks, then lo chunks), plane
        const int d_in = d_out - 1 + pl;
        if ((unsigned)d_in >= (unsigned)D) continue;
        const int part = cc >= kcp ? 1 : 0;
        const int kc = kc_base + cc - part * kcp;
        const bool second = kc >= p.c0_chunks;
        const uint4* src = p.a + (int64_t)kc * sp_in + (second ? delta1 : 0) + (part ? (second ? lo1 : lo0) : 0) +
                           ((int64_t)d_in * H + (h0 - 1 + r_lo)) * W;
        const uint32_t dst = smem_base + (uint32_t)((cc * npos + (pl * (R + 2) + r_lo) * W) * 16);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "r"(run_bytes), "r"(smem_u32(bar_slab_land))
                     : "memory");
      }
      if (lane == 0) trace_event(tbuf, 11, pass);   // slab copies issued
      for (int kdh = 0; kdh < 9; ++kdh, ++bit) {
        const int s = bit % sp.nb;
        if (lane == 0) {
          mbar_wait(&bar_bempty[s], ((uint32_t)(bit / sp.nb) & 1u) ^ 1u);
          trace_event(tbuf, 10, bit);   // weight stage free, issuing (kd,kh)
          const uint32_t bar = smem_u32(&bar_bfull[s]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"((uint32_t)b_stage), "r"(bar) : "memory");
          const uint32_t half = (uint32_t)(kcp * nst * 16);
          for (int part = 0; part <= p.x3; ++part)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_base + (uint32_t)(slab_bytes + s * b_stage) + (uint32_t)part * half),
                         "l"(wsrc + (part ? sp.wslab_lo : 0) + ((int64_t)kdh * kcs_total + kc_base) * nst), "r"(half), "r"(bar)
                         : "memory");
        }
        __syncwarp();
      }
    }
  } else if (warp == G_EPI_WARPS) {
    // =============================== MMA issuer ==================================================
    {   // all 32 lanes run the loop; the elected lane issues (see umma_bf16_ws)
      const uint32_t idesc = sp.idesc;
      // descriptors = base + (byte offset >> 4): every A start address is a multiple of W*16 >= 256 bytes
      const uint64_t a_desc0 = umma_desc(smem_base, (uint32_t)npos * 16, 128);
      const uint64_t b_desc0 = umma_desc(smem_base + (uint32_t)slab_bytes, (uint32_t)nst * 16, 128);
      const uint32_t a_ks = 2u * (uint32_t)npos, b_ks = 2u * (uint32_t)nst;    // K-step strides in 16-byte units
      const int nks = sp.kc_pass / 2;
      const uint64_t a_lo = (uint64_t)((uint32_t)kcp * (uint32_t)npos), b_lo = (uint64_t)((uint32_t)kcp * (uint32_t)nst);   // 16-byte units
      int bit = 0;
      for (int pass = 0; pass < sp.npass; ++pass) {
        mbar_wait(bar_slab_full, (uint32_t)pass & 1u);
        tc_fence_after();
        for (int kdh = 0; kdh < 9; ++kdh, ++bit) {
          const int kd = kdh / 3, kh = kdh - kd * 3;
          const uint32_t pos_row = (uint32_t)((kd * (R + 2) + kh) * W);
          const int s = bit % sp.nb;
          mbar_wait(&bar_bfull[s], (uint32_t)(bit / sp.nb) & 1u);
          tc_fence_after();
          if (lane == 0) trace_event(tbuf, 14, bit);   // weights of this (kd,kh) landed, issuing MMAs
          const uint64_t b_st = b_desc0 + (uint64_t)((uint32_t)(s * b_stage) >> 4);
          const uint32_t accum = (pass | kdh) != 0 ? 1u : 0u;
          for (int t = 0; t < sp.mt; ++t) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(t * nst);
            uint64_t ad = a_desc0 + (uint64_t)(pos_row + (uint32_t)t * 128u), bd = b_st;
            uint32_t acc_t = accum;
            for (int ks = 0; ks < nks; ++ks) {
              umma_bf16_ws(d_tmem, ad, bd, idesc, acc_t);
              if (p.x3) {                   // + a_lo * w_hi + a_hi * w_lo
                umma_bf16_ws(d_tmem, ad + a_lo, bd, idesc, 1u);
                umma_bf16_ws(d_tmem, ad, bd + b_lo, idesc, 1u);
              }
              ad += a_ks; bd += b_ks; acc_t = 1u;
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
    // =============================== transform (12 warps), then epilogue (warps 0-3) ==============
    const int xw = warp < G_EPI_WARPS ? warp : warp - 1;    // 0..11
    const bool identity = !has_norm && sp.act == ACT_NONE;
    // rows / planes outside the volume are zero padding: written once, never touched by the bulk copies
    for (int q = xw; q < kcs * 3; q += S_XF_WARPS) {
      const int c = q / 3, pl = q - c * 3;
      uint4* chunk = reinterpret_cast<uint4*>(smem + (size_t)(c * npos + pl * (R + 2) * W) * 16);
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      if ((unsigned)(d_out - 1 + pl) >= (unsigned)D) {
        for (int i = lane; i < (R + 2) * W; i += 32) chunk[i] = z;
      } else {
        if (r_lo == 1) for (int i = lane; i < W; i += 32) chunk[i] = z;
        if (r_hi == R) for (int i = lane; i < W; i += 32) chunk[(R + 1) * W + i] = z;
      }
    }
    for (int pass = 0; pass < sp.npass; ++pass) {
      const int kc_base = pass * sp.kc_pass;
      mbar_wait(bar_slab_land, (uint32_t)pass & 1u);
      if (tid == 0) trace_event(tbuf, 12, pass);   // slab landed
      if (!identity) {
        const int n_vec = (r_hi - r_lo + 1) * W;
        for (int q = xw; q < sp.kc_pass * 3; q += S_XF_WARPS) {
          const int c = q / 3, pl = q - c * 3;
          if ((unsigned)(d_out - 1 + pl) >= (unsigned)D) continue;
          const int kc = kc_base + c;
          const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + kc * 8), sc1 = *reinterpret_cast<const float4*>(s_scale + kc * 8 + 4);
          const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + kc * 8), sh1 = *reinterpret_cast<const float4*>(s_shift + kc * 8 + 4);
          const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
          const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
          uint4* base = reinterpret_cast<uint4*>(smem + (size_t)(c * npos + (pl * (R + 2) + r_lo) * W) * 16);
          if (p.x3) {
            uint4* bl = base + (size_t)kcp * npos;
            if (sp.act == ACT_RELU) slab_xf_run_x3<ACT_RELU>(base, bl, n_vec, lane, sc, sh);
            else if (sp.act == ACT_LRELU) slab_xf_run_x3<ACT_LRELU>(base, bl, n_vec, lane, sc, sh);
            else slab_xf_run_x3<ACT_NONE>(base, bl, n_vec, lane, sc, sh);
          } else if (sp.act == ACT_RELU) slab_xf_run<ACT_RELU>(base, n_vec, lane, sc, sh);
          else if (sp.act == ACT_LRELU) slab_xf_run<ACT_LRELU>(base, n_vec, lane, sc, sh);
          else slab_xf_run<ACT_NONE>(base, n_vec, lane, sc, sh);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_slab_full);
      if (tid == 0) trace_event(tbuf, 13, pass);   // slab transformed + published
    }
    // ---- epilogue: two groups of four warps (0-3 and 8-11; a warp reads the TMEM lane quarter warp % 4) take
    // alternate (tile, 16-column group) items
    if (warp < G_EPI_WARPS || (warp >= 8 && warp < 12)) {
      const int eg = warp >> 3, ew = warp & 3;
      const int ng = p.n_tile / 16;
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      if (tid == 0) trace_event(tbuf, 15, 0);     // accumulators complete
      int done = 0;
      for (int t = 0; t < sp.mt; ++t) {
        const int c0b = ((eg + t * ng) & 1) * 16;
        gemm_epilogue_tile<true>(p, tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(t * nst),
                                 m0 + t * 128 + ew * 32 + lane, m_total, n0, s_stat, s_bo, ew, lane, W,
                                 s_xchg + eg * (2 * G_EPI_WARPS * 32), done, c0b, 32, eg * 4 + ew, 1 + eg);
        done += (p.n_tile - c0b + 31) / 32;
      }
      if (p.stats != nullptr) {
        asm volatile("bar.sync 3, %0;" ::"n"(2 * G_EPI_WARPS * 32) : "memory");   // both groups
        if (warp < G_EPI_WARPS) {
          for (int c = tid; c < p.n_tile; c += G_EPI_WARPS * 32) {
            if (n0 + c < p.cout) {   // fixed summation order over the 8 partials
              float a = 0.f, q = 0.f;
#pragma unroll
              for (int sl = 0; sl < 8; ++sl) { a += s_stat[(2 * sl) * p.n_tile + c]; q += s_stat[(2 * sl + 1) * p.n_tile + c]; }
              stat_add(p.stats, n0 + c, a, q);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tid == 0) trace_event(tbuf, 16, 0);   // all roles done
  if (warp == G_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// n_tile: a multiple of 16 that divides cout_pad, <= 256, chosen so the grid has enough CTAs
static int pick_n_tile(int cout_pad, int64_t m_tiles) {
  if (m_tiles <= 4) {
    // the couplers' linear layers (129 / 258 rows): six of them run concurrently on different lanes, so a launch
    // should be a handful of light CTAs (16-wide tiles gave 64-128 CTAs of 147 KB shared memory each, one per SM,
    // queueing behind each other: ~18 us per linear)
    for (int n = 64; n >= 16; n -= 16)
      if (cout_pad % n == 0) return n;
  }
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
  p.x3 = g.x3 ? 1 : 0;
  p.w_lo = w.lo_off / 16;
  p.acc_mul = g.x3 ? w.out_mul : 1.f;
  if (w.layout != 0) { set_error("gemm_conv: weights packed for the rolling kernel"); return -1; }
  if (g.x3 && (w.lo_off == 0 || ((w.cin % 16) != 0 && g.a1 != nullptr))) { set_error("gemm_conv: split-fp16 needs split weights"); return -1; }
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

// canonical [tap][cin/8][cout_pad] -> slab order [n tile][kd,kh][cin/8][kw][n_tile] (16-byte elements)
__global__ void __launch_bounds__(256)
slab_repack_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int kcs, int cout_pad, int n_tile) {
  const int64_t total = (int64_t)27 * kcs * cout_pad;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int n = (int)(i % cout_pad);
  const int kc = (int)((i / cout_pad) % kcs);
  const int tap = (int)(i / ((int64_t)cout_pad * kcs));
  const int kdh = tap / 3, kw = tap - kdh * 3;
  const int j = n / n_tile, nn = n - j * n_tile;
  dst[((((int64_t)j * 9 + kdh) * kcs + kc) * 3 + kw) * n_tile + nn] = src[i];
}

void tc_free_weights(TcWeights* w) {
  if (w->dev) cudaFree(w->dev);
  if (w->slab_dev) cudaFree(w->slab_dev);
  w->dev = nullptr; w->slab_dev = nullptr; w->slab_ntile = 0;
}

// Tuning runs (tools/slab_sweep.py): DCL_SLAB_DEBUG=1 prints the chosen configuration and re-reads DCL_SLAB_FORCE=mt,nt,npass
// (a pinned configuration) at every launch
static int g_slab_force[3] = {0, 0, 0};
static const bool g_slab_debug = []() { const char* e = getenv("DCL_SLAB_DEBUG"); return e && e[0] == '1'; }();

int launch_slab_conv(const GemmArgs& g, const BNorm* norm, const TcWeights& w, cudaStream_t st) {
  if (!slab_conv_supported(w.cin, w.cout, g.D, g.H, g.W, g.stride, g.taps) || w.dev == nullptr || g.a0 == nullptr) {
    set_error("slab_conv: unsupported shape");
    return -1;
  }
  if (g_slab_debug) {
    g_slab_force[0] = 0;
    if (const char* f = getenv("DCL_SLAB_FORCE")) sscanf(f, "%d,%d,%d", &g_slab_force[0], &g_slab_force[1], &g_slab_force[2]);
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
  p.x3 = g.x3 ? 1 : 0;
  p.w_lo = w.lo_off / 16;
  p.acc_mul = g.x3 ? w.out_mul : 1.f;
  if (w.layout != 0) { set_error("slab_conv: weights packed for the rolling kernel"); return -1; }
  if (g.x3 && w.lo_off == 0) { set_error("slab_conv: split-fp16 needs split weights"); return -1; }
  const int xs = p.x3;
  sp.sums = norm ? norm->sums : nullptr;
  sp.inv_n = norm ? norm->inv_n : 0.f;
  sp.mean = norm ? norm->mean : nullptr;
  sp.rstd = norm ? norm->rstd : nullptr;
  sp.act = norm ? norm->act : ACT_NONE;
  const int64_t m_total = (int64_t)g.D * g.H * g.W;
  const int64_t plane = (int64_t)g.H * g.W;
  const int smem_cap = 227 * 1024;
  // configuration search: tiles per CTA (mt), output channels per CTA (nt; the MMA runs N = 3*nt <= 256),
  // input-channel passes; cost model in SM clocks = waves x (fixed + slab fetch + transform + max(MMA, weight stream))
  int best_mt = 0, best_nt = 0, best_pass = 0, best_nb = 0;
  double best_cost = 1e30;
  for (int mt = 1; mt <= 4; mt <<= 1) {
    if (plane % (mt * 128) != 0) continue;
    const int64_t m_ctas = m_total / (mt * 128);
    const int rows = mt * 128 / g.W;
    const int npos = 3 * (rows + 2) * g.W;
    for (int nt = 16; 3 * nt <= 256 && nt <= p.cout_pad; nt += 16) {
      if (p.cout_pad % nt != 0 || mt * 3 * nt > 512) continue;
      for (int npass = 1; npass <= 4; ++npass) {
        if ((cin_pad / 16) % npass != 0) continue;
        const int kc_pass = cin_pad / 8 / npass;
        const int slab = (kc_pass * npos * 16) << xs;
        const int b_stage = (kc_pass * 3 * nt * 16) << xs;
        const int fixed = (2 * 6 + 4) * 8 + 16 + 18 * nt * 4 + 2 * cin_pad * 4 + 2048 + 64;
        int nb = (smem_cap - slab - fixed) / b_stage;
        if (nb > 6) nb = 6;
        if (nb < 2) continue;
        const double ctas = (double)m_ctas * (p.cout_pad / nt);
        const double waves = ceil(ctas / 148.0);
        const double per_mma = 3 * nt <= 96 ? 50.0 : 3 * nt / 2.0 + 4.0;
        const double mma = (double)npass * 9 * mt * (kc_pass / 2) * per_mma * (xs ? 3.0 : 1.0);
        const double wstream = 27.0 * cin_pad * nt * 2 / 40.0 * (xs ? 2.0 : 1.0);
        const double cost = waves * (6000.0 + (double)npass * (slab / 40.0 + slab / 60.0) + (mma > wstream ? mma : wstream) +
                                     (nb < 3 ? 3000.0 : 0.0) + mt * nt * 6.0);
        if (g_slab_force[0] > 0 && !(mt == g_slab_force[0] && nt == g_slab_force[1] && npass == g_slab_force[2])) continue;
        if (cost < best_cost) { best_cost = cost; best_mt = mt; best_nt = nt; best_pass = npass; best_nb = nb; }
      }
    }
  }
  if (best_mt == 0) { set_error("slab_conv: tile does not fit shared memory"); return -1; }
  sp.mt = best_mt; p.n_tile = best_nt; sp.npass = best_pass; sp.nb = best_nb;
  if (g_slab_debug)
    fprintf(stderr, "slab %d->%d @%dx%dx%d x3=%d: mt %d nt %d npass %d nb %d cost %.0f\n", w.cin, w.cout, g.D, g.H, g.W, xs, best_mt, best_nt,
            best_pass, best_nb, best_cost);
  sp.idesc = umma_idesc_16(128, 3 * best_nt, xs != 0);
  sp.rows = best_mt * 128 / g.W;
  sp.kc_pass = cin_pad / 8 / best_pass;
  const int64_t n16 = (int64_t)27 * (cin_pad / 8) * p.cout_pad;
  if (w.slab_dev == nullptr || w.slab_ntile != best_nt) {   // one-time repack for this tile width (cached in w)
    if (w.slab_dev) { cudaStreamSynchronize(st); cudaFree(w.slab_dev); w.slab_dev = nullptr; }
    DCL_CUDA_OK(cudaMalloc(&w.slab_dev, (size_t)(n16 << xs) * 16));
    for (int part = 0; part <= xs; ++part)     // split-fp16: the lo image is repacked the same way, behind the hi image
      slab_repack_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>(p.w + (part ? p.w_lo : 0),
                                                                     reinterpret_cast<uint4*>(w.slab_dev) + (part ? n16 : 0),
                                                                     cin_pad / 8, p.cout_pad, best_nt);
    DCL_CUDA_OK(cudaGetLastError());
    w.slab_ntile = best_nt;
  }
  sp.wslab = reinterpret_cast<const uint4*>(w.slab_dev);
  sp.wslab_lo = n16;
  const int npos = 3 * (sp.rows + 2) * g.W;
  const int smem_bytes = ((sp.kc_pass * npos * 16 + sp.nb * sp.kc_pass * 3 * p.n_tile * 16) << xs) + (2 * sp.nb + 4) * 8 + 16 + 16 +
                         18 * p.n_tile * 4 + 2 * cin_pad * 4 + 2048 + 64;
  static bool configured = false;
  if (!configured) {
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap));
    configured = true;
  }
  dim3 grid((unsigned)(m_total / (sp.mt * 128)), p.cout_pad / p.n_tile);
  DCL_CUDA_OK(launch_pdl(conv_slab_kernel, dim3(grid), dim3(S_THREADS), (size_t)(smem_bytes), st, sp));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// fp32-NCDHW-out convenience wrapper used by the mixed pipeline (prep kernel output as the only source)
int launch_conv_gemm(const void* a_blocked, const TcWeights& w, const ConvDst& dst, int in_d, int in_h, int in_w,
                     int stride, int taps, cudaStream_t st, bool x3) {
  GemmArgs g;
  g.x3 = x3;
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
  // split-fp16 stages are twice as large: halve the K depth of a stage until three of them fit
  while (p.x3 && p.kstage % 2 == 0 && ((2 * p.kstage * 2048 + 2 * p.kstage * p.n_tile * 16) << 1) > 64 * 1024) p.kstage /= 2;
  const int stage_bytes = (2 * p.kstage * 2048 + 2 * p.kstage * p.n_tile * 16) << p.x3;
  // deep enough to cover the L2 round trip with small stages; short K loops (linears) take all stages at once
  const int total_stages = p.taps * (p.cin_pad / (16 * p.kstage));
  int ns = stage_bytes <= 12 * 1024 ? 8 : stage_bytes <= 16 * 1024 ? 6 : stage_bytes <= 26 * 1024 ? 4 : 3;
  if (total_stages <= 8 && 8 * stage_bytes <= 150 * 1024) ns = 8;
  if (m_tiles <= 4) ns = 4;      // small concurrent GEMMs: keep the footprint at two CTAs per SM (6 and 8 stages measured equal)
  const int smem_bytes = ns * stage_bytes + (2 * ns + 1) * 8 + 16 + 10 * p.n_tile * 4;
  if (smem_bytes > 220 * 1024) { set_error("conv_gemm: stage does not fit shared memory"); return -1; }
  static bool configured = false;
  if (!configured) {
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    DCL_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)m_tiles, p.cout_pad / p.n_tile);
  if (ns == 8) DCL_CUDA_OK(launch_pdl(conv_gemm_kernel<8>, dim3(grid), dim3(G_THREADS), (size_t)(smem_bytes), st, p));
  else if (ns == 6) DCL_CUDA_OK(launch_pdl(conv_gemm_kernel<6>, dim3(grid), dim3(G_THREADS), (size_t)(smem_bytes), st, p));
  else if (ns == 4) DCL_CUDA_OK(launch_pdl(conv_gemm_kernel<4>, dim3(grid), dim3(G_THREADS), (size_t)(smem_bytes), st, p));
  else DCL_CUDA_OK(launch_pdl(conv_gemm_kernel<3>, dim3(grid), dim3(G_THREADS), (size_t)(smem_bytes), st, p));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Linear layers of the couplers (nn.Linear in SelfAttention.py:62-66, ResidualNorm.py:38-44) as the same
// GEMM: rows = tokens.  prep_rows fuses the preceding nn.LayerNorm(512) (ResidualNorm.py:14-32).
// ---------------------------------------------------------------------------------------------
struct PrepRowsSrc { const float* x; const float* gamma; const float* beta; int rows; uint4* out; };
__global__ void __launch_bounds__(256)
prep_rows_kernel(PrepRowsSrc s0, PrepRowsSrc s1, int x3) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  const PrepRowsSrc& sr = blockIdx.y == 0 ? s0 : s1;
  const float* __restrict__ x = sr.x;
  const float* __restrict__ gamma = sr.gamma;
  const float* __restrict__ beta = sr.beta;
  const int rows = sr.rows;
  uint4* __restrict__ out = sr.out;
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
    if (x3) {
      uint4 l;
      split_x2(v[8 * c], v[8 * c + 1], o.x, l.x);
      split_x2(v[8 * c + 2], v[8 * c + 3], o.y, l.y);
      split_x2(v[8 * c + 4], v[8 * c + 5], o.z, l.z);
      split_x2(v[8 * c + 6], v[8 * c + 7], o.w, l.w);
      out[(int64_t)(TOKEN_DIM / 8 + lane * 2 + c) * rows + row] = l;
    } else {
      o.x = pack_bf16x2(v[8 * c], v[8 * c + 1]);
      o.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
      o.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
      o.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
    }
    out[(int64_t)(lane * 2 + c) * rows + row] = o;
  }
}

int launch_prep_rows(const float* x, const float* gamma, const float* beta, int rows, void* out, cudaStream_t st, bool x3) {
  PrepRowsSrc s0{x, gamma, beta, rows, reinterpret_cast<uint4*>(out)};
  DCL_CUDA_OK(launch_pdl(prep_rows_kernel, dim3((rows + 7) / 8), dim3(256), (size_t)(0), st, s0, s0, x3 ? 1 : 0));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// two independent (LayerNorm +) conversions in one launch (the two inputs of a DualSelfAttention block)
int launch_prep_rows2(const float* x0, const float* g0, const float* b0, int rows0, void* out0, const float* x1, const float* g1,
                      const float* b1, int rows1, void* out1, cudaStream_t st, bool x3) {
  PrepRowsSrc s0{x0, g0, b0, rows0, reinterpret_cast<uint4*>(out0)}, s1{x1, g1, b1, rows1, reinterpret_cast<uint4*>(out1)};
  const int rows = rows0 > rows1 ? rows0 : rows1;
  DCL_CUDA_OK(launch_pdl(prep_rows_kernel, dim3((rows + 7) / 8, 2), dim3(256), (size_t)(0), st, s0, s1, x3 ? 1 : 0));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_linear_tc(const void* a_blocked, const void* w_packed, const float* bias, const float* residual, float* y,
                     int m, int n, int k, bool gelu, cudaStream_t st, void* y_blocked, bool x3, float acc_mul) {
  if (n % 16 != 0 || k % 16 != 0) { set_error("linear_tc: n and k must be multiples of 16"); return -1; }
  TcWeights w;
  w.dev = const_cast<void*>(w_packed); w.cin = k; w.cout = n;
  w.lo_off = x3 ? (int64_t)n * k * 2 : 0;
  w.out_mul = acc_mul;
  GemmArgs g;
  g.x3 = x3;
  g.a0 = a_blocked; g.c0 = k;
  g.D = 1; g.H = 1; g.W = m; g.stride = 1; g.taps = 1;
  g.bias = bias; g.residual = residual; g.y = y;
  g.out_mode = 1; g.gelu = gelu ? 1 : 0;
  if (y_blocked != nullptr) {      // bf16 [n/8][m][8]: feeds the next GEMM directly (no residual in this mode)
    if (residual != nullptr) { set_error("linear_tc: blocked output takes no residual"); return -1; }
    g.y = y_blocked; g.out_mode = 2;
  }
  return launch_gemm_conv(g, w, st);
}

}  // namespace dcl
