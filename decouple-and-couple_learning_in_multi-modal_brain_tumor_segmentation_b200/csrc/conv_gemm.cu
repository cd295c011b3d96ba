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

namespace dcl {

using namespace tc;

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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

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
    const int b_items = n_chunks * p.n_tile;
    for (int it = 0; it < total; ++it) {
      const int s = it % G_NS;
      mbar_wait(&bar_empty[s], ((uint32_t)(it / G_NS) & 1u) ^ 1u);
      const int tap = it / cpt;
      const int kc0 = (it - tap * cpt) * n_chunks;
      int kd = 0, kh = 0, kw = 0;
      if (p.taps == 27) { kd = tap / 9; kh = (tap / 3) % 3; kw = tap % 3; }
      const int id = od * p.stride + kd - pad, ih = oh * p.stride + kh - pad, iw = ow * p.stride + kw - pad;
      const bool ok = row_ok && (unsigned)id < (unsigned)p.D && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
      const int64_t lin = ok ? ((int64_t)id * p.H + ih) * p.W + iw : 0;
      const uint32_t a_dst = smem_base + (uint32_t)(s * stage_bytes) + (uint32_t)(pt * 16);
      for (int c = 0; c < n_chunks; ++c) {
        const int kc = kc0 + c;
        const uint4* a_src = kc < p.c0_chunks ? p.a + (int64_t)kc * sp_in : p.a1 + (int64_t)(kc - p.c0_chunks) * sp_in;
        cp_async16(a_dst + (uint32_t)(c * 2048), a_src + lin, ok ? 16u : 0u);
      }
      const uint32_t b_dst = smem_base + (uint32_t)(s * stage_bytes + a_bytes);
      const uint4* b_src = p.w + ((int64_t)tap * (p.cin_pad / 8) + kc0) * p.cout_pad + n0;
      for (int e = pt; e < b_items; e += G_NPROD) {
        const int c = e / p.n_tile;
        const int n = e - c * p.n_tile;
        cp_async16(b_dst + (uint32_t)(e * 16), b_src + (int64_t)c * p.cout_pad + n, 16u);
      }
      cp_async_commit();
      if (it >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async();
        mbar_arrive(&bar_full[(it - LAG) % G_NS]);
      }
    }
    drain_stages<LAG - 1>(bar_full, total, G_NS);
  } else if (warp == G_EPI_WARPS) {
    // =============================== MMA issuer ==================================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.n_tile);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t b_lbo = (uint32_t)p.n_tile * 16;
      for (int it = 0; it < total; ++it) {
        const int s = it % G_NS;
        mbar_wait(&bar_full[s], (uint32_t)(it / G_NS) & 1u);
        tc_fence_after();
        const uint32_t a0 = smem_base + (uint32_t)(s * stage_bytes);
        const uint32_t b0 = a0 + (uint32_t)a_bytes;
        for (int ks = 0; ks < p.kstage; ++ks)
          umma_bf16(tmem_base, umma_desc(a0 + (uint32_t)(ks * 2 * 2048), 2048, 128),
                    umma_desc(b0 + (uint32_t)(ks * 2) * b_lbo, b_lbo, 128), idesc, (it | ks) != 0 ? 1u : 0u);
        umma_commit(&bar_empty[s]);
      }
      umma_commit(bar_acc);
    }
    __syncwarp();
  } else {
    // =============================== epilogue ====================================================
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int64_t m = m0 + warp * 32 + lane;
    const bool row_ok = m < m_total;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
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
          s_stat[(warp * 2) * p.n_tile + ch] = v[0];
          s_stat[(warp * 2 + 1) * p.n_tile + ch] = q[0];
        }
      }
    }
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
  const int ns = stage_bytes <= 12 * 1024 ? 8 : stage_bytes <= 16 * 1024 ? 6 : stage_bytes <= 26 * 1024 ? 4 : 3;
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
