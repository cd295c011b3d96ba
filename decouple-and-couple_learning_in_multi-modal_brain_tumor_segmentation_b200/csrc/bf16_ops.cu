// HBM-bound kernels of the bf16 pipeline.  Activations are "B-format": bf16, channel-blocked
// [C/8][spatial][8], i.e. one 16-byte vector per (8-channel chunk, voxel); consecutive voxels of a chunk
// are consecutive vectors, so every kernel here moves 16 bytes per thread, fully coalesced.
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace dcl {

using namespace tc;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t* p = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(p[k] << 16);
    f[2 * k + 1] = __uint_as_float(p[k] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// B-format load / store of one (chunk, voxel) vector; lo_off != 0 = split-fp16 (the lo planes sit lo_off vectors behind)
__device__ __forceinline__ void load8(const uint4* __restrict__ p, int64_t lo_off, float (&f)[8]) {
  if (lo_off != 0) {        // split mode: two fp16 halves
    unpack8_x3<false>(__ldg(p), f);
    unpack8_x3<true>(__ldg(p + lo_off), f);
  } else {
    unpack8(__ldg(p), f);
  }
}
__device__ __forceinline__ void store8(uint4* __restrict__ p, int64_t lo_off, const float (&f)[8]) {
  if (lo_off != 0) {
    uint4 h, l;
    split8(f, h, l);
    p[0] = h;
    p[lo_off] = l;
  } else {
    p[0] = pack8(f);
  }
}

// mean / rstd of the 8 channels of chunk kc into shared memory (threads 0..7 of the block)
__device__ __forceinline__ void chunk_norm_to_smem(const BNorm& n, int kc, float* s_mean, float* s_rstd) {
  if (threadIdx.x < 8) {
    const int c = kc * 8 + threadIdx.x;
    float m = 0.f, r = 1.f;
    if (n.sums != nullptr) {
      stat_mean_rstd(n.sums, c, n.inv_n, &m, &r);
    } else if (n.mean != nullptr) {
      m = n.mean[c];
      r = n.rstd[c];
    }
    s_mean[threadIdx.x] = m;
    s_rstd[threadIdx.x] = r;
  }
  __syncthreads();
}

// ---- y = act(norm(x)) (+ residual), B -> B  (EnBlock2 / DeBlock tail, cls_wise_former.py:705-713;
//      also the pre-normalised input of the GEMM kernel) -------------------------------------------
__global__ void __launch_bounds__(256)
norm_act_b_kernel(const uint4* __restrict__ x, BNorm n, const uint4* __restrict__ res, uint4* __restrict__ y,
                  int64_t spatial, int64_t lo_off) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  __shared__ float s_mean[8], s_rstd[8];
  const int kc = blockIdx.y;
  chunk_norm_to_smem(n, kc, s_mean, s_rstd);
  const int64_t base = (int64_t)kc * spatial;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < spatial; p += (int64_t)gridDim.x * 256) {
    float f[8];
    load8(x + base + p, lo_off, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = apply_act((f[k] - s_mean[k]) * s_rstd[k], n.act);
    if (res != nullptr) {
      float r[8];
      load8(res + base + p, lo_off, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
    store8(y + base + p, lo_off, f);
  }
}

int launch_norm_act_b(const void* x, const BNorm& n, const void* res, void* y, int channels, int64_t spatial,
                      cudaStream_t st, bool x3) {
  unsigned gx = (unsigned)((spatial + 255) / 256);
  if (gx > 4096) gx = 4096;
  DCL_CUDA_OK(launch_pdl(norm_act_b_kernel, dim3(dim3(gx, channels / 8)), dim3(256), (size_t)(0), st, reinterpret_cast<const uint4*>(x), n,
                                                          reinterpret_cast<const uint4*>(res),
                                                          reinterpret_cast<uint4*>(y), spatial,
                                                          x3 ? (int64_t)(channels / 8) * spatial : (int64_t)0));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- B -> fp32 NCDHW (stage read-back, tests) ---------------------------------------------------
__global__ void __launch_bounds__(256)
unblock_kernel(const uint4* __restrict__ x, float* __restrict__ y, int channels, int64_t spatial, int64_t lo_off) {
  const int kc = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= spatial) return;
  float f[8];
  load8(x + (int64_t)kc * spatial + p, lo_off, f);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (kc * 8 + k < channels) y[(int64_t)(kc * 8 + k) * spatial + p] = f[k];
}

int launch_unblock(const void* x, float* y, int channels, int64_t spatial, cudaStream_t st, bool x3) {
  unblock_kernel<<<dim3((unsigned)((spatial + 255) / 256), (channels + 7) / 8), 256, 0, st>>>(
      reinterpret_cast<const uint4*>(x), y, channels, spatial, x3 ? (int64_t)((channels + 15) / 16 * 2) * spatial : (int64_t)0);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- convert_dim (cls_wise_former.py:15-23) fused with the InstanceNorm + LeakyReLU before it -----
// tokens[(d/p0,h/p1,w/p2)][(c,p0,p1,p2)] = act(norm(x)), optionally also a dense fp32 NCDHW copy.
__global__ void __launch_bounds__(256)
tokenise_b_kernel(const uint4* __restrict__ x, BNorm n, int chunk0, float* __restrict__ tokens,
                  float* __restrict__ dense, int channels, int g, int p0, int p1, int p2, int64_t lo_off) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  __shared__ float s_mean[8], s_rstd[8];
  const int kc = blockIdx.y;                       // chunk within this region's channels
  chunk_norm_to_smem(n, chunk0 + kc, s_mean, s_rstd);
  const int64_t spatial = (int64_t)g * g * g;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= spatial) return;
  const int w = (int)(p % g);
  const int h = (int)((p / g) % g);
  const int d = (int)(p / ((int64_t)g * g));
  float f[8];
  load8(x + (int64_t)(chunk0 + kc) * spatial + p, lo_off, f);
  const int g1 = g / p1, g2 = g / p2;
  const int tok = ((d / p0) * g1 + (h / p1)) * g2 + (w / p2);
  const int sub = ((d % p0) * p1 + h % p1) * p2 + w % p2;
  const int pvol = p0 * p1 * p2;
  float* trow = tokens + (int64_t)tok * (channels * pvol) + sub;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = kc * 8 + k;
    const float v = apply_act((f[k] - s_mean[k]) * s_rstd[k], n.act);
    trow[c * pvol] = v;
    if (dense != nullptr) dense[(int64_t)c * spatial + p] = v;
  }
}

// x: B-format tensor holding (at least) channels [chunk0*8, chunk0*8 + channels) of a g^3 grid
int launch_tokenise_b(const void* x, const BNorm& n, int chunk0, float* tokens, float* dense_or_null, int channels,
                      int grid, int p0, int p1, int p2, cudaStream_t st, int x3_chunks) {
  const int64_t spatial = (int64_t)grid * grid * grid;
  DCL_CUDA_OK(launch_pdl(tokenise_b_kernel, dim3(dim3((unsigned)((spatial + 255) / 256), channels / 8)), dim3(256), (size_t)(0), st, 
      reinterpret_cast<const uint4*>(x), n, chunk0, tokens, dense_or_null, channels, grid, p0, p1, p2,
      (int64_t)x3_chunks * spatial));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- split_dim (cls_wise_former.py:26-39) of (class_token * tokens) into B-format -----------------
__global__ void __launch_bounds__(256)
untokenise_b_kernel(const float* __restrict__ tokens, const float* __restrict__ class_token, uint4* __restrict__ y,
                    int channels, int g, int p0, int p1, int p2, int64_t lo_off) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  const int kc = blockIdx.y;
  const int64_t spatial = (int64_t)g * g * g;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= spatial) return;
  const int w = (int)(p % g);
  const int h = (int)((p / g) % g);
  const int d = (int)(p / ((int64_t)g * g));
  const int g1 = g / p1, g2 = g / p2;
  const int tok = ((d / p0) * g1 + (h / p1)) * g2 + (w / p2);
  const int sub = ((d % p0) * p1 + h % p1) * p2 + w % p2;
  const int pvol = p0 * p1 * p2;
  const float* trow = tokens + (int64_t)tok * (channels * pvol) + sub;
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int fea = (kc * 8 + k) * pvol;
    f[k] = (class_token ? __ldg(class_token + fea + sub) : 1.f) * __ldg(trow + fea);
  }
  store8(y + (int64_t)kc * spatial + p, lo_off, f);
}

int launch_untokenise_b(const float* tokens, const float* class_token, void* y, int channels, int grid, int p0, int p1,
                        int p2, cudaStream_t st, bool x3) {
  const int64_t spatial = (int64_t)grid * grid * grid;
  DCL_CUDA_OK(launch_pdl(untokenise_b_kernel, dim3(dim3((unsigned)((spatial + 255) / 256), channels / 8)), dim3(256), (size_t)(0), st, 
      tokens, class_token, reinterpret_cast<uint4*>(y), channels, grid, p0, p1, p2,
      x3 ? (int64_t)(channels / 8) * spatial : (int64_t)0));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- DeUp_Cat (cls_wise_former.py:716-729) as ONE kernel ------------------------------------------
// conv1 (1x1) -> ConvTranspose3d k2 s2 -> conv3 (1x1) on cat(skip, .) is linear, and every output voxel has
// exactly one parent input voxel and one of 8 taps t = (kd,kh,kw):
//     y[v] = W3a . skip[v] + M_t . x[parent(v)] + b_t,   M_t = W3b . Wt_t^T . W1,  b_t = W3b.(Wt_t^T b1 + bt) + b3
// (M_t, b_t composed on the host in fp64).  The op is HBM bound (reads x once per tap pair from L2, skip once,
// writes y once) with ~5 GFLOP of tiny-K GEMM work, so it runs on warp-level bf16 MMAs (mma.sync m16n8k16, fp32
// accumulate) fed straight from the B-format tensors: a quad of lanes covers one 16-byte voxel vector, so every
// fragment load / store of a warp is one contiguous 128-byte (x) or 256-byte-span (skip / y, both kw) access and
// no shared-memory staging of activations is needed.  A CTA serves one (kd,kh); a warp walks tiles of 16 parent
// voxels along w and produces their 2 x 16 outputs (kw = 0, 1).  Weights: bf16 in shared memory, rows padded by
// 8 elements (conflict-free fragment reads).
// F16: fp16 operands (the split-operand mode), else bf16
template <bool F16>
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// X3: split-fp16 tensors and weights (hi image followed by the lo image): every product is hi*hi + lo*hi + hi*lo.
template <int CIN, bool X3>
__global__ void __launch_bounds__(256)
deup_mma_b_kernel(const uint4* __restrict__ x, const uint4* __restrict__ skip, const float* __restrict__ mt,
                  const float* __restrict__ w3a, const float* __restrict__ bt, uint4* __restrict__ y, int gi, float out_mul) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  constexpr int CH = CIN / 2;                       // skip / output channels
  constexpr int NT = CH / 8, KX = CIN / 16, KS = CH / 16;
  constexpr int LDM = CIN + 8, LDS_ = CH + 8;       // padded row lengths (elements)
  extern __shared__ __align__(16) uint8_t deup_smem[];
  constexpr int NIMG = X3 ? 2 : 1;
  constexpr int WM_IMG = 2 * CH * LDM, WS_IMG = CH * LDS_;                     // elements of one shared-memory image
  __nv_bfloat16* s_wm = reinterpret_cast<__nv_bfloat16*>(deup_smem);          // [hi|lo][2 kw][CH][LDM]
  __nv_bfloat16* s_ws = s_wm + NIMG * WM_IMG;                                  // [hi|lo][CH][LDS_]
  float* s_b = reinterpret_cast<float*>(s_ws + NIMG * WS_IMG);                 // [2 kw][CH]
  const int kdh = blockIdx.y;                       // kd*2 + kh
  for (int img = 0; img < NIMG; ++img) {
    // mt / w3a arrive as the bf16 padded images built at weight-load time: [hi|lo][8 taps][CH][LDM] and [hi|lo][CH][LDS_]
    const uint4* src_m = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(mt) + (size_t)img * 8 * CH * LDM +
                                                        (size_t)(kdh * 2) * CH * LDM);
    uint4* dst_m = reinterpret_cast<uint4*>(s_wm + img * WM_IMG);
    for (int e = threadIdx.x; e < 2 * CH * LDM / 8; e += 256) dst_m[e] = __ldg(src_m + e);
    const uint4* src_s = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(w3a) + (size_t)img * WS_IMG);
    uint4* dst_s = reinterpret_cast<uint4*>(s_ws + img * WS_IMG);
    for (int e = threadIdx.x; e < CH * LDS_ / 8; e += 256) dst_s[e] = __ldg(src_s + e);
  }
  for (int e = threadIdx.x; e < 2 * CH; e += 256) s_b[e] = __ldg(bt + (int64_t)(kdh * 2) * CH + e);                   // bt[t][o]
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, c = lane & 3;            // fragment row / column-pair of this lane
  const int64_t sp_in = (int64_t)gi * gi * gi;
  const int go = 2 * gi;
  const int64_t sp_out = 8 * sp_in;
  const int tiles_per_row = gi / 16;
  const int64_t n_tiles = sp_in / 16;
  const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(skip);
  uint32_t* yw = reinterpret_cast<uint32_t*>(y);
  for (int64_t tile = (int64_t)blockIdx.x * 8 + warp; tile < n_tiles; tile += (int64_t)gridDim.x * 8) {
    const int w0 = (int)(tile % tiles_per_row) * 16;
    const int64_t dh = tile / tiles_per_row;
    const int h = (int)(dh % gi), d = (int)(dh / gi);
    const int64_t p0 = (dh * gi) + w0;
    const int64_t q0 = ((int64_t)(2 * d + (kdh >> 1)) * go + (2 * h + (kdh & 1))) * go + 2 * w0;
    // A fragments of the 16 parent voxels (shared by both kw): chunk 2ks holds k 0-7, chunk 2ks+1 holds k 8-15
    uint32_t ax[NIMG][KX][4];
#pragma unroll
    for (int img = 0; img < NIMG; ++img)
#pragma unroll
    for (int ks = 0; ks < KX; ++ks) {
      const int64_t b0 = ((int64_t)(img * 2 * KX + 2 * ks) * sp_in + p0) * 4 + c, b1 = ((int64_t)(img * 2 * KX + 2 * ks + 1) * sp_in + p0) * 4 + c;
      ax[img][ks][0] = __ldg(xw + b0 + g * 4);
      ax[img][ks][1] = __ldg(xw + b0 + (g + 8) * 4);
      ax[img][ks][2] = __ldg(xw + b1 + g * 4);
      ax[img][ks][3] = __ldg(xw + b1 + (g + 8) * 4);
    }
#pragma unroll
    for (int kw = 0; kw < 2; ++kw) {
      const int64_t qa = q0 + 2 * g + kw, qb = q0 + 2 * (g + 8) + kw;     // output voxels of fragment rows g, g+8
      uint32_t as[NIMG][KS][4];
#pragma unroll
      for (int img = 0; img < NIMG; ++img)
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int kc = img * 2 * KS + 2 * ks;
        as[img][ks][0] = __ldg(sw + ((int64_t)kc * sp_out + qa) * 4 + c);
        as[img][ks][1] = __ldg(sw + ((int64_t)kc * sp_out + qb) * 4 + c);
        as[img][ks][2] = __ldg(sw + ((int64_t)(kc + 1) * sp_out + qa) * 4 + c);
        as[img][ks][3] = __ldg(sw + ((int64_t)(kc + 1) * sp_out + qb) * 4 + c);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float acc[4];
        acc[0] = acc[2] = s_b[kw * CH + nt * 8 + 2 * c];
        acc[1] = acc[3] = s_b[kw * CH + nt * 8 + 2 * c + 1];
        const __nv_bfloat16* wr = s_wm + (kw * CH + nt * 8 + g) * LDM + 2 * c;   // B fragment: n = g, k = 2c (+8)
#pragma unroll
        for (int ks = 0; ks < KX; ++ks) {
          mma_bf16_16816<X3>(acc, ax[0][ks], *reinterpret_cast<const uint32_t*>(wr + ks * 16), *reinterpret_cast<const uint32_t*>(wr + ks * 16 + 8));
          if constexpr (X3) {
            mma_bf16_16816<X3>(acc, ax[NIMG - 1][ks], *reinterpret_cast<const uint32_t*>(wr + ks * 16), *reinterpret_cast<const uint32_t*>(wr + ks * 16 + 8));
            mma_bf16_16816<X3>(acc, ax[0][ks], *reinterpret_cast<const uint32_t*>(wr + WM_IMG + ks * 16), *reinterpret_cast<const uint32_t*>(wr + WM_IMG + ks * 16 + 8));
          }
        }
        const __nv_bfloat16* wsr = s_ws + (nt * 8 + g) * LDS_ + 2 * c;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          mma_bf16_16816<X3>(acc, as[0][ks], *reinterpret_cast<const uint32_t*>(wsr + ks * 16), *reinterpret_cast<const uint32_t*>(wsr + ks * 16 + 8));
          if constexpr (X3) {
            mma_bf16_16816<X3>(acc, as[NIMG - 1][ks], *reinterpret_cast<const uint32_t*>(wsr + ks * 16), *reinterpret_cast<const uint32_t*>(wsr + ks * 16 + 8));
            mma_bf16_16816<X3>(acc, as[0][ks], *reinterpret_cast<const uint32_t*>(wsr + WS_IMG + ks * 16), *reinterpret_cast<const uint32_t*>(wsr + WS_IMG + ks * 16 + 8));
          }
        }
        // D fragment: (row g, n = 2c, 2c+1), (row g+8, same n) -> word c of the voxel's 16-byte vector of chunk nt
        if constexpr (X3) {
          // (split mode: mt / w3a / bt are stored times 2^k, out_mul = 2^-k)
          uint32_t h0, l0, h1, l1;
          tc::split_x2(acc[0] * out_mul, acc[1] * out_mul, h0, l0);
          tc::split_x2(acc[2] * out_mul, acc[3] * out_mul, h1, l1);
          yw[((int64_t)nt * sp_out + qa) * 4 + c] = h0;
          yw[((int64_t)nt * sp_out + qb) * 4 + c] = h1;
          yw[((int64_t)(NT + nt) * sp_out + qa) * 4 + c] = l0;
          yw[((int64_t)(NT + nt) * sp_out + qb) * 4 + c] = l1;
        } else {
          yw[((int64_t)nt * sp_out + qa) * 4 + c] = tc::pack_bf16x2(acc[0], acc[1]);
          yw[((int64_t)nt * sp_out + qb) * 4 + c] = tc::pack_bf16x2(acc[2], acc[3]);
        }
      }
    }
  }
}

template <int CIN, bool X3>
static int launch_deup_mma(const uint4* x, const uint4* skip, const float* mt, const float* w3a, const float* bt, uint4* y,
                           int gi, cudaStream_t st, float out_mul) {
  constexpr int CH = CIN / 2;
  constexpr int smem = (2 * CH * (CIN + 8) + CH * (CH + 8)) * 2 * (X3 ? 2 : 1) + 2 * CH * 4;
  static bool configured = false;
  if (!configured) {
    DCL_CUDA_OK(cudaFuncSetAttribute(deup_mma_b_kernel<CIN, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int64_t n_tiles = (int64_t)gi * gi * gi / 16;
  int gx = (int)((n_tiles + 7) / 8);
  if (gx > 148) gx = 148;                          // x 4 (kd,kh) CTAs of 8 warps: persistent over the parent tiles
  DCL_CUDA_OK(launch_pdl(deup_mma_b_kernel<CIN, X3>, dim3(dim3(gx, 4)), dim3(256), (size_t)(smem), st, x, skip, mt, w3a, bt, y, gi, out_mul));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_deup_fused_b(const void* x, const void* skip, const float* mt, const float* w3a, const float* bt, void* y,
                        int cin, int gi, cudaStream_t st, bool x3, float out_mul) {
  const uint4* xp = reinterpret_cast<const uint4*>(x);
  const uint4* sp = reinterpret_cast<const uint4*>(skip);
  uint4* yp = reinterpret_cast<uint4*>(y);
  if (gi % 16 != 0) { set_error("deup_fused: grid must be a multiple of 16"); return -1; }
  if (x3) {
    if (cin == 32) return launch_deup_mma<32, true>(xp, sp, mt, w3a, bt, yp, gi, st, out_mul);
    if (cin == 64) return launch_deup_mma<64, true>(xp, sp, mt, w3a, bt, yp, gi, st, out_mul);
    if (cin == 128) return launch_deup_mma<128, true>(xp, sp, mt, w3a, bt, yp, gi, st, out_mul);
  } else {
    if (cin == 32) return launch_deup_mma<32, false>(xp, sp, mt, w3a, bt, yp, gi, st, 1.f);
    if (cin == 64) return launch_deup_mma<64, false>(xp, sp, mt, w3a, bt, yp, gi, st, 1.f);
    if (cin == 128) return launch_deup_mma<128, false>(xp, sp, mt, w3a, bt, yp, gi, st, 1.f);
  }
  set_error("deup_fused: cin must be 32, 64 or 128");
  return -1;
}

// ---- endconv (1x1, 16 -> 4) + Softmax(dim=1) (cls_wise_former.py:662-663): B -> fp32 NCDHW ----------
__global__ void __launch_bounds__(256)
endconv_softmax_b_kernel(const uint4* __restrict__ x, BNorm n, const uint4* __restrict__ res, const float* __restrict__ w,
                         const float* __restrict__ b, float* __restrict__ probs_fixed, int64_t spatial,
                         const PatchDesc* __restrict__ desc, int64_t lo_off) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  // the replayed graph bakes probs_fixed in; a per-patch destination (gather-form stitch, caller's buffer) comes
  // through the device-side patch descriptor instead
  float* __restrict__ probs = (desc != nullptr && desc->probs != nullptr) ? desc->probs : probs_fixed;
  __shared__ float s_w[4][16];
  __shared__ float s_b[4];
  __shared__ float s_mean[16], s_rstd[16];
  if (threadIdx.x < 64) s_w[threadIdx.x / 16][threadIdx.x % 16] = __ldg(w + threadIdx.x);   // (4,16) row-major
  if (threadIdx.x < 4) s_b[threadIdx.x] = __ldg(b + threadIdx.x);
  const bool fused = n.sums != nullptr || n.mean != nullptr;       // input = act(norm(x)) + res (the DeBlock tail)
  if (fused && threadIdx.x >= 64 && threadIdx.x < 80) {
    const int c = threadIdx.x - 64;
    float m, r;
    if (n.sums != nullptr) stat_mean_rstd(n.sums, c, n.inv_n, &m, &r);
    else { m = n.mean[c]; r = n.rstd[c]; }
    s_mean[c] = m; s_rstd[c] = r;
  }
  __syncthreads();
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < spatial; p += (int64_t)gridDim.x * 256) {   // persistent: the
  float f[16];                                                                 // per-block prologue (fp64 statistics) amortises
#pragma unroll
  for (int kc = 0; kc < 2; ++kc) {
    float t[8];
    load8(x + kc * spatial + p, lo_off, t);
    if (fused) {
      // exactly norm_act_b_kernel's arithmetic, including the bf16 rounding of its stored result, so the fused and
      // the unfused (keep_stages) schedules give bit-identical probabilities
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = apply_act((t[k] - s_mean[kc * 8 + k]) * s_rstd[kc * 8 + k], n.act);
      if (res != nullptr) {
        float r[8];
        load8(res + kc * spatial + p, lo_off, r);
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] += r[k];
      }
      if (lo_off != 0) {       // the rounding of the stored (split) tensor
        uint4 qh, ql;
        split8(t, qh, ql);
        unpack8_x3<false>(qh, t);
        unpack8_x3<true>(ql, t);
      } else {
        const uint4 q = pack8(t);
        unpack8(q, t);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) f[kc * 8 + k] = t[k];
  }
  float l[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float a = s_b[c];
#pragma unroll
    for (int k = 0; k < 16; ++k) a = fmaf(f[k], s_w[c][k], a);
    l[c] = a;
  }
  const float m = fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3]));
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) { l[c] = expf(l[c] - m); s += l[c]; }
#pragma unroll
  for (int c = 0; c < 4; ++c) probs[(int64_t)c * spatial + p] = l[c] / s;
  }
}

int launch_endconv_softmax_b(const void* x, const float* w, const float* b, float* probs, int64_t spatial,
                             cudaStream_t st, const BNorm* norm, const void* res, const PatchDesc* desc, bool x3) {
  BNorm n;
  if (norm) n = *norm;
  unsigned gx = (unsigned)((spatial + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  DCL_CUDA_OK(launch_pdl(endconv_softmax_b_kernel, dim3(gx), dim3(256), (size_t)(0), st,
                         reinterpret_cast<const uint4*>(x), n, reinterpret_cast<const uint4*>(res), w, b, probs, spatial, desc,
                         x3 ? (int64_t)2 * spatial : (int64_t)0));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dcl
