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
                  int64_t spatial) {
  __shared__ float s_mean[8], s_rstd[8];
  const int kc = blockIdx.y;
  chunk_norm_to_smem(n, kc, s_mean, s_rstd);
  const int64_t base = (int64_t)kc * spatial;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < spatial; p += (int64_t)gridDim.x * 256) {
    float f[8];
    unpack8(__ldg(x + base + p), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = apply_act((f[k] - s_mean[k]) * s_rstd[k], n.act);
    if (res != nullptr) {
      float r[8];
      unpack8(__ldg(res + base + p), r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
    y[base + p] = pack8(f);
  }
}

int launch_norm_act_b(const void* x, const BNorm& n, const void* res, void* y, int channels, int64_t spatial,
                      cudaStream_t st) {
  unsigned gx = (unsigned)((spatial + 255) / 256);
  if (gx > 4096) gx = 4096;
  norm_act_b_kernel<<<dim3(gx, channels / 8), 256, 0, st>>>(reinterpret_cast<const uint4*>(x), n,
                                                          reinterpret_cast<const uint4*>(res),
                                                          reinterpret_cast<uint4*>(y), spatial);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- B -> fp32 NCDHW (stage read-back, tests) ---------------------------------------------------
__global__ void __launch_bounds__(256)
unblock_kernel(const uint4* __restrict__ x, float* __restrict__ y, int channels, int64_t spatial) {
  const int kc = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= spatial) return;
  float f[8];
  unpack8(__ldg(x + (int64_t)kc * spatial + p), f);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (kc * 8 + k < channels) y[(int64_t)(kc * 8 + k) * spatial + p] = f[k];
}

int launch_unblock(const void* x, float* y, int channels, int64_t spatial, cudaStream_t st) {
  unblock_kernel<<<dim3((unsigned)((spatial + 255) / 256), (channels + 7) / 8), 256, 0, st>>>(
      reinterpret_cast<const uint4*>(x), y, channels, spatial);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- convert_dim (cls_wise_former.py:15-23) fused with the InstanceNorm + LeakyReLU before it -----
// tokens[(d/p0,h/p1,w/p2)][(c,p0,p1,p2)] = act(norm(x)), optionally also a dense fp32 NCDHW copy.
__global__ void __launch_bounds__(256)
tokenise_b_kernel(const uint4* __restrict__ x, BNorm n, int chunk0, float* __restrict__ tokens,
                  float* __restrict__ dense, int channels, int g, int p0, int p1, int p2) {
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
  unpack8(__ldg(x + (int64_t)(chunk0 + kc) * spatial + p), f);
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
                      int grid, int p0, int p1, int p2, cudaStream_t st) {
  const int64_t spatial = (int64_t)grid * grid * grid;
  tokenise_b_kernel<<<dim3((unsigned)((spatial + 255) / 256), channels / 8), 256, 0, st>>>(
      reinterpret_cast<const uint4*>(x), n, chunk0, tokens, dense_or_null, channels, grid, p0, p1, p2);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- split_dim (cls_wise_former.py:26-39) of (class_token * tokens) into B-format -----------------
__global__ void __launch_bounds__(256)
untokenise_b_kernel(const float* __restrict__ tokens, const float* __restrict__ class_token, uint4* __restrict__ y,
                    int channels, int g, int p0, int p1, int p2) {
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
  y[(int64_t)kc * spatial + p] = pack8(f);
}

int launch_untokenise_b(const float* tokens, const float* class_token, void* y, int channels, int grid, int p0, int p1,
                        int p2, cudaStream_t st) {
  const int64_t spatial = (int64_t)grid * grid * grid;
  untokenise_b_kernel<<<dim3((unsigned)((spatial + 255) / 256), channels / 8), 256, 0, st>>>(
      tokens, class_token, reinterpret_cast<uint4*>(y), channels, grid, p0, p1, p2);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- DeUp_Cat (cls_wise_former.py:716-729) as ONE kernel ------------------------------------------
// conv1 (1x1) -> ConvTranspose3d k2 s2 -> conv3 (1x1) on cat(skip, .) is linear, and every output voxel has
// exactly one parent input voxel and one of 8 taps t = (kd,kh,kw):
//     y[v] = W3a . skip[v] + M_t . x[parent(v)] + b_t,   M_t = W3b . Wt_t^T . W1,  b_t = W3b.(Wt_t^T b1 + bt) + b3
// (M_t, b_t composed on the host in fp64).  Thread = one input voxel, one (kd,kh), both kw, 16 outputs.
// Weights (fp32) in shared memory: wm[kw][c][16], ws[s][16], bt[kw][16].
template <int CIN>
__global__ void __launch_bounds__(128)
deup_fused_b_kernel(const uint4* __restrict__ x, const uint4* __restrict__ skip, const float* __restrict__ mt,
                    const float* __restrict__ w3a, const float* __restrict__ bt, uint4* __restrict__ y, int gi) {
  constexpr int CH = CIN / 2;                       // skip / output channels
  __shared__ __align__(16) float s_wm[2][CIN][16];
  __shared__ __align__(16) float s_ws[CH][16];
  __shared__ float s_b[2][16];
  const int kdh = blockIdx.y;                       // kd*2 + kh
  const int og = blockIdx.z;                        // 16-output group
  for (int e = threadIdx.x; e < 2 * CIN * 16; e += 128) {
    const int o = e % 16, c = (e / 16) % CIN, kw = e / (16 * CIN);
    s_wm[kw][c][o] = __ldg(mt + ((int64_t)(kdh * 2 + kw) * CH + og * 16 + o) * CIN + c);   // mt[t][o][c]
  }
  for (int e = threadIdx.x; e < CH * 16; e += 128) {
    const int o = e % 16, s = e / 16;
    s_ws[s][o] = __ldg(w3a + (int64_t)(og * 16 + o) * CH + s);                            // w3a[o][s]
  }
  if (threadIdx.x < 32) s_b[threadIdx.x / 16][threadIdx.x % 16] = __ldg(bt + (kdh * 2 + threadIdx.x / 16) * CH + og * 16 + threadIdx.x % 16);
  __syncthreads();
  const int64_t sp_in = (int64_t)gi * gi * gi;
  const int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (p >= sp_in) return;
  const int w = (int)(p % gi);
  const int h = (int)((p / gi) % gi);
  const int d = (int)(p / ((int64_t)gi * gi));
  const int go = 2 * gi;
  const int64_t sp_out = (int64_t)go * go * go;
  const int64_t q0 = ((int64_t)(2 * d + (kdh >> 1)) * go + (2 * h + (kdh & 1))) * go + 2 * w;   // kw = 0; kw = 1 is q0 + 1
  float acc[2][16];
#pragma unroll
  for (int o = 0; o < 16; ++o) { acc[0][o] = s_b[0][o]; acc[1][o] = s_b[1][o]; }
#pragma unroll 1
  for (int kc = 0; kc < CIN / 8; ++kc) {
    float f[8];
    unpack8(__ldg(x + (int64_t)kc * sp_in + p), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int kw = 0; kw < 2; ++kw) {
        const float4* wr = reinterpret_cast<const float4*>(&s_wm[kw][kc * 8 + k][0]);
#pragma unroll
        for (int o4 = 0; o4 < 4; ++o4) {
          const float4 wv = wr[o4];
          acc[kw][4 * o4] = fmaf(f[k], wv.x, acc[kw][4 * o4]);
          acc[kw][4 * o4 + 1] = fmaf(f[k], wv.y, acc[kw][4 * o4 + 1]);
          acc[kw][4 * o4 + 2] = fmaf(f[k], wv.z, acc[kw][4 * o4 + 2]);
          acc[kw][4 * o4 + 3] = fmaf(f[k], wv.w, acc[kw][4 * o4 + 3]);
        }
      }
    }
  }
#pragma unroll 1
  for (int kc = 0; kc < CH / 8; ++kc) {
    float f0[8], f1[8];
    unpack8(__ldg(skip + (int64_t)kc * sp_out + q0), f0);
    unpack8(__ldg(skip + (int64_t)kc * sp_out + q0 + 1), f1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4* wr = reinterpret_cast<const float4*>(&s_ws[kc * 8 + k][0]);
#pragma unroll
      for (int o4 = 0; o4 < 4; ++o4) {
        const float4 wv = wr[o4];
        acc[0][4 * o4] = fmaf(f0[k], wv.x, acc[0][4 * o4]);
        acc[0][4 * o4 + 1] = fmaf(f0[k], wv.y, acc[0][4 * o4 + 1]);
        acc[0][4 * o4 + 2] = fmaf(f0[k], wv.z, acc[0][4 * o4 + 2]);
        acc[0][4 * o4 + 3] = fmaf(f0[k], wv.w, acc[0][4 * o4 + 3]);
        acc[1][4 * o4] = fmaf(f1[k], wv.x, acc[1][4 * o4]);
        acc[1][4 * o4 + 1] = fmaf(f1[k], wv.y, acc[1][4 * o4 + 1]);
        acc[1][4 * o4 + 2] = fmaf(f1[k], wv.z, acc[1][4 * o4 + 2]);
        acc[1][4 * o4 + 3] = fmaf(f1[k], wv.w, acc[1][4 * o4 + 3]);
      }
    }
  }
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    float a0[8], a1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a0[k] = acc[0][8 * hf + k]; a1[k] = acc[1][8 * hf + k]; }
    uint4* dst = y + (int64_t)(og * 2 + hf) * sp_out + q0;
    dst[0] = pack8(a0);
    dst[1] = pack8(a1);
  }
}

int launch_deup_fused_b(const void* x, const void* skip, const float* mt, const float* w3a, const float* bt, void* y,
                        int cin, int gi, cudaStream_t st) {
  const int64_t sp_in = (int64_t)gi * gi * gi;
  dim3 grid((unsigned)((sp_in + 127) / 128), 4, cin / 2 / 16);
  const uint4* xp = reinterpret_cast<const uint4*>(x);
  const uint4* sp = reinterpret_cast<const uint4*>(skip);
  uint4* yp = reinterpret_cast<uint4*>(y);
  if (cin == 32) deup_fused_b_kernel<32><<<grid, 128, 0, st>>>(xp, sp, mt, w3a, bt, yp, gi);
  else if (cin == 64) deup_fused_b_kernel<64><<<grid, 128, 0, st>>>(xp, sp, mt, w3a, bt, yp, gi);
  else if (cin == 128) deup_fused_b_kernel<128><<<grid, 128, 0, st>>>(xp, sp, mt, w3a, bt, yp, gi);
  else { set_error("deup_fused: cin must be 32, 64 or 128"); return -1; }
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- endconv (1x1, 16 -> 4) + Softmax(dim=1) (cls_wise_former.py:662-663): B -> fp32 NCDHW ----------
__global__ void __launch_bounds__(256)
endconv_softmax_b_kernel(const uint4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                         float* __restrict__ probs, int64_t spatial) {
  __shared__ float s_w[4][16];
  __shared__ float s_b[4];
  if (threadIdx.x < 64) s_w[threadIdx.x / 16][threadIdx.x % 16] = __ldg(w + threadIdx.x);   // (4,16) row-major
  if (threadIdx.x < 4) s_b[threadIdx.x] = __ldg(b + threadIdx.x);
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= spatial) return;
  float f[16];
  {
    float t[8];
    unpack8(__ldg(x + p), t);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = t[k];
    unpack8(__ldg(x + spatial + p), t);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[8 + k] = t[k];
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

int launch_endconv_softmax_b(const void* x, const float* w, const float* b, float* probs, int64_t spatial,
                             cudaStream_t st) {
  endconv_softmax_b_kernel<<<(unsigned)((spatial + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(x), w, b,
                                                                            probs, spatial);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dcl
