// InstanceNorm3d(affine=False, eps=1e-5, biased variance) statistics and the element-wise
// kernels built on them (norm + activation + residual, tokenise / untokenise).
// Reference: nn.InstanceNorm3d uses in Unet_skipconnection.py:13-14,48-55 and cls_wise_former.py:207-223,745-752.
#include "common.cuh"

namespace dcl {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Each block reduces one 8192-element chunk of one channel.  Sums are taken of (x - k) with
// k = first element of the channel, so E[(x-k)^2] - E[x-k]^2 does not cancel; block partials are
// fp32, the cross-block accumulation is fp64 atomics.
constexpr int STAT_CHUNK = 8192;

__global__ void __launch_bounds__(256)
instnorm_partial_kernel(const float* __restrict__ x, int64_t spatial, double* __restrict__ accum) {
  const int c = blockIdx.y;
  const float* xc = x + (int64_t)c * spatial;
  const float k = __ldg(xc);
  const int64_t base = (int64_t)blockIdx.x * STAT_CHUNK;
  float s = 0.f, ss = 0.f;
  if ((spatial & 3) == 0) {
#pragma unroll 4
    for (int i = threadIdx.x * 4; i < STAT_CHUNK; i += 1024) {
      int64_t p = base + i;
      if (p < spatial) {
        float4 v = __ldg(reinterpret_cast<const float4*>(xc + p));
        float a = v.x - k, b = v.y - k, cc = v.z - k, d = v.w - k;
        s += (a + b) + (cc + d);
        ss += (a * a + b * b) + (cc * cc + d * d);
      }
    }
  } else {
    for (int i = threadIdx.x; i < STAT_CHUNK; i += 256) {
      int64_t p = base + i;
      if (p < spatial) { float a = __ldg(xc + p) - k; s += a; ss += a * a; }
    }
  }
  __shared__ float red[2][8];
  s = warp_sum(s); ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float a = threadIdx.x < 8 ? red[0][threadIdx.x] : 0.f;
    float b = threadIdx.x < 8 ? red[1][threadIdx.x] : 0.f;
    a = warp_sum(a); b = warp_sum(b);
    if (threadIdx.x == 0) {
      atomicAdd(accum + 2 * c, (double)a);
      atomicAdd(accum + 2 * c + 1, (double)b);
    }
  }
}

__global__ void instnorm_finalize_kernel(const float* __restrict__ x, int64_t spatial, int channels,
                                         double* __restrict__ accum, float* __restrict__ mean,
                                         float* __restrict__ rstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  double k = (double)__ldg(x + (int64_t)c * spatial);
  double m = accum[2 * c] / (double)spatial;
  double var = accum[2 * c + 1] / (double)spatial - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)(k + m);
  rstd[c] = (float)(1.0 / sqrt(var + 1e-5));
  accum[2 * c] = 0.0;       // leave the accumulator clean for its next use
  accum[2 * c + 1] = 0.0;
}

int launch_instnorm_stats(const float* x, int channels, int64_t spatial, double* accum, float* mean, float* rstd,
                          cudaStream_t st) {
  dim3 grid((unsigned)((spatial + STAT_CHUNK - 1) / STAT_CHUNK), channels);
  instnorm_partial_kernel<<<grid, 256, 0, st>>>(x, spatial, accum);
  instnorm_finalize_kernel<<<(channels + 127) / 128, 128, 0, st>>>(x, spatial, channels, accum, mean, rstd);
  g_launches += 2;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// y = act((x - mean) * rstd) + residual, float4 over a dense (C, spatial) tensor.
__global__ void __launch_bounds__(256)
norm_act_res_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                    int act, const float* __restrict__ residual, float* __restrict__ y, int64_t spatial) {
  const int c = blockIdx.y;
  const float m = __ldg(mean + c), r = __ldg(rstd + c);
  const int64_t base = (int64_t)c * spatial;
  for (int64_t p = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; p < spatial; p += (int64_t)gridDim.x * 1024) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x + base + p));
    float4 o;
    o.x = apply_act((v.x - m) * r, act);
    o.y = apply_act((v.y - m) * r, act);
    o.z = apply_act((v.z - m) * r, act);
    o.w = apply_act((v.w - m) * r, act);
    if (residual) {
      float4 q = __ldg(reinterpret_cast<const float4*>(residual + base + p));
      o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
    }
    *reinterpret_cast<float4*>(y + base + p) = o;
  }
}

int launch_norm_act_res(const float* x, const float* mean, const float* rstd, int act, const float* residual, float* y,
                        int channels, int64_t spatial, cudaStream_t st) {
  if (spatial % 4 != 0) { set_error("norm_act_res: spatial size must be a multiple of 4"); return -1; }
  int64_t vec = spatial / 4;
  unsigned gx = (unsigned)((vec + 255) / 256);
  if (gx > 2048) gx = 2048;
  norm_act_res_kernel<<<dim3(gx, channels), 256, 0, st>>>(x, mean, rstd, act, residual, y, spatial);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// convert_dim (cls_wise_former.py:15-23) fused with the instance norm + LeakyReLU that precede it.
// Thread = one input element (coalesced reads along w).
__global__ void __launch_bounds__(256)
norm_act_tokenise_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                         int act, float* __restrict__ tokens, float* __restrict__ dense, int channels, int g, int p0,
                         int p1, int p2) {
  const int64_t n = (int64_t)channels * g * g * g;
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  int w = e % g;
  int64_t r = e / g;
  int h = r % g; r /= g;
  int d = r % g;
  int c = r / g;
  float v = apply_act((__ldg(x + e) - __ldg(mean + c)) * __ldg(rstd + c), act);
  if (dense) dense[e] = v;
  const int g1 = g / p1, g2 = g / p2;
  int tok = ((d / p0) * g1 + (h / p1)) * g2 + (w / p2);
  int fea = ((c * p0 + d % p0) * p1 + h % p1) * p2 + w % p2;
  tokens[(int64_t)tok * (channels * p0 * p1 * p2) + fea] = v;
}

int launch_norm_act_tokenise(const float* x, const float* mean, const float* rstd, int act, float* tokens,
                             float* dense_or_null, int channels, int grid, int p0, int p1, int p2, cudaStream_t st) {
  int64_t n = (int64_t)channels * grid * grid * grid;
  norm_act_tokenise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, mean, rstd, act, tokens, dense_or_null,
                                                                      channels, grid, p0, p1, p2);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// split_dim (cls_wise_former.py:26-39) of (class_token * tokens).  Thread = one output element.
__global__ void __launch_bounds__(256)
scale_untokenise_kernel(const float* __restrict__ tokens, const float* __restrict__ class_token,
                        float* __restrict__ y, int channels, int g, int p0, int p1, int p2) {
  const int64_t n = (int64_t)channels * g * g * g;
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  int w = e % g;
  int64_t r = e / g;
  int h = r % g; r /= g;
  int d = r % g;
  int c = r / g;
  const int g1 = g / p1, g2 = g / p2;
  int tok = ((d / p0) * g1 + (h / p1)) * g2 + (w / p2);
  int fea = ((c * p0 + d % p0) * p1 + h % p1) * p2 + w % p2;
  float s = class_token ? __ldg(class_token + fea) : 1.f;
  y[e] = s * __ldg(tokens + (int64_t)tok * (channels * p0 * p1 * p2) + fea);
}

int launch_scale_untokenise(const float* tokens, const float* class_token, float* y, int channels, int grid, int p0,
                            int p1, int p2, cudaStream_t st) {
  int64_t n = (int64_t)channels * grid * grid * grid;
  scale_untokenise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tokens, class_token, y, channels, grid, p0, p1,
                                                                     p2);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void fill16_kernel(float* __restrict__ dst, Floats16 v) { dst[threadIdx.x] = v.v[threadIdx.x]; }

__global__ void stamp_kernel(unsigned long long* __restrict__ dst, int idx) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  dst[idx] = t;
}
int launch_stamp(unsigned long long* dst, int idx, cudaStream_t st) {
  stamp_kernel<<<1, 1, 0, st>>>(dst, idx);
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void patch_desc_kernel(PatchDesc* __restrict__ dst, PatchDesc v) {
  if (threadIdx.x == 0) { dst->x = v.x; dst->sc = v.sc; dst->sd = v.sd; dst->sh = v.sh; dst->probs = v.probs; }
  if (threadIdx.x < 16) dst->keep[threadIdx.x] = v.keep[threadIdx.x];
  if (threadIdx.x >= 16 && threadIdx.x < 28) dst->aux[threadIdx.x - 16] = v.aux[threadIdx.x - 16];
}

int launch_patch_desc(PatchDesc* dst, const PatchDesc& v, cudaStream_t st) {
  patch_desc_kernel<<<1, 32, 0, st>>>(dst, v);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_fill16(float* dst, const Floats16& v, cudaStream_t st) {
  fill16_kernel<<<1, 16, 0, st>>>(dst, v);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dcl
