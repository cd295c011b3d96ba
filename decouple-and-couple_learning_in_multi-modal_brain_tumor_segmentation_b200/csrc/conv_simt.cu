// fp32 FFMA convolution family (parity mode DCL_FP32): 3x3x3 conv stride 1|2, pointwise conv,
// transposed conv k2 s2.  NCDHW, w (= anatomical z) contiguous.  These are the exact-arithmetic
// kernels every other implementation in this library is A/B-checked against on the device.
#include "common.cuh"

namespace dcl {

// ---------------------------------------------------------------------------------------------
// 3x3x3 convolution, padding 1.
// Block = 256 threads, output tile TD x 8 x (4*TXN) voxels x 16 output channels; each thread owns
// 4 consecutive w and 16 couts (64 fp32 accumulators).  Input channels are streamed through shared
// memory CI_T at a time together with their [27][16] weight slab; the optional instance-norm +
// activation of the *input* tensor is applied while staging (zero padding after it, as torch does).
// ---------------------------------------------------------------------------------------------
template <int STRIDE, int TXN, int CI_T>
struct ConvTile {
  static constexpr int TW = 4 * TXN;
  static constexpr int TH = 8;
  static constexpr int TD = 256 / (8 * TXN);
  static constexpr int IW = (TW - 1) * STRIDE + 3;
  static constexpr int IH = (TH - 1) * STRIDE + 3;
  static constexpr int ID = (TD - 1) * STRIDE + 3;
  static constexpr int PITCH = (IW + 3) / 4 * 4;
  static constexpr int IN_ELEMS = ID * IH * PITCH;          // per input channel
  static constexpr int NIN = 3 * STRIDE + 3;                // inputs along w a thread needs (6 | 9)
  static constexpr int SMEM_FLOATS = CI_T * IN_ELEMS + CI_T * 27 * 16;
};

template <int STRIDE, int TXN, int CI_T>
__global__ void __launch_bounds__(256, 2)
conv3d_k3_kernel(ConvSrc src, ConvDst dst, const float* __restrict__ w_packed, int cout, int cout_pad, int in_d,
                 int in_h, int in_w, int od, int oh, int ow, int tiles_h, int tiles_w) {
  using T = ConvTile<STRIDE, TXN, CI_T>;
  __shared__ __align__(16) float smem[T::SMEM_FLOATS];
  float* s_in = smem;
  float* s_w = smem + CI_T * T::IN_ELEMS;

  const int tid = threadIdx.x;
  const int tx = tid % TXN;
  const int ty = (tid / TXN) % 8;
  const int tz = tid / (TXN * 8);

  int tile = blockIdx.x;
  const int tw_i = tile % tiles_w; tile /= tiles_w;
  const int th_i = tile % tiles_h;
  const int td_i = tile / tiles_h;
  const int od0 = td_i * T::TD, oh0 = th_i * T::TH, ow0 = tw_i * T::TW;
  const int id0 = od0 * STRIDE - 1, ih0 = oh0 * STRIDE - 1, iw0 = ow0 * STRIDE - 1;
  const int co0 = blockIdx.y * 16;
  const int cin = src.c0 + src.c1;
  const int64_t dense_hw = (int64_t)in_h * in_w;
  const int64_t dense_c = (int64_t)in_d * dense_hw;

  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[i][c] = 0.f;

  for (int ci0 = 0; ci0 < cin; ci0 += CI_T) {
    __syncthreads();
    // ---- stage the input tile (norm + act fused, zero padding outside the tensor)
    for (int e = tid; e < CI_T * T::ID * T::IH * T::IW; e += 256) {
      int w = e % T::IW;
      int r = e / T::IW;
      int h = r % T::IH; r /= T::IH;
      int d = r % T::ID;
      int ci = r / T::ID;
      int c = ci0 + ci;
      int gd = id0 + d, gh = ih0 + h, gw = iw0 + w;
      float v = 0.f;
      if (c < cin && gd >= 0 && gd < in_d && gh >= 0 && gh < in_h && gw >= 0 && gw < in_w) {
        if (c < src.c0)
          v = __ldg(src.x0 + c * src.s0c + gd * src.s0d + gh * src.s0h + gw);
        else
          v = __ldg(src.x1 + (c - src.c0) * dense_c + gd * dense_hw + (int64_t)gh * in_w + gw);
        if (src.mean != nullptr) v = (v - __ldg(src.mean + c)) * __ldg(src.rstd + c);
        v = apply_act(v, src.act);
      }
      s_in[ci * T::IN_ELEMS + (d * T::IH + h) * T::PITCH + w] = v;
    }
    // ---- stage the weight slab [CI_T][27][16]
    for (int e = tid; e < CI_T * 27 * 16; e += 256) {
      int co = e % 16;
      int r = e / 16;
      int tap = r % 27;
      int ci = r / 27;
      int c = ci0 + ci;
      s_w[e] = (c < cin) ? __ldg(w_packed + ((int64_t)c * 27 + tap) * cout_pad + co0 + co) : 0.f;
    }
    __syncthreads();

#pragma unroll 1
    for (int ci = 0; ci < CI_T; ++ci) {
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float* row =
              s_in + ci * T::IN_ELEMS + ((tz * STRIDE + kd) * T::IH + (ty * STRIDE + kh)) * T::PITCH + tx * 4 * STRIDE;
          float in[T::NIN];
          {
            float4 a = *reinterpret_cast<const float4*>(row);
            in[0] = a.x; in[1] = a.y; in[2] = a.z; in[3] = a.w;
            if constexpr (STRIDE == 1) {
              float2 b = *reinterpret_cast<const float2*>(row + 4);
              in[4] = b.x; in[5] = b.y;
            } else {
              float4 b = *reinterpret_cast<const float4*>(row + 4);
              in[4] = b.x; in[5] = b.y; in[6] = b.z; in[7] = b.w;
              in[8] = row[8];
            }
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4* wp = reinterpret_cast<const float4*>(s_w + (ci * 27 + (kd * 3 + kh) * 3 + kw) * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 wv = wp[q];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float x = in[i * STRIDE + kw];
                acc[i][4 * q + 0] = fmaf(x, wv.x, acc[i][4 * q + 0]);
                acc[i][4 * q + 1] = fmaf(x, wv.y, acc[i][4 * q + 1]);
                acc[i][4 * q + 2] = fmaf(x, wv.z, acc[i][4 * q + 2]);
                acc[i][4 * q + 3] = fmaf(x, wv.w, acc[i][4 * q + 3]);
              }
            }
          }
        }
      }
    }
  }

  // ---- epilogue: bias, channel scale (dropout3d), residual
  const int d = od0 + tz, h = oh0 + ty, w0 = ow0 + tx * 4;
  if (d >= od || h >= oh || w0 >= ow) return;
  const int64_t ospatial = (int64_t)od * oh * ow;
  const bool vec = (ow % 4 == 0);
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    int co = co0 + c;
    if (co >= cout) break;
    float b = dst.bias ? __ldg(dst.bias + co) : 0.f;
    float sc = dst.out_scale ? __ldg(dst.out_scale + co) : 1.f;
    int64_t off = co * ospatial + ((int64_t)d * oh + h) * ow + w0;
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = (acc[i][c] + b) * sc;
    if (vec) {
      if (dst.residual) {
        float4 r = *reinterpret_cast<const float4*>(dst.residual + off);
        o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
      }
      *reinterpret_cast<float4*>(dst.y + off) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      for (int i = 0; i < 4 && w0 + i < ow; ++i)
        dst.y[off + i] = o[i] + (dst.residual ? dst.residual[off + i] : 0.f);
    }
  }
}

template <int STRIDE, int TXN, int CI_T>
static int launch_conv_variant(const ConvSrc& src, const ConvDst& dst, const float* w_packed, int cout, int cout_pad,
                               int in_d, int in_h, int in_w, cudaStream_t st) {
  using T = ConvTile<STRIDE, TXN, CI_T>;
  static_assert(T::SMEM_FLOATS * 4 <= 48 * 1024, "static shared memory limit");
  int od = (in_d + 2 - 3) / STRIDE + 1, oh = (in_h + 2 - 3) / STRIDE + 1, ow = (in_w + 2 - 3) / STRIDE + 1;
  int tiles_d = (od + T::TD - 1) / T::TD, tiles_h = (oh + T::TH - 1) / T::TH, tiles_w = (ow + T::TW - 1) / T::TW;
  dim3 grid(tiles_d * tiles_h * tiles_w, (cout + 15) / 16);
  conv3d_k3_kernel<STRIDE, TXN, CI_T><<<grid, 256, 0, st>>>(src, dst, w_packed, cout, cout_pad, in_d, in_h, in_w, od,
                                                            oh, ow, tiles_h, tiles_w);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_conv3d_k3(const ConvSrc& src, const ConvDst& dst, const float* w_packed, int cout, int cout_pad, int in_d,
                     int in_h, int in_w, int stride, cudaStream_t st) {
  int ow = (in_w - 1) / stride + 1;
  if (stride == 1) {
    if (ow >= 32) return launch_conv_variant<1, 8, 4>(src, dst, w_packed, cout, cout_pad, in_d, in_h, in_w, st);
    return launch_conv_variant<1, 4, 4>(src, dst, w_packed, cout, cout_pad, in_d, in_h, in_w, st);
  }
  if (stride == 2) {
    if (ow >= 32) return launch_conv_variant<2, 8, 1>(src, dst, w_packed, cout, cout_pad, in_d, in_h, in_w, st);
    return launch_conv_variant<2, 4, 1>(src, dst, w_packed, cout, cout_pad, in_d, in_h, in_w, st);
  }
  set_error("conv3d_k3: stride must be 1 or 2");
  return -1;
}

// ---------------------------------------------------------------------------------------------
// Pointwise convolution (two-source concat input), CO_T output channels per block, 4 voxels per
// thread.  SOFTMAX fuses the class softmax of Decoder.forward (cls_wise_former.py:662-663).
// ---------------------------------------------------------------------------------------------
template <int CO_T, bool SOFTMAX>
__global__ void __launch_bounds__(256)
conv1x1_kernel(ConvSrc src, ConvDst dst, const float* __restrict__ w_packed, int cout, int64_t spatial) {
  extern __shared__ float s_w1[];   // [cin][CO_T]
  const int cin = src.c0 + src.c1;
  const int co0 = blockIdx.y * CO_T;
  for (int e = threadIdx.x; e < cin * CO_T; e += 256) {
    int co = e % CO_T, ci = e / CO_T;
    s_w1[e] = (co0 + co < cout) ? __ldg(w_packed + (int64_t)ci * cout + co0 + co) : 0.f;
  }
  __syncthreads();
  const int64_t p = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (p >= spatial) return;
  float acc[4][CO_T];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[i][c] = 0.f;
  for (int ci = 0; ci < cin; ++ci) {
    const float* xp = (ci < src.c0) ? src.x0 + (int64_t)ci * src.s0c + p : src.x1 + (int64_t)(ci - src.c0) * spatial + p;
    float4 v = __ldg(reinterpret_cast<const float4*>(xp));
    float in[4] = {v.x, v.y, v.z, v.w};
    if (src.mean != nullptr) {
      float m = __ldg(src.mean + ci), r = __ldg(src.rstd + ci);
#pragma unroll
      for (int i = 0; i < 4; ++i) in[i] = apply_act((in[i] - m) * r, src.act);
    }
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
      float wv = s_w1[ci * CO_T + c];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][c] = fmaf(in[i], wv, acc[i][c]);
    }
  }
#pragma unroll
  for (int c = 0; c < CO_T; ++c) {
    float b = (dst.bias && co0 + c < cout) ? __ldg(dst.bias + co0 + c) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][c] += b;
  }
  if (SOFTMAX) {   // requires cout == CO_T (4 classes)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float m = acc[i][0];
#pragma unroll
      for (int c = 1; c < CO_T; ++c) m = fmaxf(m, acc[i][c]);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < CO_T; ++c) { acc[i][c] = expf(acc[i][c] - m); s += acc[i][c]; }
#pragma unroll
      for (int c = 0; c < CO_T; ++c) acc[i][c] = acc[i][c] / s;
    }
  }
#pragma unroll
  for (int c = 0; c < CO_T; ++c) {
    if (co0 + c >= cout) break;
    int64_t off = (int64_t)(co0 + c) * spatial + p;
    float4 o = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
    if (dst.residual) {
      float4 r = *reinterpret_cast<const float4*>(dst.residual + off);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    *reinterpret_cast<float4*>(dst.y + off) = o;
  }
}

int launch_conv1x1(const ConvSrc& src, const ConvDst& dst, const float* w_packed, int cout, int64_t spatial,
                   bool softmax, cudaStream_t st) {
  if (spatial % 4 != 0) { set_error("conv1x1: spatial size must be a multiple of 4"); return -1; }
  int cin = src.c0 + src.c1;
  unsigned gx = (unsigned)((spatial / 4 + 255) / 256);
  if (softmax) {
    if (cout != 4) { set_error("conv1x1: fused softmax needs 4 output channels"); return -1; }
    conv1x1_kernel<4, true><<<dim3(gx, 1), 256, cin * 4 * sizeof(float), st>>>(src, dst, w_packed, cout, spatial);
  } else {
    conv1x1_kernel<16, false><<<dim3(gx, (cout + 15) / 16), 256, cin * 16 * sizeof(float), st>>>(src, dst, w_packed,
                                                                                                 cout, spatial);
  }
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose3d kernel 2 stride 2 (DeUp_Cat.conv2, cls_wise_former.py:720): every input voxel
// produces a disjoint 2x2x2 output block.  Thread = one input voxel, one (kd,kh) pair, both kw,
// 16 output channels -> float2 stores that are contiguous across the warp.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
convt_k2s2_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w_packed,
                  const float* __restrict__ bias, int cin, int cout, int in_d, int in_h, int in_w) {
  extern __shared__ float s_wt[];   // [cin][2][16]
  const int kdh = blockIdx.y & 3;    // kd*2 + kh
  const int co0 = (blockIdx.y >> 2) * 16;
  for (int e = threadIdx.x; e < cin * 32; e += 256) {
    int co = e % 16, kw = (e / 16) % 2, ci = e / 32;
    s_wt[e] = (co0 + co < cout) ? __ldg(w_packed + ((int64_t)ci * 8 + kdh * 2 + kw) * cout + co0 + co) : 0.f;
  }
  __syncthreads();
  const int64_t spatial = (int64_t)in_d * in_h * in_w;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= spatial) return;
  const int w = p % in_w;
  const int h = (p / in_w) % in_h;
  const int d = p / ((int64_t)in_w * in_h);
  float acc[2][16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[0][c] = acc[1][c] = 0.f;
  for (int ci = 0; ci < cin; ++ci) {
    float v = __ldg(x + ci * spatial + p);
    const float* wp = s_wt + ci * 32;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      acc[0][c] = fmaf(v, wp[c], acc[0][c]);
      acc[1][c] = fmaf(v, wp[16 + c], acc[1][c]);
    }
  }
  const int od = 2 * in_d, oh = 2 * in_h, ow = 2 * in_w;
  const int kd = kdh >> 1, kh = kdh & 1;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    if (co0 + c >= cout) break;
    float b = bias ? __ldg(bias + co0 + c) : 0.f;
    int64_t off = (((int64_t)(co0 + c) * od + 2 * d + kd) * oh + 2 * h + kh) * ow + 2 * w;
    *reinterpret_cast<float2*>(y + off) = make_float2(acc[0][c] + b, acc[1][c] + b);
  }
}

int launch_convt_k2s2(const float* x, float* y, const float* w_packed, const float* bias, int cin, int cout, int in_d,
                      int in_h, int in_w, cudaStream_t st) {
  int64_t spatial = (int64_t)in_d * in_h * in_w;
  dim3 grid((unsigned)((spatial + 255) / 256), 4 * ((cout + 15) / 16));
  convt_k2s2_kernel<<<grid, 256, cin * 32 * sizeof(float), st>>>(x, y, w_packed, bias, cin, cout, in_d, in_h, in_w);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dcl
