// Auxiliary heads tail: F.interpolate(scale_factor=s, mode='trilinear', align_corners=False) followed
// by a 2-class softmax (SuperviseLabel.py:62-64, EdgeSuperviseLabel.py:59-60), fused so the (2,128^3)
// upsampled logits never touch HBM.
#include "common.cuh"

namespace dcl {

__device__ __forceinline__ void src_index(int dst, float inv_scale, int size, int& i0, int& i1, float& l1) {
  // area_pixel_compute_source_index with align_corners=False: max(0, (dst+0.5)/scale - 0.5)
  float s = ((float)dst + 0.5f) * inv_scale - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  i1 = i0 + (i0 < size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

__global__ void __launch_bounds__(256)
upsample_softmax2_kernel(const float* __restrict__ x, float* __restrict__ y_fixed, int g, int scale,
                         const PatchDesc* __restrict__ desc, int slot) {
  float* __restrict__ y = desc != nullptr ? desc->aux[slot] : y_fixed;
  const int og = g * scale;
  const int64_t n = (int64_t)og * og * og;
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  const int w = e % og;
  const int h = (e / og) % og;
  const int d = e / ((int64_t)og * og);
  const float inv = 1.f / (float)scale;
  int d0, d1, h0, h1, w0, w1;
  float ld, lh, lw;
  src_index(d, inv, g, d0, d1, ld);
  src_index(h, inv, g, h0, h1, lh);
  src_index(w, inv, g, w0, w1, lw);
  float v[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float* p = x + (int64_t)c * g * g * g;
    auto at = [&](int dd, int hh, int ww) { return __ldg(p + ((int64_t)dd * g + hh) * g + ww); };
    float a = (1.f - ld) * ((1.f - lh) * ((1.f - lw) * at(d0, h0, w0) + lw * at(d0, h0, w1)) +
                            lh * ((1.f - lw) * at(d0, h1, w0) + lw * at(d0, h1, w1)));
    float b = ld * ((1.f - lh) * ((1.f - lw) * at(d1, h0, w0) + lw * at(d1, h0, w1)) +
                    lh * ((1.f - lw) * at(d1, h1, w0) + lw * at(d1, h1, w1)));
    v[c] = a + b;
  }
  float m = fmaxf(v[0], v[1]);
  float e0 = expf(v[0] - m), e1 = expf(v[1] - m);
  float s = e0 + e1;
  y[e] = e0 / s;
  y[n + e] = e1 / s;
}

int launch_upsample_softmax2(const float* x, float* y, int g, int scale, cudaStream_t st, const PatchDesc* desc, int slot) {
  int64_t n = (int64_t)g * scale * g * scale * g * scale;
  upsample_softmax2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, g, scale, desc, slot);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dcl
