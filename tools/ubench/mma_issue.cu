// Micro-benchmark of MMA *issue* code shapes (the single-thread instruction stream around tcgen05.mma).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200/csrc/tc_common.cuh"
using namespace dcl::tc;

struct P { int mt, nks, n_tile, npos, W, R, nb, b_stage, b_tap_bytes, slab_bytes, stages; };

template <int SHAPE>
__global__ void __launch_bounds__(128, 1) bench(P p, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar, bfull[8];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bfull[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) mbar_arrive(&bfull[i]);   // phase 0 complete for every stage
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  if (warp == 1) {
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t idesc = umma_idesc_bf16(128, p.n_tile);
    const uint32_t a_lbo = (uint32_t)p.npos * 16, b_lbo = (uint32_t)p.n_tile * 16;
    const uint64_t a_desc0 = umma_desc(smem_base, a_lbo, 128);
    const uint64_t b_desc0 = umma_desc(smem_base + (uint32_t)p.slab_bytes, b_lbo, 128);
    const uint32_t a_ks = 2u * (uint32_t)p.npos, b_ks = 2u * (uint32_t)p.n_tile;
    const uint32_t a_tile = 128;
    const int W = p.W, R = p.R;
    uint32_t mk0[4] = {1, 1, 1, 1}, mk2[4] = {0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u};
    long long t0 = clock64();
    int bit = 0;
    for (int rep = 0; rep < p.stages / 9; ++rep)
      for (int kdh = 0; kdh < 9; ++kdh, ++bit) {
        const int kd = kdh / 3, kh = kdh - kd * 3;
        const uint32_t pos_row = (uint32_t)(1 + (kd * (R + 2) + kh) * W);
        const int s = bit % p.nb;
        mbar_wait(&bfull[s], 0);
        tc_fence_after();
        if (SHAPE == 0) {        // the slab kernel's loop as written
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int kw = i == 0 ? 1 : (i == 1 ? 0 : 2);
            const uint64_t b_tap = b_desc0 + (uint64_t)((uint32_t)(s * p.b_stage + i * p.b_tap_bytes) >> 4);
            const uint64_t a_tap = a_desc0 + (uint64_t)(pos_row + (uint32_t)(kw - 1));
            const uint32_t q0 = kw == 0 ? mk0[0] : (kw == 2 ? mk2[0] : 0u), q1 = kw == 0 ? mk0[1] : (kw == 2 ? mk2[1] : 0u);
            const uint32_t q2 = kw == 0 ? mk0[2] : (kw == 2 ? mk2[2] : 0u), q3 = kw == 0 ? mk0[3] : (kw == 2 ? mk2[3] : 0u);
            const uint32_t accum = (rep | kdh | i) != 0 ? 1u : 0u;
            for (int t = 0; t < p.mt; ++t) {
              const uint32_t d_tmem = tmem_base + (uint32_t)(t * p.n_tile);
              uint64_t ad = a_tap + (uint64_t)((uint32_t)t * a_tile), bd = b_tap;
              uint32_t acc_t = accum;
              for (int ks = 0; ks < p.nks; ++ks) {
                if (kw == 1) umma_bf16_ws(d_tmem, ad, bd, idesc, acc_t);
                else umma_bf16_masked_ws(d_tmem, ad, bd, idesc, acc_t, q0, q1, q2, q3);
                ad += a_ks; bd += b_ks; acc_t = 1u;
              }
            }
          }
        } else if (SHAPE == 1) {  // 32-bit descriptor arithmetic: only the low word (start address) changes
          const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
          const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int kw = i == 0 ? 1 : (i == 1 ? 0 : 2);
            const uint32_t b_tap = b_lo0 + ((uint32_t)(s * p.b_stage + i * p.b_tap_bytes) >> 4);
            const uint32_t a_tap = a_lo0 + pos_row + (uint32_t)(kw - 1);
            const uint32_t q0 = kw == 0 ? mk0[0] : (kw == 2 ? mk2[0] : 0u), q1 = kw == 0 ? mk0[1] : (kw == 2 ? mk2[1] : 0u);
            const uint32_t q2 = kw == 0 ? mk0[2] : (kw == 2 ? mk2[2] : 0u), q3 = kw == 0 ? mk0[3] : (kw == 2 ? mk2[3] : 0u);
            for (int t = 0; t < p.mt; ++t) {
              const uint32_t d_tmem = tmem_base + (uint32_t)(t * p.n_tile);
              uint32_t ad = a_tap + (uint32_t)t * a_tile, bd = b_tap;
              for (int ks = 0; ks < p.nks; ++ks) {
                const uint64_t ad64 = ((uint64_t)a_hi << 32) | ad, bd64 = ((uint64_t)b_hi << 32) | bd;
                umma_bf16_masked_ws(d_tmem, ad64, bd64, idesc, 1u, q0, q1, q2, q3);
                ad += a_ks; bd += b_ks;
              }
            }
          }
        }
      }
    umma_commit_ws(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  P p;
  p.mt = 2; p.nks = 4; p.n_tile = 64; p.W = 32; p.R = 8; p.npos = 962; p.nb = 3;
  p.b_tap_bytes = 8 * 64 * 16; p.b_stage = 3 * p.b_tap_bytes; p.slab_bytes = 8 * 962 * 16; p.stages = 27;
  const int mmas = p.stages * 3 * p.mt * p.nks;
  for (int shape = 0; shape < 2; ++shape) {
    for (int rep = 0; rep < 2; ++rep) {
      if (shape == 0) bench<0><<<148, 128, smem>>>(p, d);
      else bench<1><<<148, 128, smem>>>(p, d);
    }
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    printf("shape=%d : %7.1f clk/MMA (%d MMAs)\n", shape, (double)h / mmas, mmas);
  }
  return 0;
}
