// Micro-benchmark: SM clocks per tcgen05.mma (M = 128, K = 16, 16-bit operands) as a function of N, of the alignment
// of the A start address (a 3x3x3 tap shifted by one voxel = 16 bytes in the no-swizzle K-major layout) and of the A
// layout (no swizzle vs 128-byte swizzle with one voxel per 128-byte row, where a voxel shift is a whole row).
// Timing only: the operands are zeros.  Run on the GPU box: tools/ubench/mma_shape
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200/csrc/tc_common.cuh"

using namespace dcl::tc;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t base_off) {   // K-major, SWIZZLE_128B: SBO = 1024 B
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1) bench(int n, int iters, int a_off_bytes, int sw128, int f16, long long* out, int masked = 0,
                                                int a_stride16 = 8) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 1) {
    const uint32_t sb = smem_u32(smem);
    const uint32_t idesc = umma_idesc_16(128, n, f16 != 0);
    const uint32_t a_addr = sb + 4096 + (uint32_t)a_off_bytes;
    const uint64_t a0 = sw128 ? desc_sw128(a_addr, (a_addr >> 7) & 7) : umma_desc(a_addr, 20800, 128);
    const uint64_t b0 = umma_desc(sb + 128 * 1024, (uint32_t)n * 16, 128);
    long long t0 = clock64();
    if (masked) {      // the disable-output-lane form the kw = 0 / 2 taps of the rolling kernels use
      for (int i = 0; i < iters; ++i)
        umma_bf16_masked_ws(tmem + (uint32_t)((i & 1) * 256), a0 + (uint64_t)((i & 7) * a_stride16), b0, idesc, 1u, 1u, 0u, 0u, 0u);
    } else {
      for (int i = 0; i < iters; ++i) umma_bf16_ws(tmem + (uint32_t)((i & 1) * 256), a0 + (uint64_t)((i & 7) * (sw128 ? 64 : a_stride16)), b0, idesc, 1u);
    }
    umma_commit_ws(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  const int ns[] = {16, 32, 48, 64, 96, 128, 144, 192, 256};
  for (int sw = 0; sw < 2; ++sw)
    for (int off = 0; off < 2; ++off)
      for (int n : ns) {
        const int a_off = off ? (sw ? 128 : 16) : 0;      // one voxel: 16 bytes (no swizzle) or one 128-byte row (swizzle)
        for (int rep = 0; rep < 2; ++rep) bench<<<148, 128, smem>>>(n, iters, a_off, sw, 1, d);
        long long h = 0;
        cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s (sw128=%d off=%d n=%d)\n", cudaGetErrorString(e), sw, a_off, n); return 1; }
        printf("A %-10s start+%3d B  N=%3d : %7.1f clk/MMA  (%5.1f %% of the 8192 flop/clk peak)\n", sw ? "swizzle128" : "no-swizzle", a_off, n,
               (double)h / iters, 100.0 * (2.0 * 128 * n * 16) / ((double)h / iters) / 8192.0);
      }
  // masked vs unmasked, operand strides of the rolling kernel (a row = 128 positions), N = 48 / 96
  for (int masked = 0; masked < 2; ++masked)
    for (int off = 0; off < 2; ++off)
      for (int n : {48, 96}) {
        for (int rep = 0; rep < 2; ++rep) bench<<<148, 128, smem>>>(n, iters, off * 16, 0, 1, d, masked, 128);
        long long h = 0;
        if (cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
        printf("%-8s A start+%2d B, row stride 2 KB, N=%3d : %7.1f clk/MMA\n", masked ? "masked" : "unmasked", off * 16, n, (double)h / iters);
      }
  return 0;
}
