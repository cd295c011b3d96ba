// Micro-benchmark: issue rate / execution time of tcgen05.mma kind::f16 (M=128, K=16) with A and B in
// shared memory (K-major, no swizzle), for several N and issue styles.  One CTA per SM is launched on
// `nctas` SMs so per-SM contention is visible.  Prints SM clocks per MMA.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200/csrc/tc_common.cuh"

using namespace dcl::tc;

template <int STYLE>
__global__ void __launch_bounds__(128, 1) bench(int n, int iters, uint32_t a_lbo, uint32_t a_step, uint32_t b_step, int spin, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
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
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint64_t a0 = umma_desc(sb, a_lbo, 128);
    const uint64_t b0 = umma_desc(sb + 128 * 1024, (uint32_t)n * 16, 128);
    long long t0 = 0, t1 = 0;
    if (STYLE == 0) {          // single thread, loop-carried descriptors (divergent region)
      if (lane == 0) {
        t0 = clock64();
        uint64_t ad = a0, bd = b0;
        for (int i = 0; i < iters; ++i) {
          umma_bf16(tmem + (uint32_t)((i & 1) * 64), ad, bd, idesc, 1u);
          ad += a_step; bd += b_step;
          if ((i & 3) == 3) { ad = a0; bd = b0; }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
      }
    } else {                   // whole warp, elect inside the asm
      t0 = clock64();
      uint64_t ad = a0, bd = b0;
      for (int i = 0; i < iters; ++i) {
        umma_bf16_ws(tmem + (uint32_t)((i & 1) * 64), ad, bd, idesc, 1u);
        ad += a_step; bd += b_step;
        if ((i & 3) == 3) { ad = a0 + (uint64_t)(i & 31); bd = b0; }
      }
      umma_commit_ws(&bar);
      mbar_wait(&bar, 0);
      t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 2 && spin) {
    // idle roles polling an mbarrier that never completes during the measurement
    __shared__ uint64_t never;
    if (threadIdx.x == 64) mbar_init(&never, 1);
    __syncwarp();
    for (int i = 0; i < spin; ++i) mbar_try_wait(&never, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  struct Cfg { int n; uint32_t lbo, astep, bstep; int spin; const char* what; };
  Cfg cfgs[] = {
      {64, 15392, 0, 0, 0, "same A,B"},
      {64, 15392, 1924, 128, 0, "slab strides (a += 2*962, b += 2*64)"},
      {64, 15392, 1924, 128, 4000, "slab strides + 2 polling warps"},
      {64, 15392, 1924, 0, 0, "A strided only"},
      {64, 15392, 0, 128, 0, "B strided only"},
      {64, 15424, 1928, 128, 0, "npos 964 (LBO = 64 mod 128)"},
      {32, 15392, 1924, 64, 0, "N=32 slab strides"},
      {48, 20800, 0, 0, 0, "K1-like N=48"},
  };
  for (auto& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) bench<1><<<148, 128, smem>>>(c.n, iters, c.lbo, c.astep, c.bstep, c.spin, d);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    printf("N=%3d lbo=%5u %-45s : %7.1f clk/MMA\n", c.n, c.lbo, c.what, (double)h / iters);
  }
  return 0;
}
