// How long does the CTA launcher take to start N big-shared-memory CTAs, and what is the kernel-to-kernel gap?
#include <cstdio>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(448, 1) k(unsigned long long* out, int spin) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0) out[2 * blockIdx.x] = gtime();
  smem[threadIdx.x] = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    out[2 * blockIdx.x + 1] = gtime();
  }
}
int main() {
  const int reps = 6;
  for (int smem : {16 * 1024, 200 * 1024}) {
    for (int ctas : {128, 148, 296}) {
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      unsigned long long* d;
      cudaMalloc(&d, reps * 2 * ctas * 8);
      for (int r = 0; r < reps; ++r) k<<<ctas, 448, smem>>>(d + r * 2 * ctas, 10000);
      cudaDeviceSynchronize();
      std::vector<unsigned long long> h(reps * 2 * ctas);
      cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
      for (int r = 1; r < reps; ++r) {
        unsigned long long* p = h.data() + r * 2 * ctas;
        unsigned long long first = ~0ull, last_start = 0, last_end = 0, prev_end = 0;
        for (int c = 0; c < ctas; ++c) { first = std::min(first, p[2 * c]); last_start = std::max(last_start, p[2 * c]); last_end = std::max(last_end, p[2 * c + 1]); }
        unsigned long long* q = h.data() + (r - 1) * 2 * ctas;
        for (int c = 0; c < ctas; ++c) prev_end = std::max(prev_end, q[2 * c + 1]);
        printf("smem %3d KB ctas %3d: gap after previous kernel %6.2f us, start stagger %6.2f us, kernel span %6.2f us\n", smem / 1024, ctas,
               (double)((long long)first - (long long)prev_end) / 1e3, (last_start - first) / 1e3, (last_end - first) / 1e3);
      }
      cudaFree(d);
    }
  }
  return 0;
}
