// HBM-bound tail of the sliding window: crop-and-overwrite stitch (predict_overlap.py:49-56), weighted
// overlap accumulate (extension modes), and the fused normalise + arg-max + label histogram + Dice
// counters kernel (predict_overlap.py:141-153, utils/tools.py:89-109).
#include "common.cuh"

namespace dcl {

constexpr int P = 128;
constexpr int64_t P3 = (int64_t)P * P * P;

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float ld_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// out[c, dst+e] = probs[c, src+e] for e in the box.  Thread = one (x,y,z) of the box, all 4 classes;
// z is contiguous in both tensors so warps read and write full sectors.
__global__ void __launch_bounds__(256)
stitch_copy_kernel(const float* __restrict__ probs, float* __restrict__ out, StitchBox box, int Y, int Zout,
                   int64_t out_plane) {
  const int64_t n = (int64_t)box.ext[0] * box.ext[1] * box.ext[2];
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  const int z = e % box.ext[2];
  const int y = (e / box.ext[2]) % box.ext[1];
  const int x = e / ((int64_t)box.ext[2] * box.ext[1]);
  const int64_t s = ((int64_t)(box.src[0] + x) * P + (box.src[1] + y)) * P + box.src[2] + z;
  const int64_t d = ((int64_t)(box.dst[0] + x) * Y + (box.dst[1] + y)) * Zout + box.dst[2] + z;
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c * out_plane + d] = __ldg(probs + c * P3 + s);
}

int launch_stitch_copy(const float* probs, float* out, const StitchBox& box, int X, int Y, int Zout, cudaStream_t st) {
  int64_t n = (int64_t)box.ext[0] * box.ext[1] * box.ext[2];
  if (n <= 0) return 0;
  stitch_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(probs, out, box, Y, Zout, (int64_t)X * Y * Zout);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// acc[c, start+e] += w(e) * probs[c, e];  wsum[start+e] += w(e).  Patches of one volume are issued in
// order on one stream, so the read-modify-write needs no atomics.
__device__ __forceinline__ float blend_w1(int i, int gaussian) {
  if (!gaussian) return 1.f;
  float t = ((float)i - 63.5f) / 16.f;      // centre (P-1)/2, sigma P/8
  return expf(-0.5f * (t * t));
}

__global__ void __launch_bounds__(256)
accumulate_kernel(const float* __restrict__ probs, int sx, int sy, int sz, int gaussian, float* __restrict__ acc,
                  float* __restrict__ wsum, int Y, int Z, int64_t plane) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= P3) return;
  const int z = e % P;
  const int y = (e / P) % P;
  const int x = e / (P * P);
  const float w = (blend_w1(x, gaussian) * blend_w1(y, gaussian)) * blend_w1(z, gaussian);
  const int64_t d = ((int64_t)(sx + x) * Y + (sy + y)) * Z + sz + z;
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c * plane + d] += w * __ldg(probs + c * P3 + e);
  wsum[d] += w;
}

// 128-bit variant: a warp owns one z-row of the patch (128 voxels).  The accumulator row starts at an arbitrary
// 4-byte offset (Z = 155 is odd), so the warp walks the 32-33 ALIGNED float4 vectors that cover it; lanes whose vector
// sticks out of the patch add zero to the outside elements (no other warp of the launch touches those: patch rows of
// neighbouring (x,y) are Z >= 132 floats apart).  Per thread: 5 x 128-bit read-modify-write + 16 cached scalar loads
// of the (unaligned) patch probabilities, against 9 x 32-bit accesses per voxel in the scalar kernel.
__global__ void __launch_bounds__(256)
accumulate_vec_kernel(const float* __restrict__ probs, int sx, int sy, int sz, int gaussian, float* __restrict__ acc,
                      float* __restrict__ wsum, int Y, int Z, int64_t plane) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);      // (x, y) of the patch
  const int lane = threadIdx.x & 31;
  const int x = row >> 7, y = row & 127;
  const float wxy = blend_w1(x, gaussian) * blend_w1(y, gaussian);
  const int64_t d0 = ((int64_t)(sx + x) * Y + (sy + y)) * Z + sz;
  const int a = (int)(d0 & 3);
  const int64_t a0 = d0 - a;
  const float* prow = probs + (int64_t)row * P;
  for (int j = lane; 4 * j < a + P; j += 32) {              // 32 vectors, plus one more when the row is unaligned
    const int z0 = 4 * j - a;
    float w[4], p[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int z = z0 + k;
      const bool in = (unsigned)z < (unsigned)P;
      w[k] = in ? wxy * blend_w1(z, gaussian) : 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) p[c][k] = in ? __ldg(prow + c * P3 + z) : 0.f;
    }
    const int64_t v = a0 + 4 * j;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 t = *reinterpret_cast<float4*>(acc + c * plane + v);
      t.x += w[0] * p[c][0]; t.y += w[1] * p[c][1]; t.z += w[2] * p[c][2]; t.w += w[3] * p[c][3];
      *reinterpret_cast<float4*>(acc + c * plane + v) = t;
    }
    float4 t = *reinterpret_cast<float4*>(wsum + v);
    t.x += w[0]; t.y += w[1]; t.z += w[2]; t.w += w[3];
    *reinterpret_cast<float4*>(wsum + v) = t;
  }
}

int launch_accumulate(const float* probs, const int start[3], int gaussian, float* acc, float* wsum, int X, int Y,
                      int Z, cudaStream_t st) {
  const int64_t plane = (int64_t)X * Y * Z;
  const bool aligned = (plane & 3) == 0 && ((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(wsum)) & 15) == 0;
  // the vector kernel may rewrite (unchanged) up to 3 floats on either side of a row: they must exist and must not be
  // another row's patch region, i.e. rows at least 132 apart and the patch not at the very first / last floats
  const int64_t first = ((int64_t)start[0] * Y + start[1]) * Z + start[2];
  const int64_t last = ((int64_t)(start[0] + P - 1) * Y + (start[1] + P - 1)) * Z + start[2] + P;
  if (aligned && Z >= P + 4 && (first & ~(int64_t)3) >= 0 && ((last + 3) & ~(int64_t)3) <= plane) {
    accumulate_vec_kernel<<<P * P / 8, 256, 0, st>>>(probs, start[0], start[1], start[2], gaussian, acc, wsum, Y, Z, plane);
  } else {
    accumulate_kernel<<<(unsigned)((P3 + 255) / 256), 256, 0, st>>>(probs, start[0], start[1], start[2], gaussian, acc,
                                                                   wsum, Y, Z, plane);
  }
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// finalize: p_c = acc_c / wsum (or acc_c), label = argmax_c p_c (first maximum wins, as numpy),
// 13 counters: histogram[4], then (|o|,|t|,|o&t|) for WT (>0), TC ({1,3}), ET (==3).
// Persistent grid-stride kernel, 4 voxels per thread per step: four independent 128-bit streaming loads
// (one per class plane) + one 32-bit label store = 17 B / voxel of algorithmic traffic.
// ---------------------------------------------------------------------------------------------
struct LabelCounts {
  unsigned v[13];
};

__device__ __forceinline__ int argmax4(float a, float b, float c, float d) {
  int l = 0;
  float m = a;
  if (b > m) { m = b; l = 1; }
  if (c > m) { m = c; l = 2; }
  if (d > m) { l = 3; }
  return l;
}

__device__ __forceinline__ void count_label(LabelCounts& k, int l, int t, bool has_target) {
  k.v[0] += l == 0; k.v[1] += l == 1; k.v[2] += l == 2; k.v[3] += l == 3;     // no dynamic index: stays in registers
  const bool owt = l > 0, otc = (l == 1) | (l == 3), oet = l == 3;
  k.v[4] += owt; k.v[7] += otc; k.v[10] += oet;
  if (has_target) {
    const bool twt = t > 0, ttc = (t == 1) | (t == 3), tet = t == 3;
    k.v[5] += twt; k.v[6] += owt & twt;
    k.v[8] += ttc; k.v[9] += otc & ttc;
    k.v[11] += tet; k.v[12] += oet & tet;
  }
}

__global__ void __launch_bounds__(256)
finalize_labels_kernel(const float* __restrict__ acc, const float* __restrict__ wsum, int64_t total, int64_t v0,
                       int64_t nvox, float* __restrict__ probs_out, uint8_t* __restrict__ labels,
                       const uint8_t* __restrict__ target, unsigned long long* __restrict__ counts) {
  LabelCounts k;
#pragma unroll
  for (int i = 0; i < 13; ++i) k.v[i] = 0;
  // 128-bit path: plane stride, range start AND every base pointer must keep 16-byte (4-byte for the byte maps) alignment
  const bool vec_ok = ((total | v0) & 3) == 0 && (reinterpret_cast<uintptr_t>(acc) & 15) == 0 &&
                      (wsum == nullptr || (reinterpret_cast<uintptr_t>(wsum) & 15) == 0) &&
                      (probs_out == nullptr || (reinterpret_cast<uintptr_t>(probs_out) & 15) == 0) &&
                      (labels == nullptr || (reinterpret_cast<uintptr_t>(labels) & 3) == 0) &&
                      (target == nullptr || (reinterpret_cast<uintptr_t>(target) & 3) == 0);
  const int64_t nvec = vec_ok ? nvox / 4 : 0;
  for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < nvec; g += (int64_t)gridDim.x * 256) {
    const int64_t v = v0 + g * 4;
    float4 a = ld_stream4(acc + v), b = ld_stream4(acc + total + v), c = ld_stream4(acc + 2 * total + v),
           d = ld_stream4(acc + 3 * total + v);
    if (wsum) {
      float4 w = ld_stream4(wsum + v);
      a.x /= w.x; b.x /= w.x; c.x /= w.x; d.x /= w.x;
      a.y /= w.y; b.y /= w.y; c.y /= w.y; d.y /= w.y;
      a.z /= w.z; b.z /= w.z; c.z /= w.z; d.z /= w.z;
      a.w /= w.w; b.w /= w.w; c.w /= w.w; d.w /= w.w;
    }
    if (probs_out) {
      *reinterpret_cast<float4*>(probs_out + v) = a;
      *reinterpret_cast<float4*>(probs_out + total + v) = b;
      *reinterpret_cast<float4*>(probs_out + 2 * total + v) = c;
      *reinterpret_cast<float4*>(probs_out + 3 * total + v) = d;
    }
    const int l0 = argmax4(a.x, b.x, c.x, d.x), l1 = argmax4(a.y, b.y, c.y, d.y), l2 = argmax4(a.z, b.z, c.z, d.z),
              l3 = argmax4(a.w, b.w, c.w, d.w);
    if (labels) *reinterpret_cast<uchar4*>(labels + v) = make_uchar4(l0, l1, l2, l3);
    if (counts) {
      uchar4 t = target ? *reinterpret_cast<const uchar4*>(target + v) : make_uchar4(0, 0, 0, 0);
      count_label(k, l0, t.x, target != nullptr);
      count_label(k, l1, t.y, target != nullptr);
      count_label(k, l2, t.z, target != nullptr);
      count_label(k, l3, t.w, target != nullptr);
    }
  }
  // scalar tail (and the whole range when the slab is not 4-aligned)
  for (int64_t g = nvec * 4 + (int64_t)blockIdx.x * 256 + threadIdx.x; g < nvox; g += (int64_t)gridDim.x * 256) {
    const int64_t v = v0 + g;
    float a = acc[v], b = acc[total + v], c = acc[2 * total + v], d = acc[3 * total + v];
    if (wsum) { float w = wsum[v]; a /= w; b /= w; c /= w; d /= w; }
    if (probs_out) { probs_out[v] = a; probs_out[total + v] = b; probs_out[2 * total + v] = c; probs_out[3 * total + v] = d; }
    const int l = argmax4(a, b, c, d);
    if (labels) labels[v] = (uint8_t)l;
    if (counts) count_label(k, l, target ? target[v] : 0, target != nullptr);
  }
  if (counts) {
    __shared__ unsigned s_cnt[13];
    if (threadIdx.x < 13) s_cnt[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 13; ++i) {
      unsigned r = __reduce_add_sync(0xffffffffu, k.v[i]);
      if ((threadIdx.x & 31) == 0 && r) atomicAdd(&s_cnt[i], r);
    }
    __syncthreads();
    if (threadIdx.x < 13 && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
  }
}

int launch_finalize_labels(const float* acc, const float* wsum, int64_t total, int64_t v0, int64_t nvox,
                           float* probs_out, uint8_t* labels, const uint8_t* target, unsigned long long* counts,
                           cudaStream_t st) {
  if (nvox <= 0) return 0;
  int64_t work = (nvox + 3) / 4;
  int64_t blocks = (work + 255) / 256;
  const int64_t persistent = 148 * 8;   // one wave of 8 resident CTAs per SM
  if (blocks > persistent) blocks = persistent;
  finalize_labels_kernel<<<(unsigned)blocks, 256, 0, st>>>(acc, wsum, total, v0, nvox, probs_out, labels, target,
                                                          counts);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Gather form of the weighted stitch (extension modes): every patch keeps its probabilities in its own slot of
// `pp` (P x 4 x 128^3) and ONE kernel forms, per output voxel, sum_i w_i p_i / sum_i w_i over the patches that
// cover it, the arg-max label and the counters.  Against accumulate + finalize this drops the accumulator round
// trip: P x 33.5 MB read + 1 B / voxel written (613 MB for 18 patches) instead of P x 117 MB + 196 MB (2.3 GB).
// The sums run in patch order with the same FMA as accumulate_kernel, so the result is bit-identical to it.
// A WARP owns one (x,y) row of the volume at a time (persistent grid striding over the rows): it lists the patches
// that cover (x,y) with one ballot per 32 plan entries (patch order preserved), then lane l handles z = l, l+32, ...
// testing only the z range; loads are coalesced along z in every patch slot, four patches (16 loads) in flight per
// lane.  Counters stay in registers until the CTA's single flush.
// ---------------------------------------------------------------------------------------------
constexpr int GATHER_WARPS = 8;
__global__ void __launch_bounds__(GATHER_WARPS * 32)
gather_finalize_kernel(GatherSlots slots, GatherPlan plan, int gaussian, int row0, int row1, int rows, int Y, int Z,
                       float* __restrict__ probs_out, uint8_t* __restrict__ labels, const uint8_t* __restrict__ target,
                       unsigned long long* __restrict__ counts) {
  // this launch covers the (x,y) rows [row0, row1) of the volume; `rows` = X * Y of the whole volume (plane stride)
  __shared__ const float* s_row[GATHER_WARPS][GatherPlan::MAX];   // address of (x, y, z = 0) in the covering patch's slot
                                                                   // (local or peer memory; may point before the slot: z >= sz only)
  __shared__ short s_sz[GATHER_WARPS][GatherPlan::MAX];     // z origin of the patch
  __shared__ float s_wxy[GATHER_WARPS][GatherPlan::MAX];
  __shared__ unsigned s_cnt[13];
  __shared__ float s_wz[P];                                 // blend_w1 along z, computed once (same expression, same bits)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x < 13) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x < P) s_wz[threadIdx.x] = blend_w1(threadIdx.x, gaussian);
  __syncthreads();
  const int64_t plane = (int64_t)rows * Z;          // voxels of the volume
  LabelCounts k;
#pragma unroll
  for (int i = 0; i < 13; ++i) k.v[i] = 0;
  for (int row = row0 + blockIdx.x * GATHER_WARPS + w; row < row1; row += gridDim.x * GATHER_WARPS) {
    const int x = row / Y, y = row - x * Y;
    int n = 0;
    for (int i0 = 0; i0 < plan.n; i0 += 32) {
      const int i = i0 + lane;
      bool hit = false;
      int lx = 0, ly = 0;
      if (i < plan.n) {
        lx = x - plan.start[i][0]; ly = y - plan.start[i][1];
        hit = (unsigned)lx < (unsigned)P && (unsigned)ly < (unsigned)P;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int j = n + __popc(m & ((1u << lane) - 1u));
        s_row[w][j] = slots.ptr[i] + ((lx * P + ly) * P - plan.start[i][2]);
        s_sz[w][j] = (short)plan.start[i][2];
        s_wxy[w][j] = blend_w1(lx, gaussian) * blend_w1(ly, gaussian);
      }
      n += __popc(m);
    }
    __syncwarp();
    for (int z = lane; z < Z; z += 32) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, ws = 0.f;
      // four patches per trip, branch-free: a patch that does not cover z contributes w = 0, p = 0 (fma(0, 0, a) == a,
      // so the sums keep accumulate_kernel's bits); the 16 loads of a trip are in flight before the first FMA
      for (int j0 = 0; j0 < n; j0 += 4) {
        float wt[4], p[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          wt[u] = 0.f;
          p[u][0] = p[u][1] = p[u][2] = p[u][3] = 0.f;
          if (j < n) {
            const int lz = z - s_sz[w][j];
            if ((unsigned)lz < (unsigned)P) {
              wt[u] = s_wxy[w][j] * s_wz[lz];
              const float* q = s_row[w][j] + z;
              p[u][0] = ld_stream1(q); p[u][1] = ld_stream1(q + P3); p[u][2] = ld_stream1(q + 2 * P3); p[u][3] = ld_stream1(q + 3 * P3);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a0 = __fmaf_rn(wt[u], p[u][0], a0); a1 = __fmaf_rn(wt[u], p[u][1], a1);
          a2 = __fmaf_rn(wt[u], p[u][2], a2); a3 = __fmaf_rn(wt[u], p[u][3], a3);
          ws += wt[u];
        }
      }
      a0 /= ws; a1 /= ws; a2 /= ws; a3 /= ws;
      const int64_t v = (int64_t)row * Z + z;
      if (probs_out) { probs_out[v] = a0; probs_out[plane + v] = a1; probs_out[2 * plane + v] = a2; probs_out[3 * plane + v] = a3; }
      const int l = argmax4(a0, a1, a2, a3);
      if (labels) labels[v] = (uint8_t)l;
      if (counts) count_label(k, l, target ? target[v] : 0, target != nullptr);
    }
    __syncwarp();
  }
  if (counts) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 13; ++i) {
      unsigned r = __reduce_add_sync(0xffffffffu, k.v[i]);
      if (lane == 0 && r) atomicAdd(&s_cnt[i], r);
    }
    __syncthreads();
    if (threadIdx.x < 13 && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
  }
}

int launch_gather_finalize_slots(const GatherSlots& slots, const GatherPlan& plan, int gaussian, int X, int Y, int Z, int x0,
                                 int x1, float* probs_out, uint8_t* labels, const uint8_t* target, unsigned long long* counts,
                                 cudaStream_t st) {
  if (plan.n < 1 || plan.n > GatherPlan::MAX || plan.slot_planes < 4 || x0 < 0 || x1 > X || x0 > x1) {
    set_error("gather_finalize: 1..128 patches, 0 <= x0 <= x1 <= X");
    return -1;
  }
  if (x0 == x1) return 0;
  const int rows = (x1 - x0) * Y;
  int blocks = (rows + GATHER_WARPS - 1) / GATHER_WARPS;
  static const int resident = [] {                 // persistent: exactly one wave of resident CTAs
    int per_sm = 0, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_finalize_kernel, GATHER_WARPS * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    return per_sm * sms;
  }();
  if (blocks > resident) blocks = resident;
  gather_finalize_kernel<<<blocks, GATHER_WARPS * 32, 0, st>>>(slots, plan, gaussian, x0 * Y, x1 * Y, X * Y, Y, Z, probs_out,
                                                              labels, target, counts);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_gather_finalize(const float* patch_probs, const GatherPlan& plan, int gaussian, int X, int Y, int Z,
                           float* probs_out, uint8_t* labels, const uint8_t* target, unsigned long long* counts,
                           cudaStream_t st) {
  GatherSlots sl;
  for (int i = 0; i < GatherPlan::MAX; ++i)
    sl.ptr[i] = i < plan.n ? patch_probs + (int64_t)i * plan.slot_planes * P3 : nullptr;
  return launch_gather_finalize_slots(sl, plan, gaussian, X, Y, Z, 0, X, probs_out, labels, target, counts, st);
}

// ---------------------------------------------------------------------------------------------
// 8-flip test-time augmentation around the tiling (predict_cls.py:180-203):
//   logit  = softmax(T(x));  logit += softmax(flip_f(T(flip_f(x))))  for the 7 flips;  output = logit / 8
// where T = tailor_and_concat.  flip_volume_kernel builds flip_f(x[..., :155]); tta_accumulate_kernel adds the
// un-flipped softmax (note: a softmax of what already are probabilities - the reference does exactly that).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
flip_volume_kernel(const float* __restrict__ src, float* __restrict__ dst, int X, int Y, int Zs, int Zd, int fx, int fy,
                   int fz) {
  const int64_t n = (int64_t)X * Y * Zd;
  const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= n) return;
  const int z = (int)(v % Zd);
  const int y = (int)((v / Zd) % Y);
  const int x = (int)(v / ((int64_t)Zd * Y));
  const int sx = fx ? X - 1 - x : x, sy = fy ? Y - 1 - y : y, sz = fz ? Zd - 1 - z : z;
  const int64_t s = ((int64_t)sx * Y + sy) * Zs + sz;
  const int64_t sp_s = (int64_t)X * Y * Zs;
#pragma unroll
  for (int c = 0; c < 4; ++c) dst[c * n + v] = __ldg(src + c * sp_s + s);
}

int launch_flip_volume(const float* src, float* dst, int X, int Y, int Zs, int Zd, int fx, int fy, int fz, cudaStream_t st) {
  const int64_t n = (int64_t)X * Y * Zd;
  flip_volume_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, X, Y, Zs, Zd, fx, fy, fz);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256)
tta_accumulate_kernel(const float* __restrict__ yf, float* __restrict__ out, int X, int Y, int Z, int fx, int fy, int fz,
                      int first, int last) {
  const int64_t n = (int64_t)X * Y * Z;
  const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= n) return;
  const int z = (int)(v % Z);
  const int y = (int)((v / Z) % Y);
  const int x = (int)(v / ((int64_t)Z * Y));
  const int sx = fx ? X - 1 - x : x, sy = fy ? Y - 1 - y : y, sz = fz ? Z - 1 - z : z;
  const int64_t s = ((int64_t)sx * Y + sy) * Z + sz;
  float p[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) p[c] = __ldg(yf + c * n + s);
  const float m = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[2], p[3]));
  float e[4], sum = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) { e[c] = expf(p[c] - m); sum += e[c]; }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float o = e[c] / sum;
    if (!first) o += out[c * n + v];
    if (last) o = o / 8.0f;
    out[c * n + v] = o;
  }
}

int launch_tta_accumulate(const float* yf, float* out, int X, int Y, int Z, int fx, int fy, int fz, int first, int last,
                          cudaStream_t st) {
  const int64_t n = (int64_t)X * Y * Z;
  tta_accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(yf, out, X, Y, Z, fx, fy, fz, first, last);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dcl
