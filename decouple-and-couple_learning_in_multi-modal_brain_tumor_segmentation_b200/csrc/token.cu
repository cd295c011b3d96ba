// Token path of the region couplers: score + top-k selection, sequence assembly, LayerNorm, linear
// layers, 8-head attention over 129-token sequences, row scatter.  All fp32 (the discrete
// top-k makes low precision fragile here).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include "common.cuh"

namespace dcl {

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// score[i] = <token, feats[i]>  (cls_wise_former.py:345: e_token @ feats^T).  One warp per row.
__global__ void __launch_bounds__(256)
score_kernel(const float* __restrict__ token, const float* __restrict__ feats, int n_tokens,
             float* __restrict__ score) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_tokens) return;
  const float4* f = reinterpret_cast<const float4*>(feats + (int64_t)row * TOKEN_DIM);
  const float4* t = reinterpret_cast<const float4*>(token);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < TOKEN_DIM / 128; ++j) {
    float4 a = __ldg(f + lane + 32 * j), b = __ldg(t + lane + 32 * j);
    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
  }
  s = warp_sum_t(s);
  if (lane == 0) score[row] = s;
}

// topk(128, largest, sorted) of up to 2048 scores (cls_wise_former.py:346) by rank counting: element i
// lands at position #{j : s_j > s_i or (s_j == s_i and j < i)}; positions < 128 are the sorted top-k
// (ties broken by index, as a stable descending sort would).  128 elements per block, all scores in smem.
// Block = 8 warps x 4 elements; the 32 lanes of a warp split the scan over all scores (each lane keeps its
// n/32 scores in registers), so an element's rank is 64 compares per lane plus one warp reduction.
constexpr int TOPK_EPW = 4;                         // elements per warp
constexpr int TOPK_EPB = 8 * TOPK_EPW;              // elements per block
__global__ void __launch_bounds__(256)
topk_kernel(const float* __restrict__ score, int n, int* __restrict__ idx_out) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float key[64];                                    // key[t] = score[lane + 32 t]
#pragma unroll
  for (int t = 0; t < 64; ++t) {
    const int j = lane + 32 * t;
    key[t] = j < n ? __ldg(score + j) : -FLT_MAX;
  }
  const int nt = (n + 31) / 32;
#pragma unroll
  for (int e = 0; e < TOPK_EPW; ++e) {
    const int i = blockIdx.x * TOPK_EPB + warp * TOPK_EPW + e;
    if (i >= n) break;                              // uniform per warp
    const float si = __ldg(score + i);
    int rank = 0;
#pragma unroll
    for (int t = 0; t < 64; ++t) {
      if (t < nt) {
        const int j = lane + 32 * t;
        rank += (key[t] > si || (key[t] == si && j < i)) ? 1 : 0;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (lane == 0 && rank < TOP_NUM) idx_out[rank] = i;
  }
}

int launch_select_topk(const float* token, const float* feats, int n_tokens, float* score_scratch, int* idx_out,
                       cudaStream_t st) {
  if (n_tokens > 2048 || n_tokens < TOP_NUM) { set_error("select_topk: 128 <= n_tokens <= 2048 required"); return -1; }
  DCL_CUDA_OK(launch_pdl(score_kernel, dim3((n_tokens + 7) / 8), dim3(256), (size_t)(0), st, token, feats, n_tokens, score_scratch));
  DCL_CUDA_OK(launch_pdl(topk_kernel, dim3((n_tokens + TOPK_EPB - 1) / TOPK_EPB), dim3(256), (size_t)(0), st, score_scratch, n_tokens, idx_out));
  g_launches += 2;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// seq[0] = class token, seq[1+i] = feats[idx[i]] + pe_row  (index_select + positional encoding + cat,
// cls_wise_former.py:347-350; PositionalEncoding.py:20-22 adds row 0 of `pe` to every token).
__global__ void __launch_bounds__(128)
build_sequence_kernel(const float* __restrict__ class_token, const float* __restrict__ feats,
                      const int* __restrict__ idx, const float* __restrict__ pe_row, float* __restrict__ seq) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  const int r = blockIdx.x;
  float4 v;
  if (r == 0) {
    v = __ldg(reinterpret_cast<const float4*>(class_token) + threadIdx.x);
  } else {
    v = __ldg(reinterpret_cast<const float4*>(feats + (int64_t)idx[r - 1] * TOKEN_DIM) + threadIdx.x);
    float4 p = __ldg(reinterpret_cast<const float4*>(pe_row) + threadIdx.x);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
  }
  reinterpret_cast<float4*>(seq + (int64_t)r * TOKEN_DIM)[threadIdx.x] = v;
}

int launch_build_sequence(const float* class_token, const float* feats, const int* idx, const float* pe_row,
                          float* seq, cudaStream_t st) {
  DCL_CUDA_OK(launch_pdl(build_sequence_kernel, dim3(SEQ), dim3(128), (size_t)(0), st, class_token, feats, idx, pe_row, seq));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// nn.LayerNorm(512), eps 1e-5, one warp per row, two-pass variance in registers.
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float* __restrict__ y, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * TOKEN_DIM);
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[j] = __ldg(xr + lane + 32 * j); s += (v[j].x + v[j].y) + (v[j].z + v[j].w); }
  const float mean = warp_sum_t(s) * (1.f / TOKEN_DIM);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum_t(q) * (1.f / TOKEN_DIM) + 1e-5f);
  float4* yr = reinterpret_cast<float4*>(y + (int64_t)row * TOKEN_DIM);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
    float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
    float4 o;
    o.x = (v[j].x - mean) * rstd * g.x + b.x;
    o.y = (v[j].y - mean) * rstd * g.y + b.y;
    o.z = (v[j].z - mean) * rstd * g.z + b.z;
    o.w = (v[j].w - mean) * rstd * g.w + b.w;
    yr[lane + 32 * j] = o;
  }
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, int rows, cudaStream_t st) {
  layernorm_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, gamma, beta, y, rows);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// y = x W^T + b (+GELU erf) (+residual).  32x64 output tile, BK = 32, 256 threads x (2 x 4) outputs.
constexpr int LBM = 32, LBN = 64, LBK = 32;

__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              const float* __restrict__ residual, float* __restrict__ y, int m, int n, int k, int gelu) {
  __shared__ float xs[LBK][LBM + 1];
  __shared__ __align__(16) float ws[LBK][LBN + 4];
  const int m0 = blockIdx.y * LBM, n0 = blockIdx.x * LBN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int k0 = 0; k0 < k; k0 += LBK) {
    for (int e = threadIdx.x; e < LBM * LBK; e += 256) {
      int c = e % LBK, r = e / LBK;
      xs[c][r] = (m0 + r < m && k0 + c < k) ? __ldg(x + (int64_t)(m0 + r) * k + k0 + c) : 0.f;
    }
    for (int e = threadIdx.x; e < LBN * LBK; e += 256) {
      int c = e % LBK, r = e / LBK;
      ws[c][r] = (n0 + r < n && k0 + c < k) ? __ldg(w + (int64_t)(n0 + r) * k + k0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LBK; ++kk) {
      float a0 = xs[kk][ty * 2], a1 = xs[kk][ty * 2 + 1];
      float4 b = *reinterpret_cast<const float4*>(&ws[kk][tx * 4]);
      acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
      acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
      acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
      acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int r = m0 + ty * 2 + i;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = n0 + tx * 4 + j;
      if (c >= n) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + c) : 0.f);
      if (gelu) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
      if (residual) v += __ldg(residual + (int64_t)r * n + c);
      y[(int64_t)r * n + c] = v;
    }
  }
}

int launch_linear(const float* x, const float* w, const float* bias, const float* residual, float* y, int m, int n,
                  int k, bool gelu, cudaStream_t st) {
  dim3 grid((n + LBN - 1) / LBN, (m + LBM - 1) / LBM);
  linear_kernel<<<grid, 256, 0, st>>>(x, w, bias, residual, y, m, n, k, gelu ? 1 : 0);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// DualSelfAttention core (SelfAttention.py:94-99): softmax(q k^T / 8) v per head.
// Block = (16-query chunk, head), 4 warps x 4 query rows processed together; K (row pitch 68 floats so
// 16-byte row reads of 8 consecutive keys hit 32 distinct banks) and V of the head live in shared memory,
// filled with 16-byte loads.
constexpr int ATT_HD = 64;
constexpr int ATT_QCHUNK = 16;
constexpr int ATT_MAXK = 160;
constexpr int ATT_KP = ATT_HD + 4;
constexpr int ATT_KG = ATT_MAXK / 32;     // key groups per lane

__global__ void __launch_bounds__(128)
attention_kernel(const float* __restrict__ q, const float* __restrict__ kv, float* __restrict__ out,
                 unsigned short* __restrict__ out_blocked, int mq, int mk, int x3) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;                              // [mk][68]
  float* vs = ks + mk * ATT_KP;                // [mk][64]
  float* qs = vs + mk * ATT_HD;                // [16][64]
  float* ps = qs + ATT_QCHUNK * ATT_HD;        // [16][ATT_MAXK]
  const int head = blockIdx.y;
  const int row0 = blockIdx.x * ATT_QCHUNK;
  for (int e = threadIdx.x; e < mk * (ATT_HD / 4); e += 128) {
    const int d4 = e % (ATT_HD / 4), j = e / (ATT_HD / 4);
    const float4 k4 = __ldg(reinterpret_cast<const float4*>(kv + (int64_t)j * 1024 + head * ATT_HD) + d4);
    const float4 v4 = __ldg(reinterpret_cast<const float4*>(kv + (int64_t)j * 1024 + 512 + head * ATT_HD) + d4);
    *reinterpret_cast<float4*>(ks + j * ATT_KP + 4 * d4) = k4;
    *reinterpret_cast<float4*>(vs + j * ATT_HD + 4 * d4) = v4;
  }
  for (int e = threadIdx.x; e < ATT_QCHUNK * (ATT_HD / 4); e += 128) {
    const int d4 = e % (ATT_HD / 4), r = e / (ATT_HD / 4);
    float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < mq) q4 = __ldg(reinterpret_cast<const float4*>(q + (int64_t)(row0 + r) * TOKEN_DIM + head * ATT_HD) + d4);
    *reinterpret_cast<float4*>(qs + r * ATT_HD + 4 * d4) = q4;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* qw = qs + warp * 4 * ATT_HD;     // this warp's 4 query rows
  float* pw = ps + warp * 4 * ATT_MAXK;
  float s[4][ATT_KG];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int g = 0; g < ATT_KG; ++g) s[r][g] = 0.f;
#pragma unroll 4
  for (int d4 = 0; d4 < ATT_HD / 4; ++d4) {
    float4 qv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) qv[r] = *reinterpret_cast<const float4*>(qw + r * ATT_HD + 4 * d4);
#pragma unroll
    for (int g = 0; g < ATT_KG; ++g) {
      const int j = lane + 32 * g;
      if (j < mk) {
        const float4 kq = *reinterpret_cast<const float4*>(ks + j * ATT_KP + 4 * d4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          s[r][g] = fmaf(qv[r].x, kq.x, s[r][g]);
          s[r][g] = fmaf(qv[r].y, kq.y, s[r][g]);
          s[r][g] = fmaf(qv[r].z, kq.z, s[r][g]);
          s[r][g] = fmaf(qv[r].w, kq.w, s[r][g]);
        }
      }
    }
  }
  float inv[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float mx = -FLT_MAX;
#pragma unroll
    for (int g = 0; g < ATT_KG; ++g)
      if (lane + 32 * g < mk) { s[r][g] *= 0.125f; mx = fmaxf(mx, s[r][g]); }   // head_dim ** -0.5
    mx = warp_max_t(mx);
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < ATT_KG; ++g)
      if (lane + 32 * g < mk) { const float e = expf(s[r][g] - mx); pw[r * ATT_MAXK + lane + 32 * g] = e; sum += e; }
    inv[r] = 1.f / warp_sum_t(sum);
  }
  __syncwarp();
  float o[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) o[r][0] = o[r][1] = 0.f;
#pragma unroll 4
  for (int j = 0; j < mk; ++j) {
    const float v0 = vs[j * ATT_HD + lane], v1 = vs[j * ATT_HD + lane + 32];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float p = pw[r * ATT_MAXK + j];
      o[r][0] = fmaf(p, v0, o[r][0]);
      o[r][1] = fmaf(p, v1, o[r][1]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = row0 + warp * 4 + r;
    if (row < mq) {
      if (out_blocked != nullptr) {
        // bf16, channel-blocked [64 chunks][mq rows][8]: exactly the A operand of the out-projection GEMM, so the
        // separate fp32 -> blocked conversion pass is skipped
        const int c0 = head * ATT_HD + lane, c1 = c0 + 32;
        const float v0 = o[r][0] * inv[r], v1 = o[r][1] * inv[r];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        out_blocked[((int64_t)(c0 >> 3) * mq + row) * 8 + (c0 & 7)] = __bfloat16_as_ushort(h0);
        out_blocked[((int64_t)(c1 >> 3) * mq + row) * 8 + (c1 & 7)] = __bfloat16_as_ushort(h1);
        if (x3) {      // split mode: fp16 hi + fp16 lo, the lo planes follow the 64 hi chunks
          const __half g0 = __float2half_rn(v0), g1 = __float2half_rn(v1);
          out_blocked[((int64_t)(c0 >> 3) * mq + row) * 8 + (c0 & 7)] = __half_as_ushort(g0);
          out_blocked[((int64_t)(c1 >> 3) * mq + row) * 8 + (c1 & 7)] = __half_as_ushort(g1);
          out_blocked[((int64_t)(64 + (c0 >> 3)) * mq + row) * 8 + (c0 & 7)] = __half_as_ushort(__float2half_rn(v0 - __half2float(g0)));
          out_blocked[((int64_t)(64 + (c1 >> 3)) * mq + row) * 8 + (c1 & 7)] = __half_as_ushort(__float2half_rn(v1 - __half2float(g1)));
        }
      } else {
        out[(int64_t)row * TOKEN_DIM + head * ATT_HD + lane] = o[r][0] * inv[r];
        out[(int64_t)row * TOKEN_DIM + head * ATT_HD + lane + 32] = o[r][1] * inv[r];
      }
    }
  }
}

int launch_attention(const float* q, const float* kv, float* out, int mq, int mk, cudaStream_t st, void* out_blocked, bool x3) {
  if (mk > ATT_MAXK) { set_error("attention: at most 160 keys"); return -1; }
  size_t smem = (size_t)(mk * ATT_KP + mk * ATT_HD + ATT_QCHUNK * ATT_HD + ATT_QCHUNK * ATT_MAXK) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    DCL_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  DCL_CUDA_OK(launch_pdl(attention_kernel, dim3(dim3((mq + ATT_QCHUNK - 1) / ATT_QCHUNK, 8)), dim3(128), (size_t)(smem), st, q, kv, out, reinterpret_cast<unsigned short*>(out_blocked), mq, mk, x3 ? 1 : 0));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

// feats[idx[i]] = rows[i]: the whole-row scatter_ of cls_wise_former.py:467 (fix_index.txt rows are [k]*512).
__global__ void __launch_bounds__(128)
scatter_rows_kernel(float* __restrict__ feats, const int* __restrict__ idx, const float* __restrict__ rows,
                    int row_stride) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  const int i = blockIdx.x;
  float4 v = __ldg(reinterpret_cast<const float4*>(rows + (int64_t)i * row_stride) + threadIdx.x);
  reinterpret_cast<float4*>(feats + (int64_t)idx[i] * TOKEN_DIM)[threadIdx.x] = v;
}

int launch_scatter_rows(float* feats, const int* idx, const float* rows, int row_stride, cudaStream_t st) {
  DCL_CUDA_OK(launch_pdl(scatter_rows_kernel, dim3(TOP_NUM), dim3(128), (size_t)(0), st, feats, idx, rows, row_stride));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256)
add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
            float* __restrict__ y, int64_t n4) {
  pdl_wait();      // programmatic dependent launch: the predecessor kernel has completed past this point
  pdl_trigger();
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  float4 u = __ldg(reinterpret_cast<const float4*>(a) + i);
  float4 v = __ldg(reinterpret_cast<const float4*>(b) + i);
  float4 w = __ldg(reinterpret_cast<const float4*>(c) + i);
  reinterpret_cast<float4*>(y)[i] = make_float4((u.x + v.x) + w.x, (u.y + v.y) + w.y, (u.z + v.z) + w.z,
                                                (u.w + v.w) + w.w);
}

int launch_add3(const float* a, const float* b, const float* c, float* y, int64_t n, cudaStream_t st) {
  if (n % 4 != 0) { set_error("add3: n must be a multiple of 4"); return -1; }
  DCL_CUDA_OK(launch_pdl(add3_kernel, dim3((unsigned)((n / 4 + 255) / 256)), dim3(256), (size_t)(0), st, a, b, c, y, n / 4));
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dcl
