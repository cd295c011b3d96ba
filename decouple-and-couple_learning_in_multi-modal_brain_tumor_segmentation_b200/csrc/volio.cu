// The two ends of the sliding-window path.
//
// Output side -- the label-map export the reference's drivers ask for through savepath / save_format / snapshot
// (predict.py:310-350, the block validate_softmax's arguments feed; predict_simple.py:186-278 for the per-slice
// pictures and tables): labels {0,1,2,3} -> BraTS labels {0,1,2,4}, the 1/2/4 and WT/TC/ET voxel counts of the
// verbose print, the NIfTI (x fastest) ordering of the same volume, RGB snapshot frames, per-slice Dice counters;
// host writers for .nii / .nii.gz (NIfTI-1 single file), .npy (np.save of the int64 arg-max map) and .png.
//
// Input side -- what data/ClsWiseBraTS128Test.BraDataSet128 (test_overlap.py:14,94-97; NOT shipped by the reference)
// must produce for predict_overlap.py:132-135: four NIfTI modalities -> brain mask (sum over modalities > 0) ->
// per-modality z-score over the mask -> (4, X, Y, Z padded 155 -> 160) with Z contiguous.  The recipe is the TransBTS
// one the reference's predict scripts descend from; with the loader absent its parity is UNPINNED.
//
// All kernels here are HBM-bound byte shuffles: coalesced reads along the contiguous axis, a shared-memory tile to turn
// the axis order, coalesced writes, integer counters reduced per warp and per CTA before the global atomics.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

#include "../../include/dcl_b200.h"
#include "common.cuh"

namespace dcl {

namespace {

// ---------------------------------------------------------------------------------------------
// export: one pass over the labels.  Tile = 32 (a = slow axis of the input) x 32 (c = contiguous axis) for a fixed
// middle index b; in[(a*B + b)*Cn + c] -> plain[(same)] = relabelled, turned[(c*B + b)*A + a] = relabelled.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
export_labels_kernel(const uint8_t* __restrict__ lab, int A, int B, int Cn, uint8_t* __restrict__ plain,
                     uint8_t* __restrict__ turned, unsigned long long* __restrict__ counts) {
  __shared__ uint8_t tile[32][33];
  __shared__ unsigned s_cnt[3];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  unsigned n1 = 0, n2 = 0, n4 = 0;
  // persistent CTAs walk the tiles (c fastest), so the label counters cost six global atomics per CTA, not per tile
  const int tc = (Cn + 31) / 32, ta = (A + 31) / 32;
  const int64_t tiles = (int64_t)tc * ta * B;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int c0 = (int)(t % tc) * 32, a0 = (int)((t / tc) % ta) * 32, b = (int)(t / ((int64_t)tc * ta));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = a0 + ty + 8 * i, c = c0 + tx;
      uint8_t v = 0;
      if (a < A && c < Cn) {
        const int64_t idx = ((int64_t)a * B + b) * Cn + c;
        const uint8_t l = lab[idx];
        v = l == 3 ? (uint8_t)4 : l;                       // seg_img[output == 3] = 4 (predict.py:322-324)
        n1 += v == 1; n2 += v == 2; n4 += v == 4;
        if (plain) plain[idx] = v;
      }
      tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
    if (turned) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, a = a0 + tx;
        if (a < A && c < Cn) turned[((int64_t)c * B + b) * A + a] = tile[tx][ty + 8 * i];
      }
    }
    __syncthreads();
  }
  if (counts) {
    n1 = __reduce_add_sync(0xffffffffu, n1);
    n2 = __reduce_add_sync(0xffffffffu, n2);
    n4 = __reduce_add_sync(0xffffffffu, n4);
    __syncthreads();
    if (tx == 0) {
      if (n1) atomicAdd(&s_cnt[0], n1);
      if (n2) atomicAdd(&s_cnt[1], n2);
      if (n4) atomicAdd(&s_cnt[2], n4);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long c1 = s_cnt[0], c2 = s_cnt[1], c4 = s_cnt[2];
      if (c1) atomicAdd(counts + 0, c1);
      if (c2) atomicAdd(counts + 1, c2);
      if (c4) atomicAdd(counts + 2, c4);
      if (c1 + c2 + c4) atomicAdd(counts + 3, c1 + c2 + c4);   // WT = 1|2|4   (predict.py:326-328)
      if (c1 + c4) atomicAdd(counts + 4, c1 + c4);             // TC = 1|4
      if (c4) atomicAdd(counts + 5, c4);                       // ET = 4
    }
  }
}

// frames[z][x][y][rgb] = palette[label[x][y][z]]: tile over (y, z) for a fixed x
struct Palette { uint8_t rgb[4][3]; };
__global__ void __launch_bounds__(256)
snapshot_kernel(const uint8_t* __restrict__ lab, int X, int Y, int Z, Palette pal, uint8_t* __restrict__ frames) {
  __shared__ uint8_t tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int z0 = blockIdx.x * 32, y0 = blockIdx.y * 32, x = blockIdx.z;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int y = y0 + ty + 8 * i, z = z0 + tx;
    tile[ty + 8 * i][tx] = (y < Y && z < Z) ? lab[((int64_t)x * Y + y) * Z + z] : 0;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int z = z0 + ty + 8 * i, y = y0 + tx;
    if (y < Y && z < Z) {
      const int l = tile[tx][ty + 8 * i] & 3;
      uint8_t* p = frames + (((int64_t)z * X + x) * Y + y) * 3;
      p[0] = pal.rgb[l][0]; p[1] = pal.rgb[l][1]; p[2] = pal.rgb[l][2];
    }
  }
}

// per z-slice (|o|, |t|, |o&t|) for WT, TC, ET: the counters behind output_excel's per-frame Dice
// (predict_simple.py:224-243).  A warp owns whole (x,y) rows: lane l always sees z = l, l+32, ... so its counters
// live in registers (NZ slices x 9) while the persistent grid strides over the rows; one atomic flush per warp.
constexpr int SLICE_MAX = 256;
template <int NZ>
__global__ void __launch_bounds__(256)
slice_counts_kernel(const uint8_t* __restrict__ lab, const uint8_t* __restrict__ tgt, int64_t rows, int Z,
                    unsigned long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  unsigned cnt[NZ][9];
#pragma unroll
  for (int k = 0; k < NZ; ++k)
#pragma unroll
    for (int j = 0; j < 9; ++j) cnt[k][j] = 0;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const uint8_t* lr = lab + r * Z;
    const uint8_t* tr = tgt + r * Z;
#pragma unroll
    for (int k = 0; k < NZ; ++k) {
      const int z = 32 * k + lane;
      if (z < Z) {
        const int l = lr[z], t = tr[z];
        const unsigned o0 = l > 0, o1 = (l == 1) | (l == 3), o2 = l == 3;
        const unsigned t0 = t > 0, t1 = (t == 1) | (t == 3), t2 = t == 3;
        cnt[k][0] += o0; cnt[k][1] += t0; cnt[k][2] += o0 & t0;
        cnt[k][3] += o1; cnt[k][4] += t1; cnt[k][5] += o1 & t1;
        cnt[k][6] += o2; cnt[k][7] += t2; cnt[k][8] += o2 & t2;
      }
    }
  }
  __shared__ unsigned s[NZ * 32 * 9];                     // CTA-level sum first: one global atomic per slot and CTA
  for (int i = threadIdx.x; i < NZ * 32 * 9; i += 256) s[i] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NZ; ++k) {
    const int z = 32 * k + lane;
#pragma unroll
    for (int j = 0; j < 9; ++j)
      if (cnt[k][j]) atomicAdd(&s[z * 9 + j], cnt[k][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Z * 9; i += 256)
    if (s[i]) atomicAdd(out + i, (unsigned long long)s[i]);
}

// ---------------------------------------------------------------------------------------------
// input side
// ---------------------------------------------------------------------------------------------
// stats[0..3] += sum over the mask of modality c, stats[4..7] += sum of squares, stats[8] += mask voxels
__global__ void __launch_bounds__(256)
mask_stats_kernel(const float* __restrict__ img, int64_t n, double* __restrict__ stats) {
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  double cnt = 0;
  for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < n; v += (int64_t)gridDim.x * 256) {
    const float m0 = img[v], m1 = img[n + v], m2 = img[2 * n + v], m3 = img[3 * n + v];
    if (((m0 + m1) + m2) + m3 > 0.f) {                     // mask = images.sum(-1) > 0, float32, left to right
      s[0] += m0; s[1] += m1; s[2] += m2; s[3] += m3;
      q[0] += (double)m0 * m0; q[1] += (double)m1 * m1; q[2] += (double)m2 * m2; q[3] += (double)m3 * m3;
      cnt += 1.0;
    }
  }
  __shared__ double sh[9];
  if (threadIdx.x < 9) sh[threadIdx.x] = 0;
  __syncthreads();
  double vals[9] = {s[0], s[1], s[2], s[3], q[0], q[1], q[2], q[3], cnt};
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    double x = vals[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(&sh[i], x);
  }
  __syncthreads();
  if (threadIdx.x < 9 && sh[threadIdx.x] != 0.0) atomicAdd(stats + threadIdx.x, sh[threadIdx.x]);
}

// stats (sums) -> stats[9..12] = mean, stats[13..16] = population std (numpy .std()), as float32-rounded doubles
__global__ void finish_stats_kernel(double* stats) {
  const int c = threadIdx.x;
  if (c >= 4) return;
  const double n = stats[8];
  double mean = 0, sd = 1;
  if (n > 0) {
    mean = stats[c] / n;
    const double var = stats[4 + c] / n - mean * mean;
    sd = sqrt(var > 0 ? var : 0);
  }
  stats[9 + c] = (double)(float)mean;
  stats[13 + c] = (double)(float)sd;
}

// in (4, Z, Y, X) x fastest -> out (4, X, Y, Zp) z fastest, z-scored over the mask, zero padded for z in [Z, Zp)
__global__ void __launch_bounds__(256)
normalise_reorder_kernel(const float* __restrict__ img, int X, int Y, int Z, int Zp, const double* __restrict__ stats,
                         float* __restrict__ out) {
  __shared__ float tile[4][32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32, y = blockIdx.z;
  const int64_t n = (int64_t)X * Y * Z;
  float mean[4], sd[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { mean[c] = (float)stats[9 + c]; sd[c] = (float)stats[13 + c]; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int z = z0 + ty + 8 * i, x = x0 + tx;
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    if (z < Z && x < X) {
      const int64_t v = ((int64_t)z * Y + y) * X + x;
#pragma unroll
      for (int c = 0; c < 4; ++c) m[c] = img[c * n + v];
      if (((m[0] + m[1]) + m[2]) + m[3] > 0.f) {
#pragma unroll
        for (int c = 0; c < 4; ++c) m[c] = (m[c] - mean[c]) / sd[c];     // x[mask] -= mean; x[mask] /= std
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) tile[c][ty + 8 * i][tx] = m[c];
  }
  __syncthreads();
  const int64_t np = (int64_t)X * Y * Zp;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x0 + ty + 8 * i, z = z0 + tx;
    if (x < X && z < Zp) {
#pragma unroll
      for (int c = 0; c < 4; ++c) out[c * np + ((int64_t)x * Y + y) * Zp + z] = tile[c][tx][ty + 8 * i];
    }
  }
}

// seg (Z, Y, X) x fastest -> target (X, Y, Zp) z fastest, optional 4 -> 3 (predict_overlap.py:150-152), zero padded
__global__ void __launch_bounds__(256)
reorder_labels_kernel(const uint8_t* __restrict__ seg, int X, int Y, int Z, int Zp, int map4to3,
                      uint8_t* __restrict__ out) {
  __shared__ uint8_t tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32, y = blockIdx.z;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int z = z0 + ty + 8 * i, x = x0 + tx;
    uint8_t v = 0;
    if (z < Z && x < X) v = seg[((int64_t)z * Y + y) * X + x];
    if (map4to3 && v == 4) v = 3;
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x0 + ty + 8 * i, z = z0 + tx;
    if (x < X && z < Zp) out[((int64_t)x * Y + y) * Zp + z] = tile[tx][ty + 8 * i];
  }
}

bool bad_shape(const int32_t shape[3]) {
  return !shape || shape[0] < 1 || shape[1] < 1 || shape[2] < 1 || shape[1] > 65535;
}

bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// ---------------------------------------------------------------------------------------------
// NIfTI-1 single-file header (348 bytes + 4 bytes of extension flags, data at 352); little endian
// ---------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct Nifti1Header {
  int32_t sizeof_hdr;
  char data_type[10];
  char db_name[18];
  int32_t extents;
  int16_t session_error;
  char regular;
  char dim_info;
  int16_t dim[8];
  float intent_p1, intent_p2, intent_p3;
  int16_t intent_code;
  int16_t datatype;
  int16_t bitpix;
  int16_t slice_start;
  float pixdim[8];
  float vox_offset;
  float scl_slope, scl_inter;
  int16_t slice_end;
  char slice_code;
  char xyzt_units;
  float cal_max, cal_min;
  float slice_duration;
  float toffset;
  int32_t glmax, glmin;
  char descrip[80];
  char aux_file[24];
  int16_t qform_code, sform_code;
  float quatern_b, quatern_c, quatern_d;
  float qoffset_x, qoffset_y, qoffset_z;
  float srow_x[4], srow_y[4], srow_z[4];
  char intent_name[16];
  char magic[4];
};
#pragma pack(pop)
static_assert(sizeof(Nifti1Header) == 348, "NIfTI-1 header is 348 bytes");

int bytes_of_datatype(int dt) {
  switch (dt) {
    case 2: return 1;      // uint8
    case 4: return 2;      // int16
    case 8: return 4;      // int32
    case 16: return 4;     // float32
    case 64: return 8;     // float64
    case 256: return 1;    // int8
    case 512: return 2;    // uint16
    case 768: return 4;    // uint32
    default: return 0;
  }
}

int read_header(gzFile f, Nifti1Header* h, const char* path) {
  if (gzread(f, h, 348) != 348) { set_error(std::string("nifti: short header in ") + path); return DCL_ERR_ARG; }
  if (h->sizeof_hdr != 348) { set_error(std::string("nifti: not a little-endian NIfTI-1 file: ") + path); return DCL_ERR_ARG; }
  if (strncmp(h->magic, "n+1", 3) != 0) { set_error(std::string("nifti: only single-file n+1 is supported: ") + path); return DCL_ERR_ARG; }
  if (h->dim[0] < 3 || h->dim[0] > 7) { set_error(std::string("nifti: need at least 3 dimensions: ") + path); return DCL_ERR_ARG; }
  for (int i = 4; i <= h->dim[0]; ++i)
    if (h->dim[i] > 1) { set_error(std::string("nifti: more than 3 non-singleton dimensions: ") + path); return DCL_ERR_ARG; }
  if (!bytes_of_datatype(h->datatype)) { set_error(std::string("nifti: unsupported datatype in ") + path); return DCL_ERR_ARG; }
  return DCL_OK;
}

}  // namespace
}  // namespace dcl

using namespace dcl;

extern "C" {

DCL_API int dcl_export_labels(const uint8_t* labels_dev, const int32_t shape[3], uint8_t* seg_out_dev,
                              uint8_t* seg_nifti_dev, uint64_t* counts_out_dev, void* stream) {
  if (!labels_dev || bad_shape(shape)) { set_error("dcl_export_labels: bad argument"); return DCL_ERR_ARG; }
  const int X = shape[0], Y = shape[1], Z = shape[2];
  cudaStream_t st = (cudaStream_t)stream;
  if (counts_out_dev) DCL_CUDA_OK(cudaMemsetAsync(counts_out_dev, 0, 6 * sizeof(uint64_t), st));
  int64_t tiles = (int64_t)((Z + 31) / 32) * ((X + 31) / 32) * Y;
  if (tiles > 148 * 8) tiles = 148 * 8;
  export_labels_kernel<<<(unsigned)tiles, 256, 0, st>>>(labels_dev, X, Y, Z, seg_out_dev, seg_nifti_dev,
                                                        (unsigned long long*)counts_out_dev);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return DCL_OK;
}

DCL_API int dcl_snapshot_frames(const uint8_t* labels_dev, const int32_t shape[3], const uint8_t palette[12],
                                uint8_t* frames_out_dev, void* stream) {
  if (!labels_dev || !palette || !frames_out_dev || bad_shape(shape) || shape[0] > 65535) {
    set_error("dcl_snapshot_frames: bad argument");
    return DCL_ERR_ARG;
  }
  const int X = shape[0], Y = shape[1], Z = shape[2];
  Palette pal;
  memcpy(pal.rgb, palette, 12);
  snapshot_kernel<<<dim3((Z + 31) / 32, (Y + 31) / 32, X), 256, 0, (cudaStream_t)stream>>>(labels_dev, X, Y, Z, pal,
                                                                                          frames_out_dev);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return DCL_OK;
}

DCL_API int dcl_slice_counts(const uint8_t* labels_dev, const uint8_t* target_dev, const int32_t shape[3],
                             uint64_t* counts_out_dev, void* stream) {
  if (!labels_dev || !target_dev || !counts_out_dev || bad_shape(shape) || shape[2] > SLICE_MAX) {
    set_error("dcl_slice_counts: bad argument (Z <= 256)");
    return DCL_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  DCL_CUDA_OK(cudaMemsetAsync(counts_out_dev, 0, (size_t)shape[2] * 9 * sizeof(uint64_t), st));
  const int64_t rows = (int64_t)shape[0] * shape[1];
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 2) blocks = 148 * 2;
  unsigned long long* out = (unsigned long long*)counts_out_dev;
  const int nz = (shape[2] + 31) / 32;
  if (nz <= 5) slice_counts_kernel<5><<<(unsigned)blocks, 256, 0, st>>>(labels_dev, target_dev, rows, shape[2], out);
  else slice_counts_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(labels_dev, target_dev, rows, shape[2], out);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return DCL_OK;
}

DCL_API int dcl_preprocess_volume(const float* modalities_dev, const int32_t shape[3], int32_t z_pad, float* vol_out_dev,
                                  double* stats_dev, void* stream) {
  if (!modalities_dev || !vol_out_dev || !stats_dev || bad_shape(shape) || z_pad < shape[2]) {
    set_error("dcl_preprocess_volume: bad argument");
    return DCL_ERR_ARG;
  }
  const int X = shape[0], Y = shape[1], Z = shape[2];
  const int64_t n = (int64_t)X * Y * Z;
  cudaStream_t st = (cudaStream_t)stream;
  DCL_CUDA_OK(cudaMemsetAsync(stats_dev, 0, 17 * sizeof(double), st));
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  mask_stats_kernel<<<(unsigned)blocks, 256, 0, st>>>(modalities_dev, n, stats_dev);
  finish_stats_kernel<<<1, 32, 0, st>>>(stats_dev);
  normalise_reorder_kernel<<<dim3((X + 31) / 32, (z_pad + 31) / 32, Y), 256, 0, st>>>(modalities_dev, X, Y, Z, z_pad,
                                                                                     stats_dev, vol_out_dev);
  g_launches += 3;
  DCL_CUDA_OK(cudaGetLastError());
  return DCL_OK;
}

DCL_API int dcl_reorder_labels(const uint8_t* seg_nifti_dev, const int32_t shape[3], int32_t z_pad, int32_t map4to3,
                               uint8_t* target_out_dev, void* stream) {
  if (!seg_nifti_dev || !target_out_dev || bad_shape(shape) || z_pad < shape[2]) {
    set_error("dcl_reorder_labels: bad argument");
    return DCL_ERR_ARG;
  }
  const int X = shape[0], Y = shape[1], Z = shape[2];
  reorder_labels_kernel<<<dim3((X + 31) / 32, (z_pad + 31) / 32, Y), 256, 0, (cudaStream_t)stream>>>(
      seg_nifti_dev, X, Y, Z, z_pad, map4to3, target_out_dev);
  ++g_launches;
  DCL_CUDA_OK(cudaGetLastError());
  return DCL_OK;
}

// ---- host-side files (no device needed) --------------------------------------------------------

DCL_API int dcl_write_nifti(const char* path, const void* data_host, int32_t datatype, const int32_t shape[3]) {
  if (!path || !data_host || bad_shape(shape)) { set_error("dcl_write_nifti: bad argument"); return DCL_ERR_ARG; }
  const int bytes = bytes_of_datatype(datatype);
  if (!bytes) { set_error("dcl_write_nifti: unsupported datatype code"); return DCL_ERR_ARG; }
  if (shape[0] > 32767 || shape[1] > 32767 || shape[2] > 32767) { set_error("dcl_write_nifti: dimension too large"); return DCL_ERR_ARG; }
  Nifti1Header h;
  memset(&h, 0, sizeof h);
  h.sizeof_hdr = 348;
  h.regular = 'r';
  h.dim[0] = 3; h.dim[1] = (int16_t)shape[0]; h.dim[2] = (int16_t)shape[1]; h.dim[3] = (int16_t)shape[2];
  for (int i = 4; i < 8; ++i) h.dim[i] = 1;
  h.datatype = (int16_t)datatype;
  h.bitpix = (int16_t)(8 * bytes);
  for (int i = 0; i < 8; ++i) h.pixdim[i] = 1.f;          // Nifti1Image(data, None): unit voxels, no transform codes
  h.vox_offset = 352.f;
  h.scl_slope = NAN; h.scl_inter = NAN;                    // "no scaling"
  h.srow_x[0] = 1.f; h.srow_y[1] = 1.f; h.srow_z[2] = 1.f;
  memcpy(h.magic, "n+1", 4);
  const char ext[4] = {0, 0, 0, 0};
  const size_t nbytes = (size_t)shape[0] * shape[1] * shape[2] * bytes;
  const std::string p(path);
  if (ends_with(p, ".gz")) {
    gzFile f = gzopen(path, "wb6");
    if (!f) { set_error(std::string("dcl_write_nifti: cannot open ") + path); return DCL_ERR_ARG; }
    bool ok = gzwrite(f, &h, 348) == 348 && gzwrite(f, ext, 4) == 4;
    const char* d = (const char*)data_host;
    for (size_t off = 0; ok && off < nbytes;) {
      const unsigned chunk = (unsigned)((nbytes - off) > (1u << 30) ? (1u << 30) : (nbytes - off));
      ok = gzwrite(f, d + off, chunk) == (int)chunk;
      off += chunk;
    }
    ok = (gzclose(f) == Z_OK) && ok;
    if (!ok) { set_error(std::string("dcl_write_nifti: write failed: ") + path); return DCL_ERR_ARG; }
  } else {
    FILE* f = fopen(path, "wb");
    if (!f) { set_error(std::string("dcl_write_nifti: cannot open ") + path); return DCL_ERR_ARG; }
    bool ok = fwrite(&h, 1, 348, f) == 348 && fwrite(ext, 1, 4, f) == 4 && fwrite(data_host, 1, nbytes, f) == nbytes;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error(std::string("dcl_write_nifti: write failed: ") + path); return DCL_ERR_ARG; }
  }
  return DCL_OK;
}

DCL_API int dcl_read_nifti_header(const char* path, int32_t shape_out[3], int32_t* datatype_out, float pixdim_out[3]) {
  if (!path || !shape_out) { set_error("dcl_read_nifti_header: bad argument"); return DCL_ERR_ARG; }
  gzFile f = gzopen(path, "rb");                           // reads plain and gzip files alike
  if (!f) { set_error(std::string("nifti: cannot open ") + path); return DCL_ERR_ARG; }
  Nifti1Header h;
  const int rc = read_header(f, &h, path);
  gzclose(f);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) shape_out[i] = h.dim[1 + i];
  if (datatype_out) *datatype_out = h.datatype;
  if (pixdim_out) for (int i = 0; i < 3; ++i) pixdim_out[i] = h.pixdim[1 + i];
  return DCL_OK;
}

// Reads the voxel data as float32 in storage order (x fastest), scl_slope / scl_inter applied like nibabel's
// get_fdata (slope 0 or NaN = unscaled).  Returns the element count.
DCL_API int64_t dcl_read_nifti_f32(const char* path, float* data_out_host, int64_t capacity) {
  if (!path || !data_out_host) { set_error("dcl_read_nifti_f32: bad argument"); return DCL_ERR_ARG; }
  gzFile f = gzopen(path, "rb");
  if (!f) { set_error(std::string("nifti: cannot open ") + path); return DCL_ERR_ARG; }
  gzbuffer(f, 1 << 20);
  Nifti1Header h;
  int rc = read_header(f, &h, path);
  if (rc) { gzclose(f); return rc; }
  const int64_t n = (int64_t)h.dim[1] * h.dim[2] * h.dim[3];
  if (n > capacity) { gzclose(f); set_error("dcl_read_nifti_f32: output buffer too small"); return DCL_ERR_ARG; }
  const int64_t skip = (int64_t)h.vox_offset - 348;
  if (skip < 0 || gzseek(f, (z_off_t)h.vox_offset, SEEK_SET) < 0) { gzclose(f); set_error("nifti: bad vox_offset"); return DCL_ERR_ARG; }
  const int bytes = bytes_of_datatype(h.datatype);
  std::vector<char> raw((size_t)n * bytes);
  int64_t got = 0;
  while (got < (int64_t)raw.size()) {
    const unsigned chunk = (unsigned)(((int64_t)raw.size() - got) > (1 << 30) ? (1 << 30) : ((int64_t)raw.size() - got));
    const int r = gzread(f, raw.data() + got, chunk);
    if (r <= 0) break;
    got += r;
  }
  gzclose(f);
  if (got != (int64_t)raw.size()) { set_error(std::string("nifti: truncated data in ") + path); return DCL_ERR_ARG; }
  const bool scaled = h.scl_slope != 0.f && !isnan(h.scl_slope) && !(h.scl_slope == 1.f && (h.scl_inter == 0.f || isnan(h.scl_inter)));
  const double slope = h.scl_slope, inter = isnan(h.scl_inter) ? 0.0 : h.scl_inter;
  for (int64_t i = 0; i < n; ++i) {
    double v;
    const char* p = raw.data() + i * bytes;
    switch (h.datatype) {
      case 2: v = *(const uint8_t*)p; break;
      case 4: { int16_t t; memcpy(&t, p, 2); v = t; break; }
      case 8: { int32_t t; memcpy(&t, p, 4); v = t; break; }
      case 16: { float t; memcpy(&t, p, 4); v = t; break; }
      case 64: { double t; memcpy(&t, p, 8); v = t; break; }
      case 256: v = *(const int8_t*)p; break;
      case 512: { uint16_t t; memcpy(&t, p, 2); v = t; break; }
      default: { uint32_t t; memcpy(&t, p, 4); v = t; break; }
    }
    data_out_host[i] = (float)(scaled ? v * slope + inter : v);
  }
  return n;
}

// np.save(path, output) for the int64 arg-max map (predict.py:313-314): NPY format 1.0, C order, '<i8'
DCL_API int dcl_write_npy_labels(const char* path, const uint8_t* labels_host, const int32_t shape[3]) {
  if (!path || !labels_host || bad_shape(shape)) { set_error("dcl_write_npy_labels: bad argument"); return DCL_ERR_ARG; }
  char dict[256];
  snprintf(dict, sizeof dict, "{'descr': '<i8', 'fortran_order': False, 'shape': (%d, %d, %d), }", shape[0], shape[1], shape[2]);
  std::string header(dict);
  char first[32];
  snprintf(first, sizeof first, "%d", shape[0]);
  header.append(21 - strlen(first), ' ');                  // numpy's GROWTH_AXIS_MAX_DIGITS spare room
  const size_t unpadded = 10 + header.size() + 1;          // magic(6) + version(2) + length(2) + header + '\n'
  header.append((64 - unpadded % 64) % 64, ' ');
  header.push_back('\n');
  FILE* f = fopen(path, "wb");
  if (!f) { set_error(std::string("dcl_write_npy_labels: cannot open ") + path); return DCL_ERR_ARG; }
  const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  const uint16_t hlen = (uint16_t)header.size();
  bool ok = fwrite(magic, 1, 8, f) == 8 && fwrite(&hlen, 1, 2, f) == 2 && fwrite(header.data(), 1, header.size(), f) == header.size();
  const int64_t n = (int64_t)shape[0] * shape[1] * shape[2];
  std::vector<int64_t> buf(1 << 16);
  for (int64_t off = 0; ok && off < n; off += (int64_t)buf.size()) {
    const int64_t m = (n - off) < (int64_t)buf.size() ? (n - off) : (int64_t)buf.size();
    for (int64_t i = 0; i < m; ++i) buf[i] = labels_host[off + i];
    ok = fwrite(buf.data(), 8, (size_t)m, f) == (size_t)m;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) { set_error(std::string("dcl_write_npy_labels: write failed: ") + path); return DCL_ERR_ARG; }
  return DCL_OK;
}

// 8-bit RGB PNG (imageio.imwrite of an (H, W, 3) uint8 frame, predict.py:350 / predict_simple.py:198)
DCL_API int dcl_write_png_rgb(const char* path, const uint8_t* rgb_host, int32_t height, int32_t width) {
  if (!path || !rgb_host || height < 1 || width < 1) { set_error("dcl_write_png_rgb: bad argument"); return DCL_ERR_ARG; }
  const size_t row = (size_t)width * 3;
  std::vector<unsigned char> raw((row + 1) * (size_t)height);
  for (int y = 0; y < height; ++y) {
    raw[(row + 1) * y] = 0;                                // filter type 0 (None)
    memcpy(&raw[(row + 1) * y + 1], rgb_host + row * y, row);
  }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<unsigned char> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) { set_error("dcl_write_png_rgb: deflate failed"); return DCL_ERR_ARG; }
  FILE* f = fopen(path, "wb");
  if (!f) { set_error(std::string("dcl_write_png_rgb: cannot open ") + path); return DCL_ERR_ARG; }
  auto be32 = [](unsigned char* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; };
  bool ok = true;
  auto chunk = [&](const char* type, const unsigned char* data, uint32_t len) {
    unsigned char hdr[8];
    be32(hdr, len);
    memcpy(hdr + 4, type, 4);
    uLong crc = crc32(0L, hdr + 4, 4);
    if (len) crc = crc32(crc, data, len);
    unsigned char tail[4];
    be32(tail, (uint32_t)crc);
    ok = ok && fwrite(hdr, 1, 8, f) == 8 && (len == 0 || fwrite(data, 1, len, f) == len) && fwrite(tail, 1, 4, f) == 4;
  };
  const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  ok = fwrite(sig, 1, 8, f) == 8;
  unsigned char ihdr[13];
  be32(ihdr, (uint32_t)width); be32(ihdr + 4, (uint32_t)height);
  ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;     // 8 bit, colour type 2 (RGB)
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", comp.data(), (uint32_t)clen);
  chunk("IEND", nullptr, 0);
  ok = (fclose(f) == 0) && ok;
  if (!ok) { set_error(std::string("dcl_write_png_rgb: write failed: ") + path); return DCL_ERR_ARG; }
  return DCL_OK;
}

}  // extern "C"
