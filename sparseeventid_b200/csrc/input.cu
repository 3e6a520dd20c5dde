// larcv batch-filler array -> SparseConvNet input tuple, on the device.
// Replaces the host numpy pass of the reference (src/io/data_transforms.py:21-49 larcvsparse_to_scnsparse_3d,
// :198-252 larcvsparse_to_scnsparse_2d): the pinned [B][P][V][D+1] fp32 buffer is copied to the
// GPU as it is, padding rows (value == -999) are dropped by an ORDER-PRESERVING compaction, and the rows come out
// exactly in the reference's order -- 3-D: numpy.where order (batch, voxel); 2-D: plane-major (plane, batch, voxel).
// HBM-bound: one pass to count, one pass to compact, 16*V bytes read per (sample, plane) each.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// counts[p][b] = number of rows of (sample b, plane p) whose value column differs from pad
__global__ void __launch_bounds__(kThreads) k_larcv_count(const float* __restrict__ a, int B, int P, int V, int ncol,
                                                           float pad, int32_t* __restrict__ counts) {
  const int p = blockIdx.x / B, b = blockIdx.x % B;
  const float* base = a + ((size_t)b * P + p) * (size_t)V * ncol;
  int c = 0;
  for (int v = threadIdx.x; v < V; v += kThreads) c += base[(size_t)v * ncol + (ncol - 1)] != pad;
  __shared__ int s[kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += s[w];
    counts[blockIdx.x] = t;
  }
}

// layout 0 (3-D): coords = (col0, col1, col2, b);  layout 1 (2-D): coords = (plane, col1 (y), col0 (x), b)
__global__ void __launch_bounds__(kThreads) k_larcv_compact(const float* __restrict__ a, int B, int P, int V, int ncol,
                                                             float pad, const int64_t* __restrict__ offs, int layout,
                                                             int32_t* __restrict__ coords, float* __restrict__ feats) {
  const int p = blockIdx.x / B, b = blockIdx.x % B;
  const float* base = a + ((size_t)b * P + p) * (size_t)V * ncol;
  __shared__ int s_warp[kThreads / 32];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int64_t row0 = offs[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int v0 = 0; v0 < V; v0 += kThreads) {
    const int v = v0 + threadIdx.x;
    float val = pad;
    const float* r = base + (size_t)v * ncol;
    if (v < V) val = r[ncol - 1];
    const bool keep = v < V && val != pad;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = s_base, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      const int c = s_warp[w];
      if (w < warp) before += c;
      total += c;
    }
    if (keep) {
      const int64_t row = row0 + before + __popc(m & ((1u << lane) - 1u));
      int4 c;
      if (layout == 0) c = make_int4((int)r[0], (int)r[1], ncol > 3 ? (int)r[2] : 0, b);
      else c = make_int4(p, (int)r[1], (int)r[0], b);
      *reinterpret_cast<int4*>(coords + row * 4) = c;
      feats[row] = val;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
}

}  // namespace

extern "C" int scn_larcv_count(const float* larcv, int B, int P, int V, int ncol, float pad_value, int32_t* counts,
                               void* stream) {
  if (!larcv || !counts || B < 1 || P < 1 || V < 0 || ncol < 2) return SCN_ERR_ARG;
  k_larcv_count<<<(unsigned)(B * P), kThreads, 0, (cudaStream_t)stream>>>(larcv, B, P, V, ncol, pad_value, counts);
  SCN_LAUNCH_CHECK();
  return SCN_OK;
}

extern "C" int scn_larcv_compact(const float* larcv, int B, int P, int V, int ncol, float pad_value,
                                 const int64_t* row_offsets, int layout, int32_t* coords4, float* features,
                                 void* stream) {
  if (!larcv || !row_offsets || B < 1 || P < 1 || V < 0 || ncol < 2 || (layout != 0 && layout != 1)) return SCN_ERR_ARG;
  if (layout == 1 && ncol < 3) return SCN_ERR_ARG;
  if (!coords4 || !features) return SCN_ERR_ARG;
  k_larcv_compact<<<(unsigned)(B * P), kThreads, 0, (cudaStream_t)stream>>>(larcv, B, P, V, ncol, pad_value, row_offsets,
                                                                            layout, coords4, features);
  SCN_LAUNCH_CHECK();
  return SCN_OK;
}
