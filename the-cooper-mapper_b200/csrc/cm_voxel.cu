// cm_voxel.cu -- K3: batched voxel-grid downsampling with pcl::VoxelGrid<PointXYZI> semantics.
//
// Replaces the pcl::VoxelGrid::filter calls at ScanRegistration.cpp:390-399 (per ring, leaf 0.2),
// LaserMatcher.cpp:293-300 (frame, filter_corner / filter_surf) and ScanMatch.cpp:381-394; indexing as stated
// in-tree by util/voxel_grid_partition.hpp:91-272: bounding box -> min_b = floor(min * inv_leaf) -> ijk =
// floor(p * inv_leaf) - min_b -> idx = i + j*dx + k*dx*dy -> sort by idx -> one centroid (all four fields) per
// occupied cell, output ordered by idx.  Canonical choice shared with the oracle: points of one cell are summed
// in INPUT order (stable sort; PCL's std::sort is unstable).  Non-finite points are skipped; when dx*dy*dz
// overflows int32 the input is passed through unchanged, like PCL.
//
// Segments ("streams") are independent clouds laid out as in[s * cap_in .. s * cap_in + n_in[s]).  The sort is a
// library primitive (cub::DeviceRadixSort, stable LSD) over 64-bit keys (segment << 32 | idx); everything else is
// hand-written.
#include "cm_host.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <float.h>

namespace cm {


__global__ void __launch_bounds__(1024) vox_bbox_kernel(const float4* __restrict__ in, const int* __restrict__ n_in, int cap_in,
                                                       float inv, VoxBox* __restrict__ box) {
  const int s = blockIdx.x;
  const int n = n_in[s];
  const float4* p = in + (size_t)s * cap_in;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int nf = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float4 q = p[i];
    if (isfinite(q.x) && isfinite(q.y) && isfinite(q.z)) {
      nf++;
      mn[0] = fminf(mn[0], q.x); mn[1] = fminf(mn[1], q.y); mn[2] = fminf(mn[2], q.z);
      mx[0] = fmaxf(mx[0], q.x); mx[1] = fmaxf(mx[1], q.y); mx[2] = fmaxf(mx[2], q.z);
    }
  }
  __shared__ float smn[32][3], smx[32][3];
  __shared__ int snf[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      mn[k] = fminf(mn[k], __shfl_down_sync(0xffffffffu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_down_sync(0xffffffffu, mx[k], o));
    }
    nf += __shfl_down_sync(0xffffffffu, nf, o);
  }
  if (lane == 0) { for (int k = 0; k < 3; k++) { smn[warp][k] = mn[k]; smx[warp][k] = mx[k]; } snf[warp] = nf; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
      for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], smn[w][k]); mx[k] = fmaxf(mx[k], smx[w][k]); }
      nf += snf[w];
    }
    VoxBox b;
    b.nfinite = nf; b.passthrough = 0; b.cells = 1;
    b.minb[0] = b.minb[1] = b.minb[2] = 0; b.mul1 = b.mul2 = 0;
    if (nf > 0) {
      // voxel_grid_partition.hpp:108-137
      long long dx = (long long)((mx[0] - mn[0]) * inv) + 1;
      long long dy = (long long)((mx[1] - mn[1]) * inv) + 1;
      long long dz = (long long)((mx[2] - mn[2]) * inv) + 1;
      if (dx * dy * dz > 2147483647LL) b.passthrough = 1;
      b.cells = b.passthrough ? (long long)n : dx * dy * dz;   // upper bound of the sort index (identity order when passing through)
      int maxb[3];
      for (int k = 0; k < 3; k++) { b.minb[k] = (int)floorf(mn[k] * inv); maxb[k] = (int)floorf(mx[k] * inv); }
      int d0 = maxb[0] - b.minb[0] + 1, d1 = maxb[1] - b.minb[1] + 1;
      b.mul1 = d0; b.mul2 = d0 * d1;
    }
    box[s] = b;
  }
}

#define CM_VOX_PAD 0xFFFFFFFFFFFFFFFFull

__global__ void vox_key_kernel(const float4* __restrict__ in, const int* __restrict__ n_in, int cap_in, int max_n, int nseg, float inv,
                               const VoxBox* __restrict__ box, int shift, unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (size_t)nseg * max_n) return;
  int s = (int)(g / max_n), i = (int)(g - (size_t)s * max_n);
  const size_t src = (size_t)s * cap_in + i;
  unsigned long long key = CM_VOX_PAD;
  if (i < n_in[s]) {
    const VoxBox b = box[s];
    float4 q = in[src];
    if (b.passthrough) {
      key = ((unsigned long long)s << shift) | (unsigned int)i;   // identity order
    } else if (isfinite(q.x) && isfinite(q.y) && isfinite(q.z)) {
      // voxel_grid_partition.hpp:212-226
      int ijk0 = (int)(floorf(q.x * inv) - (float)b.minb[0]);
      int ijk1 = (int)(floorf(q.y * inv) - (float)b.minb[1]);
      int ijk2 = (int)(floorf(q.z * inv) - (float)b.minb[2]);
      int idx = ijk0 + ijk1 * b.mul1 + ijk2 * b.mul2;
      key = ((unsigned long long)s << shift) | (unsigned int)idx;
    }
  }
  keys[g] = key;
  vals[g] = (unsigned int)src;
}

// head flag = first element of a (segment, idx) run among the sorted keys
__global__ void vox_head_kernel(const unsigned long long* __restrict__ keys, size_t n, int* __restrict__ flags) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  unsigned long long k = keys[g];
  flags[g] = (k != CM_VOX_PAD && (g == 0 || keys[g - 1] != k)) ? 1 : 0;
}

// seg_first[s] = position of the segment's first sorted element = number of valid elements of the segments before it
__global__ void vox_segstart_kernel(const int* __restrict__ n_in, const VoxBox* __restrict__ box, int nseg, int* __restrict__ seg_first) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int acc = 0;
    for (int s = 0; s < nseg; s++) { seg_first[s] = acc; acc += box[s].passthrough ? n_in[s] : box[s].nfinite; }
    seg_first[nseg] = acc;
  }
}

__global__ void vox_centroid_kernel(const float4* __restrict__ in, const unsigned long long* __restrict__ keys,
                                    const unsigned int* __restrict__ vals, const int* __restrict__ flags,
                                    const int* __restrict__ rank, const int* __restrict__ seg_first, size_t n, int nseg, int shift,
                                    float4* __restrict__ out, int cap_out, int* __restrict__ n_out, int* __restrict__ overflow) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  if (!flags[g]) return;
  unsigned long long k = keys[g];
  int s = (int)(k >> shift);
  int first = seg_first[s];
  int pos = rank[g] - rank[first];
  // segment total = heads in [first, seg_first[s+1]); the last head of the segment publishes it
  size_t end = (size_t)seg_first[s + 1];
  float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;
  size_t j = g;
  for (; j < end && keys[j] == k; j++) {
    float4 q = in[vals[j]];
    cx += q.x; cy += q.y; cz += q.z; ci += q.w;   // Eigen::VectorXf centroid += point, in sorted (= input) order
  }
  float cnt = (float)(j - g);
  if (j == end) n_out[s] = (pos + 1 <= cap_out) ? pos + 1 : cap_out;
  if (pos < cap_out) out[(size_t)s * cap_out + pos] = make_float4(cx / cnt, cy / cnt, cz / cnt, ci / cnt);
  else if (overflow) atomicExch(overflow, 1);
}

__global__ void vox_zero_counts_kernel(const int* __restrict__ n_in, const VoxBox* __restrict__ box, int nseg, int* __restrict__ n_out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) n_out[s] = 0;
}

void launch_vox_bbox(int nseg, const float4* d_in, const int* d_n_in, int cap_in, float leaf, VoxBox* d_box, cudaStream_t stream) {
  if (nseg <= 0 || cap_in <= 0) return;
  CM_LAUNCH(vox_bbox_kernel, nseg, 1024, 0, stream, d_in, d_n_in, cap_in, 1.0f / leaf, d_box);
}

// number of key bits that index `cells` distinct voxel indices
int vox_index_bits(long long cells) { int b = 1; while (b < 32 && (1LL << b) < cells) b++; return b; }

void VoxelFilter::run(int nseg, const float4* d_in, const int* d_n_in, int cap_in, int max_n, float leaf, float4* d_out, int* d_n_out,
                      int cap_out, int* d_overflow, cudaStream_t stream, const VoxBox* d_box_ready, int idx_bits) {
  if (nseg <= 0 || cap_in <= 0) return;
  if (max_n <= 0 || max_n > cap_in) max_n = cap_in;   // host-known upper bound of n_in[s]: only that many slots per segment are sorted
  const size_t n = (size_t)nseg * max_n;
  const float inv = 1.0f / leaf;   // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
  box.reserve(sizeof(VoxBox) * nseg);
  keys_a.reserve(n * 8); keys_b.reserve(n * 8); vals_a.reserve(n * 4); vals_b.reserve(n * 4);
  flags.reserve(n * 4); rank.reserve(n * 4); seg_first.reserve(sizeof(int) * (nseg + 1));
  // key = segment << shift | voxel index; the all-ones segment code is left to the padding key so that padding can never tie
  // with a real key inside the sorted bit range
  if (idx_bits < 1 || idx_bits > 32 || !d_box_ready) idx_bits = 32;
  const int shift = idx_bits;
  int sbits = 0;
  while ((1 << sbits) < nseg + 1) sbits++;
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (unsigned int*)nullptr,
                                  (unsigned int*)nullptr, (long long)n, 0, shift + sbits, stream);
  cub::DeviceScan::ExclusiveSum(nullptr, t2, (int*)nullptr, (int*)nullptr, (long long)n, stream);
  temp.reserve(t1 > t2 ? t1 : t2);
  const int T = 256;
  const unsigned int nb = (unsigned int)((n + T - 1) / T);
  const VoxBox* d_box = d_box_ready;
  if (!d_box) { CM_LAUNCH(vox_bbox_kernel, nseg, 1024, 0, stream, d_in, d_n_in, cap_in, inv, (VoxBox*)box.p); d_box = (const VoxBox*)box.p; }
  CM_LAUNCH(vox_zero_counts_kernel, (nseg + 63) / 64, 64, 0, stream, d_n_in, d_box, nseg, d_n_out);
  CM_LAUNCH(vox_key_kernel, nb, T, 0, stream, d_in, d_n_in, cap_in, max_n, nseg, inv, d_box, shift, (unsigned long long*)keys_a.p,
            (unsigned int*)vals_a.p);
  size_t tb = temp.cap;
  CM_TIMED("cub_radix_sort(voxel)", stream,
           cub::DeviceRadixSort::SortPairs(temp.p, tb, (const unsigned long long*)keys_a.p, (unsigned long long*)keys_b.p,
                                           (const unsigned int*)vals_a.p, (unsigned int*)vals_b.p, (long long)n, 0, shift + sbits, stream));
  g_launch_count += (shift + sbits + 7) / 8 + 1;   // onesweep: one histogram + one pass per 8 bits (library kernels)
  CM_LAUNCH(vox_head_kernel, nb, T, 0, stream, (const unsigned long long*)keys_b.p, n, (int*)flags.p);
  tb = temp.cap;
  CM_TIMED("cub_scan(voxel)", stream, cub::DeviceScan::ExclusiveSum(temp.p, tb, (const int*)flags.p, (int*)rank.p, (long long)n, stream));
  g_launch_count += 2;
  CM_LAUNCH(vox_segstart_kernel, 1, 32, 0, stream, d_n_in, d_box, nseg, (int*)seg_first.p);
  CM_LAUNCH(vox_centroid_kernel, nb, T, 0, stream, d_in, (const unsigned long long*)keys_b.p, (const unsigned int*)vals_b.p,
            (const int*)flags.p, (const int*)rank.p, (const int*)seg_first.p, n, nseg, shift, d_out, cap_out, d_n_out, d_overflow);
}

}  // namespace cm
