// cm_voxel.cu -- K3: batched voxel-grid downsampling with pcl::VoxelGrid<PointXYZI> semantics, ONE launch.
//
// Replaces the pcl::VoxelGrid::filter calls at ScanRegistration.cpp:390-399 (per ring, leaf 0.2: done inside sr_ring_kernel),
// LaserMatcher.cpp:293-300 (frame, filter_corner / filter_surf) and ScanMatch.cpp:381-394; indexing as stated
// in-tree by util/voxel_grid_partition.hpp:91-272: bounding box -> min_b = floor(min * inv_leaf) -> ijk =
// floor(p * inv_leaf) - min_b -> idx = i + j*dx + k*dx*dy -> sort by idx -> one centroid (all four fields) per
// occupied cell, output ordered by idx.  Canonical choice shared with the oracle: points of one cell are summed
// in INPUT order (stable sort; PCL's std::sort is unstable).  Non-finite points are skipped; when dx*dy*dz
// overflows int32 the input is passed through unchanged, like PCL.
//
// Segments ("streams") are independent clouds laid out as in[s * cap_in .. s * cap_in + n_in[s]).  One CTA of 1024
// threads owns one segment from bounding box to centroids, so a whole batch (both feature classes of every stream) is
// a single launch whose sizes live on the device only:
//   1. bounding box (block reduction);
//   2. voxel index per point, RUNS of consecutive points with equal index (a LiDAR feature cloud follows the scan: ~2.6
//      points per run), ordered compaction of the runs by ballot prefix sums, digit histograms;
//   3. stable LSD radix sort of the runs by voxel index, 8 bits per pass, only as many passes as the segment's index space
//      needs (3 for a 240 m frame at 0.8 m): per tile of 1024 runs the rank of an item is (digit base) + (runs of the
//      same digit in earlier warps, a 32 x 256 counter matrix in shared memory) + (earlier lanes of its warp,
//      __match_any_sync); runs with equal index stay in input order, which is what makes the centroid sums reproducible;
//   4. first run of every voxel -> output rank (ballot prefix sums) -> the head's thread adds up the voxel's runs in order.
// (Round 1 used cub::DeviceRadixSort over all segments + six helper launches: 486 us per step at 64 HDL-64 streams.)
#include "cm_host.h"
#include <float.h>
#include <cooperative_groups.h>

namespace cm {

#define VS_T 1024
#define VS_WARPS (VS_T / 32)
#define VS_PAD 0xFFFFFFFFu
#define VS_RC 16384          // runs whose sort buffers fit in shared memory (2 x u32 keys + 2 x u16 values = 192 KB)
#define VS_DYN_SMEM ((size_t)VS_RC * 12)   // dynamic shared memory: the sort buffers, later the point stage of the centroid phase
#define VS_ARRAYS 7          // scratch arrays per segment: key A / B, value A / B, run start, run end, point order

struct VoxClass {            // one batch of segments filtered with one leaf
  const float4* in; const int* n_in; int cap_in;
  float inv;                 // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
  float4* out; int* n_out; int cap_out;
  unsigned int* scratch;     // [nseg][VS_ARRAYS][stride]
};
struct VoxSegArgs {
  VoxClass cls[2];
  int nseg, ncls;
  unsigned int stride;       // scratch entries per array and segment (> the largest n_in)
  int* overflow;             // optional flag: an output exceeded cap_out (or an input the scratch stride)
  VoxBox* box_out;           // optional [ncls][nseg]: the bounding boxes (diagnostics)
};

// voxel_grid_partition.hpp:212-226
__device__ __forceinline__ unsigned int vox_index_of(const float4& q, const VoxBox& b, float inv, unsigned int i) {
  if (b.passthrough) return i;   // identity order
  if (!(isfinite(q.x) && isfinite(q.y) && isfinite(q.z))) return VS_PAD;
  int ijk0 = (int)(floorf(q.x * inv) - (float)b.minb[0]);
  int ijk1 = (int)(floorf(q.y * inv) - (float)b.minb[1]);
  int ijk2 = (int)(floorf(q.z * inv) - (float)b.minb[2]);
  return (unsigned int)(ijk0 + ijk1 * b.mul1 + ijk2 * b.mul2);
}

struct VoxShared {
  VoxBox box;
  float red[VS_WARPS][6];
  int redn[VS_WARPS];
  unsigned int idx[VS_T + 2];
  unsigned int carry;
  int wt[2][VS_WARPS];
  int wl[2][VS_WARPS];
  unsigned int hist[4][256];
  unsigned int digit[256];
  unsigned int wsum[8];
  unsigned short gsum[4][256];
  unsigned char wcnt[VS_WARPS][256];
};

// Stable LSD radix sort of (key, value) pairs by the low 8 * npass key bits, ping-pong between buffers 0 and 1 (in shared or
// global memory: generic pointers).  Pass 0 takes value r for item r (the identity), so the caller fills key[0] only.
// Returns the buffer that holds the result.  hist[p] = digit histogram of pass p (filled by the caller).
template <typename V>
__device__ __forceinline__ int vox_lsd_sort(VoxShared& sh, unsigned int* key0, unsigned int* key1, V* val0, V* val1, unsigned int nruns, int npass) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int lt = (1u << lane) - 1u;
  int cur = 0;
  for (int p = 0; p < npass; p++) {
    const int shift = 8 * p;
    if (tid < 256) {   // exclusive scan of the digit histogram
      const unsigned int v = sh.hist[p][tid];
      unsigned int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      if (lane == 31) sh.wsum[warp] = x;
      sh.digit[tid] = x - v;
    }
    __syncthreads();
    if (tid < 256) {
      unsigned int add = 0;
      for (int w = 0; w < warp; w++) add += sh.wsum[w];
      sh.digit[tid] += add;
    }
    // (the first barrier of the tile loop orders these writes before their first use)
    const unsigned int* sk = cur ? key1 : key0; const V* sv = cur ? val1 : val0;
    unsigned int* dk = cur ? key0 : key1; V* dv = cur ? val0 : val1;
    for (unsigned int t0 = 0; t0 < nruns; t0 += VS_T) {
      const unsigned int r = t0 + tid;
      const bool ok = r < nruns;
      const unsigned int key = ok ? sk[r] : 0u;
      const unsigned int val = ok ? (p == 0 ? r : (unsigned int)sv[r]) : 0u;
      const unsigned int d = ok ? ((key >> shift) & 255u) : (256u + (unsigned int)lane);
#pragma unroll
      for (int q = 0; q < 2; q++) reinterpret_cast<unsigned int*>(&sh.wcnt[0][0])[tid + q * VS_T] = 0u;
      __syncthreads();
      const unsigned int m = __match_any_sync(0xffffffffu, d);
      const unsigned int rank_w = __popc(m & lt);
      if (ok && rank_w == 0) sh.wcnt[warp][d] = (unsigned char)__popc(m);   // <= 32
      __syncthreads();
      {   // per digit: exclusive prefix over the 32 warps, four threads per digit (8 warps each) + their group totals
        const int d2 = tid & 255, g = tid >> 8;
        unsigned int acc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const unsigned int c = sh.wcnt[g * 8 + w][d2]; sh.wcnt[g * 8 + w][d2] = (unsigned char)acc; acc += c; }   // acc <= 224 before the last add
        sh.gsum[g][d2] = (unsigned short)acc;
      }
      __syncthreads();
      if (ok) {
        unsigned int pos = sh.digit[d] + sh.wcnt[warp][d] + rank_w;
        const int g = warp >> 3;
        if (g > 0) pos += sh.gsum[0][d];
        if (g > 1) pos += sh.gsum[1][d];
        if (g > 2) pos += sh.gsum[2][d];
        dk[pos] = key; dv[pos] = (V)val;
      }
      __syncthreads();
      if (tid < 256) sh.digit[tid] += (unsigned int)sh.gsum[0][tid] + sh.gsum[1][tid] + sh.gsum[2][tid] + sh.gsum[3][tid];
    }
    __syncthreads();
    cur ^= 1;
  }
  return cur;
}

extern __shared__ __align__(16) unsigned char vox_dyn_smem[];

__global__ void __launch_bounds__(VS_T, 1) vox_segment_kernel(VoxSegArgs a) {
  const int s = blockIdx.x, ci = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const VoxClass& k = a.cls[ci];
  const float inv = k.inv;
  const float4* in = k.in + (size_t)s * k.cap_in;
  int n = k.n_in[s];
  if (n < 0) n = 0;
  if ((unsigned int)n >= a.stride) { n = (int)a.stride - 1; if (tid == 0 && a.overflow) atomicExch(a.overflow, 1); }
  unsigned int* base = k.scratch + (size_t)s * VS_ARRAYS * a.stride;
  unsigned int* gkey0 = base; unsigned int* gkey1 = base + a.stride;
  unsigned int* gval0 = base + 2 * (size_t)a.stride; unsigned int* gval1 = base + 3 * (size_t)a.stride;
  unsigned int* run_start = base + 4 * (size_t)a.stride;
  unsigned int* run_end = base + 5 * (size_t)a.stride;
  unsigned int* order = base + 6 * (size_t)a.stride;

  __shared__ VoxShared sh;
  const unsigned int lt = (1u << lane) - 1u;

  // ---- 1. bounding box (voxel_grid_partition.hpp:108-137) ------------------------------------------------------------
  {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int nf = 0;
    for (int i = tid; i < n; i += VS_T) {
      const float4 q = in[i];
      if (isfinite(q.x) && isfinite(q.y) && isfinite(q.z)) {
        nf++;
        mn[0] = fminf(mn[0], q.x); mn[1] = fminf(mn[1], q.y); mn[2] = fminf(mn[2], q.z);
        mx[0] = fmaxf(mx[0], q.x); mx[1] = fmaxf(mx[1], q.y); mx[2] = fmaxf(mx[2], q.z);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
        mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
      }
      nf += __shfl_xor_sync(0xffffffffu, nf, o);
    }
    if (lane == 0) { for (int c = 0; c < 3; c++) { sh.red[warp][c] = mn[c]; sh.red[warp][3 + c] = mx[c]; } sh.redn[warp] = nf; }
    for (int q = tid; q < 4 * 256; q += VS_T) (&sh.hist[0][0])[q] = 0u;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < VS_WARPS; w++) {
        for (int c = 0; c < 3; c++) { mn[c] = fminf(mn[c], sh.red[w][c]); mx[c] = fmaxf(mx[c], sh.red[w][3 + c]); }
        nf += sh.redn[w];
      }
      VoxBox b;
      b.nfinite = nf; b.passthrough = 0; b.cells = 1;
      b.minb[0] = b.minb[1] = b.minb[2] = 0; b.mul1 = b.mul2 = 0;
      if (nf > 0) {
        long long dx = (long long)((mx[0] - mn[0]) * inv) + 1;
        long long dy = (long long)((mx[1] - mn[1]) * inv) + 1;
        long long dz = (long long)((mx[2] - mn[2]) * inv) + 1;
        if (dx * dy * dz > 2147483647LL) b.passthrough = 1;
        b.cells = b.passthrough ? (long long)n : dx * dy * dz;   // upper bound of the sort index (identity order when passing through)
        int maxb[3];
        for (int c = 0; c < 3; c++) { b.minb[c] = (int)floorf(mn[c] * inv); maxb[c] = (int)floorf(mx[c] * inv); }
        int d0 = maxb[0] - b.minb[0] + 1, d1 = maxb[1] - b.minb[1] + 1;
        b.mul1 = d0; b.mul2 = d0 * d1;
        // the index space actually addressed (floor of the scaled extrema) can exceed the truncated estimate by a cell per axis
        const long long span = (long long)d0 * d1 * (long long)(maxb[2] - b.minb[2] + 1);
        if (!b.passthrough && span > b.cells) b.cells = span;
      }
      sh.box = b;
      if (a.box_out) a.box_out[ci * a.nseg + s] = b;
      sh.carry = VS_PAD;
    }
    __syncthreads();
  }
  const VoxBox box = sh.box;
  int bits = 1;
  while (bits < 32 && (1LL << bits) < box.cells) bits++;
  const int npass = (bits + 7) / 8;
  if (box.nfinite == 0 && !box.passthrough) { if (tid == 0) k.n_out[s] = 0; return; }

  // ---- 2. runs of equal voxel index, in input order ---------------------------------------------------------------------
  unsigned int nruns = 0, nends = 0;   // a run that spans a tile boundary has started but not ended: two counters
  {
    int buf = 0;
    for (int t0 = 0; t0 < n; t0 += VS_T) {
      const int i = t0 + tid;
      unsigned int idx = VS_PAD;
      if (i < n) idx = vox_index_of(in[i], box, inv, (unsigned int)i);
      sh.idx[1 + tid] = idx;
      if (tid == 0) {
        sh.idx[0] = sh.carry;
        const int j = t0 + VS_T;
        sh.idx[1 + VS_T] = j < n ? vox_index_of(in[j], box, inv, (unsigned int)j) : VS_PAD;
      }
      __syncthreads();
      const bool valid = idx != VS_PAD;
      const bool st = valid && sh.idx[tid] != idx, en = valid && sh.idx[tid + 2] != idx;
      const unsigned int b0 = __ballot_sync(0xffffffffu, st), b1 = __ballot_sync(0xffffffffu, en);
      if (lane == 0) sh.wt[buf][warp] = __popc(b0) | (__popc(b1) << 16);
      if (tid == VS_T - 1) sh.carry = idx;
      __syncthreads();
      const int v = sh.wt[buf][lane];
      const int before = __reduce_add_sync(0xffffffffu, lane < warp ? v : 0);
      const int tot = __reduce_add_sync(0xffffffffu, v);
      const unsigned int ps = nruns + (before & 0xFFFF) + __popc(b0 & lt), pe = nends + (before >> 16) + __popc(b1 & lt);
      if (st) {
        gkey0[ps] = idx; run_start[ps] = (unsigned int)i;
        for (int p = 0; p < npass; p++) atomicAdd(&sh.hist[p][(idx >> (8 * p)) & 255u], 1u);
      }
      if (en) run_end[pe] = (unsigned int)i + 1u;
      nruns += (unsigned int)(tot & 0xFFFF); nends += (unsigned int)(tot >> 16);
      buf ^= 1;
    }
  }
  __syncthreads();

  // ---- 3. stable LSD radix sort of the runs by voxel index: in shared memory when they fit, else in the global scratch -----
  const unsigned int* sk; const void* sv; bool v16;
  unsigned int* voff;   // start of every voxel's points in `order` (a free key buffer)
  if (nruns <= VS_RC) {
    unsigned int* k0 = reinterpret_cast<unsigned int*>(vox_dyn_smem); unsigned int* k1 = k0 + VS_RC;
    unsigned short* v0 = reinterpret_cast<unsigned short*>(k1 + VS_RC); unsigned short* v1 = v0 + VS_RC;
    for (unsigned int r = tid; r < nruns; r += VS_T) k0[r] = gkey0[r];
    __syncthreads();
    const int cur = vox_lsd_sort<unsigned short>(sh, k0, k1, v0, v1, nruns, npass);
    sk = cur ? k1 : k0; sv = cur ? (const void*)v1 : (const void*)v0; v16 = true;
    if (npass == 0) { for (unsigned int r = tid; r < nruns; r += VS_T) v0[r] = (unsigned short)r; __syncthreads(); }
    voff = gkey1;
  } else {
    const int cur = vox_lsd_sort<unsigned int>(sh, gkey0, gkey1, gval0, gval1, nruns, npass);
    sk = cur ? gkey1 : gkey0; sv = cur ? (const void*)gval1 : (const void*)gval0; v16 = false;
    if (npass == 0) { for (unsigned int r = tid; r < nruns; r += VS_T) gval0[r] = r; __syncthreads(); }
    voff = cur ? gkey0 : gkey1;
  }

  // ---- 4. flatten: the points of every voxel, in (voxel, input) order, as one index list; voxel starts ------------------------
  unsigned int nvox = 0, npts = 0;
  {
    int buf = 0;
    for (unsigned int t0 = 0; t0 < nruns; t0 += VS_T) {
      const unsigned int r = t0 + tid;
      const bool ok = r < nruns;
      const unsigned int key = ok ? sk[r] : VS_PAD;
      const bool head = ok && (r == 0 || sk[r - 1] != key);
      unsigned int i0 = 0, len = 0;
      if (ok) {
        const unsigned int id = v16 ? (unsigned int)reinterpret_cast<const unsigned short*>(sv)[r] : reinterpret_cast<const unsigned int*>(sv)[r];
        i0 = run_start[id]; len = run_end[id] - i0;
      }
      const unsigned int b0 = __ballot_sync(0xffffffffu, head);
      unsigned int x = len;   // inclusive warp scan of the run lengths
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      if (lane == 0) sh.wt[buf][warp] = __popc(b0);
      if (lane == 31) sh.wl[buf][warp] = (int)x;
      __syncthreads();
      const int vh = sh.wt[buf][lane], vl = sh.wl[buf][lane];
      const int hbefore = __reduce_add_sync(0xffffffffu, lane < warp ? vh : 0), htot = __reduce_add_sync(0xffffffffu, vh);
      const int lbefore = __reduce_add_sync(0xffffffffu, lane < warp ? vl : 0), ltot = __reduce_add_sync(0xffffffffu, vl);
      const unsigned int off = npts + (unsigned int)lbefore + (x - len);
      if (head) voff[nvox + (unsigned int)hbefore + __popc(b0 & lt)] = off;
      for (unsigned int j = 0; j < len; j++) order[off + j] = i0 + j;
      nvox += (unsigned int)htot; npts += (unsigned int)ltot;
      buf ^= 1;
    }
    if (tid == 0) voff[nvox] = npts;
  }
  __syncthreads();

  // ---- 5. centroids.  The points of up to 1024 consecutive voxels (contiguous in `order`) are gathered into shared memory by the
  //      whole CTA -- independent loads, coalesced over `order` -- then ONE thread per voxel adds its points up sequentially out of
  //      shared memory: Eigen::VectorXf centroid += point in sorted (= input) order.  (The sort buffers are free by now.) ------------
  float4* out = k.out + (size_t)s * k.cap_out;
  {
    float4* stage = reinterpret_cast<float4*>(vox_dyn_smem);
    const unsigned int PB = (unsigned int)(VS_DYN_SMEM / sizeof(float4));   // points per round
    for (unsigned int v0 = 0; v0 < nvox;) {
      const unsigned int pbase = voff[v0];
      const unsigned int v = v0 + (unsigned int)tid;
      const bool fits = v < nvox && voff[v + 1] - pbase <= PB;              // monotone in tid: the voxels of this round are a prefix
      const unsigned int nv = (unsigned int)__syncthreads_count(fits ? 1 : 0);
      if (nv == 0) {   // one voxel with more points than the stage holds: its sum straight from global memory
        if (tid == 0) {
          const unsigned int pe = voff[v0 + 1];
          float cx = 0.f, cy = 0.f, cz = 0.f, cw = 0.f;
          for (unsigned int p = pbase; p < pe; p++) { const float4 q = in[order[p]]; cx += q.x; cy += q.y; cz += q.z; cw += q.w; }
          const float c = (float)(pe - pbase);
          if (v0 < (unsigned int)k.cap_out) out[v0] = make_float4(cx / c, cy / c, cz / c, cw / c);
          else if (a.overflow) atomicExch(a.overflow, 1);
        }
        v0 += 1;
        continue;
      }
      const unsigned int npnt = voff[v0 + nv] - pbase;
      for (unsigned int p = tid; p < npnt; p += VS_T) stage[p] = in[order[pbase + p]];
      __syncthreads();
      if ((unsigned int)tid < nv) {
        const unsigned int pa = voff[v] - pbase, pe = voff[v + 1] - pbase;
        float cx = 0.f, cy = 0.f, cz = 0.f, cw = 0.f;
        for (unsigned int p = pa; p < pe; p++) { const float4 q = stage[p]; cx += q.x; cy += q.y; cz += q.z; cw += q.w; }
        const float c = (float)(pe - pa);
        if (v < (unsigned int)k.cap_out) out[v] = make_float4(cx / c, cy / c, cz / c, cw / c);
        else if (a.overflow) atomicExch(a.overflow, 1);
      }
      __syncthreads();   // the next round overwrites the stage
      v0 += nv;
    }
  }
  if (tid == 0) k.n_out[s] = nvox <= (unsigned int)k.cap_out ? (int)nvox : k.cap_out;
}

// ---- the same filter for a FEW segments: a cluster of VC CTAs per segment ---------------------------------------------------------
// One CTA per segment leaves 146 SMs idle when a single sweep is filtered (one stream, latency mode: 177 us of a 0.75 ms step).
// Here VC = 8 CTAs of one thread-block cluster share a segment: every phase of vox_segment_kernel is cut into VC slices in INPUT
// order, the slices are stitched together through distributed shared memory (per-CTA counts, read by every CTA after a cluster
// barrier) and the run / key / order arrays in global memory (L2).  Same arithmetic, same stable order, same output bytes:
//   1. bounding box: per-CTA extrema -> every CTA combines the VC partial boxes;
//   2. runs: a counting pass over the slice, cluster-wide exclusive prefix of (starts, ends), a writing pass;
//   3. LSD radix sort over the cluster: per pass every CTA histograms ITS chunk of the current key array, the digit bases are
//      (digits below, all CTAs) + (same digit, lower-ranked CTAs) -- chunks are in array order, so the pass stays stable -- then the
//      usual tile ranking, scattering into the other global buffer;
//   4. flatten: count heads / points per chunk, prefix, write voxel starts and the point order;
//   5. centroids: the voxels are split over the CTAs.
#define VC 8
namespace cg = cooperative_groups;
struct VoxXch { float mn[3], mx[3]; int nf; unsigned int nruns, nends, nvox, npts; unsigned int hist[256]; };

__global__ void __cluster_dims__(VC, 1, 1) __launch_bounds__(VS_T, 1) vox_cluster_kernel(VoxSegArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int s = blockIdx.x / VC, ci = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const VoxClass& k = a.cls[ci];
  const float inv = k.inv;
  const float4* in = k.in + (size_t)s * k.cap_in;
  int n = k.n_in[s];
  if (n < 0) n = 0;
  if ((unsigned int)n >= a.stride) { n = (int)a.stride - 1; if (tid == 0 && rank == 0 && a.overflow) atomicExch(a.overflow, 1); }
  unsigned int* base = k.scratch + (size_t)s * VS_ARRAYS * a.stride;
  unsigned int* gkey[2] = {base, base + a.stride};
  unsigned int* gval[2] = {base + 2 * (size_t)a.stride, base + 3 * (size_t)a.stride};
  unsigned int* run_start = base + 4 * (size_t)a.stride;
  unsigned int* run_end = base + 5 * (size_t)a.stride;
  unsigned int* order = base + 6 * (size_t)a.stride;

  __shared__ VoxShared sh;
  __shared__ VoxXch xc;
  const VoxXch* peer[VC];
#pragma unroll
  for (int c = 0; c < VC; c++) peer[c] = cluster.map_shared_rank(&xc, c);
  const unsigned int lt = (1u << lane) - 1u;
  // input slice of this CTA: whole tiles
  const int per = (((n + VC - 1) / VC) + VS_T - 1) & ~(VS_T - 1);
  const int i_lo = min(n, rank * per), i_hi = min(n, i_lo + per);

  // ---- 1. bounding box ----------------------------------------------------------------------------------------------------------
  {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int nf = 0;
    for (int i = i_lo + tid; i < i_hi; i += VS_T) {
      const float4 q = in[i];
      if (isfinite(q.x) && isfinite(q.y) && isfinite(q.z)) {
        nf++;
        mn[0] = fminf(mn[0], q.x); mn[1] = fminf(mn[1], q.y); mn[2] = fminf(mn[2], q.z);
        mx[0] = fmaxf(mx[0], q.x); mx[1] = fmaxf(mx[1], q.y); mx[2] = fmaxf(mx[2], q.z);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
        mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
      }
      nf += __shfl_xor_sync(0xffffffffu, nf, o);
    }
    if (lane == 0) { for (int c = 0; c < 3; c++) { sh.red[warp][c] = mn[c]; sh.red[warp][3 + c] = mx[c]; } sh.redn[warp] = nf; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < VS_WARPS; w++) {
        for (int c = 0; c < 3; c++) { mn[c] = fminf(mn[c], sh.red[w][c]); mx[c] = fmaxf(mx[c], sh.red[w][3 + c]); }
        nf += sh.redn[w];
      }
      for (int c = 0; c < 3; c++) { xc.mn[c] = mn[c]; xc.mx[c] = mx[c]; }
      xc.nf = nf;
    }
    cluster.sync();
    if (tid == 0) {
      for (int c = 0; c < 3; c++) { mn[c] = FLT_MAX; mx[c] = -FLT_MAX; }
      nf = 0;
      for (int r = 0; r < VC; r++) {
        for (int c = 0; c < 3; c++) { mn[c] = fminf(mn[c], peer[r]->mn[c]); mx[c] = fmaxf(mx[c], peer[r]->mx[c]); }
        nf += peer[r]->nf;
      }
      VoxBox b;
      b.nfinite = nf; b.passthrough = 0; b.cells = 1;
      b.minb[0] = b.minb[1] = b.minb[2] = 0; b.mul1 = b.mul2 = 0;
      if (nf > 0) {
        long long dx = (long long)((mx[0] - mn[0]) * inv) + 1;
        long long dy = (long long)((mx[1] - mn[1]) * inv) + 1;
        long long dz = (long long)((mx[2] - mn[2]) * inv) + 1;
        if (dx * dy * dz > 2147483647LL) b.passthrough = 1;
        b.cells = b.passthrough ? (long long)n : dx * dy * dz;
        int maxb[3];
        for (int c = 0; c < 3; c++) { b.minb[c] = (int)floorf(mn[c] * inv); maxb[c] = (int)floorf(mx[c] * inv); }
        int d0 = maxb[0] - b.minb[0] + 1, d1 = maxb[1] - b.minb[1] + 1;
        b.mul1 = d0; b.mul2 = d0 * d1;
        const long long span = (long long)d0 * d1 * (long long)(maxb[2] - b.minb[2] + 1);
        if (!b.passthrough && span > b.cells) b.cells = span;
      }
      sh.box = b;
      if (a.box_out && rank == 0) a.box_out[ci * a.nseg + s] = b;
    }
    __syncthreads();
  }
  const VoxBox box = sh.box;
  int bits = 1;
  while (bits < 32 && (1LL << bits) < box.cells) bits++;
  const int npass = (bits + 7) / 8;
  if (box.nfinite == 0 && !box.passthrough) {
    if (tid == 0 && rank == 0) k.n_out[s] = 0;
    cluster.sync();   // nobody leaves while a peer may still read its shared memory
    return;
  }

  // ---- 2. runs of equal voxel index: count, prefix over the cluster, write ----------------------------------------------------------
  unsigned int run_base = 0, end_base = 0, nruns = 0;
  for (int phase = 0; phase < 2; phase++) {
    unsigned int cr = 0, ce = 0;
    int buf = 0;
    if (tid == 0) sh.carry = i_lo > 0 ? vox_index_of(in[i_lo - 1], box, inv, (unsigned int)(i_lo - 1)) : VS_PAD;
    __syncthreads();
    for (int t0 = i_lo; t0 < i_hi; t0 += VS_T) {
      const int i = t0 + tid;
      unsigned int idx = VS_PAD;
      if (i < n) idx = vox_index_of(in[i], box, inv, (unsigned int)i);
      sh.idx[1 + tid] = idx;
      if (tid == 0) {
        sh.idx[0] = sh.carry;
        const int j = t0 + VS_T;
        sh.idx[1 + VS_T] = j < n ? vox_index_of(in[j], box, inv, (unsigned int)j) : VS_PAD;
      }
      __syncthreads();
      const bool valid = idx != VS_PAD;
      const bool st = valid && sh.idx[tid] != idx, en = valid && sh.idx[tid + 2] != idx;
      const unsigned int b0 = __ballot_sync(0xffffffffu, st), b1 = __ballot_sync(0xffffffffu, en);
      if (lane == 0) sh.wt[buf][warp] = __popc(b0) | (__popc(b1) << 16);
      if (tid == VS_T - 1) sh.carry = idx;
      __syncthreads();
      const int v = sh.wt[buf][lane];
      const int before = __reduce_add_sync(0xffffffffu, lane < warp ? v : 0);
      const int tot = __reduce_add_sync(0xffffffffu, v);
      if (phase == 1) {
        const unsigned int ps = run_base + cr + (before & 0xFFFF) + __popc(b0 & lt), pe = end_base + ce + (before >> 16) + __popc(b1 & lt);
        if (st) { gkey[0][ps] = idx; run_start[ps] = (unsigned int)i; }
        if (en) run_end[pe] = (unsigned int)i + 1u;
      }
      cr += (unsigned int)(tot & 0xFFFF); ce += (unsigned int)(tot >> 16);
      buf ^= 1;
    }
    if (phase == 0) {
      if (tid == 0) { xc.nruns = cr; xc.nends = ce; }
      cluster.sync();
      for (int r = 0; r < VC; r++) {
        const unsigned int pr = peer[r]->nruns, pe = peer[r]->nends;
        if (r < rank) { run_base += pr; end_base += pe; }
        nruns += pr;
      }
    }
    __syncthreads();
  }
  cluster.sync();   // the run arrays are complete

  // ---- 3. stable LSD radix sort of the runs over the cluster ------------------------------------------------------------------------
  const unsigned int rper = (nruns + VC - 1) / VC;
  const unsigned int r_lo = min(nruns, (unsigned int)rank * rper), r_hi = min(nruns, r_lo + rper);
  int cur = 0;
  for (int p = 0; p < npass; p++) {
    const int shift = 8 * p;
    const unsigned int* sk = gkey[cur]; const unsigned int* sv = gval[cur];
    unsigned int* dk = gkey[cur ^ 1]; unsigned int* dv = gval[cur ^ 1];
    if (tid < 256) xc.hist[tid] = 0u;
    __syncthreads();
    for (unsigned int r = r_lo + tid; r < r_hi; r += VS_T) atomicAdd(&xc.hist[(sk[r] >> shift) & 255u], 1u);
    cluster.sync();
    if (tid < 256) {   // digit base = (all CTAs, lower digits) + (lower-ranked CTAs, this digit)
      unsigned int tot = 0, mine = 0;
#pragma unroll
      for (int r = 0; r < VC; r++) { const unsigned int h = peer[r]->hist[tid]; tot += h; if (r < rank) mine += h; }
      unsigned int x = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      if (lane == 31) sh.wsum[warp] = x;
      sh.digit[tid] = x - tot + mine;
    }
    __syncthreads();
    if (tid < 256) {
      unsigned int add = 0;
      for (int w = 0; w < warp; w++) add += sh.wsum[w];
      sh.digit[tid] += add;
    }
    for (unsigned int t0 = r_lo; t0 < r_hi; t0 += VS_T) {
      const unsigned int r = t0 + tid;
      const bool ok = r < r_hi;
      const unsigned int key = ok ? sk[r] : 0u;
      const unsigned int val = ok ? (p == 0 ? r : sv[r]) : 0u;
      const unsigned int d = ok ? ((key >> shift) & 255u) : (256u + (unsigned int)lane);
#pragma unroll
      for (int q = 0; q < 2; q++) reinterpret_cast<unsigned int*>(&sh.wcnt[0][0])[tid + q * VS_T] = 0u;
      __syncthreads();
      const unsigned int m = __match_any_sync(0xffffffffu, d);
      const unsigned int rank_w = __popc(m & lt);
      if (ok && rank_w == 0) sh.wcnt[warp][d] = (unsigned char)__popc(m);
      __syncthreads();
      {
        const int d2 = tid & 255, g = tid >> 8;
        unsigned int acc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const unsigned int c = sh.wcnt[g * 8 + w][d2]; sh.wcnt[g * 8 + w][d2] = (unsigned char)acc; acc += c; }
        sh.gsum[g][d2] = (unsigned short)acc;
      }
      __syncthreads();
      if (ok) {
        unsigned int pos = sh.digit[d] + sh.wcnt[warp][d] + rank_w;
        const int g = warp >> 3;
        if (g > 0) pos += sh.gsum[0][d];
        if (g > 1) pos += sh.gsum[1][d];
        if (g > 2) pos += sh.gsum[2][d];
        dk[pos] = key; dv[pos] = val;
      }
      __syncthreads();
      if (tid < 256) sh.digit[tid] += (unsigned int)sh.gsum[0][tid] + sh.gsum[1][tid] + sh.gsum[2][tid] + sh.gsum[3][tid];
    }
    cluster.sync();   // the pass is complete everywhere (and every peer has read this CTA's histogram)
    cur ^= 1;
  }
  const unsigned int* sk = gkey[cur]; const unsigned int* sv = gval[cur];
  unsigned int* voff = gkey[cur ^ 1];

  // ---- 4. flatten: voxel starts and the point order ------------------------------------------------------------------------------
  unsigned int nvox = 0, npts = 0, vox_base = 0, pts_base = 0;
  for (int phase = 0; phase < 2; phase++) {
    unsigned int cv = 0, cp = 0;
    int buf = 0;
    for (unsigned int t0 = r_lo; t0 < r_hi; t0 += VS_T) {
      const unsigned int r = t0 + tid;
      const bool ok = r < r_hi;
      const unsigned int key = ok ? sk[r] : VS_PAD;
      const bool head = ok && (r == 0 || sk[r - 1] != key);
      unsigned int i0 = 0, len = 0;
      if (ok) { const unsigned int id = sv[r]; i0 = run_start[id]; len = run_end[id] - i0; }
      const unsigned int b0 = __ballot_sync(0xffffffffu, head);
      unsigned int x = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      if (lane == 0) sh.wt[buf][warp] = __popc(b0);
      if (lane == 31) sh.wl[buf][warp] = (int)x;
      __syncthreads();
      const int vh = sh.wt[buf][lane], vl = sh.wl[buf][lane];
      const int hbefore = __reduce_add_sync(0xffffffffu, lane < warp ? vh : 0), htot = __reduce_add_sync(0xffffffffu, vh);
      const int lbefore = __reduce_add_sync(0xffffffffu, lane < warp ? vl : 0), ltot = __reduce_add_sync(0xffffffffu, vl);
      if (phase == 1) {
        const unsigned int off = pts_base + cp + (unsigned int)lbefore + (x - len);
        if (head) voff[vox_base + cv + (unsigned int)hbefore + __popc(b0 & lt)] = off;
        for (unsigned int j = 0; j < len; j++) order[off + j] = i0 + j;
      }
      cv += (unsigned int)htot; cp += (unsigned int)ltot;
      buf ^= 1;
    }
    if (phase == 0) {
      if (tid == 0) { xc.nvox = cv; xc.npts = cp; }
      cluster.sync();
      for (int r = 0; r < VC; r++) {
        const unsigned int pv = peer[r]->nvox, pp = peer[r]->npts;
        if (r < rank) { vox_base += pv; pts_base += pp; }
        nvox += pv; npts += pp;
      }
    }
    __syncthreads();
  }
  if (tid == 0 && rank == 0) voff[nvox] = npts;
  cluster.sync();

  // ---- 5. centroids of this CTA's share of the voxels -----------------------------------------------------------------------------
  float4* out = k.out + (size_t)s * k.cap_out;
  {
    float4* stage = reinterpret_cast<float4*>(vox_dyn_smem);
    const unsigned int PB = (unsigned int)(VS_DYN_SMEM / sizeof(float4));
    const unsigned int vper = (nvox + VC - 1) / VC;
    const unsigned int v_lo = min(nvox, (unsigned int)rank * vper), v_hi = min(nvox, v_lo + vper);
    for (unsigned int v0 = v_lo; v0 < v_hi;) {
      const unsigned int pbase = voff[v0];
      const unsigned int v = v0 + (unsigned int)tid;
      const bool fits = v < v_hi && voff[v + 1] - pbase <= PB;
      const unsigned int nv = (unsigned int)__syncthreads_count(fits ? 1 : 0);
      if (nv == 0) {
        if (tid == 0) {
          const unsigned int pe = voff[v0 + 1];
          float cx = 0.f, cy = 0.f, cz = 0.f, cw = 0.f;
          for (unsigned int p = pbase; p < pe; p++) { const float4 q = in[order[p]]; cx += q.x; cy += q.y; cz += q.z; cw += q.w; }
          const float c = (float)(pe - pbase);
          if (v0 < (unsigned int)k.cap_out) out[v0] = make_float4(cx / c, cy / c, cz / c, cw / c);
          else if (a.overflow) atomicExch(a.overflow, 1);
        }
        v0 += 1;
        continue;
      }
      const unsigned int npnt = voff[v0 + nv] - pbase;
      for (unsigned int p = tid; p < npnt; p += VS_T) stage[p] = in[order[pbase + p]];
      __syncthreads();
      if ((unsigned int)tid < nv) {
        const unsigned int pa = voff[v] - pbase, pe = voff[v + 1] - pbase;
        float cx = 0.f, cy = 0.f, cz = 0.f, cw = 0.f;
        for (unsigned int p = pa; p < pe; p++) { const float4 q = stage[p]; cx += q.x; cy += q.y; cz += q.z; cw += q.w; }
        const float c = (float)(pe - pa);
        if (v < (unsigned int)k.cap_out) out[v] = make_float4(cx / c, cy / c, cz / c, cw / c);
        else if (a.overflow) atomicExch(a.overflow, 1);
      }
      __syncthreads();
      v0 += nv;
    }
  }
  if (tid == 0 && rank == 0) k.n_out[s] = nvox <= (unsigned int)k.cap_out ? (int)nvox : k.cap_out;
}

// number of key bits that index `cells` distinct voxel indices
int vox_index_bits(long long cells) { int b = 1; while (b < 32 && (1LL << b) < cells) b++; return b; }

static void vox_configure() {
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(vox_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VS_DYN_SMEM);
    cudaFuncSetAttribute(vox_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VS_DYN_SMEM);
    done = true;
  }
}
// a few clouds only (the latency case): a cluster of CTAs per cloud while the clusters fit in two waves (measured on one HDL-64E
// sweep: 65 us against 185 us for the single CTA); COOPERMAP_VOX_CLUSTER=0 / 1 forces the choice
static bool vox_use_cluster(int nclouds) {
  if (const char* e = getenv("COOPERMAP_VOX_CLUSTER")) return atoi(e) != 0;   // read per call: the tests run both kernels in one process
  return nclouds * VC <= 2 * 148;
}

static void fill_class(VoxClass& c, const float4* d_in, const int* d_n_in, int cap_in, float leaf, float4* d_out, int* d_n_out, int cap_out,
                       unsigned int* scratch) {
  c.in = d_in; c.n_in = d_n_in; c.cap_in = cap_in; c.inv = 1.0f / leaf; c.out = d_out; c.n_out = d_n_out; c.cap_out = cap_out; c.scratch = scratch;
}

void VoxelFilter::run(int nseg, const float4* d_in, const int* d_n_in, int cap_in, int max_n, float leaf, float4* d_out, int* d_n_out,
                      int cap_out, int* d_overflow, cudaStream_t stream) {
  if (nseg <= 0 || cap_in <= 0) return;
  if (max_n <= 0 || max_n > cap_in) max_n = cap_in;   // host-known upper bound of n_in[s]: sizes the scratch
  const size_t stride = ((size_t)max_n + 1 + 255) & ~(size_t)255;
  scratch.reserve((size_t)nseg * VS_ARRAYS * stride * sizeof(unsigned int));
  VoxSegArgs a;
  fill_class(a.cls[0], d_in, d_n_in, cap_in, leaf, d_out, d_n_out, cap_out, (unsigned int*)scratch.p);
  a.cls[1] = a.cls[0];
  a.nseg = nseg; a.ncls = 1; a.stride = (unsigned int)stride; a.overflow = d_overflow; a.box_out = nullptr;
  vox_configure();
  if (vox_use_cluster(nseg)) CM_LAUNCH(vox_cluster_kernel, dim3(nseg * VC, 1), VS_T, VS_DYN_SMEM, stream, a);
  else CM_LAUNCH(vox_segment_kernel, dim3(nseg, 1), VS_T, VS_DYN_SMEM, stream, a);
}

void VoxelFilter::run2(int nseg, const float4* d_in0, const int* d_n_in0, int cap_in0, float leaf0, float4* d_out0, int* d_n_out0, int cap_out0,
                       const float4* d_in1, const int* d_n_in1, int cap_in1, float leaf1, float4* d_out1, int* d_n_out1, int cap_out1,
                       int max_n, int* d_overflow, cudaStream_t stream) {
  if (nseg <= 0 || cap_in0 <= 0 || cap_in1 <= 0) return;
  const int capmax = cap_in0 > cap_in1 ? cap_in0 : cap_in1;
  if (max_n <= 0 || max_n > capmax) max_n = capmax;
  const size_t stride = ((size_t)max_n + 1 + 255) & ~(size_t)255;
  scratch.reserve((size_t)2 * nseg * VS_ARRAYS * stride * sizeof(unsigned int));
  VoxSegArgs a;
  fill_class(a.cls[0], d_in0, d_n_in0, cap_in0, leaf0, d_out0, d_n_out0, cap_out0, (unsigned int*)scratch.p);
  fill_class(a.cls[1], d_in1, d_n_in1, cap_in1, leaf1, d_out1, d_n_out1, cap_out1, (unsigned int*)scratch.p + (size_t)nseg * VS_ARRAYS * stride);
  a.nseg = nseg; a.ncls = 2; a.stride = (unsigned int)stride; a.overflow = d_overflow; a.box_out = nullptr;
  vox_configure();
  if (vox_use_cluster(2 * nseg)) CM_LAUNCH(vox_cluster_kernel, dim3(nseg * VC, 2), VS_T, VS_DYN_SMEM, stream, a);
  else CM_LAUNCH(vox_segment_kernel, dim3(nseg, 2), VS_T, VS_DYN_SMEM, stream, a);
}

}  // namespace cm
