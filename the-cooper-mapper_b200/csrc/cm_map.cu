// cm_map.cu -- K7: device-resident local feature map (a GPU voxel hash) with on-device insertion + voxel merge.
//
// Replaces FeatureMap<PointT>::addFeatureCloud / pushCornerPoint / pushSurfPoint / downsizeValidCloud
// (L_SLAM/src/util/FeatureMap.h:189-230, 289-306) and the map side of getSurroundFeature (:256-265).
// Reference semantics restated: a point pushed into the map goes to the 50 m cube round(p / 50) + origin
// (:475-487); after every insert each valid cube is re-filtered with pcl::VoxelGrid, i.e. within a cube all points
// that share the voxel floor(p * inv_leaf) are replaced by their mean (old centroid + new points, equal weight,
// old first).  Voxels that receive no new point keep their single point unchanged, so only voxels hit by the new
// cloud need work: O(inserted points) instead of the reference's O(map) re-sort.
//
// Layout: one open-addressing table of cells (cell = kdiv^3 voxels, CellEntry{key, start, count}) over a point pool
// with a bump allocator; a cell owns a contiguous block of `cellcap` slots and is moved to a larger block when it
// overflows (the old block is abandoned; compaction is left to a rebuild).  The same table is what the matcher's
// 5-NN search reads (cm_device.cuh), so an insert is immediately searchable.
//
// Cubes outside the valid window (needs returns beyond ~106 m: range + half a cube diagonal > 150 m): the reference pushes points
// into them but downsizeValidCloud (:289-306) only filters VALID cubes, so such a cube holds raw points -- several per voxel --
// until it is valid during a later insert, when its whole cloud is filtered at once: per voxel the centroid of [the point left by
// the last filter, then the raw points in push order].  Here: a group that lands in a cube that is invalid (or still holds raw
// points) is appended RAW, members in push order in consecutive slots, and the cube is marked dirty; a cell block keeps its
// append order, so at the end of an insert the dirty cubes that are valid now are re-filtered cell by cell IN BLOCK ORDER, which
// is the reference's summation order.  Raw points are ordinary pool points: the search sees them like the reference's surround
// cloud does.
// FeatureMap::shift (the grid re-centring when the sensor comes within 3 cubes of the grid border, :232-245, 354-376) swaps cube
// POINTERS in place while it iterates upwards.  Worked out (and checked against the literal loop for every |d| <= 2): the contents
// move by T = -sigma * d, sigma = sign of the first non-zero component of d in (i, j, k) order, contents that leave the grid are
// cleared.  The origin moves by +d, so for sigma < 0 this is the proper shift and for sigma > 0 every cube ends up 2 d away from
// where the coordinates of its points say.  The map here is keyed by coordinates; a byte per pool slot (`epoch`) and a displacement
// per epoch (`eoff`) record which cube STORES a point, which is what the surround selection, the per-cube voxel merge, the cube
// counts and the exported clouds go by.  Without a wrong-way shift every displacement is zero and none of this is looked at.
#include "cm_host.h"
#include "cm_math.h"
#include <algorithm>

namespace cm {

#define CM_MAP_PAD 0xFFFFFFFFFFFFFFFFull
#define CM_VOX_BIAS 65536   // voxel coordinates must stay within +-65536 (17 bits per axis)

__device__ __forceinline__ int world_to_cube_axis(float x, float cube_size, int origin) { return (int)(roundf(x / cube_size) + (float)origin); }

// ---- 1. transform + key + grouping -----------------------------------------------------------------------------------
// key = stream (8 bits) | cube parity (3 bits) | voxel z, y, x (17 bits each, biased).  Points with equal keys form one
// merge group.  Groups are found WITHOUT sorting: every point inserts its key into a scratch hash table whose entry keeps
// the smallest point index of the group (its head) and a linked list of the members; the head's thread later walks the
// list in ascending index order, which reproduces the input-order sum of a stable sort.  (A frame inserts ~1 point per
// map voxel, so the lists are almost always of length 1; the 62-bit radix sort this replaces cost 8 passes per class.)
struct GroupEntry { unsigned long long key; int head; int tail; };

// The launch sizes of an insert come from an ESTIMATE of the per-stream counts (they live on the device).  If a stream has
// more points than the estimate, the whole insert is skipped (every kernel below tests the flag first) and the host, which sees
// the flag with the step's results, repeats it with exact sizes: nothing is ever inserted partially.
__global__ void map_insert_guard_kernel(const int* __restrict__ n_pts, int nstreams, int max_n, int* __restrict__ skip,
                                        const int* __restrict__ step_skip) {
  int over = (step_skip && *step_skip) ? 1 : 0;   // the whole step is being repeated: insert nothing
  for (int s = threadIdx.x; s < nstreams; s += blockDim.x) over |= n_pts[s] > max_n ? 1 : 0;
  over = __syncthreads_or(over);
  if (threadIdx.x == 0) *skip = over;
}

// flags the step when a stream's filtered clouds exceed what the step's launches were sized for (cm_mapping.cu)
__global__ void step_guard_kernel(const int* __restrict__ n2, int nstreams, int max_c, int max_s, int* __restrict__ flag) {
  int over = 0;
  for (int s = threadIdx.x; s < nstreams; s += blockDim.x) over |= (n2[s] > max_c || n2[nstreams + s] > max_s) ? 1 : 0;
  over = __syncthreads_or(over);
  if (threadIdx.x == 0) *flag = over;
}
void launch_step_guard(const int* d_n2, int nstreams, int max_c, int max_s, int* d_flag, cudaStream_t stream) {
  CM_LAUNCH(step_guard_kernel, 1, 256, 0, stream, d_n2, nstreams, max_c, max_s, d_flag);
}

__global__ void map_group_clear_kernel(GroupEntry* tab, unsigned int cap, const int* __restrict__ skip) {
  if (*skip) return;
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { tab[i].key = CM_MAP_PAD; tab[i].head = 0x7fffffff; tab[i].tail = -1; }
}

// Sharded map: does this rank hold the point?  Yes if it owns the point's cube, or a neighbouring cube whose box is within the
// halo of the point: a query evaluated by the owner of its cube finds every map point within sqrt(5) m (the 5-NN gate) locally.
__device__ __forceinline__ bool shard_keeps(const MapClassDev& m, float x, float y, float z, int ci, int cj, int ck, int rank, int nranks) {
  if (cube_owner(ci, cj, ck, nranks) == rank) return true;
  const float HALO = 2.2360680f * 1.001f + 0.01f;
  const float half = 0.5f * m.cube_size;
  const float c[3] = {x, y, z};
  const int idx[3] = {ci, cj, ck};
  int lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float centre = m.cube_size * (float)(idx[a] - m.origin[a]);
    lo[a] = (c[a] - (centre - half)) < HALO ? -1 : 0;
    hi[a] = ((centre + half) - c[a]) < HALO ? 1 : 0;
  }
  for (int di = lo[0]; di <= hi[0]; di++)
    for (int dj = lo[1]; dj <= hi[1]; dj++)
      for (int dk = lo[2]; dk <= hi[2]; dk++) {
        const int ni = ci + di, nj = cj + dj, nk = ck + dk;
        if (ni < 0 || ni >= m.dims[0] || nj < 0 || nj >= m.dims[1] || nk < 0 || nk >= m.dims[2]) continue;
        if (cube_owner(ni, nj, nk, nranks) == rank) return true;
      }
  return false;
}

__global__ void map_key_kernel(const float4* __restrict__ pts, const int* __restrict__ n_pts, int cap, int max_n, int nstreams,
                               const MatchState* __restrict__ state, const float* __restrict__ tf_override, MapClassDev* maps,
                               float4* __restrict__ world, unsigned long long* __restrict__ keys, unsigned int* __restrict__ slot_of,
                               int* __restrict__ next, GroupEntry* __restrict__ gtab, unsigned int gmask, int* __restrict__ flags,
                               const int* __restrict__ skip, int shard_rank, int shard_nranks) {
  if (*skip) return;
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (size_t)nstreams * max_n) return;
  int s = (int)(g / max_n), i = (int)(g - (size_t)s * max_n);
  unsigned long long key = CM_MAP_PAD;
  if (i < n_pts[s]) {
    const MapClassDev& m = maps[s];
    float4 p = pts[(size_t)s * cap + i];
    float R[9], t[3];
    if (tf_override) { for (int k = 0; k < 9; k++) R[k] = tf_override[s * 12 + k]; for (int k = 0; k < 3; k++) t[k] = tf_override[s * 12 + 9 + k]; }
    else { for (int k = 0; k < 9; k++) R[k] = state[s].R[k]; for (int k = 0; k < 3; k++) t[k] = state[s].pose[3 + k]; }
    float4 w;
    transform_point(R, t, p.x, p.y, p.z, &w.x, &w.y, &w.z);   // transformPointCloud, transform_utils.h:601-614
    w.w = p.w;
    world[g] = w;
    // non-finite points: PCL's filter drops them from the cube (is_dense == false branch); dropped here at insert
    if (isfinite(w.x) && isfinite(w.y) && isfinite(w.z)) {
      int ci = world_to_cube_axis(w.x, m.cube_size, m.origin[0]);
      int cj = world_to_cube_axis(w.y, m.cube_size, m.origin[1]);
      int ck = world_to_cube_axis(w.z, m.cube_size, m.origin[2]);
      if (ci >= 0 && ci < m.dims[0] && cj >= 0 && cj < m.dims[1] && ck >= 0 && ck < m.dims[2] &&   // isIndexValid, :102-108
          (shard_nranks <= 1 || shard_keeps(m, w.x, w.y, w.z, ci, cj, ck, shard_rank, shard_nranks))) {
        float vx = floorf(w.x * m.inv_leaf), vy = floorf(w.y * m.inv_leaf), vz = floorf(w.z * m.inv_leaf);
        if (fabsf(vx) < (float)CM_VOX_BIAS && fabsf(vy) < (float)CM_VOX_BIAS && fabsf(vz) < (float)CM_VOX_BIAS) {
          unsigned long long kx = (unsigned long long)((int)vx + CM_VOX_BIAS), ky = (unsigned long long)((int)vy + CM_VOX_BIAS),
                             kz = (unsigned long long)((int)vz + CM_VOX_BIAS);
          unsigned long long par = (unsigned long long)((ci & 1) | ((cj & 1) << 1) | ((ck & 1) << 2));
          key = ((unsigned long long)s << 54) | (par << 51) | (kz << 34) | (ky << 17) | kx;
        } else {
          atomicExch(flags, 1);   // outside the supported voxel range
        }
      }
    }
  }
  keys[g] = key;
  if (key != CM_MAP_PAD) {
    unsigned int h = hash_cell(key) & gmask;
    while (true) {
      const unsigned long long prev = atomicCAS(&gtab[h].key, CM_MAP_PAD, key);
      if (prev == CM_MAP_PAD || prev == key) break;
      h = (h + 1) & gmask;
    }
    atomicMin(&gtab[h].head, (int)g);
    next[g] = atomicExch(&gtab[h].tail, (int)g);
    slot_of[g] = h;
  }
}

// ---- 2. one thread per (stream, cube, voxel) group: merge with the resident point or queue an append ---------------
struct PendingAdd { float4 p; unsigned int entry; int stream; int cube; int raw_tail; int raw_n; int pad[3]; };   // raw_n > 0: append the group's members unmerged

__device__ __forceinline__ unsigned int map_find_or_create_cell(MapClassDev& m, int cx, int cy, int cz) {
  unsigned long long key = pack_cell(cx, cy, cz);
  unsigned int h = hash_cell(key) & m.mask;
  for (unsigned int probes = 0; probes <= m.mask; probes++) {   // bounded: a full table is a capacity error, not a hang
    unsigned long long prev = atomicCAS(&m.entries[h].key, CM_EMPTY_KEY, key);
    if (prev == CM_EMPTY_KEY || prev == key) return h;
    h = (h + 1) & m.mask;
  }
  return 0xFFFFFFFFu;
}

__global__ void map_merge_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ slot_of,
                                 const int* __restrict__ next, const GroupEntry* __restrict__ gtab, size_t n,
                                 const float4* __restrict__ world, MapClassDev* maps, PendingAdd* __restrict__ pending,
                                 unsigned int* __restrict__ n_pending, unsigned int pending_cap, int* __restrict__ flags,
                                 const int* __restrict__ skip, const CubeWindow* __restrict__ windows) {
  if (*skip) return;
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const unsigned long long key = keys[g];
  if (key == CM_MAP_PAD) return;
  const GroupEntry ge = gtab[slot_of[g]];
  if (ge.head != (int)g) return;   // not the head (smallest index) of its group
  const int s = (int)(key >> 54);
  MapClassDev& m = maps[s];
  const float4 first = world[g];
  const int vx = (int)floorf(first.x * m.inv_leaf), vy = (int)floorf(first.y * m.inv_leaf), vz = (int)floorf(first.z * m.inv_leaf);
  const int ci = world_to_cube_axis(first.x, m.cube_size, m.origin[0]), cj = world_to_cube_axis(first.y, m.cube_size, m.origin[1]),
            ck = world_to_cube_axis(first.z, m.cube_size, m.origin[2]);
  const unsigned int e = map_find_or_create_cell(m, floor_div(vx, m.kdiv), floor_div(vy, m.kdiv), floor_div(vz, m.kdiv));
  if (e == 0xFFFFFFFFu) { atomicExch(flags + 2, 1); return; }   // cell table full: reported as CM_ERR_CAPACITY
  const int cube_lin = ci + cj * m.dims[0] + ck * m.dims[0] * m.dims[1];
  if (windows) {   // is the cube in _cubeValidInd, and free of raw points?  (windows == NULL: filter everywhere)
    const CubeWindow& w = windows[s];
    const int wi = ci - w.w0[0], wj = cj - w.w0[1], wk = ck - w.w0[2];
    const bool valid = wi >= 0 && wi < 7 && wj >= 0 && wj < 7 && wk >= 0 && wk < 7 && w.active[(wi * 7 + wj) * 7 + wk];
    if (!valid || m.dirty[cube_lin]) {
      int nmem = 0;
      for (int j = ge.tail; j >= 0; j = next[j]) nmem++;
      m.dirty[cube_lin] = 1;
      unsigned int slot = atomicAdd(n_pending, 1u);
      if (slot < pending_cap) {
        PendingAdd pa; pa.p = first; pa.entry = e; pa.stream = s; pa.cube = cube_lin; pa.raw_tail = ge.tail; pa.raw_n = nmem;
        pending[slot] = pa;
        atomicAdd(&m.pending[e], (unsigned int)nmem);
      } else atomicExch(flags + 2, 1);
      return;
    }
  }
  // resident point(s) of this voxel in this cube come first in the sum (they precede the pushed points in the cube cloud)
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  int cnt = 0, keep = -1;
  const unsigned int start = m.entries[e].start, count = m.entries[e].count;
  for (unsigned int j = 0; j < count; j++) {
    float4 q = m.pts[start + j];
    int o0 = 0, o1 = 0, o2 = 0;   // the resident point is merged only if the same CUBE stores it (a displaced epoch is another cube)
    if (m.cur_epoch) { const int* o = m.eoff + 3 * (int)m.epoch[start + j]; o0 = o[0]; o1 = o[1]; o2 = o[2]; }
    if ((int)floorf(q.x * m.inv_leaf) == vx && (int)floorf(q.y * m.inv_leaf) == vy && (int)floorf(q.z * m.inv_leaf) == vz &&
        world_to_cube_axis(q.x, m.cube_size, m.origin[0]) + o0 == ci && world_to_cube_axis(q.y, m.cube_size, m.origin[1]) + o1 == cj &&
        world_to_cube_axis(q.z, m.cube_size, m.origin[2]) + o2 == ck) {
      sx += q.x; sy += q.y; sz += q.z; si += q.w; cnt++;
      if (keep < 0) keep = (int)j;
      else atomicExch(flags + 1, 1);   // two resident points in one voxel (rounding drift): not merged here, only reported
    }
  }
  // members in ascending index order (= the order of the pushed cloud): repeated minimum over the (short) list
  for (int last = -1;;) {
    int best = 0x7fffffff;
    for (int j = ge.tail; j >= 0; j = next[j]) if (j > last && j < best) best = j;
    if (best == 0x7fffffff) break;
    const float4 q = world[best];
    sx += q.x; sy += q.y; sz += q.z; si += q.w; cnt++;
    last = best;
  }
  const float c = (float)cnt;
  const float4 cen = make_float4(sx / c, sy / c, sz / c, si / c);
  if (keep >= 0) {
    m.pts[start + keep] = cen;
  } else {
    unsigned int slot = atomicAdd(n_pending, 1u);
    if (slot < pending_cap) {
      PendingAdd pa; pa.p = cen; pa.entry = e; pa.stream = s; pa.cube = cube_lin; pa.raw_tail = -1; pa.raw_n = 0;
      pending[slot] = pa;
      atomicAdd(&m.pending[e], 1u);
    } else atomicExch(flags + 2, 1);
  }
}

// ---- 3. grow the cells that would overflow (first pending item of a cell does it) -----------------------------------
__global__ void map_grow_kernel(const PendingAdd* __restrict__ pending, const unsigned int* __restrict__ n_pending,
                                unsigned int pending_cap, MapClassDev* maps, int* __restrict__ flags) {
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int np = *n_pending; if (np > pending_cap) np = pending_cap;
  if (i >= np) return;
  const PendingAdd pa = pending[i];
  MapClassDev& m = maps[pa.stream];
  const unsigned int e = pa.entry;
  const unsigned int want = atomicExch(&m.pending[e], 0u);   // only one thread per cell sees a non-zero value
  if (want == 0) return;
  const unsigned int count = m.entries[e].count, cap = m.cellcap[e];
  if (count + want <= cap) return;
  unsigned int ncap = cap ? cap : 8u;
  while (ncap < count + want) ncap <<= 1;
  const unsigned int nstart = atomicAdd(m.cursor, ncap);
  if (nstart + ncap > m.pool_cap) { atomicExch(flags + 3, 1); return; }   // pool exhausted: the appends are dropped
  const unsigned int ostart = m.entries[e].start;
  for (unsigned int j = 0; j < count; j++) { m.pts[nstart + j] = m.pts[ostart + j]; m.epoch[nstart + j] = m.epoch[ostart + j]; }
  m.entries[e].start = nstart;
  m.cellcap[e] = ncap;
}

// ---- 4. append ---------------------------------------------------------------------------------------------------------
__global__ void map_append_kernel(const PendingAdd* __restrict__ pending, const unsigned int* __restrict__ n_pending,
                                  unsigned int pending_cap, MapClassDev* maps, const int* __restrict__ next, const float4* __restrict__ world) {
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int np = *n_pending; if (np > pending_cap) np = pending_cap;
  if (i >= np) return;
  const PendingAdd pa = pending[i];
  MapClassDev& m = maps[pa.stream];
  const unsigned int e = pa.entry;
  if (pa.raw_n > 0) {   // a group for an invalid / unfiltered cube: its members, unmerged, in push order, in consecutive slots
    const unsigned int j0 = atomicAdd(&m.entries[e].count, (unsigned int)pa.raw_n);
    if (j0 + (unsigned int)pa.raw_n > m.cellcap[e]) { atomicSub(&m.entries[e].count, (unsigned int)pa.raw_n); return; }   // growth failed (flagged)
    unsigned int w = m.entries[e].start + j0;
    for (int last = -1;;) {
      int best = 0x7fffffff;
      for (int j = pa.raw_tail; j >= 0; j = next[j]) if (j > last && j < best) best = j;
      if (best == 0x7fffffff) break;
      m.pts[w] = world[best]; m.epoch[w] = (unsigned char)m.cur_epoch; w++;
      last = best;
    }
    atomicAdd(&m.cube_count[pa.cube], pa.raw_n);
    atomicAdd(m.total, pa.raw_n);
    return;
  }
  if (m.entries[e].count >= m.cellcap[e]) return;   // growth failed (pool exhausted, already flagged)
  const unsigned int j = atomicAdd(&m.entries[e].count, 1u);
  if (j < m.cellcap[e]) {
    m.pts[m.entries[e].start + j] = pa.p;
    m.epoch[m.entries[e].start + j] = (unsigned char)m.cur_epoch;
    atomicAdd(&m.cube_count[pa.cube], 1);
    atomicAdd(m.total, 1);
  } else {
    atomicSub(&m.entries[e].count, 1u);
  }
}

// ---- 5. downsizeValidCloud for the cubes that hold raw points and are valid now --------------------------------------------------
__device__ __forceinline__ bool cube_valid_now(const CubeWindow& w, int ci, int cj, int ck) {
  const int wi = ci - w.w0[0], wj = cj - w.w0[1], wk = ck - w.w0[2];
  return wi >= 0 && wi < 7 && wj >= 0 && wj < 7 && wk >= 0 && wk < 7 && w.active[(wi * 7 + wj) * 7 + wk];
}
// need[s] = some cube of stream s is dirty and valid (one warp per stream)
__global__ void map_need_kernel(const MapClassDev* maps, const CubeWindow* __restrict__ windows, int nstreams, int* __restrict__ need,
                                const int* __restrict__ skip) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nstreams) return;
  int any = 0;
  if (!*skip) {
    const MapClassDev& m = maps[s];
    const CubeWindow& w = windows[s];
    // 11 cubes of the 7x7x7 window per lane: all `active` bytes first, then all `dirty` bytes (two load latencies, not twenty-two)
    unsigned char act[11];
#pragma unroll
    for (int q = 0; q < 11; q++) { const int a = lane + 32 * q; act[q] = a < 343 ? w.active[a] : (unsigned char)0; }
    unsigned char dr[11];
#pragma unroll
    for (int q = 0; q < 11; q++) {
      const int a = lane + 32 * q;
      const int i = a / 49 + w.w0[0], j = (a / 7) % 7 + w.w0[1], k = a % 7 + w.w0[2];
      const bool ok = act[q] && i >= 0 && i < m.dims[0] && j >= 0 && j < m.dims[1] && k >= 0 && k < m.dims[2];
      dr[q] = ok ? m.dirty[i + j * m.dims[0] + k * m.dims[0] * m.dims[1]] : (unsigned char)0;
    }
#pragma unroll
    for (int q = 0; q < 11; q++) any |= dr[q];
  }
  any = __any_sync(0xffffffffu, any);
  if (lane == 0) need[s] = any ? 1 : 0;
}
// one thread per cell of a stream that needs it: the points of dirty valid cubes, voxel by voxel in block order (= push order),
// become one centroid at the voxel's first slot; the block is compacted in place
__global__ void map_refilter_kernel(MapClassDev* maps, const CubeWindow* __restrict__ windows, int nstreams, const int* __restrict__ need) {
  for (int s0 = 0; s0 < nstreams; s0 += 32) {   // 32 streams' flags per load: the usual case (nothing to do) costs nstreams / 32 loads
   const int lane_ = threadIdx.x & 31;
   unsigned int todo = __ballot_sync(0xffffffffu, s0 + lane_ < nstreams && need[s0 + lane_] != 0);
   while (todo) {
    const int s = s0 + __ffs(todo) - 1;
    todo &= todo - 1;
    MapClassDev& m = maps[s];
    const CubeWindow& w = windows[s];
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e <= m.mask; e += gridDim.x * blockDim.x) {
      if (m.entries[e].key == CM_EMPTY_KEY) continue;
      const unsigned int start = m.entries[e].start, count = m.entries[e].count;
      unsigned int kept = 0;
      for (unsigned int j = 0; j < count; j++) {
        float4 q = m.pts[start + j];
        if (isnan(q.x)) continue;   // consumed by an earlier voxel of this pass (a stored point is never NaN: the insert drops those)
        const unsigned char ep = m.epoch[start + j];
        const int* o = m.eoff + 3 * (int)ep;
        const int ci = world_to_cube_axis(q.x, m.cube_size, m.origin[0]) + o[0], cj = world_to_cube_axis(q.y, m.cube_size, m.origin[1]) + o[1],
                  ck = world_to_cube_axis(q.z, m.cube_size, m.origin[2]) + o[2];
        const int lin = ci + cj * m.dims[0] + ck * m.dims[0] * m.dims[1];
        const bool inside = ci >= 0 && ci < m.dims[0] && cj >= 0 && cj < m.dims[1] && ck >= 0 && ck < m.dims[2];
        if (inside && m.dirty[lin] && cube_valid_now(w, ci, cj, ck)) {
          const int vx = (int)floorf(q.x * m.inv_leaf), vy = (int)floorf(q.y * m.inv_leaf), vz = (int)floorf(q.z * m.inv_leaf);
          float sx = q.x, sy = q.y, sz = q.z, si = q.w;
          int cnt = 1;
          for (unsigned int t = j + 1; t < count; t++) {
            const float4 r = m.pts[start + t];
            if (isnan(r.x)) continue;
            if ((int)floorf(r.x * m.inv_leaf) != vx || (int)floorf(r.y * m.inv_leaf) != vy || (int)floorf(r.z * m.inv_leaf) != vz) continue;
            const int* o2 = m.eoff + 3 * (int)m.epoch[start + t];
            if (world_to_cube_axis(r.x, m.cube_size, m.origin[0]) + o2[0] != ci || world_to_cube_axis(r.y, m.cube_size, m.origin[1]) + o2[1] != cj ||
                world_to_cube_axis(r.z, m.cube_size, m.origin[2]) + o2[2] != ck)
              continue;
            sx += r.x; sy += r.y; sz += r.z; si += r.w; cnt++;
            const float qn = __int_as_float(0x7fc00000);
            m.pts[start + t] = make_float4(qn, qn, qn, qn);   // consumed
          }
          if (cnt > 1) {
            const float c = (float)cnt;
            q = make_float4(sx / c, sy / c, sz / c, si / c);
            atomicSub(&m.cube_count[lin], cnt - 1);
            atomicSub(m.total, cnt - 1);
          }
        }
        m.pts[start + kept] = q; m.epoch[start + kept] = ep;
        kept++;
      }
      if (kept != count) m.entries[e].count = kept;
    }
   }
  }
}
// the valid cubes of the streams that were re-filtered are clean again
__global__ void map_dirty_clear_kernel(MapClassDev* maps, const CubeWindow* __restrict__ windows, int nstreams, const int* __restrict__ need) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nstreams || !need[s]) return;
  MapClassDev& m = maps[s];
  const CubeWindow& w = windows[s];
  for (int a = lane; a < 343; a += 32) {
    if (!w.active[a]) continue;
    const int i = a / 49 + w.w0[0], j = (a / 7) % 7 + w.w0[1], k = a % 7 + w.w0[2];
    if (i < 0 || i >= m.dims[0] || j < 0 || j >= m.dims[1] || k < 0 || k >= m.dims[2]) continue;
    m.dirty[i + j * m.dims[0] + k * m.dims[0] * m.dims[1]] = 0;
  }
}

__global__ void map_clear_kernel(CellEntry* e, unsigned int* cellcap, unsigned int* pend, unsigned int cap) {
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { e[i].key = CM_EMPTY_KEY; e[i].start = 0; e[i].count = 0; cellcap[i] = 0; pend[i] = 0; }
}

// searchable size of the surround map (sum of the valid cubes) -> GridView.npts, and the rest of the view.
// One warp per stream.
__global__ void map_view_kernel(MapClassDev* maps, MapClassDev* maps1, const CubeWindow* windows, GridView* views, GridView* views1, int nstreams,
                                float gate, int shard_rank, int shard_nranks) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nstreams) return;
  if (blockIdx.y) { maps = maps1; views = views1; }   // both classes in one launch
  MapClassDev& m = maps[s];
  const CubeWindow& w = windows[s];
  int total = 0;
  unsigned char act[11];
#pragma unroll
  for (int u = 0; u < 11; u++) { const int a = lane + 32 * u; act[u] = a < 343 ? w.active[a] : 0; }   // independent loads, then the counts
#pragma unroll
  for (int u = 0; u < 11; u++) {
    const int a = lane + 32 * u;
    if (!act[u]) continue;
    int i = a / 49 + w.w0[0], j = (a / 7) % 7 + w.w0[1], k = a % 7 + w.w0[2];
    if (shard_nranks > 1 && cube_owner(i, j, k, shard_nranks) != shard_rank) continue;   // halo copies are counted by their owner
    total += m.cube_count[i + j * m.dims[0] + k * m.dims[0] * m.dims[1]];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if (lane != 0) return;
  m.origin[0] = w.origin[0]; m.origin[1] = w.origin[1]; m.origin[2] = w.origin[2];
  GridView v;
  v.entries = m.entries; v.pts = m.pts; v.mask = m.mask; v.inv_leaf = m.inv_leaf; v.kdiv = m.kdiv; v.cell = (float)m.kdiv * m.leaf;
  v.npts = total;
  int L = (int)ceilf(sqrtf(gate) / (0.98f * v.cell) - 0.5f);
  v.max_level = L < 0 ? 0 : L;
  v.window = windows + s; v.cube_count = m.cube_count;
  v.epoch = m.epoch; v.eoff = m.eoff; v.displaced = m.cur_epoch > 0 ? 1 : 0;
  views[s] = v;
}

// export every resident point with its cube index (diagnostics, publishing, tests)
__global__ void map_export_kernel(const MapClassDev* maps, int s, float4* out, int* cube_out, unsigned int* n_out, unsigned int cap) {
  const MapClassDev& m = maps[s];
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > m.mask) return;
  if (m.entries[i].key == CM_EMPTY_KEY) return;
  const unsigned int start = m.entries[i].start, count = m.entries[i].count;
  for (unsigned int j = 0; j < count; j++) {
    float4 p = m.pts[start + j];
    unsigned int o = atomicAdd(n_out, 1u);
    if (o < cap) {
      out[o] = p;
      int ci = world_to_cube_axis(p.x, m.cube_size, m.origin[0]), cj = world_to_cube_axis(p.y, m.cube_size, m.origin[1]),
          ck = world_to_cube_axis(p.z, m.cube_size, m.origin[2]);
      if (m.cur_epoch) { const int* eo = m.eoff + 3 * (int)m.epoch[start + j]; ci += eo[0]; cj += eo[1]; ck += eo[2]; }
      cube_out[o] = ci + cj * m.dims[0] + ck * m.dims[0] * m.dims[1];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static unsigned int pow2_at_least(size_t v) { unsigned int p = 1024; while (p < v) p <<= 1; return p; }

void DeviceMap::create(int nstreams_, const MapConfig& c, cudaStream_t stream) {
  nstreams = nstreams_; cfg = c;
  const int ncubes = c.dims[0] * c.dims[1] * c.dims[2];
  h_eoff.assign((size_t)nstreams * 256 * 3, 0);
  cur_epoch.assign(nstreams, 0);
  eoff.reserve(h_eoff.size() * sizeof(int));
  cudaMemsetAsync(eoff.p, 0, h_eoff.size() * sizeof(int), stream);
  for (int cls = 0; cls < 2; cls++) {
    const size_t maxpts = cls == 0 ? c.max_corner : c.max_surf;
    const unsigned int tcap = pow2_at_least(maxpts);            // cells <= points; load factor <= 0.5 when cells hold >= 2 points
    const unsigned int pool = (unsigned int)(maxpts * 4 + 1024);   // blocks are powers of two >= 8 slots: slack for sparse cells
    table_cap[cls] = tcap; pool_cap[cls] = pool;
    entries[cls].reserve((size_t)nstreams * tcap * sizeof(CellEntry));
    cellcap[cls].reserve((size_t)nstreams * tcap * sizeof(unsigned int));
    pending_cnt[cls].reserve((size_t)nstreams * tcap * sizeof(unsigned int));
    pts[cls].reserve((size_t)nstreams * pool * sizeof(float4));
    cube_count[cls].reserve((size_t)nstreams * ncubes * sizeof(int));
    epoch[cls].reserve((size_t)nstreams * pool);
    cudaMemsetAsync(epoch[cls].p, 0, (size_t)nstreams * pool, stream);
    cube_dirty[cls].reserve((size_t)nstreams * ncubes);
    cudaMemsetAsync(cube_dirty[cls].p, 0, (size_t)nstreams * ncubes, stream);
    need[cls].reserve((size_t)nstreams * sizeof(int));
    cursor[cls].reserve((size_t)nstreams * 2 * sizeof(unsigned int));
    dev[cls].reserve((size_t)nstreams * sizeof(MapClassDev));
    views[cls].reserve((size_t)nstreams * sizeof(GridView));
    cudaMemsetAsync(cube_count[cls].p, 0, (size_t)nstreams * ncubes * sizeof(int), stream);
    cudaMemsetAsync(cursor[cls].p, 0, (size_t)nstreams * 2 * sizeof(unsigned int), stream);
    const size_t tot = (size_t)nstreams * tcap;
    CM_LAUNCH(map_clear_kernel, (unsigned int)((tot + 255) / 256), 256, 0, stream, (CellEntry*)entries[cls].p,
              (unsigned int*)cellcap[cls].p, (unsigned int*)pending_cnt[cls].p, (unsigned int)tot);
    std::vector<MapClassDev> h(nstreams);
    const float leaf = cls == 0 ? c.leaf_corner : c.leaf_surf;
    for (int s = 0; s < nstreams; s++) {
      MapClassDev& m = h[s];
      m.entries = (CellEntry*)entries[cls].p + (size_t)s * tcap;
      m.cellcap = (unsigned int*)cellcap[cls].p + (size_t)s * tcap;
      m.pending = (unsigned int*)pending_cnt[cls].p + (size_t)s * tcap;
      m.mask = tcap - 1;
      m.pts = (float4*)pts[cls].p + (size_t)s * pool;
      m.pool_cap = pool;
      m.cursor = (unsigned int*)cursor[cls].p + s * 2;
      m.total = (int*)cursor[cls].p + s * 2 + 1;
      m.cube_count = (int*)cube_count[cls].p + (size_t)s * ncubes;
      m.leaf = leaf; m.inv_leaf = 1.0f / leaf;
      m.kdiv = cls == 0 ? c.kdiv_corner : c.kdiv_surf;
      m.cube_size = c.cube_size;
      for (int k = 0; k < 3; k++) { m.dims[k] = c.dims[k]; m.origin[k] = c.origin[k]; }
      m.epoch = (unsigned char*)epoch[cls].p + (size_t)s * pool;
      m.eoff = (int*)eoff.p + (size_t)s * 256 * 3;
      m.cur_epoch = 0;
      m.dirty = (unsigned char*)cube_dirty[cls].p + (size_t)s * ncubes;
    }
    cudaMemcpyAsync(dev[cls].p, h.data(), sizeof(MapClassDev) * nstreams, cudaMemcpyHostToDevice, stream);
    hdev[cls] = h;
  }
  windows.reserve(sizeof(CubeWindow) * nstreams);
  flags.reserve(sizeof(int) * 8);
  for (int c = 0; c < 2; c++) n_pending[c].reserve(sizeof(unsigned int));
  cudaMemsetAsync(flags.p, 0, sizeof(int) * 8, stream);
  cudaStreamSynchronize(stream);   // h goes out of scope
}

__global__ void stage_copy_kernel(const unsigned int* __restrict__ src, unsigned int* __restrict__ dst, size_t nwords) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
void staged_upload(void* d_dst, const void* pinned, size_t bytes, cudaStream_t stream) {
  const size_t nwords = (bytes + 3) / 4;   // both buffers are allocated in multiples of 4 bytes
  if (!nwords) return;
  const unsigned int nb = (unsigned int)std::min<size_t>((nwords + 255) / 256, 64);
  CM_LAUNCH(stage_copy_kernel, nb, 256, 0, stream, (const unsigned int*)pinned, (unsigned int*)d_dst, nwords);
}

void DeviceMap::set_windows(const CubeWindow* h_windows, float gate, cudaStream_t stream, bool staged) {
  windows_valid = true;
  if (!h_windows) {}
  else if (staged) staged_upload(windows.p, h_windows, sizeof(CubeWindow) * nstreams, stream);
  else cudaMemcpyAsync(windows.p, h_windows, sizeof(CubeWindow) * nstreams, cudaMemcpyHostToDevice, stream);
  CM_LAUNCH(map_view_kernel, dim3((nstreams * 32 + 127) / 128, 2), 128, 0, stream, (MapClassDev*)dev[0].p, (MapClassDev*)dev[1].p,
            (const CubeWindow*)windows.p, (GridView*)views[0].p, (GridView*)views[1].p, nstreams, gate, shard_rank, shard_nranks);
}

__global__ void map_npts_pack_kernel(const GridView* vc, const GridView* vs, int nstreams, double* vec) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nstreams) { vec[s] = (double)vc[s].npts; vec[nstreams + s] = (double)vs[s].npts; }
}
__global__ void map_npts_unpack_kernel(GridView* vc, GridView* vs, int nstreams, const double* vec) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nstreams) { vc[s].npts = (int)vec[s]; vs[s].npts = (int)vec[nstreams + s]; }
}
void DeviceMap::pack_npts(double* d_vec, cudaStream_t stream) {
  CM_LAUNCH(map_npts_pack_kernel, (nstreams + 63) / 64, 64, 0, stream, (const GridView*)views[0].p, (const GridView*)views[1].p, nstreams, d_vec);
}
void DeviceMap::unpack_npts(const double* d_vec, cudaStream_t stream) {
  CM_LAUNCH(map_npts_unpack_kernel, (nstreams + 63) / 64, 64, 0, stream, (GridView*)views[0].p, (GridView*)views[1].p, nstreams, d_vec);
}

void DeviceMap::insert(int cls, const float4* d_pts, const int* d_n, int cap, int max_n, const MatchState* d_state, const float* d_tf,
                       cudaStream_t stream, const int* step_skip, bool filter_all) {
  const CubeWindow* wins = (filter_all || !windows_valid) ? nullptr : (const CubeWindow*)windows.p;   // no update() yet: every cube counts as valid
  if (cap <= 0) return;
  if (max_n <= 0 || max_n > cap) max_n = cap;   // host-known upper bound of d_n[s]
  const size_t n = (size_t)nstreams * max_n;
  world[cls].reserve(n * sizeof(float4));
  unsigned int gcap = 1024;
  while ((size_t)gcap < 2 * n) gcap <<= 1;
  keys_a[cls].reserve(n * 8); vals_a[cls].reserve(n * 4); vals_b[cls].reserve(n * 4); keys_b[cls].reserve((size_t)gcap * sizeof(GroupEntry));
  pending[cls].reserve(n * sizeof(PendingAdd));
  const unsigned int nb = (unsigned int)((n + 255) / 256);
  const int* skip = (const int*)flags.p + 4 + cls;   // flags[4 + cls]: this class' insert was skipped (a count exceeded max_n)
  cudaMemsetAsync(n_pending[cls].p, 0, sizeof(unsigned int), stream);
  CM_LAUNCH(map_insert_guard_kernel, 1, 256, 0, stream, d_n, nstreams, max_n, (int*)flags.p + 4 + cls, step_skip);
  CM_LAUNCH(map_group_clear_kernel, (gcap + 255) / 256, 256, 0, stream, (GroupEntry*)keys_b[cls].p, gcap, skip);
  CM_LAUNCH(map_key_kernel, nb, 256, 0, stream, d_pts, d_n, cap, max_n, nstreams, d_state, d_tf, (MapClassDev*)dev[cls].p, (float4*)world[cls].p,
            (unsigned long long*)keys_a[cls].p, (unsigned int*)vals_a[cls].p, (int*)vals_b[cls].p, (GroupEntry*)keys_b[cls].p, gcap - 1, (int*)flags.p, skip, shard_rank, shard_nranks);
  CM_LAUNCH(map_merge_kernel, nb, 256, 0, stream, (const unsigned long long*)keys_a[cls].p, (const unsigned int*)vals_a[cls].p, (const int*)vals_b[cls].p,
            (const GroupEntry*)keys_b[cls].p, n, (const float4*)world[cls].p, (MapClassDev*)dev[cls].p, (PendingAdd*)pending[cls].p,
            (unsigned int*)n_pending[cls].p, (unsigned int)n, (int*)flags.p, skip, wins);
  CM_LAUNCH(map_grow_kernel, nb, 256, 0, stream, (const PendingAdd*)pending[cls].p, (const unsigned int*)n_pending[cls].p, (unsigned int)n,
            (MapClassDev*)dev[cls].p, (int*)flags.p);
  CM_LAUNCH(map_append_kernel, nb, 256, 0, stream, (const PendingAdd*)pending[cls].p, (const unsigned int*)n_pending[cls].p, (unsigned int)n,
            (MapClassDev*)dev[cls].p, (const int*)vals_b[cls].p, (const float4*)world[cls].p);
  if (wins) {   // downsizeValidCloud for cubes that hold raw points and are valid now (normally none: two tiny launches and an empty one)
    CM_LAUNCH(map_need_kernel, (nstreams * 32 + 127) / 128, 128, 0, stream, (const MapClassDev*)dev[cls].p, wins, nstreams, (int*)need[cls].p, skip);
    CM_LAUNCH(map_refilter_kernel, 296, 256, 0, stream, (MapClassDev*)dev[cls].p, wins, nstreams, (const int*)need[cls].p);
    CM_LAUNCH(map_dirty_clear_kernel, (nstreams * 32 + 127) / 128, 128, 0, stream, (MapClassDev*)dev[cls].p, wins, nstreams, (const int*)need[cls].p);
  }
}

// ---- FeatureMap::shift ----------------------------------------------------------------------------------------------------------
// One thread per cell: keeps the points whose storage cube (coordinates under the NEW origin + displacement of the epoch) is still
// inside the grid, in their order, and counts them into cube_count (zeroed by the caller).  m->origin / eoff already hold the
// state after the shift.
// drop (optional): [W*H*D] bytes, non-zero = the cube is evicted (DynamicFeatureMap paging, cm_mapio.cu)
__global__ void map_shift_kernel(MapClassDev* mp, const unsigned char* __restrict__ drop) {
  MapClassDev& m = *mp;
  const unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > m.mask || m.entries[e].key == CM_EMPTY_KEY) return;
  const unsigned int start = m.entries[e].start, count = m.entries[e].count;
  unsigned int kept = 0;
  for (unsigned int j = 0; j < count; j++) {
    const float4 q = m.pts[start + j];
    const unsigned char ep = m.epoch[start + j];
    const int* o = m.eoff + 3 * (int)ep;
    const int ci = world_to_cube_axis(q.x, m.cube_size, m.origin[0]) + o[0], cj = world_to_cube_axis(q.y, m.cube_size, m.origin[1]) + o[1],
              ck = world_to_cube_axis(q.z, m.cube_size, m.origin[2]) + o[2];
    if (ci < 0 || ci >= m.dims[0] || cj < 0 || cj >= m.dims[1] || ck < 0 || ck >= m.dims[2]) continue;
    if (drop && drop[ci + cj * m.dims[0] + ck * m.dims[0] * m.dims[1]]) continue;
    if (kept != j) { m.pts[start + kept] = q; m.epoch[start + kept] = ep; }
    kept++;
    atomicAdd(&m.cube_count[ci + cj * m.dims[0] + ck * m.dims[0] * m.dims[1]], 1);
  }
  if (kept != count) { m.entries[e].count = kept; atomicSub(m.total, (int)(count - kept)); }
}

bool DeviceMap::shift(int s, const int d[3], const int new_origin[3], cudaStream_t stream, bool literal, const unsigned char* d_drop) {
  // contents move by T = -sigma d (see the header); the displacement of a stored point changes by T - d
  // literal == false: a plain re-centring (every cube follows the origin), d_drop: cubes to evict on the way
  int sigma = 0;
  for (int k = 0; k < 3 && !sigma && literal; k++) sigma = d[k] > 0 ? 1 : (d[k] < 0 ? -1 : 0);
  int* row = h_eoff.data() + (size_t)s * 256 * 3;
  if (sigma > 0) {
    if (cur_epoch[s] >= 255) return false;
    for (int e = 0; e <= cur_epoch[s]; e++) for (int k = 0; k < 3; k++) row[3 * e + k] -= 2 * d[k];
    cur_epoch[s]++;                                   // points inserted from now on are stored where their coordinates say
    for (int k = 0; k < 3; k++) row[3 * cur_epoch[s] + k] = 0;
    cudaMemcpyAsync((int*)eoff.p + (size_t)s * 256 * 3, row, sizeof(int) * 3 * (cur_epoch[s] + 1), cudaMemcpyHostToDevice, stream);
  }
  const int ncubes = cfg.dims[0] * cfg.dims[1] * cfg.dims[2];
  for (int cls = 0; cls < 2; cls++) {
    MapClassDev& h = hdev[cls][s];
    for (int k = 0; k < 3; k++) h.origin[k] = new_origin[k];
    h.cur_epoch = cur_epoch[s];
    MapClassDev* dm = (MapClassDev*)dev[cls].p + s;
    cudaMemcpyAsync(dm, &h, sizeof(MapClassDev), cudaMemcpyHostToDevice, stream);
    cudaMemsetAsync(h.cube_count, 0, sizeof(int) * ncubes, stream);
    CM_LAUNCH(map_shift_kernel, (table_cap[cls] + 255) / 256, 256, 0, stream, dm, d_drop);
  }
  return true;
}

size_t DeviceMap::export_points(int cls, int s, float4* d_out, int* d_cube, unsigned int* d_n, unsigned int cap, cudaStream_t stream) {
  cudaMemsetAsync(d_n, 0, sizeof(unsigned int), stream);
  CM_LAUNCH(map_export_kernel, (table_cap[cls] + 255) / 256, 256, 0, stream, (const MapClassDev*)dev[cls].p, s, d_out, d_cube, d_n, cap);
  return 0;
}

}  // namespace cm
