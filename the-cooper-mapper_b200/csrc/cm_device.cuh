// cm_device.cuh -- device-side building blocks: the voxel-cell hash ("grid") that replaces the reference's
// KD-tree, and the exact 5-nearest-neighbour search over it.
//
// Replaces nanoflann::KdTreeFLANN::setInputCloud / nearestKSearch (nanoflann_pcl.h:132-162, core
// nanoflann.hpp:931-1030,1433-1497) as used by ScanMatch.cpp:68-76,100-101,119.  Exactness argument: the
// reference discards a query whose 5th neighbour has d2 >= 5.0 (ScanMatch.cpp:102,120), so the search radius is
// bounded; cells are visited in growing cubic shells around the query and the search stops as soon as the 5th
// best distance is provably smaller than the distance to any unvisited cell.  Distances are accumulated exactly
// like L2_Simple_Adaptor::evalMetric (nanoflann.hpp:364-372): ((0 + dx*dx) + dy*dy) + dz*dz in float, unfused.
// Ties are broken by (d2, index) -- the canonical order of the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace cm {

#define CM_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

// One hash slot: packed cell coordinates -> [start, start + count) in the point pool.  16 bytes, one LDG.128.
struct __align__(16) CellEntry {
  unsigned long long key;
  unsigned int start;
  unsigned int count;
};

// The reference's cube lattice around the sensor (FeatureMap.h:232-254, 308-352, 475-487): which 50 m cubes are
// "valid" (searched).  Uploaded by the host every frame; NULL for stateless clouds (everything is searched).
struct CubeWindow {
  int origin[3];              // _cubeOriginWidth/Height/Depth
  int dims[3];                // _cubeWidth/Height/Depth
  int w0[3];                  // lowest cube index of the 7x7x7 window (_curCube* - 3)
  float cube_size;            // _worldCubeSize
  unsigned char active[343];  // cube in _cubeValidInd
  unsigned char interior[343];// cube and its 26 neighbours all active -> no per-point test needed
};

// Read-only view of one grid (one map cloud of one stream).  A point p lives in voxel v = floor(p * inv_leaf)
// (PCL's voxel lattice) and in cell floor_div(v, kdiv); stateless clouds use kdiv = 1 and inv_leaf = 1 / cell.
struct GridView {
  const CellEntry* entries;   // open addressing, linear probing, capacity = mask + 1 (power of two)
  const float4* pts;          // x, y, z, w  (w = original index bits for stateless clouds, intensity for the map)
  unsigned int mask;
  float inv_leaf;             // 1 / leaf (float, computed once on the host like PCL's inverse_leaf_size_)
  int kdiv;                   // cell edge = kdiv * leaf
  float cell;                 // cell edge in metres
  int npts;                   // number of searchable points (the reference gates on it, ScanMatch.cpp:57-58)
  int max_level;              // last shell to visit so that (max_level + 0.48) * cell >= sqrt(gate)
  const CubeWindow* window;   // NULL: no cube filtering
  const int* cube_count;      // points per 50 m cube [W*H*D] (map grids; NULL for stateless clouds)
  // FeatureMap::shift (FeatureMap.h:354-376) moves the cube POINTERS; when it moves them the wrong way (see cm_map.cu) a point is
  // stored in another cube than its coordinates say.  epoch[slot] names the shift history of a pool slot, eoff[3 * epoch] the
  // displacement (in cubes) of its storage cube from its coordinate cube; displaced == 0: no such point exists (the common case)
  const unsigned char* epoch; const int* eoff; int displaced;
};

__host__ __device__ __forceinline__ int floor_div(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

__host__ __device__ __forceinline__ unsigned long long pack_cell(int x, int y, int z) {
  const unsigned long long B = 1ull << 20;
  return ((unsigned long long)(x + (long long)B) & 0x1FFFFF) | (((unsigned long long)(y + (long long)B) & 0x1FFFFF) << 21) |
         (((unsigned long long)(z + (long long)B) & 0x1FFFFF) << 42);
}
// 32-bit mix of the two key halves (the GPU has no 64-bit multiplier: a 64-bit finaliser costs ~3x the instructions and
// the probe sequence is 15 % of the search kernel); lattice coordinates differ in their low bits, which both rounds spread
__host__ __device__ __forceinline__ unsigned int hash_cell(unsigned long long k) {
  unsigned int h = ((unsigned int)k * 0x9E3779B1u) ^ (((unsigned int)(k >> 32) + 0x7F4A7C15u) * 0x85EBCA77u);
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}

__device__ __forceinline__ bool grid_probe(const GridView& g, int x, int y, int z, unsigned int* start, unsigned int* count) {
  unsigned long long key = pack_cell(x, y, z);
  unsigned int h = hash_cell(key) & g.mask;
  for (unsigned int probes = 0; probes <= g.mask; probes++) {   // bounded even if the table were completely full
    uint4 e = __ldg(reinterpret_cast<const uint4*>(g.entries + h));
    unsigned long long k = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
    if (k == key) { *start = e.z; *count = e.w; return true; }
    if (k == CM_EMPTY_KEY) return false;
    h = (h + 1) & g.mask;
  }
  return false;
}

// Sorted 5-best list by (d2, idx), held as packed 64-bit keys: (bits of d2) << 32 | idx.  d2 >= 0, so the bit pattern of
// the float orders like its value and ONE unsigned 64-bit comparison is the lexicographic (d2, idx) comparison.
// slot = position in the point pool (to fetch the coordinates afterwards); for map grids idx == slot and the slot
// array is dead code.
struct Top5 {
  unsigned long long key[5];
  int slot[5];
  __device__ __forceinline__ float d(int k) const { return __uint_as_float((unsigned int)(key[k] >> 32)); }
  __device__ __forceinline__ int idx(int k) const { return (int)(unsigned int)(key[k] & 0xFFFFFFFFull); }
};
#define CM_TOP5_EMPTY ((((unsigned long long)0x7f7fffffu) << 32) | 0x7fffffffull)   // (FLT_MAX, INT_MAX)
__device__ __forceinline__ unsigned long long top5_key(float d, int idx) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)idx; }
__device__ __forceinline__ void top5_init(Top5& t) {
#pragma unroll
  for (int k = 0; k < 5; k++) { t.key[k] = CM_TOP5_EMPTY; t.slot[k] = -1; }
}
__device__ __forceinline__ void top5_insert_key(Top5& t, unsigned long long key, int slot) {
  if (!(key < t.key[4])) return;
  const bool c0 = key < t.key[0], c1 = key < t.key[1], c2 = key < t.key[2], c3 = key < t.key[3];
  // new slot k = c_{k-1} ? old[k-1] : (c_k ? new : old[k])
  t.key[4] = c3 ? t.key[3] : key;                    t.slot[4] = c3 ? t.slot[3] : slot;
  t.key[3] = c2 ? t.key[2] : (c3 ? key : t.key[3]);  t.slot[3] = c2 ? t.slot[2] : (c3 ? slot : t.slot[3]);
  t.key[2] = c1 ? t.key[1] : (c2 ? key : t.key[2]);  t.slot[2] = c1 ? t.slot[1] : (c2 ? slot : t.slot[2]);
  t.key[1] = c0 ? t.key[0] : (c1 ? key : t.key[1]);  t.slot[1] = c0 ? t.slot[0] : (c1 ? slot : t.slot[1]);
  t.key[0] = c0 ? key : t.key[0];                    t.slot[0] = c0 ? slot : t.slot[0];
}
// A NaN or infinite distance never enters the list (nor the reference's result set: nanoflann compares dist < worst with worst =
// FLT_MAX): the bit patterns of NaN and +inf are larger than FLT_MAX's, the empty key, and d is a sum of squares, never negative.
__device__ __forceinline__ void top5_insert(Top5& t, float d, int idx, int slot) {
  top5_insert_key(t, top5_key(d, idx), slot);
}
// the same for a list that may already hold the candidate (warm start from the previous iteration's neighbours)
__device__ __forceinline__ void top5_insert_key_unique(Top5& t, unsigned long long key, int slot) {
  if (!(key < t.key[4])) return;
  if (key == t.key[0] || key == t.key[1] || key == t.key[2] || key == t.key[3]) return;
  top5_insert_key(t, key, slot);
}
// Insert that also tracks d6, the smallest distance among the candidates that were turned away at the door or pushed out of the
// list (only candidates at or below the current 5th distance ever get here, and only those can tie with the final 5th).
// At the end  d6 == d(4)  or two equal neighbours in the list  <=>  the list depends on how exact distance ties were broken.
__device__ __forceinline__ void top5_insert_track(Top5& t, unsigned long long key, int slot, bool unique, float& d6) {
  if (!(key < t.key[4])) {   // turned away (key == key[4]: a warm list meeting its own 5th entry again)
    if (key != t.key[4]) d6 = fminf(d6, __uint_as_float((unsigned int)(key >> 32)));
    return;
  }
  if (unique && (key == t.key[0] || key == t.key[1] || key == t.key[2] || key == t.key[3])) return;
  d6 = fminf(d6, t.d(4));
  top5_insert_key(t, key, slot);
}
__device__ __forceinline__ bool top5_has_tie(const Top5& t, float d6) {
  const float d4 = t.d(4);
  return (d6 == d4 && d4 < FLT_MAX) || t.d(0) == t.d(1) || t.d(1) == t.d(2) || t.d(2) == t.d(3) || (t.d(3) == d4 && d4 < FLT_MAX);
}

// worldToCube (FeatureMap.h:475-487) of a map point: the cube its coordinates name plus the displacement of its epoch (see GridView)
__device__ __forceinline__ void label_cube(const GridView& g, const CubeWindow& w, const float4& p, int slot, int* i, int* j, int* k) {
  *i = (int)(roundf(p.x / w.cube_size) + (float)w.origin[0]);
  *j = (int)(roundf(p.y / w.cube_size) + (float)w.origin[1]);
  *k = (int)(roundf(p.z / w.cube_size) + (float)w.origin[2]);
  if (g.displaced) { const int* o = g.eoff + 3 * (int)g.epoch[slot]; *i += o[0]; *j += o[1]; *k += o[2]; }
}
// -> index into the 7x7x7 window, -1 outside it
__device__ __forceinline__ int window_index(const GridView& g, const CubeWindow& w, const float4& p, int slot) {
  int i, j, k;
  label_cube(g, w, p, slot, &i, &j, &k);
  i -= w.w0[0]; j -= w.w0[1]; k -= w.w0[2];
  if (i < 0 || i > 6 || j < 0 || j > 6 || k < 0 || k > 6) return -1;
  return (i * 7 + j) * 7 + k;
}

// Sharded map (cm_dist.cu): the rank that owns cube (i, j, k) of the FeatureMap lattice.  Neighbouring cubes go to different
// ranks on purpose: the cubes around a sensor -- where all of a sweep's queries fall -- are spread over every rank.
__host__ __device__ __forceinline__ int cube_owner(int i, int j, int k, int nranks) {
  return (int)((unsigned int)(i + 3 * j + 5 * k) % (unsigned int)nranks);
}

// Candidate filter codes (KnnGeom::filt): CM_FILT_NONE: every point of the grid is searched; CM_FILT_WINDOW: only points
// whose cube is in the active window (FeatureMap::getSurroundFeature, FeatureMap.h:256-265); >= 0: only points of the
// cube with that linear index -- the localisation matcher searches the query's own cube (FeatureMap.h:521-527).
// slot: the candidate's pool slot (its epoch decides which cube STORES it).
#define CM_FILT_NONE (-1)
#define CM_FILT_WINDOW (-2)
__device__ __forceinline__ bool cand_ok(const GridView& g, int filt, const float4& p, int slot) {
  if (filt == CM_FILT_NONE) return true;
  const CubeWindow& w = *g.window;
  if (filt == CM_FILT_WINDOW) { const int wi = window_index(g, w, p, slot); return wi >= 0 && w.active[wi]; }
  int i, j, k;
  label_cube(g, w, p, slot, &i, &j, &k);
  return i >= 0 && i < w.dims[0] && j >= 0 && j < w.dims[1] && k >= 0 && k < w.dims[2] && (i + j * w.dims[0] + k * w.dims[0] * w.dims[1]) == filt;
}

// ---- exact distance ties ----------------------------------------------------------------------------------------------------
// Two map points at EXACTLY the same float distance from a query (about once per 30 sweeps and stream) must be ordered the way the
// reference's clouds order them, or the neighbour list -- and with it the last bit of a pose -- would depend on the pool slots, i.e.
// on the order in which atomics happened to append points.  For clouds handed in by the caller the list keys carry the original
// index (the oracle's rule: ties by index).  For the device-resident map the rule is the order of the reference's surround cloud
// (FeatureMap::getSurroundFeature, FeatureMap.h:256-265): cubes in (i, j, k) loop order, inside a cube pcl::VoxelGrid's output
// order, i.e. by voxel index (z, then y, then x).  Evaluated only when a tie actually occurs.
__device__ inline bool canon_less_map(const GridView& g, const float4 pa, int slot_a, const float4 pb, int slot_b) {
  if (g.window) {
    const CubeWindow& w = *g.window;
    int ia, ja, ka, ib, jb, kb;
    label_cube(g, w, pa, slot_a, &ia, &ja, &ka);
    label_cube(g, w, pb, slot_b, &ib, &jb, &kb);
    if (ia != ib) return ia < ib;
    if (ja != jb) return ja < jb;
    if (ka != kb) return ka < kb;
  }
  const float za = floorf(pa.z * g.inv_leaf), zb = floorf(pb.z * g.inv_leaf);
  if (za != zb) return za < zb;
  const float ya = floorf(pa.y * g.inv_leaf), yb = floorf(pb.y * g.inv_leaf);
  if (ya != yb) return ya < yb;
  const float xa = floorf(pa.x * g.inv_leaf), xb = floorf(pb.x * g.inv_leaf);
  if (xa != xb) return xa < xb;
  return slot_a < slot_b;   // two points of one voxel and cube (a drifted centroid next to the resident one)
}

// Insert with canonical tie handling.  kOrigIdx lists are canonical by construction (the key's low word is the original index).
// unique: the list may already hold this very point (warm start).
template <bool kOrigIdx>
__device__ __forceinline__ void top5_insert_canon(const GridView& g, Top5& t, float d, int idx, int slot, const float4& p, bool unique) {
  const unsigned long long key = top5_key(d, idx);
  if (kOrigIdx) {
    if (unique) top5_insert_key_unique(t, key, slot); else top5_insert_key(t, key, slot);
    return;
  }
  const unsigned int db = __float_as_uint(d);
  const bool tie = db == (unsigned int)(t.key[0] >> 32) || db == (unsigned int)(t.key[1] >> 32) || db == (unsigned int)(t.key[2] >> 32) ||
                   db == (unsigned int)(t.key[3] >> 32) || db == (unsigned int)(t.key[4] >> 32);
  if (!tie) { top5_insert_key(t, key, slot); return; }
  if (key == t.key[0] || key == t.key[1] || key == t.key[2] || key == t.key[3] || key == t.key[4]) return;   // the same point
  bool c[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const unsigned int dk = (unsigned int)(t.key[k] >> 32);
    c[k] = db < dk;
    if (db == dk) { const int sk = t.idx(k); c[k] = canon_less_map(g, p, slot, __ldg(g.pts + sk), sk); }
  }
  if (!c[4]) return;
  // new slot k = c_{k-1} ? old[k-1] : (c_k ? new : old[k])   (c is monotone: c_k implies c_{k+1})
  t.key[4] = c[3] ? t.key[3] : key;
  t.key[3] = c[2] ? t.key[2] : (c[3] ? key : t.key[3]);
  t.key[2] = c[1] ? t.key[1] : (c[2] ? key : t.key[2]);
  t.key[1] = c[0] ? t.key[0] : (c[1] ? key : t.key[1]);
  t.key[0] = c[0] ? key : t.key[0];
}

// Squared distance from u (voxel units) to the slab [lo, hi) of a cell along one axis, 0 inside.
__device__ __forceinline__ float slab_dist(float u, float lo, float hi) { float d = fmaxf(fmaxf(lo - u, u - hi), 0.f); return d; }

__device__ __forceinline__ unsigned long long entry_key(const uint4& e) { return (unsigned long long)e.x | ((unsigned long long)e.y << 32); }

// Candidates of one cell [start, start + count): loads are issued four at a time so they overlap.
// kOrigIdx: tie-break index = original cloud index stored in pts[].w (stateless clouds); otherwise the pool slot.
// filter: test every candidate's cube against the active window (only for queries near an inactive cube).
template <bool kOrigIdx>
__device__ __forceinline__ void scan_points(const GridView& g, unsigned int start, unsigned int count, float qx, float qy, float qz,
                                            int filt, Top5& best) {
  for (unsigned int j0 = 0; j0 < count; j0 += 4) {
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (j0 + u < count) p[u] = __ldg(g.pts + start + j0 + u);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (j0 + u < count) {
        if (!cand_ok(g, filt, p[u], (int)(start + j0 + u))) continue;
        float dx = qx - p[u].x, dy = qy - p[u].y, dz = qz - p[u].z;
        float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const int j = (int)(start + j0 + u);
        top5_insert(best, d, kOrigIdx ? __float_as_int(p[u].w) : j, j);
      }
    }
  }
}

template <bool kOrigIdx>
__device__ __forceinline__ void scan_cell(const GridView& g, int x, int y, int z, float qx, float qy, float qz, int filt,
                                          Top5& best) {
  unsigned int start, count;
  if (!grid_probe(g, x, y, z, &start, &count)) return;
  scan_points<kOrigIdx>(g, start, count, qx, qy, qz, filt, best);
}

// ------------------------------------------------------------------------------------------------------------------
// Exact 5-NN of one query per thread, executed by FULL WARPS (every lane must call, `valid` masks idle lanes).
// Exact provided the 5th neighbour lies within sqrt(gate) (the reference's 5.0 gate); beyond that the returned 5th
// distance is only an upper bound that is >= gate.
//
// Level 0 (per thread): the 2x2x2 cells nearest to the query -- 8 probes in two batches of 4 independent loads, the
// hit cells' point ranges are staged in shared memory and scanned in ONE flattened loop (a warp iterates
// max-over-lanes of the total candidate count, not the sum over cells of the per-cell maxima).
// Level L >= 1 (per warp): the few queries whose 5th distance is not yet provably final are finished one at a time
// by the whole warp: the lanes split the cells of the shell that extends the block by L cells on every side, keep
// lane-local 5-best lists, and a shuffle merge feeds the owning lane.  After level L every unseen point is farther than
//   r_L = 0.98 * leaf * (min over axes of the distance, in voxels, from the query to the faces of the level-0 block
//                        + L * kdiv)
// (cells are unions of PCL voxels and the voxel index floor(p * inv_leaf) is monotone in p; the 2 % margin covers
// the float rounding of p * inv_leaf).  Shell cells whose box lies farther than the current 5th distance (or the
// gate) are skipped without a probe.

// Geometry of one query relative to the grid: low cell of its 2x2x2 level-0 block, distance to the block faces.
struct KnnGeom {
  float fx, fy, fz;     // query in voxel units
  int lx, ly, lz;       // low cell of the level-0 block
  int own;              // position of the query's own cell inside the block (bit per axis)
  float m0;             // distance (voxel units) from the query to the nearest face of the block
  int filt;             // candidate filter code (CM_FILT_*)
};

// returns false when the coordinates are too large for the cell arithmetic (the query then has no neighbours)
__device__ __forceinline__ bool knn5_geom(const GridView& g, float qx, float qy, float qz, float gate, KnnGeom& c, bool own_cube = false) {
  const int k = g.kdiv;
  c.fx = qx * g.inv_leaf; c.fy = qy * g.inv_leaf; c.fz = qz * g.inv_leaf;
  const float flx = floorf(c.fx), fly = floorf(c.fy), flz = floorf(c.fz);
  c.lx = c.ly = c.lz = 0; c.own = 0; c.m0 = 0.f; c.filt = CM_FILT_NONE;
  // keep the casts defined for absurd coordinates
  if (!(fabsf(flx) < 1.0e6f && fabsf(fly) < 1.0e6f && fabsf(flz) < 1.0e6f)) return false;
  const int vx = (int)flx, vy = (int)fly, vz = (int)flz;
  const int cx = floor_div(vx, k), cy = floor_div(vy, k), cz = floor_div(vz, k);
  const float half = 0.5f * (float)k;
  // low cell of the 2-cell span that keeps the query >= cell/2 away from both ends
  c.lx = cx + (((float)(vx - cx * k) + (c.fx - flx)) < half ? -1 : 0);
  c.ly = cy + (((float)(vy - cy * k) + (c.fy - fly)) < half ? -1 : 0);
  c.lz = cz + (((float)(vz - cz * k) + (c.fz - flz)) < half ? -1 : 0);
  c.own = (cx - c.lx) | ((cy - c.ly) << 1) | ((cz - c.lz) << 2);
  if (own_cube) {
    // localisation: worldToIndex of the query (FeatureMap.h:464-487); no cube or fewer than 5 points in it -> no match (:522-523)
    if (!g.window || !g.cube_count) return false;
    const CubeWindow& w = *g.window;
    const int i = (int)(roundf(qx / w.cube_size) + (float)w.origin[0]), j = (int)(roundf(qy / w.cube_size) + (float)w.origin[1]),
              kk = (int)(roundf(qz / w.cube_size) + (float)w.origin[2]);
    if (!(i >= 0 && i < w.dims[0] && j >= 0 && j < w.dims[1] && kk >= 0 && kk < w.dims[2])) return false;
    c.filt = i + j * w.dims[0] + kk * w.dims[0] * w.dims[1];
    if (g.cube_count[c.filt] < 5) return false;
  } else if (g.window) {
    // candidates can only lie within sqrt(gate) of the query: per-point cube tests are needed only if one of the
    // (at most 8) cubes touched by that box is not searched
    const CubeWindow& w = *g.window;
    const float rg = sqrtf(gate) * 1.0001f;
    int i0 = (int)(roundf((qx - rg) / w.cube_size) + (float)w.origin[0]) - w.w0[0], i1 = (int)(roundf((qx + rg) / w.cube_size) + (float)w.origin[0]) - w.w0[0];
    int j0 = (int)(roundf((qy - rg) / w.cube_size) + (float)w.origin[1]) - w.w0[1], j1 = (int)(roundf((qy + rg) / w.cube_size) + (float)w.origin[1]) - w.w0[1];
    int k0 = (int)(roundf((qz - rg) / w.cube_size) + (float)w.origin[2]) - w.w0[2], k1 = (int)(roundf((qz + rg) / w.cube_size) + (float)w.origin[2]) - w.w0[2];
    bool f = false;
    if (i0 < 0 || i1 > 6 || j0 < 0 || j1 > 6 || k0 < 0 || k1 > 6) f = true;
    else {
      for (int i = i0; i <= i1; i++)
        for (int j = j0; j <= j1; j++)
          for (int kk = k0; kk <= k1; kk++) f = f || !w.active[(i * 7 + j) * 7 + kk];
    }
    if (f || g.displaced) c.filt = CM_FILT_WINDOW;   // displaced content: a nearby point may be stored in an inactive cube
  }
  // distance (voxel units) from the query to the nearest face of the level-0 block
  const float lox = (float)(c.lx * k), loy = (float)(c.ly * k), loz = (float)(c.lz * k);
  const float span = (float)(2 * k);
  c.m0 = fminf(fminf(c.fx - lox, lox + span - c.fx), fminf(fminf(c.fy - loy, loy + span - c.fy), fminf(c.fz - loz, loz + span - c.fz)));
  return true;
}

#ifndef CM_KNN_UNROLL
#define CM_KNN_UNROLL 4      // candidate loads in flight per thread and loop iteration
#endif
// Level 0 of one query (per thread).  Returns true when the list is not yet provably final (levels >= 1 needed).
// The 8 cells are probed together; their point ranges are staged in shared memory with the squared lower bound of
// the distance from the query to the cell's box, own cell first, then face / edge / corner neighbours.  A range is
// skipped when that bound already exceeds the current 5th distance (every point of the cell is then strictly
// farther than five known points) -- on a 0.4 m map this drops about half of the candidate loads.
// rng: shared memory, 8 * blockDim.x uint4 (this thread uses rng[c * blockDim.x + threadIdx.x]).
// Warm start: prev (optional) = the pool slots of the query's 5 neighbours in the PREVIOUS Gauss-Newton iteration.  The pose moves
// by centimetres between iterations, so those five real map points give the list a near-final 5th distance before the first
// cell is opened: most cells are pruned by their lower bound and almost no candidate passes the insert test, which is where
// the search spent a third of its instructions at 5 of 32 lanes (the divergent sorted insert).  Any five valid points are a
// correct start: the result is the exact top-5 of (cells scanned) U (start points), the same set either way.
// KTH: the list entry whose distance the search has to prove final -- 4 for the 5 nearest neighbours, 0 when the caller only uses the
// nearest one (the odometry's correspondences): cells and shells beyond the current KTH-th distance are skipped, the entries behind
// KTH are then not the true runners-up.
template <bool kOrigIdx, int KTH = 4>
__device__ __forceinline__ bool knn5_level0(const GridView& g, const KnnGeom& c, float qx, float qy, float qz, uint4* rng, Top5& best,
                                            unsigned int* ncand = nullptr, const int* prev = nullptr, bool* tie_out = nullptr) {
  float d6 = FLT_MAX;   // smallest distance turned away from / pushed out of the list (map grids: exact-tie detection, top5_insert_track)
  const int k = g.kdiv;
  const float kf = (float)k;
  const float leaf98 = 0.98f * (g.cell / kf);
  int nr = 0;
  uint4* my = rng + threadIdx.x;
  const int stride = blockDim.x;
  bool warm = false;
  if (prev) {
    int sl[5];
#pragma unroll
    for (int u = 0; u < 5; u++) sl[u] = prev[u];
    if (sl[0] >= 0) {
      float4 q[5];
#pragma unroll
      for (int u = 0; u < 5; u++) q[u] = __ldg(g.pts + sl[u]);
#pragma unroll
      for (int u = 0; u < 5; u++) {
        if (cand_ok(g, c.filt, q[u], sl[u])) {
          float dx = qx - q[u].x, dy = qy - q[u].y, dz = qz - q[u].z;
          float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          top5_insert_key_unique(best, top5_key(d, kOrigIdx ? __float_as_int(q[u].w) : sl[u]), sl[u]);
        }
      }
      warm = true;
    }
  }
#pragma unroll
  for (int b = 0; b < 2; b++) {
    unsigned long long key[4]; uint4 e[4];
#pragma unroll
    for (int cI = 0; cI < 4; cI++) {
      const int cc = c.own ^ ((0x76534210 >> (4 * (b * 4 + cI))) & 7);   // xor masks 0, 1, 2, 4, 3, 5, 6, 7
      key[cI] = pack_cell(c.lx + (cc & 1), c.ly + ((cc >> 1) & 1), c.lz + (cc >> 2));
      e[cI] = __ldg(reinterpret_cast<const uint4*>(g.entries + (hash_cell(key[cI]) & g.mask)));
    }
#pragma unroll
    for (int cI = 0; cI < 4; cI++) {
      unsigned long long kk = entry_key(e[cI]);
      if (kk != key[cI] && kk != CM_EMPTY_KEY) {   // linear probing (rare)
        unsigned int h = hash_cell(key[cI]) & g.mask, probes = 0;
        do { h = (h + 1) & g.mask; e[cI] = __ldg(reinterpret_cast<const uint4*>(g.entries + h)); kk = entry_key(e[cI]); }
        while (kk != key[cI] && kk != CM_EMPTY_KEY && ++probes <= g.mask);
      }
      if (kk == key[cI] && e[cI].w > 0) {
        const int cc = c.own ^ ((0x76534210 >> (4 * (b * 4 + cI))) & 7);
        const float x0 = (float)((c.lx + (cc & 1)) * k), y0 = (float)((c.ly + ((cc >> 1) & 1)) * k), z0 = (float)((c.lz + (cc >> 2)) * k);
        const float dxv = slab_dist(c.fx, x0, x0 + kf), dyv = slab_dist(c.fy, y0, y0 + kf), dzv = slab_dist(c.fz, z0, z0 + kf);
        const float lb = leaf98 * sqrtf(dxv * dxv + dyv * dyv + dzv * dzv);
        my[nr * stride] = make_uint4(e[cI].z, e[cI].w, __float_as_uint(lb * lb), 0u);
        nr++;
      }
    }
  }
  // ---- one flattened candidate loop ----
  unsigned int scanned = 0;
  int ci = 0; unsigned int j0 = 0;
  uint4 r = nr ? my[0] : make_uint4(0u, 0u, 0u, 0u);
  const bool nofilt = c.filt == CM_FILT_NONE;
  while (ci < nr) {
    if (j0 == 0 && __uint_as_float(r.z) > best.d(KTH)) {   // the whole cell is farther than the current 5th neighbour
      ci++; if (ci < nr) r = my[ci * stride];
      continue;
    }
    const unsigned int left = r.y - j0;
    float4 p[CM_KNN_UNROLL];
#pragma unroll
    for (int u = 0; u < CM_KNN_UNROLL; u++)
      if ((unsigned int)u < left) p[u] = __ldg(g.pts + r.x + j0 + u);
    // keys of the batch, then ONE divergent region that inserts the (few) candidates below the current 5th distance, lowest
    // position first: the warp iterates max-over-lanes of the number of passing candidates instead of once per position
    unsigned long long kk[CM_KNN_UNROLL];
    unsigned int pass = 0;
    const float d5 = best.d(KTH);   // candidates at or below the 5th distance go to the exact (d2, index) test of the insert
#pragma unroll
    for (int u = 0; u < CM_KNN_UNROLL; u++) {
      kk[u] = CM_TOP5_EMPTY;
      if ((unsigned int)u < left && (nofilt || cand_ok(g, c.filt, p[u], (int)(r.x + j0 + u)))) {
        float dx = qx - p[u].x, dy = qy - p[u].y, dz = qz - p[u].z;
        float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        kk[u] = top5_key(d, kOrigIdx ? __float_as_int(p[u].w) : (int)(r.x + j0 + u));
        pass |= (d <= d5 ? 1u : 0u) << u;
      }
    }
    while (pass) {
      const int u = __ffs(pass) - 1;
      pass &= pass - 1;
      unsigned long long kq = kk[0];
#pragma unroll
      for (int v = 1; v < CM_KNN_UNROLL; v++) kq = (u == v) ? kk[v] : kq;
      if (kOrigIdx) { if (warm) top5_insert_key_unique(best, kq, (int)(r.x + j0 + u)); else top5_insert_key(best, kq, (int)(r.x + j0 + u)); }
      else top5_insert_track(best, kq, (int)(r.x + j0 + u), warm, d6);   // map grids: remember what was turned away (exact ties)
    }
    if (ncand) scanned += left < (unsigned int)CM_KNN_UNROLL ? left : (unsigned int)CM_KNN_UNROLL;
    j0 += CM_KNN_UNROLL;
    if (j0 >= r.y) { ci++; j0 = 0; if (ci < nr) r = my[ci * stride]; }
  }
  if (ncand) *ncand = scanned;
  if (tie_out) *tie_out = !kOrigIdx && top5_has_tie(best, d6);   // the caller re-runs such a query canonically (top5_insert_canon)
  if (g.max_level < 1) return false;
  const float r0 = leaf98 * c.m0;
  return !(best.d(KTH) < r0 * r0);
}

// Levels >= 1 of ONE query, executed by a whole warp.  The query (position, geometry, current list) lives in lane h;
// on return lane h's list is final.
// L0 = 0 repeats the level-0 block as well (a query whose per-thread pass met an exact distance tie starts over from an empty list:
// this path orders ties canonically).
template <bool kOrigIdx, int KTH = 4>
__device__ __forceinline__ void knn5_warp_finish(const GridView& g, int h, const KnnGeom& c, float qx, float qy, float qz, float gate,
                                                 Top5& best, int L0 = 1) {
  const unsigned int FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int k = g.kdiv;
  const float leaf = g.cell / (float)k;
  const float bqx = __shfl_sync(FULL, qx, h), bqy = __shfl_sync(FULL, qy, h), bqz = __shfl_sync(FULL, qz, h);
  const float bfx = __shfl_sync(FULL, c.fx, h), bfy = __shfl_sync(FULL, c.fy, h), bfz = __shfl_sync(FULL, c.fz, h);
  const int blx = __shfl_sync(FULL, c.lx, h), bly = __shfl_sync(FULL, c.ly, h), blz = __shfl_sync(FULL, c.lz, h);
  const float bm0 = __shfl_sync(FULL, c.m0, h);
  const int bfilter = __shfl_sync(FULL, c.filt, h);
  float bd5 = __shfl_sync(FULL, best.d(KTH), h);
  const float kf = (float)k;
  for (int L = L0; L <= g.max_level; L++) {
    const float rr = 0.98f * leaf * (bm0 + (float)((L - 1) * k));   // radius guaranteed by the previous level
    if (L > 0 && bd5 < rr * rr) break;
    const int n = 2 + 2 * L, ncells = n * n * n;
    const float bound = fminf(bd5, gate);
    Top5 loc;
    top5_init(loc);
    // the lanes probe 32 shell cells at a time; the non-empty ones are then scanned by the WHOLE warp, one cell after the
    // other, lane j taking every 32nd point (coalesced loads).  Hard queries live where the map is sparse: few of the 56
    // shell cells hold points, so "every lane scans its own cell" would leave 2-5 lanes busy.
    for (int t0 = 0; t0 < ncells; t0 += 32) {
      const int t = t0 + lane;
      unsigned int start = 0, count = 0;
      if (t < ncells) {
        const int dz = t / (n * n), dy = (t / n) % n, dx = t % n;
        const bool interior = dx > 0 && dx < n - 1 && dy > 0 && dy < n - 1 && dz > 0 && dz < n - 1;   // visited at the previous levels
        if (!interior) {
          const float cxl = (float)((blx - L + dx) * k), cyl = (float)((bly - L + dy) * k), czl = (float)((blz - L + dz) * k);
          const float dxv = slab_dist(bfx, cxl, cxl + kf), dyv = slab_dist(bfy, cyl, cyl + kf), dzv = slab_dist(bfz, czl, czl + kf);
          const float lb = 0.98f * leaf * sqrtf(dxv * dxv + dyv * dyv + dzv * dzv);   // lower bound of the distance to this cell
          if (lb * lb < bound && !grid_probe(g, blx - L + dx, bly - L + dy, blz - L + dz, &start, &count)) count = 0;
        }
      }
      unsigned int full = __ballot_sync(FULL, count > 0);
      while (full) {
        const int src = __ffs(full) - 1;
        full &= full - 1;
        const unsigned int st = __shfl_sync(FULL, start, src), ct = __shfl_sync(FULL, count, src);
        for (unsigned int j = lane; j < ct; j += 32) {
          const float4 p = __ldg(g.pts + st + j);
          if (!cand_ok(g, bfilter, p, (int)(st + j))) continue;
          float ddx = bqx - p.x, ddy = bqy - p.y, ddz = bqz - p.z;
          float d = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
          const int jj = (int)(st + j);
          if (d <= loc.d(4)) top5_insert_canon<kOrigIdx>(g, loc, d, kOrigIdx ? __float_as_int(p.w) : jj, jj, p, false);
        }
      }
    }
    // merge the lane-local lists into the owner's list: at most 5 winners can enter
    for (int round = 0; round < 5; round++) {
      const unsigned int dbits = (unsigned int)(loc.key[0] >> 32), ibits = (unsigned int)(loc.key[0] & 0xFFFFFFFFull);
      const unsigned int mind = __reduce_min_sync(FULL, dbits);
      if (mind == __float_as_uint(FLT_MAX)) break;
      const unsigned int cand = (dbits == mind) ? ibits : 0xFFFFFFFFu;
      unsigned int mini = __reduce_min_sync(FULL, cand);
      unsigned int tied = __ballot_sync(FULL, dbits == mind);
      int src = __ffs(__ballot_sync(FULL, dbits == mind && ibits == mini)) - 1;
      const int myslot = kOrigIdx ? loc.slot[0] : (int)ibits;
      // the winner's point travels to the owner; on an exact distance tie between lanes the canonical order decides (map grids)
      float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dbits == mind && myslot >= 0) hp = __ldg(g.pts + myslot);
      if (!kOrigIdx && (tied & (tied - 1))) {
        src = __ffs(tied) - 1; tied &= tied - 1;
        while (tied) {
          const int o = __ffs(tied) - 1; tied &= tied - 1;
          const float4 pa = make_float4(__shfl_sync(FULL, hp.x, o), __shfl_sync(FULL, hp.y, o), __shfl_sync(FULL, hp.z, o), 0.f);
          const float4 pb = make_float4(__shfl_sync(FULL, hp.x, src), __shfl_sync(FULL, hp.y, src), __shfl_sync(FULL, hp.z, src), 0.f);
          const int sa = __shfl_sync(FULL, myslot, o), sb = __shfl_sync(FULL, myslot, src);
          if (canon_less_map(g, pa, sa, pb, sb)) src = o;
        }
        mini = __shfl_sync(FULL, ibits, src);
      }
      const int wslot = __shfl_sync(FULL, myslot, src);
      const float4 wp = make_float4(__shfl_sync(FULL, hp.x, src), __shfl_sync(FULL, hp.y, src), __shfl_sync(FULL, hp.z, src), __shfl_sync(FULL, hp.w, src));
      // unique: a list that was warm-started from the previous iteration's neighbours may already hold points of this shell
      if (lane == h) top5_insert_canon<kOrigIdx>(g, best, __uint_as_float(mind), (int)mini, wslot, wp, true);
      if (lane == src) {
#pragma unroll
        for (int u = 0; u < 4; u++) { loc.key[u] = loc.key[u + 1]; loc.slot[u] = loc.slot[u + 1]; }
        loc.key[4] = CM_TOP5_EMPTY; loc.slot[4] = -1;
      }
    }
    bd5 = __shfl_sync(FULL, best.d(KTH), h);
  }
}

// Level 0 per thread, then the warp finishes its unresolved queries one at a time (callers without a deferred pass).
template <bool kOrigIdx, int KTH = 4>
__device__ __forceinline__ void knn5_search(const GridView& g, bool valid, float qx, float qy, float qz, float gate, uint4* rng,
                                            Top5& best) {
  top5_init(best);
  KnnGeom c;
  valid = knn5_geom(g, qx, qy, qz, gate, c, false) && valid;
  bool need = false;
  if (valid) need = knn5_level0<kOrigIdx, KTH>(g, c, qx, qy, qz, rng, best);
  unsigned int hard = __ballot_sync(0xffffffffu, need);
  while (hard) {
    const int h = __ffs(hard) - 1;
    hard &= hard - 1;
    knn5_warp_finish<kOrigIdx, KTH>(g, h, c, qx, qy, qz, gate, best);
  }
}

// Column-parallel form of cm_math.h::colpiv_qr_solve<6, 6> for one warp: the SAME float operations on the same operands in the same
// order as the sequential routine the oracle runs (every dot product and norm is still summed row by row inside one lane; -fmad=false),
// only spread over the lanes: lane c < 6 holds column c of A in a[0..5], lane 6 holds b (the reflectors are applied to it as to a
// seventh trailing column, step by step, instead of in a second sweep), the other lanes carry zeros.  The pivot scan, the reflector
// of column k and the back substitution are evaluated redundantly by every lane from shuffled values, so the control flow is
// uniform.  One lane running the unrolled sequential routine took 35 us per Gauss-Newton iteration (~5,000 dependent instructions,
// 80 KB of straight-line code fetched once); this form is ~7x shorter.  Every lane returns the whole solution X.
__device__ __forceinline__ void warp_qr_solve6(float (&a)[6], const int lane, float (&X)[6]) {
  const unsigned int FULL = 0xffffffffu;
  float nu, nd;
  int perm = lane;
  {
    float sq = 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r) sq += a[r] * a[r];
    nd = sqrtf(sq); nu = nd;
  }
  float maxnorm = 0.f;
#pragma unroll
  for (int k = 0; k < 6; ++k) { const float v = __shfl_sync(FULL, nu, k); if (v > maxnorm) maxnorm = v; }
  const float te = maxnorm * FLT_EPSILON;
  const float threshold_helper = (te * te) / 6.0f;
  const float norm_downdate_threshold = sqrtf(FLT_EPSILON);
  int nzp = 6;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int big = k;
    float bigv = __shfl_sync(FULL, nu, k);
#pragma unroll
    for (int j = k + 1; j < 6; ++j) { const float v = __shfl_sync(FULL, nu, j); if (v > bigv) { bigv = v; big = j; } }
    const float big_sq = bigv * bigv;
    if (nzp == 6 && big_sq < threshold_helper * (float)(6 - k)) nzp = k;
    {   // column swap k <-> big (whole columns, with their norms and permutation entry)
      const int src = lane == k ? big : (lane == big ? k : lane);
#pragma unroll
      for (int r = 0; r < 6; ++r) a[r] = __shfl_sync(FULL, a[r], src);
      nu = __shfl_sync(FULL, nu, src); nd = __shfl_sync(FULL, nd, src); perm = __shfl_sync(FULL, perm, src);
    }
    // Householder reflector of column k, rows k..5
    float ck[6], v[6];
#pragma unroll
    for (int r = k; r < 6; ++r) ck[r] = __shfl_sync(FULL, a[r], k);
    float tailSqNorm = 0.f;
#pragma unroll
    for (int r = k + 1; r < 6; ++r) tailSqNorm += ck[r] * ck[r];
    const float c0 = ck[k];
    float tau, beta;
    if (k == 5 || tailSqNorm <= FLT_MIN) {
      tau = 0.f; beta = c0;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) v[r] = 0.f;
    } else {
      float bb = sqrtf(c0 * c0 + tailSqNorm);
      if (c0 >= 0.f) bb = -bb;
      const float d = c0 - bb;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) v[r] = ck[r] / d;
      tau = (bb - c0) / bb;
      beta = bb;
    }
    if (lane == k) {
      a[k] = beta;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) a[r] = v[r];
    }
    // H_k on the trailing columns and on b (b only while k < nonzero_pivots: colpiv_qr_solve's second sweep stops there)
    if ((lane > k && lane < 6) || (lane == 6 && k < nzp)) {
      float tmp = a[k];
#pragma unroll
      for (int r = k + 1; r < 6; ++r) tmp += v[r] * a[r];
      a[k] -= tau * tmp;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) a[r] -= tau * v[r] * tmp;
    }
    if (lane > k && lane < 6 && nu != 0.f) {   // LAPACK-style norm downdate
      float temp = fabsf(a[k]) / nu;
      temp = (1.f + temp) * (1.f - temp);
      temp = temp < 0.f ? 0.f : temp;
      const float ratio = nu / nd;
      const float temp2 = temp * (ratio * ratio);
      if (temp2 <= norm_downdate_threshold) {
        float sq = 0.f;
#pragma unroll
        for (int r = k + 1; r < 6; ++r) sq += a[r] * a[r];
        nd = sqrtf(sq); nu = nd;
      } else {
        nu *= sqrtf(temp);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) X[i] = 0.f;
  if (nzp == 0) return;
  float bq[6], Rm[6][6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    bq[i] = __shfl_sync(FULL, a[i], 6);
#pragma unroll
    for (int j = i; j < 6; ++j) Rm[i][j] = __shfl_sync(FULL, a[i], j);
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    if (i < nzp) {
      float sb = bq[i];
#pragma unroll
      for (int j = i + 1; j < 6; ++j)
        if (j < nzp) sb -= Rm[i][j] * bq[j];
      bq[i] = sb / Rm[i][i];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int pi = __shfl_sync(FULL, perm, i);
    if (i < nzp) {
#pragma unroll
      for (int c = 0; c < 6; ++c)
        if (pi == c) X[c] = bq[i];
    }
  }
}


// true when the smallest eigenvalue of the symmetric 6x6 float matrix A is PROVABLY above `threshold` as the restated Eigen solver
// would compute it: A - shift I is positive definite (LDL^T in double, every pivot clearly positive) for shift = threshold + 1e-4
// trace(A).  The float eigen-solver's error is a small multiple of eps_float * ||A|| (<= 1e-5 trace with room to spare), so above the
// shift its smallest eigenvalue cannot come out below the threshold and the 6x6 eigen-decomposition of evaluation 0 (one lane,
// ~5,000 dependent instructions, 20 us) is skipped for the usual, well-conditioned frame.  Anything else -- degenerate, borderline,
// non-finite -- takes the full solve as before.
static __device__ __noinline__ bool dev_min_eig_above6(const float* A, float threshold) {
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < 6; i++) tr += (double)A[i * 6 + i];
  if (!(tr > 0.0) || !(tr < 1e300)) return false;
  const double shift = (double)threshold + 1e-4 * tr, tiny = 1e-9 * tr;
  double L[36], D[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double d = (double)A[j * 6 + j] - shift;
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
    if (!(d > tiny)) return false;
    D[j] = d;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double v = (double)A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; k++) v -= L[i * 6 + k] * L[j * 6 + k] * D[k];
      L[i * 6 + j] = v / d;
    }
  }
  return true;
}


}  // namespace cm
