// cm_match.cu -- scan-to-map registration kernels (K4 grid build, K5 fused correspondence, K6 reduce + solve).
//
// Replaces ScanMatch::scanMatchScan (L_SLAM/src/scan_to_scan_match/ScanMatch.cpp:51-347) and the helpers it
// calls (util/feature_utils.h:17-26,63-75,97-204; util/transform_utils.h:288-311,476-482; util/Angle.h).
// Compiled with -fmad=false: every float operation below is an IEEE operation in source order, which is what
// makes rows, neighbour sets and poses bit-identical to the CPU oracle.
#include "cm_match.cuh"
#include "cm_math.h"
#include "cm_host.h"
#include <vector>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

namespace cm {

// ============================================================================================================
// K4: grid build for a stateless reference cloud (the reference rebuilds two KD-trees per call,
// ScanMatch.cpp:75-76).  count -> offsets -> scatter; O(M), no sort.
// ============================================================================================================
__global__ void grid_clear_kernel(CellEntry* e, unsigned int cap) {
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { e[i].key = CM_EMPTY_KEY; e[i].start = 0; e[i].count = 0; }
}

__global__ void grid_count_kernel(const float4* __restrict__ pts, int n, CellEntry* entries, unsigned int mask, float inv,
                                  int* __restrict__ cell_of) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) { cell_of[i] = -1; return; }
  float fx = floorf(p.x * inv), fy = floorf(p.y * inv), fz = floorf(p.z * inv);
  if (!(fabsf(fx) < 1.0e6f && fabsf(fy) < 1.0e6f && fabsf(fz) < 1.0e6f)) { cell_of[i] = -1; return; }
  unsigned long long key = pack_cell((int)fx, (int)fy, (int)fz);
  unsigned int h = hash_cell(key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&entries[h].key, CM_EMPTY_KEY, key);
    if (prev == CM_EMPTY_KEY || prev == key) break;
    h = (h + 1) & mask;
  }
  atomicAdd(&entries[h].count, 1u);
  cell_of[i] = (int)h;
}

__global__ void grid_offsets_kernel(CellEntry* entries, unsigned int cap, unsigned int* cursor) {
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  unsigned int c = entries[i].count;
  if (c) { entries[i].start = atomicAdd(cursor, c); entries[i].count = 0; }
}

__global__ void grid_scatter_kernel(const float4* __restrict__ pts, int n, CellEntry* entries, const int* __restrict__ cell_of,
                                    float4* __restrict__ out, int keep_w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int h = cell_of[i];
  if (h < 0) return;
  float4 p = pts[i];
  unsigned int j = atomicAdd(&entries[h].count, 1u);
  if (!keep_w) p.w = __int_as_float(i);
  out[entries[h].start + j] = p;
}

// ---- the same for a batch of independent clouds (blockIdx.y = stream; on[s] == 0: that stream keeps its old grid) -----------------
__global__ void gridb_clear_kernel(CellEntry* e, unsigned int tcap, unsigned int* cursor, const int* __restrict__ on) {
  const int s = blockIdx.y;
  if (!on[s]) return;
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < tcap) { CellEntry& c = e[(size_t)s * tcap + i]; c.key = CM_EMPTY_KEY; c.start = 0; c.count = 0; }
  if (i == 0) cursor[s] = 0u;
}
__global__ void gridb_count_kernel(const float4* __restrict__ pts, const int* __restrict__ n, int cap, CellEntry* entries, unsigned int tcap,
                                   float inv, int* __restrict__ cell_of, const int* __restrict__ on) {
  const int s = blockIdx.y;
  if (!on[s]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n[s]) return;
  CellEntry* ent = entries + (size_t)s * tcap;
  const unsigned int mask = tcap - 1;
  int* co = cell_of + (size_t)s * cap;
  float4 p = pts[(size_t)s * cap + i];
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) { co[i] = -1; return; }
  float fx = floorf(p.x * inv), fy = floorf(p.y * inv), fz = floorf(p.z * inv);
  if (!(fabsf(fx) < 1.0e6f && fabsf(fy) < 1.0e6f && fabsf(fz) < 1.0e6f)) { co[i] = -1; return; }
  unsigned long long key = pack_cell((int)fx, (int)fy, (int)fz);
  unsigned int h = hash_cell(key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&ent[h].key, CM_EMPTY_KEY, key);
    if (prev == CM_EMPTY_KEY || prev == key) break;
    h = (h + 1) & mask;
  }
  atomicAdd(&ent[h].count, 1u);
  co[i] = (int)h;
}
__global__ void gridb_offsets_kernel(CellEntry* entries, unsigned int tcap, unsigned int* cursor, const int* __restrict__ on) {
  const int s = blockIdx.y;
  if (!on[s]) return;
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tcap) return;
  CellEntry& c = entries[(size_t)s * tcap + i];
  const unsigned int cnt = c.count;
  if (cnt) { c.start = atomicAdd(&cursor[s], cnt); c.count = 0; }
}
__global__ void gridb_scatter_kernel(const float4* __restrict__ pts, const int* __restrict__ n, int cap, CellEntry* entries, unsigned int tcap,
                                     const int* __restrict__ cell_of, float4* __restrict__ out, const int* __restrict__ on) {
  const int s = blockIdx.y;
  if (!on[s]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n[s]) return;
  const int h = cell_of[(size_t)s * cap + i];
  if (h < 0) return;
  CellEntry& c = entries[(size_t)s * tcap + h];
  float4 p = pts[(size_t)s * cap + i];
  const unsigned int j = atomicAdd(&c.count, 1u);
  p.w = __int_as_float(i);           // original index: tie-break + reported neighbour
  out[(size_t)s * cap + c.start + j] = p;
}
__global__ void gridb_view_kernel(GridView* views, const CellEntry* entries, unsigned int tcap, const float4* pts, int cap, const int* __restrict__ n,
                                  float inv, float cell, int max_level, const int* __restrict__ on, int nstreams) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nstreams || !on[s]) return;
  GridView v;
  v.entries = entries + (size_t)s * tcap; v.pts = pts + (size_t)s * cap; v.mask = tcap - 1; v.inv_leaf = inv; v.kdiv = 1; v.cell = cell;
  v.npts = n[s]; v.max_level = max_level; v.window = nullptr; v.cube_count = nullptr; v.epoch = nullptr; v.eoff = nullptr; v.displaced = 0;
  views[s] = v;
}

// ============================================================================================================
// Pose bookkeeping (Twist + Angle caches + Isometry3f rotation)
// ============================================================================================================
__device__ void state_set_pose(MatchState& st, const float pose[6]) {
  for (int k = 0; k < 6; k++) st.pose[k] = pose[k];
  for (int k = 0; k < 3; k++) cm_sincosf(pose[k], &st.sn[k], &st.cs[k]);   // Angle(float), Angle.h:19-20
  pose_to_matrix(pose, st.R);                                             // convertTransform, transform_utils.h:308-311
}

__global__ void match_init_kernel(MatchState* states, const float* __restrict__ poses, const GridView* gc, const GridView* gs,
                                  MatchParamsDev prm, int nstreams) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nstreams) return;
  MatchState& st = states[s];
  state_set_pose(st, poses + 6 * s);
  for (int k = 0; k < 36; k++) st.P[k] = 0.f;
  st.flags = 0; st.iterations = 0; st.rows = st.line = st.plane = 0; st.score = 0.0; st.done = 0;
  if (gc[s].npts < prm.min_ref_corner || gs[s].npts < prm.min_ref_surf) { st.flags = CM_F_TOO_FEW_REF; st.done = 1; }   // ScanMatch.cpp:57-61
}

// ============================================================================================================
// K5: correspondence (ScanMatch.cpp:97-132 and :154-204): K5a transform -> exact 5-NN (memory-latency bound, divergent,
// wants occupancy) and K5b line / plane fit -> residual -> Jacobian row (uniform arithmetic) are separate launches so
// that each gets the register budget and occupancy it needs; the 20 bytes per query between them stay in L2.
// ============================================================================================================
struct CorrArgs {
  const float4* corner; const float4* surf;   // [nstreams][cap*]
  const int* n_corner; const int* n_surf;     // per stream (device)
  int cap_corner, cap_surf;
  const GridView* grid_corner; const GridView* grid_surf;
  const MatchState* state;
  RowOut* rows;                               // [nstreams][cap_corner + cap_surf]
  int* nn_slot;                               // [nstreams][cap_corner + cap_surf][5] pool slots of the neighbours, -1: gated out
  int* nn;                                    // optional [nstreams][cap_corner + cap_surf][5], -1 where gated out
  const float* own_box;                       // optional {lo[3], hi[3]}: only queries whose map-frame position is inside are evaluated (sharded map)
  unsigned long long* dbg;                    // optional per-warp trace of search_kernel (development aid, cm_debug_search_trace)
  const int* iter_dev;                        // optional: the evaluation index lives in device memory (graph WHILE loop); hard_count is then the row base
  int iter;                                   // the evaluation index when iter_dev is NULL
  int warm;                                   // 1: warm-start the 5-NN lists from the previous iteration (COOPERMAP_NO_WARM=1: off)
  int spread;                                 // search_kernel: a warp takes 32 >> spread queries (its first lanes), see search_spread()
  const int* skip;                            // optional: non-zero = do nothing (MatchLaunch::skip)
  int dist_rank, dist_nranks;                 // sharded map: only the queries whose map-frame cube this rank owns are evaluated
  void* hard; int* hard_count; int hard_cap;  // optional device-wide list of the queries that need levels >= 1 (this evaluation's counter)
  MatchParamsDev prm;
};

struct PoseCoef {   // pose-only factors of the Jacobian text at ScanMatch.cpp:185-195, same association
  float x1, x2, x3, x4, x5, x6;
  float y1, y2, y3, y4, y5, y6, y7, y8, y9;
  float z1, z3, z4, z5, z6, z7;
};
__device__ __forceinline__ void make_pose_coef(const MatchState& st, PoseCoef& k) {
  float srx = st.sn[0], crx = st.cs[0], sry = st.sn[1], cry = st.cs[1], srz = st.sn[2], crz = st.cs[2];
  k.x1 = crz * sry * crx + srz * srx;  k.x2 = srz * crx - crz * sry * srx;
  k.x3 = srz * sry * crx - crz * srx;  k.x4 = srz * sry * srx + crz * crx;
  k.x5 = cry * crx;                    k.x6 = cry * srx;
  k.y1 = -crz * sry;  k.y2 = crz * cry * srx;  k.y3 = crz * cry * crx;
  k.y4 = -srz * sry;  k.y5 = srz * cry * srx;  k.y6 = srz * cry * crx;
  k.y7 = -cry;        k.y8 = sry * srx;        k.y9 = sry * crx;
  k.z1 = -srz * cry;  k.z3 = crz * srx - srz * sry * crx;
  k.z4 = crz * cry;   k.z5 = crz * sry * srx - srz * crx;  k.z6 = crz * sry * crx;  k.z7 = srz * srx;
}

__device__ __forceinline__ float norm3f(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

// findLine (feature_utils.h:108-154) + getCornerFeatureCoefficients (:17-26, 63-75).  Returns bit0 kept, bit1 counted.
__device__ __forceinline__ int corner_row(const float4 nb[5], float sx, float sy, float sz, float co[4]) {
  float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
  for (int j = 0; j < 5; j++) { cx += nb[j].x; cy += nb[j].y; cz += nb[j].z; }
  cx /= 5.0f; cy /= 5.0f; cz /= 5.0f;
  float a00 = 0.f, a10 = 0.f, a20 = 0.f, a11 = 0.f, a21 = 0.f, a22 = 0.f;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    float ax = nb[j].x - cx, ay = nb[j].y - cy, az = nb[j].z - cz;
    a00 += ax * ax; a10 += ax * ay; a20 += ax * az; a11 += ay * ay; a21 += ay * az; a22 += az * az;
  }
  float A[6] = {a00 / 5.0f, a10 / 5.0f, a20 / 5.0f, a11 / 5.0f, a21 / 5.0f, a22 / 5.0f};
  float w[3], V[9];
  eig3_sym(A, w, V);
  if (!(w[2] > 5.f * w[1])) return 0;
  float vx = V[2], vy = V[5], vz = V[8];
  float Ax = cx - vx * 0.1f, Ay = cy - vy * 0.1f, Az = cz - vz * 0.1f;
  float Bx = cx + vx * 0.1f, By = cy + vy * 0.1f, Bz = cz + vz * 0.1f;
  float bx = sx - Bx, by = sy - By, bz = sz - Bz;
  float ax = sx - Ax, ay = sy - Ay, az = sz - Az;
  float kx = by * az - bz * ay, ky = bz * ax - bx * az, kz = bx * ay - by * ax;
  float knorm = norm3f(kx, ky, kz);
  float lengthAB = norm3f(Ax - Bx, Ay - By, Az - Bz);
  float ex = Bx - Ax, ey = By - Ay, ez = Bz - Az;
  float ux = ky * ez - kz * ey, uy = kz * ex - kx * ez, uz = kx * ey - ky * ex;
  float den = knorm * lengthAB;
  float dirx = -ux / den, diry = -uy / den, dirz = -uz / den;
  float distance = knorm / lengthAB;
  float weight = (float)(1 - 0.9f * fabs((double)distance));
  co[0] = dirx * weight; co[1] = diry * weight; co[2] = dirz * weight; co[3] = distance * weight;
  return ((double)weight > 0.1) ? 3 : 2;
}

// findPlane (feature_utils.h:157-204) + getSurfaceFeatureCoefficients (:97-106).
__device__ __forceinline__ int surf_row(const float4 nb[5], float sx, float sy, float sz, float max_dist, float co[4]) {
  float A[15], B[5], X[3];
  float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    cx += nb[j].x; cy += nb[j].y; cz += nb[j].z;
    A[j * 3 + 0] = nb[j].x; A[j * 3 + 1] = nb[j].y; A[j * 3 + 2] = nb[j].z;
    B[j] = -1.f;
  }
  cx /= 5.0f; cy /= 5.0f; cz /= 5.0f;
  colpiv_qr_solve<5, 3>(A, B, X);
  float p0 = X[0], p1 = X[1], p2 = X[2], p3 = 0.f;
  float norm = sqrtf(p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3);
  p0 /= norm; p1 /= norm; p2 /= norm; p3 /= norm;
  p3 = -(p0 * cx + p1 * cy + p2 * cz);
#pragma unroll
  for (int j = 0; j < 5; j++) {
    float distance = (p0 * nb[j].x + p1 * nb[j].y + p2 * nb[j].z) + p3;
    if (fabs((double)distance) > max_dist) return 0;
  }
  float distance = ((p0 * sx + p1 * sy) + p2 * sz) + p3;
  float xn = norm3f(sx, sy, sz);
  float weight = (float)(1 - 0.9 * fabs((double)distance) / sqrt((double)xn));
  co[0] = p0 * weight; co[1] = p1 * weight; co[2] = p2 * weight; co[3] = distance * weight;
  return ((double)weight > 0.1) ? 3 : 2;
}

// Query index space of one stream: corner queries first, padded to a multiple of 32 so that a warp never mixes the
// corner grid with the surf grid (the search finishes hard queries warp-cooperatively on ONE grid).
__device__ __forceinline__ bool decode_query(int t, int nC, int nS, bool* isCorner, int* src, int* row) {
  const int nCpad = (nC + 31) & ~31;
  if (t < nC) { *isCorner = true; *src = t; *row = t; return true; }
  if (t >= nCpad && t - nCpad < nS) { *isCorner = false; *src = t - nCpad; *row = nC + (t - nCpad); return true; }
  *isCorner = t < nCpad; *src = 0; *row = 0;
  return false;
}

// K5a: transform + exact 5-NN.  Writes the pool slots of the 5 neighbours (or -1 when the 5.0 gate rejects the
// query, ScanMatch.cpp:102,120).  Level 0 (the 2x2x2 cells around the query) runs one query per thread.  Queries whose
// list is not provably final after level 0 ("hard": sparse surroundings, typically corner queries far from any map
// edge) need the next shell of cells, which a whole warp scans cooperatively.  With a hard queue (a.hard != NULL) they
// are appended to a device-wide list and finished by search_hard_kernel, one warp per query, spread over the whole
// GPU -- otherwise a warp made of 32 hard queries would serialise 32 shell scans while the rest of the SMs idle (ncu:
// 15 % warps active, issue slots 6 % busy over the kernel's duration before this split).
struct HardItem { int s, t; float d[5]; int idx[5]; int slot[5]; int pad; };   // 72 bytes

template <bool kOrigIdx>
__device__ __forceinline__ void search_store(const CorrArgs& a, int s, int row, bool gate, const Top5& best) {
  const int capQ = a.cap_corner + a.cap_surf;
  int* out = a.nn_slot + ((size_t)s * capQ + row) * 5;
#pragma unroll
  for (int k = 0; k < 5; k++) out[k] = gate ? (kOrigIdx ? best.slot[k] : best.idx(k)) : -1;   // map grids: idx == slot
  if (a.nn) {
    int* nn = a.nn + ((size_t)s * capQ + row) * 5;
#pragma unroll
    for (int k = 0; k < 5; k++) nn[k] = gate ? best.idx(k) : -1;
  }
}

// the query of thread slot t of stream s in the map frame; false: idle slot / outside this rank's box
__device__ __forceinline__ bool search_query(const CorrArgs& a, int s, int t, const float* sR, const float* sT, bool* in_range,
                                             bool* isCorner, int* row, float* sx, float* sy, float* sz) {
  const int nC = a.n_corner[s], nS = a.n_surf[s];
  int src;
  *in_range = decode_query(t, nC, nS, isCorner, &src, row);
  bool valid = *in_range;
  *sx = *sy = *sz = 0.f;
  if (valid) {
    const float4 p = *isCorner ? a.corner[(size_t)s * a.cap_corner + src] : a.surf[(size_t)s * a.cap_surf + src];
    transform_point(sR, sT, p.x, p.y, p.z, sx, sy, sz);   // pointAssociateToMap
    if (a.own_box) {
      const float* b = a.own_box;
      valid = *sx >= b[0] && *sx < b[3] && *sy >= b[1] && *sy < b[4] && *sz >= b[2] && *sz < b[5];
    }
    if (a.dist_nranks > 1) {   // worldToCube (FeatureMap.h:475-487) of the query, clamped into the grid
      const CubeWindow& w = *a.grid_surf[s].window;
      int ci = (int)(roundf(*sx / w.cube_size) + (float)w.origin[0]), cj = (int)(roundf(*sy / w.cube_size) + (float)w.origin[1]),
          ck = (int)(roundf(*sz / w.cube_size) + (float)w.origin[2]);
      ci = min(max(ci, 0), w.dims[0] - 1); cj = min(max(cj, 0), w.dims[1] - 1); ck = min(max(ck, 0), w.dims[2] - 1);
      valid = cube_owner(ci, cj, ck, a.dist_nranks) == a.dist_rank;
    }
  }
  return valid;
}

#ifndef CM_SEARCH_THREADS
#define CM_SEARCH_THREADS 128
#endif
#ifndef CM_SEARCH_MINB
#define CM_SEARCH_MINB 6
#endif
// out of line on purpose: search_kernel only gets here without a hard queue (never in the mapping stage), and the shell pass
// inlined into it costs the per-thread pass 120 bytes of spills
template <bool kOrigIdx>
__device__ __noinline__ void knn5_warp_finish_cold(const GridView* g, int h, const KnnGeom* c, float qx, float qy, float qz, float gate, Top5* best) {
  Top5 b = *best;
  knn5_warp_finish<kOrigIdx>(*g, h, *c, qx, qy, qz, gate, b);
  *best = b;
}

template <bool kOrigIdx>
__global__ void __launch_bounds__(CM_SEARCH_THREADS, CM_SEARCH_MINB) search_kernel(CorrArgs a) {
  const int s = blockIdx.y;
  const MatchState& st = a.state[s];
  if (st.done || (a.skip && *a.skip)) return;
  __shared__ float sR[9], sT[3];
  __shared__ uint4 rng[8 * CM_SEARCH_THREADS];
  if (threadIdx.x < 9) sR[threadIdx.x] = st.R[threadIdx.x];
  if (threadIdx.x < 3) sT[threadIdx.x] = st.pose[3 + threadIdx.x];
  __syncthreads();
  const int nC = a.n_corner[s], nS = a.n_surf[s];
  const int nT = ((nC + 31) & ~31) + nS;
  const int it = a.iter_dev ? *a.iter_dev : a.iter;
  const unsigned int FULL = 0xffffffffu;
  // (the grid may be sized from an ESTIMATE of the filtered feature counts: the caller checks on the device that no stream
  // exceeded it and repeats the match otherwise -- a grid-stride loop here cost 25 % of the kernel in spills)
  // a.spread > 0 (few queries in the whole launch): a warp takes only 32 >> spread consecutive query slots, in its first lanes --
  // the per-thread pass is divergent (every lane walks its own cells), so a warp's instruction count grows with its active lanes;
  // with most SMs idle anyway, thinner warps on more SMs finish sooner
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = 32 >> a.spread;
  const int t = (gt >> 5) * per + (gt & 31);
  if ((gt >> 5) * per >= nT) return;   // whole warp idle
  const bool lane_on = (gt & 31) < per;
  unsigned long long t0 = 0;
  if (a.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  bool in_range, isCorner; int row;
  float sx, sy, sz;
  bool valid = search_query(a, s, lane_on ? t : nT + 32, sR, sT, &in_range, &isCorner, &row, &sx, &sy, &sz);
  if (!lane_on) isCorner = (gt >> 5) * per < ((nC + 31) & ~31);   // (only picks the grid the idle lane's dummy geometry reads)
  const GridView& g = isCorner ? a.grid_corner[s] : a.grid_surf[s];
  Top5 best;
  top5_init(best);
  KnnGeom c;
  valid = knn5_geom(g, sx, sy, sz, a.prm.knn_gate, c, a.prm.own_cube_only != 0) && valid;
  bool need = false;
  unsigned int ncand = 0;
  // iterations >= 1 start from the neighbours found one iteration ago (written by search_store for every query of the stream)
  const int* prev = (valid && it > 0 && a.warm) ? a.nn_slot + ((size_t)s * (a.cap_corner + a.cap_surf) + row) * 5 : nullptr;
  bool tie = false;
  if (valid) need = knn5_level0<kOrigIdx>(g, c, sx, sy, sz, rng, best, a.dbg ? &ncand : nullptr, prev, &tie);
  // an exact distance tie (about one query in 10^5): the query goes to the warp-cooperative pass, which starts it over with the
  // canonical tie order -- needs the hard queue (always there for map grids)
  tie = tie && a.hard != nullptr;
  need = need || tie;
  unsigned int hard = __ballot_sync(FULL, need);
  if (a.dbg) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    const unsigned int mx = __reduce_max_sync(FULL, ncand), sm = __reduce_add_sync(FULL, ncand);
    if ((threadIdx.x & 31) == 0) {
      unsigned long long* d = a.dbg + 4 * ((size_t)(s * gridDim.x + blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5));
      unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      d[0] = t0; d[1] = t1; d[2] = ((unsigned long long)mx << 32) | sm;
      d[3] = ((unsigned long long)__popc(hard) << 32) | ((unsigned long long)smid << 16) | (isCorner ? 1u : 0u);
    }
  }
  if (a.hard) {
    // deferred: one warp-aggregated reservation, every hard lane writes its own item
    if (hard) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(a.iter_dev ? a.hard_count + it : a.hard_count, __popc(hard));
      base = __shfl_sync(FULL, base, 0);
      if (need) {
        const int pos = base + __popc(hard & ((1u << lane) - 1));
        if (pos < a.hard_cap) {
          HardItem item;
          item.s = s; item.t = t; item.pad = tie ? 1 : 0;
#pragma unroll
          for (int k = 0; k < 5; k++) { item.d[k] = best.d(k); item.idx[k] = best.idx(k); item.slot[k] = kOrigIdx ? best.slot[k] : best.idx(k); }
          reinterpret_cast<HardItem*>(a.hard)[pos] = item;
        }
      }
    }
    if (!in_range || need) return;
  } else {
    while (hard) {
      const int h = __ffs(hard) - 1;
      hard &= hard - 1;
      knn5_warp_finish_cold<kOrigIdx>(&g, h, &c, sx, sy, sz, a.prm.knn_gate, &best);
    }
    if (!in_range) return;
  }
  search_store<kOrigIdx>(a, s, row, valid && best.d(4) < a.prm.knn_gate, best);
}

// K5a': the hard queries of one Gauss-Newton evaluation, one warp per query (grid-stride over the list).
template <bool kOrigIdx>
__global__ void __launch_bounds__(256) search_hard_kernel(CorrArgs a) {
  if (a.skip && *a.skip) return;
  const int n = min(a.iter_dev ? a.hard_count[*a.iter_dev] : *a.hard_count, a.hard_cap);
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  // (taking the queries off a shared cursor instead of this static split was measured: same kernel time, 9 % of the samples on the atomic)
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
    const HardItem* item = reinterpret_cast<const HardItem*>(a.hard) + i;
    const int s = item->s, t = item->t;
    const MatchState& st = a.state[s];
    float R[9], T[3];
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = st.R[k];
#pragma unroll
    for (int k = 0; k < 3; k++) T[k] = st.pose[3 + k];
    bool in_range, isCorner; int row;
    float sx, sy, sz;
    search_query(a, s, t, R, T, &in_range, &isCorner, &row, &sx, &sy, &sz);
    const GridView& g = isCorner ? a.grid_corner[s] : a.grid_surf[s];
    KnnGeom c;
    knn5_geom(g, sx, sy, sz, a.prm.knn_gate, c, a.prm.own_cube_only != 0);
    Top5 best;
    const int redo = item->pad;   // 1: exact distance tie in the per-thread pass -> start over from the level-0 block, canonical ties
#pragma unroll
    for (int k = 0; k < 5; k++) { best.key[k] = redo ? CM_TOP5_EMPTY : top5_key(item->d[k], item->idx[k]); best.slot[k] = redo ? -1 : item->slot[k]; }
    knn5_warp_finish<kOrigIdx>(g, 0, c, sx, sy, sz, a.prm.knn_gate, best, redo ? 0 : 1);
    if (lane == 0) search_store<kOrigIdx>(a, s, row, best.d(4) < a.prm.knn_gate, best);
  }
}

// K5b: line / plane fit on the 5 neighbours, residual and Jacobian row (ScanMatch.cpp:103-112, 121-130, 154-204).
__device__ __forceinline__ void fit_row(const CorrArgs& a, int s, const PoseCoef& kc, const float* sR, const float* sT, bool isCorner, int src,
                                        int row, float4* o0, float4* o1) {
  const int capQ = a.cap_corner + a.cap_surf;
  const float4 p = isCorner ? a.corner[(size_t)s * a.cap_corner + src] : a.surf[(size_t)s * a.cap_surf + src];
  const int* slots = a.nn_slot + ((size_t)s * capQ + row) * 5;
  RowOut rowv;
#pragma unroll
  for (int k = 0; k < 6; k++) rowv.a[k] = 0.f;
  rowv.b = 0.f; rowv.flag = 0;
  int sl[5];
#pragma unroll
  for (int k = 0; k < 5; k++) sl[k] = slots[k];
  if (sl[0] >= 0) {
    const GridView& g = isCorner ? a.grid_corner[s] : a.grid_surf[s];
    float sx, sy, sz;
    transform_point(sR, sT, p.x, p.y, p.z, &sx, &sy, &sz);
    float4 nb[5];
#pragma unroll
    for (int k = 0; k < 5; k++) nb[k] = __ldg(g.pts + sl[k]);
    float co[4];
    int f = isCorner ? corner_row(nb, sx, sy, sz, co) : surf_row(nb, sx, sy, sz, a.prm.plane_max_dist, co);
    rowv.flag = f;
    if (f & 1) {
      const float x = p.x, y = p.y, z = p.z;
      float arx = (kc.x1 * y + kc.x2 * z) * co[0] + (kc.x3 * y - kc.x4 * z) * co[1] + (kc.x5 * y - kc.x6 * z) * co[2];
      float ary = (kc.y1 * x + kc.y2 * y + kc.y3 * z) * co[0] + (kc.y4 * x + kc.y5 * y + kc.y6 * z) * co[1] +
                  (kc.y7 * x - kc.y8 * y - kc.y9 * z) * co[2];
      float arz = (kc.z1 * x - kc.x4 * y + kc.z3 * z) * co[0] + (kc.z4 * x + kc.z5 * y + kc.z6 + kc.z7 * z) * co[1] +
                  0 * co[2];
      rowv.a[0] = arx; rowv.a[1] = ary; rowv.a[2] = arz; rowv.a[3] = co[0]; rowv.a[4] = co[1]; rowv.a[5] = co[2];
      rowv.b = -co[3];
    }
  }
  *o0 = make_float4(rowv.a[0], rowv.a[1], rowv.a[2], rowv.a[3]);
  *o1 = make_float4(rowv.a[4], rowv.a[5], rowv.b, __int_as_float(rowv.flag));
}

__global__ void __launch_bounds__(256) fit_kernel(CorrArgs a) {
  const int s = blockIdx.y;
  const MatchState& st = a.state[s];
  if (st.done) return;
  __shared__ PoseCoef kc;
  __shared__ float sR[9], sT[3];
  if (threadIdx.x == 0) make_pose_coef(st, kc);
  if (threadIdx.x < 9) sR[threadIdx.x] = st.R[threadIdx.x];
  if (threadIdx.x < 3) sT[threadIdx.x] = st.pose[3 + threadIdx.x];
  __syncthreads();
  const int nC = a.n_corner[s], nS = a.n_surf[s];
  const int capQ = a.cap_corner + a.cap_surf;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  bool isCorner; int src, row;
  if (!decode_query(t, nC, nS, &isCorner, &src, &row)) return;
  float4 r0, r1;
  fit_row(a, s, kc, sR, sT, isCorner, src, row, &r0, &r1);
  float4* dst = reinterpret_cast<float4*>(a.rows + (size_t)s * capQ + row);
  dst[0] = r0; dst[1] = r1;
}

// ============================================================================================================
// K6: reduce A^T A / A^T b over the rows of one stream (K6a), then solve + degeneracy projection + pose update +
// convergence test (K6b)  (ScanMatch.cpp:134-260).
// ============================================================================================================
struct SolveArgs {
  const RowOut* rows;
  const int* n_corner; const int* n_surf;
  int cap_corner, cap_surf;
  MatchState* state;
  IterTrace* trace;   // optional [nstreams][max_iterations]
  int iter;
  const int* iter_dev;   // optional: overrides iter (graph WHILE loop)
  const int* skip;       // optional: non-zero = do nothing (MatchLaunch::skip)
  MatchParamsDev prm;
};

#define CM_NACC 31   // 21 (AtA upper) + 6 (AtB) + rows + line / plane counted + score

// K6a: deterministic reduction of one stream's rows.  One CTA per stream.  Float products are exact in double and
// are accumulated in double, so any summation order rounds to the same float A^T A as the oracle's.
__device__ __noinline__ void solve_stream(const SolveArgs& a_in, int s, const double* tot);
// kSolve: the 6x6 step of the stream right behind its reduction, by warp 0 of the same CTA (the odometry loop: one launch less per
// evaluation; the sharded-map path keeps them apart, its exchange sits in between)
template <bool kSolve>
__global__ void __launch_bounds__(512) reduce_rows_kernel(SolveArgs a, double* __restrict__ sums) {
  const int s = blockIdx.x;
  if (a.state[s].done) return;
  const int nC = a.n_corner[s], nS = a.n_surf[s];
  const RowOut* rows = a.rows + (size_t)s * (a.cap_corner + a.cap_surf);
  double acc[CM_NACC];
#pragma unroll
  for (int k = 0; k < CM_NACC; k++) acc[k] = 0.0;
  for (int q = threadIdx.x; q < nC + nS; q += blockDim.x) {
    const float4* src = reinterpret_cast<const float4*>(rows + q);
    float4 r0 = src[0], r1 = src[1];
    int flag = __float_as_int(r1.w);
    if (flag & 2) { if (q < nC) acc[28] += 1.0; else acc[29] += 1.0; }
    if (flag & 1) {
      double v[7] = {(double)r0.x, (double)r0.y, (double)r0.z, (double)r0.w, (double)r1.x, (double)r1.y, (double)r1.z};
      int t = 0;
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = r; c < 6; c++) acc[t++] += v[r] * v[c];
#pragma unroll
      for (int r = 0; r < 6; r++) acc[21 + r] += v[r] * v[6];
      acc[27] += 1.0;
      acc[30] += exp(-fabs(v[6]));
    }
  }
  __shared__ double sm[16][CM_NACC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < CM_NACC; k++) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < CM_NACC) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += sm[w][threadIdx.x];
    sums[(size_t)s * 32 + threadIdx.x] = v;
  }
  if (kSolve) {
    __syncthreads();   // the sums of this stream are in global memory, visible to the whole CTA
    if (threadIdx.x < 32) solve_stream(a, s, sums + (size_t)s * 32);
  }
}

// The 6x6 routines are kept out of line on purpose: each gets its own small stack frame.  Inlined into one large
// kernel body, nvcc 12.9 (sm_100a) produced a wrong 6x6 solve (see the note in cm_math.h::colpiv_qr_solve).
__device__ __noinline__ void dev_eig6_values(const float* A, float* w) {
  float Aw[36], ww[6];
  for (int k = 0; k < 36; k++) Aw[k] = A[k];
  eig_sym<6>(Aw, ww, (float*)nullptr);
  for (int k = 0; k < 6; k++) w[k] = ww[k];
}
__device__ __noinline__ void dev_eig6_full(const float* A, float* w, float* V) {
  float Aw[36], ww[6], Vw[36];
  for (int k = 0; k < 36; k++) Aw[k] = A[k];
  eig_sym<6>(Aw, ww, Vw);
  for (int k = 0; k < 6; k++) w[k] = ww[k];
  for (int k = 0; k < 36; k++) V[k] = Vw[k];
}
__device__ __noinline__ bool dev_inverse6(const float* A, float* inv) {
  float Aw[36], iw[36];
  for (int k = 0; k < 36; k++) Aw[k] = A[k];
  bool ok = inverse_lu<6>(Aw, iw);
  for (int k = 0; k < 36; k++) inv[k] = iw[k];
  return ok;
}

// K6b: solve, degeneracy projection, pose update, convergence (ScanMatch.cpp:134-260).  One WARP per stream: the 6x6 solve is
// column-parallel (warp_qr_solve6), the rest is lane 0's; the 6x6 eigen work goes through cm_math.h, i.e. the same instruction
// sequence as the oracle's.
__device__ __noinline__ void solve_stream(const SolveArgs& a_in, int s, const double* tot) {
  const unsigned int FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  SolveArgs a = a_in;
  if (a.iter_dev) a.iter = *a.iter_dev;
  MatchState& st = a.state[s];
  IterTrace* tr = a.trace ? a.trace + (size_t)s * a.prm.max_iterations + a.iter : nullptr;
  int go = 0;
  if (lane == 0 && !st.done) {
    const int nrows = (int)tot[27];
    const int nline = (int)tot[28], nplane = (int)tot[29];
    st.rows = nrows; st.line = nline; st.plane = nplane; st.score = tot[30];
    if (tr) {
      for (int k = 0; k < 6; k++) tr->pose_in[k] = st.pose[k];
      tr->rows = nrows; tr->line = nline; tr->plane = nplane; tr->degenerate = (st.flags & CM_F_DEGENERATE) ? 1 : 0;
      for (int k = 0; k < 36; k++) tr->AtA[k] = 0.f;
      for (int k = 0; k < 6; k++) { tr->AtB[k] = 0.f; tr->x[k] = 0.f; }
    }
    if (nrows >= a.prm.min_rows) go = 1;
    else if (!a.prm.few_rows_continue) { st.flags |= CM_F_TOO_FEW_MATCHES; st.done = 1; }   // ScanMatch.cpp:141-145: leave the loop
    // (few_rows_continue, LaserOdometry.cpp:501-503: skip this iteration)
  }
  go = __shfl_sync(FULL, go, 0);
  if (!go) return;
  float X[6];
  {
    // lane c < 6: column c of A^T A (symmetric: element (r, c) is entry (min, max) of the upper triangle in tot[0..21)); lane 6: A^T b
    float col[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const int lo = r < lane ? r : lane, hi = r < lane ? lane : r;
      const int t = lane < 6 ? lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo) : 21 + r;
      col[r] = lane < 7 ? (float)tot[t] : 0.f;
    }
    warp_qr_solve6(col, lane, X);   // ScanMatch.cpp:209
  }
  if (lane != 0) return;
  float AtA[36], AtB[6];
  if (a.iter == 0 || tr) {
    int t = 0;
    for (int r = 0; r < 6; r++)
      for (int c = r; c < 6; c++) { float v = (float)tot[t++]; AtA[r * 6 + c] = v; AtA[c * 6 + r] = v; }
    for (int r = 0; r < 6; r++) AtB[r] = (float)tot[21 + r];
  }
  if (a.iter == 0 && !dev_min_eig_above6(AtA, a.prm.eig_threshold)) {   // ScanMatch.cpp:211-235 (skipped when provably not degenerate)
    float E[6];
    dev_eig6_values(AtA, E);
    if (E[0] < a.prm.eig_threshold) {
      float V[36], V2[36], Vinv[36];
      dev_eig6_full(AtA, E, V);
      for (int k = 0; k < 36; k++) V2[k] = V[k];
      for (int i = 0; i < 6; i++) {
        if (E[i] < a.prm.eig_threshold) { for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0.f; }
        else break;
      }
      if (!dev_inverse6(V, Vinv)) { for (int k = 0; k < 36; k++) Vinv[k] = nanf(""); }
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
          float sum = 0.f;
          for (int k = 0; k < 6; k++) sum += Vinv[r * 6 + k] * V2[k * 6 + c];
          st.P[r * 6 + c] = sum;
        }
      st.flags |= CM_F_DEGENERATE;
    }
  }
  if (st.flags & CM_F_DEGENERATE) {   // ScanMatch.cpp:237-240
    float x2[6];
    for (int k = 0; k < 6; k++) x2[k] = X[k];
    for (int r = 0; r < 6; r++) {
      float sum = 0.f;
      for (int k = 0; k < 6; k++) sum += st.P[r * 6 + k] * x2[k];
      X[r] = sum;
    }
  }
  float np[6];
  for (int k = 0; k < 6; k++) np[k] = st.pose[k] + X[k];   // ScanMatch.cpp:242-247
  if (a.prm.nan_guard) for (int k = 0; k < 6; k++) if (!isfinite(np[k])) np[k] = 0.f;   // LaserOdometry.cpp:622-634
  state_set_pose(st, np);
  st.iterations = a.iter + 1;
  if (tr) {
    for (int k = 0; k < 36; k++) tr->AtA[k] = AtA[k];
    for (int k = 0; k < 6; k++) { tr->AtB[k] = AtB[k]; tr->x[k] = X[k]; }
    tr->degenerate = (st.flags & CM_F_DEGENERATE) ? 1 : 0;
  }
  // ScanMatch.cpp:249-260: rad2deg(float) = (float)(r*180.0/M_PI); pow(., 2) in double; sqrt; narrowed to float
  const double PI = 3.14159265358979323846;
  double d0 = (double)(float)((double)X[0] * 180.0 / PI), d1 = (double)(float)((double)X[1] * 180.0 / PI),
         d2 = (double)(float)((double)X[2] * 180.0 / PI);
  float deltaR = (float)sqrt(d0 * d0 + d1 * d1 + d2 * d2);
  double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
  float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
  if (deltaR < a.prm.delta_r_abort && deltaT < a.prm.delta_t_abort) { st.flags |= CM_F_CONVERGED; st.done = 1; }
}

// K5b + K6 in one launch: fit_kernel's rows are reduced inside the CTA (double, fixed order), the per-CTA partial sums
// go to global memory, and the LAST CTA of a stream to finish (ticket counter) adds the partials in CTA order and runs the
// 6x6 step -- streams solve concurrently on different SMs instead of one after the other in a single warp, and two
// launches per Gauss-Newton iteration disappear.  Determinism: every sum has a fixed order (row -> 32-row group -> CTA).
struct FusedArgs { double* partials; int* tickets; double* sums; int ptiles; /* partial slots per stream */ };

#ifndef CM_FIT_MINB
#define CM_FIT_MINB 4
#endif
__global__ void __launch_bounds__(256, CM_FIT_MINB) fit_solve_kernel(CorrArgs a, FusedArgs f) {
  const int s = blockIdx.y;
  const MatchState& st = a.state[s];
  if (st.done || (a.skip && *a.skip)) return;
  const int nC = a.n_corner[s], nS = a.n_surf[s];
  const int capQ = a.cap_corner + a.cap_surf;
  // tiles of 256 query slots.  The grid comes from an estimate of the filtered counts; a CTA takes tiles blockIdx.x,
  // blockIdx.x + gridDim.x, ..  Partial sums are kept PER TILE and added in tile order by the last CTA to finish, so the
  // result does not depend on the grid size.
  const int nT = ((nC + 31) & ~31) + nS;
  int ntiles = (nT + 255) >> 8;
  if (ntiles < 1) ntiles = 1;             // a stream without queries still has to reach its solve ("matched cloud points too few")
  if (ntiles > f.ptiles) ntiles = f.ptiles;   // cannot happen: ptiles is sized from the INPUT counts, an upper bound
  if ((int)blockIdx.x >= ntiles) return;
  __shared__ PoseCoef kc;
  __shared__ float sR[9], sT[3];
  __shared__ float4 srow[2 * 256];
  __shared__ double sacc[8][32];
  __shared__ double sexp[256];
  __shared__ int s_last;
  if (threadIdx.x == 0) make_pose_coef(st, kc);
  if (threadIdx.x < 9) sR[threadIdx.x] = st.R[threadIdx.x];
  if (threadIdx.x < 3) sT[threadIdx.x] = st.pose[3 + threadIdx.x];
  __syncthreads();
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int t = tile * 256 + threadIdx.x;
    bool isCorner; int src, row;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    if (decode_query(t, nC, nS, &isCorner, &src, &row)) {
      fit_row(a, s, kc, sR, sT, isCorner, src, row, &r0, &r1);
      float4* dst = reinterpret_cast<float4*>(a.rows + (size_t)s * capQ + row);
      dst[0] = r0; dst[1] = r1;
      if (isCorner) r1.w = __int_as_float(__float_as_int(r1.w) | 4);
    }
    // score term of the row (ScanMatch.cpp:277-286), evaluated by the row's own thread
    sexp[threadIdx.x] = (__float_as_int(r1.w) & 1) ? exp(-fabs((double)r1.z)) : 0.0;
    srow[2 * threadIdx.x] = r0; srow[2 * threadIdx.x + 1] = r1;
    __syncthreads();
    // accumulator k of row group grp: 0..20 upper triangle of A^T A, 21..26 A^T b, 27 rows, 28 / 29 counted corner / surf, 30 score
    const int k = threadIdx.x & 31, grp = threadIdx.x >> 5;
    {
      int ra = 0, rb = 0;
      if (k < 21) { int tt = k; while (tt >= 6 - ra) { tt -= 6 - ra; ra++; } rb = ra + tt; }
      else if (k < 27) { ra = k - 21; rb = 6; }
      // branch-free inner loop: every lane adds term(i) when (flag & need) == want, with per-lane constants
      //   k < 27: term = row[ra] * row[rb] (kept rows);  27: 1 (kept rows);  28 / 29: 1 (counted corner / surf rows);  30: score term
      const bool prod = k < 27;
      const int ia = prod ? ra : 7, ib = prod ? rb : 7;
      const int need = (k == 28 || k == 29) ? 6 : 1, want = (k == 28) ? 6 : (k == 29 ? 2 : 1);
      double acc = 0.0;
      const float* rows = reinterpret_cast<const float*>(srow);
      for (int i = grp * 32; i < grp * 32 + 32; i++) {
        const float* rv = rows + 8 * i;
        const int flag = __float_as_int(rv[7]);
        const float fa = rv[ia], fb = rv[ib];
        double term = prod ? (double)fa * (double)fb : 1.0;
        if (k == 30) term = sexp[i];
        if (k < 31 && (flag & need) == want) acc += term;
      }
      sacc[grp][k] = acc;
    }
    __syncthreads();
    double* mine = f.partials + ((size_t)s * f.ptiles + tile) * 32;
    if (threadIdx.x < 32) {
      double v = 0.0;
#pragma unroll
      for (int g = 0; g < 8; g++) v += sacc[g][threadIdx.x];
      __stcg(mine + threadIdx.x, v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(f.tickets + s, 1) == ntiles - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (threadIdx.x < 32) {
        const double* p = f.partials + (size_t)s * f.ptiles * 32 + threadIdx.x;
        double v = 0.0;
        for (int b = 0; b < ntiles; b++) v += __ldcg(p + (size_t)b * 32);
        f.sums[(size_t)s * 32 + threadIdx.x] = v;
      }
      __syncthreads();
      // (the 6x6 step runs in solve_warp_kernel right behind this launch: inside this kernel it was capped at 64 registers and
      // ran out of local memory, ~45 us per iteration)
      if (threadIdx.x == 0) f.tickets[s] = 0;
    }
  }
}

// K6b: one warp per stream
__global__ void __launch_bounds__(32) solve_warp_kernel(SolveArgs a, const double* __restrict__ sums) {
  if (a.skip && *a.skip) return;
  solve_stream(a, blockIdx.x, sums + (size_t)blockIdx.x * 32);
}

#include "cm_odom.inl"

// ============================================================================================================
// Stand-alone exact 5-NN (test hook and operator): queries already in the map frame.
// ============================================================================================================
__global__ void __launch_bounds__(128) knn5_kernel(GridView g, const float* __restrict__ q, int nq, float gate, int* __restrict__ idx,
                                                   float* __restrict__ d2) {
  __shared__ uint4 rng[8 * 128];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < nq;
  Top5 best;
  knn5_search<true>(g, valid, valid ? q[3 * i] : 0.f, valid ? q[3 * i + 1] : 0.f, valid ? q[3 * i + 2] : 0.f, gate, rng, best);
  if (!valid) return;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    idx[5 * i + k] = best.slot[k] < 0 ? -1 : best.idx(k);
    d2[5 * i + k] = best.d(k);
  }
}

// ============================================================================================================
// Host-side launchers
// ============================================================================================================
static inline unsigned int next_pow2(unsigned int v) { unsigned int p = 64; while (p < v) p <<= 1; return p; }

// last shell so that 0.98 * cell * (0.5 + L) >= sqrt(gate) (worst case: query in the middle of the level-0 block)
int grid_max_level(float cell, float gate) {
  int L = (int)ceilf(sqrtf(gate) / (0.98f * cell) - 0.5f);
  return L < 0 ? 0 : L;
}

void GridStorage::build(const float4* d_pts, int n, float cell_size, float gate, int keep_w, cudaStream_t stream) {
  unsigned int cap = next_pow2((unsigned int)(2 * (n > 0 ? n : 1)));
  entries.reserve((size_t)cap * sizeof(CellEntry));
  pts.reserve((size_t)(n > 0 ? n : 1) * sizeof(float4));
  cell_of.reserve((size_t)(n > 0 ? n : 1) * sizeof(int));
  cursor.reserve(sizeof(unsigned int));
  float inv = 1.0f / cell_size;
  CM_LAUNCH(grid_clear_kernel, (cap + 255) / 256, 256, 0, stream, (CellEntry*)entries.p, cap);
  cudaMemsetAsync(cursor.p, 0, sizeof(unsigned int), stream);
  if (n > 0) {
    int nb = (n + 255) / 256;
    CM_LAUNCH(grid_count_kernel, nb, 256, 0, stream, d_pts, n, (CellEntry*)entries.p, cap - 1, inv, (int*)cell_of.p);
    CM_LAUNCH(grid_offsets_kernel, (cap + 255) / 256, 256, 0, stream, (CellEntry*)entries.p, cap, (unsigned int*)cursor.p);
    CM_LAUNCH(grid_scatter_kernel, nb, 256, 0, stream, d_pts, n, (CellEntry*)entries.p, (const int*)cell_of.p, (float4*)pts.p, keep_w);
  }
  view.entries = (const CellEntry*)entries.p;
  view.pts = (const float4*)pts.p;
  view.mask = cap - 1;
  view.inv_leaf = inv; view.kdiv = 1; view.cell = cell_size;
  view.npts = n;
  view.window = nullptr; view.cube_count = nullptr;
  view.epoch = nullptr; view.eoff = nullptr; view.displaced = 0;
  view.max_level = grid_max_level(cell_size, gate);
}

void launch_knn5(const GridView& g, const float* d_q, int nq, float gate, int* d_idx, float* d_d2, cudaStream_t stream) {
  if (nq > 0) CM_LAUNCH(knn5_kernel, (nq + 127) / 128, 128, 0, stream, g, d_q, nq, gate, d_idx, d_d2);
}

static void fill_args(const MatchLaunch& m, CorrArgs& ca, SolveArgs& sa) {
  ca.corner = m.corner; ca.surf = m.surf; ca.n_corner = m.n_corner; ca.n_surf = m.n_surf;
  ca.cap_corner = m.cap_corner; ca.cap_surf = m.cap_surf; ca.grid_corner = m.grid_corner; ca.grid_surf = m.grid_surf;
  ca.state = m.state; ca.rows = m.rows; ca.nn_slot = m.nn_slot; ca.nn = nullptr; ca.own_box = m.own_box; ca.prm = m.prm;
  ca.hard = m.hard; ca.hard_count = m.hard_count; ca.hard_cap = m.hard_cap; ca.dbg = nullptr; ca.iter_dev = nullptr; sa.iter_dev = nullptr;
  static const int warm = getenv("COOPERMAP_NO_WARM") ? 0 : 1;
  ca.iter = 0; ca.warm = warm; ca.spread = 0; ca.skip = m.skip; sa.skip = m.skip;
  ca.dist_rank = m.dist_rank; ca.dist_nranks = m.dist_nranks;
  sa.rows = m.rows; sa.n_corner = m.n_corner; sa.n_surf = m.n_surf; sa.cap_corner = m.cap_corner; sa.cap_surf = m.cap_surf;
  sa.state = m.state; sa.trace = m.trace; sa.prm = m.prm; sa.iter = 0;
}

void launch_match_init(const MatchLaunch& m, cudaStream_t stream) {
  if (m.hard) cudaMemsetAsync(m.hard_count, 0, sizeof(int) * CM_MAX_EVALS * 8, stream);
  if (m.tickets) cudaMemsetAsync(m.tickets, 0, sizeof(int) * m.nstreams, stream);
  CM_LAUNCH(match_init_kernel, (m.nstreams + 63) / 64, 64, 0, stream, m.state, m.pose_in, m.grid_corner, m.grid_surf, m.prm, m.nstreams);
}

// one Gauss-Newton evaluation: correspondences + rows + (partial) normal-equation sums into m.sums
// CTAs of search_hard_kernel (grid-strided, one warp per hard query): latency-bound shell scans, so as many warps as the register
// file holds -- 3 CTAs of 256 threads per SM at 80 registers
static int hard_blocks_default() {
  static const int n = []() { const char* e = getenv("COOPERMAP_HARD_BLOCKS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 444; }();
  return n;
}

// search_kernel's spread for a launch of `threads` query slots in all: thin the warps out while the launch leaves SM sub-partitions
// without a warp of their own (B200: 148 SMs x 4)
static int search_spread(long long threads) {
  static const int forced = getenv("COOPERMAP_SEARCH_SPREAD") ? atoi(getenv("COOPERMAP_SEARCH_SPREAD")) : -1;
  if (forced >= 0) return forced > 3 ? 3 : forced;
  const long long warps = threads / 32;
  if (warps * 4 <= 148 * 4 * 2) return 2;
  if (warps * 2 <= 148 * 4 * 2) return 1;
  return 0;
}
static dim3 search_grid(int bx, int nstreams, int spread) {
  return dim3((unsigned int)((((long long)bx * 256) << spread) + CM_SEARCH_THREADS - 1) / CM_SEARCH_THREADS, nstreams);
}

void launch_match_partial(const MatchLaunch& m, int it, cudaStream_t stream, KernelProfiler* prof, bool fused, bool defer_solve) {
  CorrArgs ca; SolveArgs sa;
  fill_args(m, ca, sa);
  const int capQ = m.cap_corner + m.cap_surf;
  int bx = ((m.max_queries > 0 ? m.max_queries : capQ) + 32 + 255) / 256;   // + 32: the corner block is padded to a warp
  if (bx < 1) bx = 1;
  dim3 grid(bx, m.nstreams);
  ca.nn = m.nn ? m.nn + (size_t)it * m.nstreams * capQ * 5 : nullptr;
  ca.iter = it;
  if (prof) prof->begin(stream);
  if (m.dbg && it == m.dbg_iter) { ca.dbg = m.dbg; cudaMemsetAsync(m.dbg, 0, (size_t)bx * m.nstreams * 8 * 4 * sizeof(unsigned long long), stream);   /* bx * 8 warps per stream */ }
  if (it >= CM_MAX_EVALS) ca.hard = nullptr;       // no counter left: finish hard queries inside their own warp
  if (ca.hard) ca.hard_count = m.hard_count + it;  // one counter per evaluation, zeroed by launch_match_init
  ca.spread = m.dbg ? 0 : search_spread((long long)bx * 256 * m.nstreams);   // (the search trace is laid out for whole warps)
  const dim3 sgrid = search_grid(bx, m.nstreams, ca.spread);
  if (m.orig_idx) CM_LAUNCH(search_kernel<true>, sgrid, CM_SEARCH_THREADS, 0, stream, ca);
  else CM_LAUNCH(search_kernel<false>, sgrid, CM_SEARCH_THREADS, 0, stream, ca);
  if (ca.hard) {
    const int hb = m.hard_blocks > 0 ? m.hard_blocks : hard_blocks_default();
    if (m.orig_idx) CM_LAUNCH(search_hard_kernel<true>, hb, 256, 0, stream, ca);
    else CM_LAUNCH(search_hard_kernel<false>, hb, 256, 0, stream, ca);
  }
  if (prof) prof->end(stream);
  sa.iter = it;
  if (fused) {
    FusedArgs f; f.partials = m.partials; f.tickets = m.tickets; f.sums = m.sums; f.ptiles = m.partial_blocks;
    CM_LAUNCH(fit_solve_kernel, grid, 256, 0, stream, ca, f);
    if (!defer_solve) CM_LAUNCH(solve_warp_kernel, m.nstreams, 32, 0, stream, sa, (const double*)m.sums);
    return;
  }
  CM_LAUNCH(fit_kernel, grid, 256, 0, stream, ca);
  CM_LAUNCH(reduce_rows_kernel<false>, m.nstreams, 512, 0, stream, sa, m.sums);
}
// reduction + 6x6 step in one launch (it: the evaluation index, or read from d_iter)
void launch_match_reduce_solve(const MatchLaunch& m, int it, cudaStream_t stream, const int* d_iter) {
  CorrArgs ca; SolveArgs sa;
  fill_args(m, ca, sa);
  sa.iter = it; sa.iter_dev = d_iter;
  CM_LAUNCH(reduce_rows_kernel<true>, m.nstreams, 512, 0, stream, sa, m.sums);
}

// solve + pose update + convergence test from the (complete) sums
// reduce the rows already written to m.rows (used by the odometry, whose correspondence kernel is its own)
void launch_match_reduce(const MatchLaunch& m, int it, cudaStream_t stream) {
  CorrArgs ca; SolveArgs sa;
  fill_args(m, ca, sa);
  sa.iter = it;
  CM_LAUNCH(reduce_rows_kernel<false>, m.nstreams, 512, 0, stream, sa, m.sums);
}

void launch_match_solve_warp(const MatchLaunch& m, int it, cudaStream_t stream) {
  CorrArgs ca; SolveArgs sa;
  fill_args(m, ca, sa);
  sa.iter = it;
  CM_LAUNCH(solve_warp_kernel, m.nstreams, 32, 0, stream, sa, (const double*)m.sums);
}

void launch_match_solve(const MatchLaunch& m, int it, const double* sums, cudaStream_t stream) {
  CorrArgs ca; SolveArgs sa;
  fill_args(m, ca, sa);
  sa.iter = it;
  CM_LAUNCH(solve_warp_kernel, m.nstreams, 32, 0, stream, sa, sums);
}

// odometry correspondence kernels: a warp takes 32 >> spread query slots; thinned out, down to one query per warp, while the launch
// stays below a few warps per SM sub-partition (see launch_odom_corr_batch)
static int odom_spread(long long slots) {
  static const int forced = getenv("COOPERMAP_ODOM_SPREAD") ? atoi(getenv("COOPERMAP_ODOM_SPREAD")) : -1;
  if (forced >= 0) return forced > 5 ? 5 : forced;
  // (slots are padded launch sizes, about three per real query; measured on VLP-16 chains, correspondence kernels per step at 1 / 2 /
  // 4 / 8 / 16 / 32 queries per warp: 32 streams 1722 / 1124 / 863 / 820 / 978 us from 32 down to 2, 8 streams 638 / 532 at 8 / 4 and 594 at 1)
  const long long warps = slots / 32;
  if ((warps << 5) <= 4096) return 5;   // one sweep: one query per warp (235 us per sweep against 259 / 304 at 2 / 4 queries per warp)
  int spread = 0;
  while (spread < 3 && (warps << (spread + 1)) <= 148 * 4 * 64 + 4096) spread++;   // batches: 4 queries per warp while the launch stays below ~64 warps per sub-partition
  return spread;
}
void launch_odom_corr(const OdomLaunch& o, int iter, cudaStream_t stream) {
  OdomArgs a;
  a.sharp = o.sharp; a.flat = o.flat; a.n_sharp = o.n_sharp; a.n_flat = o.n_flat;
  a.last_corner = o.last_corner; a.last_surf = o.last_surf; a.bound_corner = o.bound_corner; a.bound_surf = o.bound_surf;
  a.grid_corner = o.grid_corner; a.grid_surf = o.grid_surf; a.state = o.state; a.ind = o.ind; a.rows = o.rows; a.iter = iter;
  a.box_corner = nullptr; a.box_surf = nullptr;
  const int nT = ((o.n_sharp + 31) & ~31) + o.n_flat;
  a.spread = odom_spread((long long)nT);
  const long long nthreads = (long long)((nT + 31) / 32) * 32 << a.spread;
  CM_LAUNCH(odom_corr_kernel, (unsigned int)((nthreads + 127) / 128 > 0 ? (nthreads + 127) / 128 : 1), 128, 0, stream, a);
}
void GridBatch::create(int nstreams_, int cap_, cudaStream_t stream) {
  nstreams = nstreams_; cap = cap_ > 0 ? cap_ : 1;
  tcap = next_pow2((unsigned int)(2 * cap));
  entries.reserve((size_t)nstreams * tcap * sizeof(CellEntry)); pts.reserve((size_t)nstreams * cap * sizeof(float4));
  cell_of.reserve((size_t)nstreams * cap * sizeof(int)); cursor.reserve((size_t)nstreams * sizeof(unsigned int));
  views.reserve((size_t)nstreams * sizeof(GridView));
  cudaMemsetAsync(views.p, 0, (size_t)nstreams * sizeof(GridView), stream);   // npts = 0 until a stream's first build
}
void GridBatch::build(const float4* d_pts, const int* d_n, int max_n, const int* d_on, float cell_size, float gate, cudaStream_t stream) {
  const float inv = 1.0f / cell_size;
  if (max_n > cap) max_n = cap;
  const dim3 gt((tcap + 255) / 256, nstreams), gp((max_n + 255) / 256 > 0 ? (max_n + 255) / 256 : 1, nstreams);
  CM_LAUNCH(gridb_clear_kernel, gt, 256, 0, stream, (CellEntry*)entries.p, tcap, (unsigned int*)cursor.p, d_on);
  CM_LAUNCH(gridb_count_kernel, gp, 256, 0, stream, d_pts, d_n, cap, (CellEntry*)entries.p, tcap, inv, (int*)cell_of.p, d_on);
  CM_LAUNCH(gridb_offsets_kernel, gt, 256, 0, stream, (CellEntry*)entries.p, tcap, (unsigned int*)cursor.p, d_on);
  CM_LAUNCH(gridb_scatter_kernel, gp, 256, 0, stream, d_pts, d_n, cap, (CellEntry*)entries.p, tcap, (const int*)cell_of.p, (float4*)pts.p, d_on);
  CM_LAUNCH(gridb_view_kernel, (nstreams + 63) / 64, 64, 0, stream, (GridView*)views.p, (const CellEntry*)entries.p, tcap, (const float4*)pts.p, cap,
            d_n, inv, cell_size, grid_max_level(cell_size, gate), d_on, nstreams);
}

void launch_odom_corr_batch(const OdomBatchLaunch& o, int iter, cudaStream_t stream, const int* d_iter) {
  OdomBatchArgs b;
  b.iter_dev = d_iter;
  b.sharp = o.sharp; b.flat = o.flat; b.cap_sharp = o.cap_sharp; b.cap_flat = o.cap_flat; b.n_sharp = o.n_sharp; b.n_flat = o.n_flat;
  b.last_corner = o.last_corner; b.last_surf = o.last_surf; b.cap_last_corner = o.cap_last_corner; b.cap_last_surf = o.cap_last_surf;
  b.bound_corner = o.bound_corner; b.bound_surf = o.bound_surf; b.grid_corner = o.grid_corner; b.grid_surf = o.grid_surf;
  b.state = o.state; b.ind = o.ind; b.rows = o.rows; b.iter = iter;
  b.box_corner = (const ChunkBox*)o.box_corner; b.box_surf = (const ChunkBox*)o.box_surf; b.box_cap_corner = o.box_cap_corner; b.box_cap_surf = o.box_cap_surf;
  const int nT = ((o.max_sharp + 31) & ~31) + o.max_flat;
  // Every fifth evaluation a query walks several rings of the last cloud (thousands of points, its warp working on one query at a
  // time): with 32 queries per warp a single sweep keeps 10 SMs busy for 440 us.  Thin the warps out -- down to one query per warp --
  // while the launch stays below a few warps per SM sub-partition.
  const int spread = odom_spread((long long)nT * o.nstreams);
  b.spread = spread;
  const long long nthreads = (long long)((nT + 31) / 32) * 32 << spread;
  CM_LAUNCH(odom_corr_batch_kernel, dim3((unsigned int)((nthreads + 127) / 128 > 0 ? (nthreads + 127) / 128 : 1), o.nstreams), 128, 0, stream, b);
}
void launch_odom_boxes_batch(const float4* d_cloud, int cap, const int* d_n, int max_n, int nstreams, void* d_boxes, int box_cap, cudaStream_t stream) {
  const int nb = (max_n + 31) / 32;   // chunks that can hold points (the others are never looked at: the walks stop at n)
  if (nb > 0) CM_LAUNCH(odom_boxes_batch_kernel, dim3((nb * 32 + 127) / 128, nstreams), 128, 0, stream, d_cloud, cap, d_n, (ChunkBox*)d_boxes, box_cap);
}
void launch_odom_gate(MatchState* d_state, const int* d_active, int nstreams, cudaStream_t stream) {
  CM_LAUNCH(odom_gate_kernel, (nstreams + 63) / 64, 64, 0, stream, d_state, d_active, nstreams);
}
void launch_odom_to_end_batch(float4* d_cloud, int cap, const int* d_n, int max_n, int nstreams, const float* d_tf6, const float* d_inv12,
                              const int* d_on, cudaStream_t stream) {
  if (max_n > 0) CM_LAUNCH(odom_to_end_batch_kernel, dim3((max_n + 255) / 256, nstreams), 256, 0, stream, d_cloud, cap, d_n, d_tf6, d_inv12, d_on);
}
void launch_odom_to_end(float4* d_cloud, int n, const float* d_tf6, const float* d_inv12, cudaStream_t stream) {
  if (n > 0) CM_LAUNCH(odom_to_end_kernel, (n + 255) / 256, 256, 0, stream, d_cloud, n, d_tf6, d_inv12);
}

void launch_match(const MatchLaunch& m, cudaStream_t stream, KernelProfiler* prof) {
  launch_match_init(m, stream);
  const bool fused = m.partials && m.tickets && (((m.max_queries > 0 ? m.max_queries : m.cap_corner + m.cap_surf) + 32 + 255) / 256) <= m.partial_blocks;
  for (int it = 0; it < m.prm.max_iterations; it++) {
    launch_match_partial(m, it, stream, prof, fused);
    if (!fused) launch_match_solve(m, it, (const double*)m.sums, stream);
  }
}

// Last node of the WHILE body: next evaluation index, loop again while some stream is still iterating.
__global__ void gn_advance_kernel(cudaGraphConditionalHandle handle, int* iter, const MatchState* state, int nstreams, int max_iterations,
                                  const int* skip) {
  const int it = *iter + 1;
  *iter = it;
  int active = 0;
  for (int s = 0; s < nstreams; s++) active |= state[s].done ? 0 : 1;
  if (skip && *skip) active = 0;
  cudaGraphSetConditional(handle, (active && it < max_iterations) ? 1u : 0u);
}

// One Gauss-Newton evaluation with the evaluation index read from device memory (body of the WHILE node).
static void launch_match_body(const MatchLaunch& m, const int* d_iter, cudaStream_t stream) {
  CorrArgs ca; SolveArgs sa;
  fill_args(m, ca, sa);
  ca.iter_dev = d_iter; sa.iter_dev = d_iter; sa.iter = 0;
  const int capQ = m.cap_corner + m.cap_surf;
  int bx = ((m.max_queries > 0 ? m.max_queries : capQ) + 32 + 255) / 256;
  if (bx < 1) bx = 1;
  dim3 grid(bx, m.nstreams);
  ca.spread = search_spread((long long)bx * 256 * m.nstreams);
  const dim3 sgrid = search_grid(bx, m.nstreams, ca.spread);
  if (m.orig_idx) CM_LAUNCH(search_kernel<true>, sgrid, CM_SEARCH_THREADS, 0, stream, ca);
  else CM_LAUNCH(search_kernel<false>, sgrid, CM_SEARCH_THREADS, 0, stream, ca);
  const int hb = m.hard_blocks > 0 ? m.hard_blocks : hard_blocks_default();
  if (m.orig_idx) CM_LAUNCH(search_hard_kernel<true>, hb, 256, 0, stream, ca);
  else CM_LAUNCH(search_hard_kernel<false>, hb, 256, 0, stream, ca);
  FusedArgs f; f.partials = m.partials; f.tickets = m.tickets; f.sums = m.sums; f.ptiles = m.partial_blocks;
  CM_LAUNCH(fit_solve_kernel, grid, 256, 0, stream, ca, f);
  CM_LAUNCH(solve_warp_kernel, m.nstreams, 32, 0, stream, sa, (const double*)m.sums);
}

// init -> WHILE { body, advance }: as many evaluations as the slowest stream needs, one submission.  init / body enqueue their
// launches on `stream` (captured); the body reads the evaluation index from d_iter.
template <typename InitFn, typename BodyFn>
static cudaGraphExec_t build_while_graph_fn(cudaStream_t stream, int* d_iter, const MatchState* state, int nstreams, int max_iterations,
                                            const int* skip, InitFn init, BodyFn body, unsigned long long* launches_per_eval) {
  cudaGraph_t g = nullptr, gi = nullptr, tmp = nullptr;
  cudaGraphExec_t exec = nullptr;
  const unsigned long long before = g_launch_count;
  bool ok = cudaGraphCreate(&g, 0) == cudaSuccess;
  // init part as a child graph
  if (ok) ok = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  if (ok) {
    init();
    cudaMemsetAsync(d_iter, 0, sizeof(int), stream);
    ok = cudaStreamEndCapture(stream, &gi) == cudaSuccess && gi;
  }
  cudaGraphNode_t n_init = nullptr, n_cond = nullptr;
  if (ok) ok = cudaGraphAddChildGraphNode(&n_init, g, nullptr, 0, gi) == cudaSuccess;
  cudaGraphConditionalHandle handle = 0;
  if (ok) ok = cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault) == cudaSuccess;
  cudaGraphNodeParams prm = {cudaGraphNodeTypeConditional};
  if (ok) {
    prm.conditional.handle = handle; prm.conditional.type = cudaGraphCondTypeWhile; prm.conditional.size = 1;
    ok = cudaGraphAddNode(&n_cond, g, &n_init, 1, &prm) == cudaSuccess && prm.conditional.phGraph_out;
  }
  if (ok) {
    cudaGraph_t bodyg = prm.conditional.phGraph_out[0];
    ok = cudaStreamBeginCaptureToGraph(stream, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      const unsigned long long b0 = g_launch_count;
      body();
      CM_LAUNCH(gn_advance_kernel, 1, 1, 0, stream, handle, d_iter, state, nstreams, max_iterations, skip);
      *launches_per_eval = g_launch_count - b0;
      ok = cudaStreamEndCapture(stream, &tmp) == cudaSuccess;
    }
  }
  if (ok) ok = cudaGraphInstantiate(&exec, g, 0) == cudaSuccess;
  if (gi) cudaGraphDestroy(gi);
  if (g) cudaGraphDestroy(g);
  g_launch_count = before;
  if (!ok) { cudaGetLastError(); if (exec) cudaGraphExecDestroy(exec); return nullptr; }
  return exec;
}

// the mapping stage's loop: init -> WHILE { search, hard search, fit + solve, advance }
static cudaGraphExec_t build_while_graph(const MatchLaunch& m, int* d_iter, cudaStream_t stream, unsigned long long* launches_per_eval) {
  if (!(m.partials && m.tickets && m.hard) || m.prm.max_iterations > CM_MAX_EVALS) return nullptr;
  const int maxq = m.max_queries > 0 ? m.max_queries : m.cap_corner + m.cap_surf;
  if ((maxq + 32 + 255) / 256 > m.partial_blocks) return nullptr;
  return build_while_graph_fn(stream, d_iter, (const MatchState*)m.state, m.nstreams, m.prm.max_iterations, m.skip,
                              [&]() { launch_match_init(m, stream); }, [&]() { launch_match_body(m, d_iter, stream); }, launches_per_eval);
}

// the batch odometry's loop: init (+ the gate that parks the streams without a usable last frame) -> WHILE { correspondences + rows,
// row reduction, 6x6 step, advance }
bool OdomGraphCache::launch(const MatchLaunch& m, const OdomBatchLaunch& o, const int* d_active, cudaStream_t stream) {
  if (!usable || g_timeline.on) return false;
  std::vector<unsigned long long> key;
  auto P = [&](const void* p) { key.push_back((unsigned long long)(uintptr_t)p); };
  auto I = [&](long long v) { key.push_back((unsigned long long)v); };
  I(m.nstreams); P(m.corner); P(m.surf); P(m.n_corner); P(m.n_surf); I(m.cap_corner); I(m.cap_surf); P(m.grid_corner); P(m.grid_surf);
  P(m.pose_in); P(m.state); P(m.rows); P(m.sums); I(m.prm.max_iterations); P(d_active);
  P(o.last_corner); P(o.last_surf); I(o.cap_last_corner); I(o.cap_last_surf); P(o.bound_corner); P(o.bound_surf); P(o.ind);
  I(o.max_sharp); I(o.max_flat); P(o.box_corner); P(o.box_surf); I(o.box_cap_corner); I(o.box_cap_surf);
  for (size_t i = 0; i < entries.size(); i++) {
    Entry& e = entries[i];
    if (e.key == key) {
      if (e.gen != g_alloc_generation) { cudaGraphExecDestroy(e.exec); entries.erase(entries.begin() + i); break; }
      if (cudaGraphLaunch(e.exec, stream) != cudaSuccess) { cudaGetLastError(); return false; }
      g_launch_count += e.launches;
      return true;
    }
  }
  if (entries.size() >= 16) clear();
  if (!d_iter && cudaMalloc(&d_iter, sizeof(int)) != cudaSuccess) { cudaGetLastError(); d_iter = nullptr; usable = false; return false; }
  Entry e; e.key = key; e.gen = g_alloc_generation; unsigned long long per_eval = 4;
  int* di = d_iter;
  e.exec = build_while_graph_fn(stream, di, (const MatchState*)m.state, m.nstreams, m.prm.max_iterations, nullptr,
                                [&]() { launch_match_init(m, stream); launch_odom_gate(m.state, d_active, m.nstreams, stream); },
                                [&]() {
                                  launch_odom_corr_batch(o, 0, stream, di);
                                  launch_match_reduce_solve(m, 0, stream, di);
                                }, &per_eval);
  if (!e.exec) { usable = false; return false; }
  e.launches = 2 + per_eval;
  entries.push_back(e);
  if (cudaGraphLaunch(e.exec, stream) != cudaSuccess) { cudaGetLastError(); return false; }
  g_launch_count += e.launches;
  return true;
}
void OdomGraphCache::clear() {
  for (Entry& e : entries) if (e.exec) cudaGraphExecDestroy(e.exec);
  entries.clear();
}
OdomGraphCache::~OdomGraphCache() { clear(); if (d_iter) cudaFree(d_iter); }

static std::vector<unsigned long long> match_graph_key(const MatchLaunch& m) {
  std::vector<unsigned long long> k;
  auto P = [&](const void* p) { k.push_back((unsigned long long)(uintptr_t)p); };
  auto I = [&](long long v) { k.push_back((unsigned long long)v); };
  auto F = [&](float v) { unsigned int u; memcpy(&u, &v, 4); k.push_back(u); };
  I(m.nstreams); P(m.corner); P(m.surf); P(m.n_corner); P(m.n_surf); I(m.cap_corner); I(m.cap_surf); P(m.grid_corner); P(m.grid_surf);
  P(m.pose_in); P(m.state); P(m.rows); P(m.nn_slot); P(m.sums); P(m.trace); P(m.nn); I(m.orig_idx);
  const int maxq = m.max_queries > 0 ? m.max_queries : m.cap_corner + m.cap_surf;
  I((maxq + 32 + 255) / 256);   // the grids depend on max_queries only through this
  P(m.own_box); I(m.dist_rank); I(m.dist_nranks); P(m.skip); P(m.hard); P(m.hard_count); I(m.hard_cap); I(m.hard_blocks); P(m.partials); P(m.tickets); I(m.partial_blocks);
  I(m.prm.max_iterations); F(m.prm.delta_t_abort); F(m.prm.delta_r_abort); F(m.prm.knn_gate); F(m.prm.plane_max_dist);
  I(m.prm.min_ref_corner); I(m.prm.min_ref_surf); I(m.prm.min_rows); F(m.prm.eig_threshold); I(m.prm.few_rows_continue);
  I(m.prm.own_cube_only); I(m.prm.nan_guard);
  return k;
}

bool MatchGraphCache::launch(const MatchLaunch& m, cudaStream_t stream) {
  if (m.dbg || m.nn || m.trace) return false;   // diagnostics are not replayable
  const std::vector<unsigned long long> key = match_graph_key(m);
  for (size_t i = 0; i < entries.size(); i++) {
    Entry& e = entries[i];
    if (e.key == key) {
      if (e.gen != g_alloc_generation) {   // a device buffer moved since the capture (the key only holds this launch's own pointers)
        cudaGraphExecDestroy(e.exec);
        entries.erase(entries.begin() + i);
        break;
      }
      if (cudaGraphLaunch(e.exec, stream) != cudaSuccess) return false;
      g_launch_count += e.launches;
      return true;
    }
  }
  if (entries.size() >= 32) clear();   // buffers were re-allocated many times: start over
  if (use_while) {
    if (!d_iter) { if (cudaMalloc(&d_iter, sizeof(int)) != cudaSuccess) { cudaGetLastError(); d_iter = nullptr; use_while = false; } }
    if (use_while) {
      Entry e; e.key = key; e.launches = 2; e.gen = g_alloc_generation; unsigned long long per_eval = 4;
      e.exec = build_while_graph(m, d_iter, stream, &per_eval);
      if (e.exec) {
        ++builds;
        e.launches = 1 + per_eval;   // at least one evaluation runs; the real count is data dependent (a lower bound for gpu_launches)
        entries.push_back(e);
        if (cudaGraphLaunch(e.exec, stream) != cudaSuccess) { cudaGetLastError(); return false; }
        g_launch_count += e.launches;
        return true;
      }
      use_while = false;   // conditional nodes unavailable on this driver: fall back to the unrolled graph for good
    }
  }
  const unsigned long long before = g_launch_count;
  if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
  launch_match(m, stream, nullptr);
  cudaGraph_t graph = nullptr;
  if (cudaStreamEndCapture(stream, &graph) != cudaSuccess || !graph) { cudaGetLastError(); return false; }
  Entry e; e.key = key; e.exec = nullptr; e.launches = g_launch_count - before; e.gen = g_alloc_generation;
  g_launch_count = before; ++builds;
  const cudaError_t rc = cudaGraphInstantiate(&e.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (rc != cudaSuccess) { cudaGetLastError(); return false; }
  entries.push_back(e);
  if (cudaGraphLaunch(e.exec, stream) != cudaSuccess) return false;
  g_launch_count += e.launches;
  return true;
}
void MatchGraphCache::clear() {
  for (Entry& e : entries) if (e.exec) cudaGraphExecDestroy(e.exec);
  entries.clear();
}
MatchGraphCache::~MatchGraphCache() { clear(); if (d_iter) cudaFree(d_iter); }

// Streams [s0, s1) of m as a launch of its own (every per-stream array is indexed by blockIdx.y).
static MatchLaunch match_slice(const MatchLaunch& m, int g, int s0, int s1) {
  MatchLaunch r = m;
  const int capQ = m.cap_corner + m.cap_surf;
  r.nstreams = s1 - s0;
  r.corner = m.corner + (size_t)s0 * m.cap_corner; r.surf = m.surf + (size_t)s0 * m.cap_surf;
  r.n_corner = m.n_corner + s0; r.n_surf = m.n_surf + s0;
  r.grid_corner = m.grid_corner + s0; r.grid_surf = m.grid_surf + s0;
  r.pose_in = m.pose_in + 6 * (size_t)s0; r.state = m.state + s0; r.rows = m.rows + (size_t)s0 * capQ;
  r.nn_slot = m.nn_slot ? m.nn_slot + (size_t)s0 * capQ * 5 : nullptr;
  r.sums = m.sums + (size_t)s0 * 32;
  r.trace = m.trace ? m.trace + (size_t)s0 * m.prm.max_iterations : nullptr;
  r.nn = nullptr;
  if (m.hard) {
    const size_t per_stream = (size_t)m.hard_cap / (size_t)m.nstreams;
    r.hard = (char*)m.hard + (size_t)s0 * per_stream * CM_HARD_ITEM_BYTES;
    r.hard_cap = (int)(per_stream * (size_t)(s1 - s0));
    r.hard_count = m.hard_count + g * CM_MAX_EVALS;
  }
  if (m.partials) { r.partials = m.partials + (size_t)s0 * m.partial_blocks * 32; r.tickets = m.tickets + s0; }
  if (g != 0) r.dbg = nullptr;
  return r;
}

void launch_match_groups(const MatchLaunch& m, cudaStream_t stream, int ngroups, cudaStream_t* gs, cudaEvent_t fork, cudaEvent_t* join,
                         KernelProfiler* prof) {
  if (ngroups <= 1 || m.nstreams < 2 || m.nn) { launch_match(m, stream, prof); return; }
  if (ngroups > m.nstreams) ngroups = m.nstreams;
  if (ngroups > 8) ngroups = 8;
  launch_match_init(m, stream);
  cudaEventRecord(fork, stream);
  const bool fused = m.partials && m.tickets && (((m.max_queries > 0 ? m.max_queries : m.cap_corner + m.cap_surf) + 32 + 255) / 256) <= m.partial_blocks;
  std::vector<MatchLaunch> part;
  for (int g = 0; g < ngroups; g++) {
    const int s0 = (int)((long long)m.nstreams * g / ngroups), s1 = (int)((long long)m.nstreams * (g + 1) / ngroups);
    part.push_back(match_slice(m, g, s0, s1));
    cudaStreamWaitEvent(gs[g], fork, 0);
  }
  // issue iteration by iteration, round-robin over the groups, so that the host feeds all group streams evenly
  for (int it = 0; it < m.prm.max_iterations; it++)
    for (int g = 0; g < ngroups; g++) {
      launch_match_partial(part[g], it, gs[g], prof, fused);
      if (!fused) launch_match_solve(part[g], it, (const double*)part[g].sums, gs[g]);
    }
  for (int g = 0; g < ngroups; g++) { cudaEventRecord(join[g], gs[g]); cudaStreamWaitEvent(stream, join[g], 0); }
}

}  // namespace cm
