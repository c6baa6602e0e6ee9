// cm_scanreg.cu -- K1/K2: scan registration (ring compaction, unreliable-point mask, curvature, feature selection,
// per-ring less-flat voxel filter) for organised sweeps, one CTA per (stream, ring).
//
// Replaces OrganisedScanRegistration::process (OrganizedScanRegistration.cpp:82-150) and
// ScanRegistration::extractFeatures with its helpers (ScanRegistration.cpp:190-666, ScanRegistration.h:280-311).
// The reference's loops are sequential greedy sweeps; the parallel formulation used here produces the same
// results, cell for cell:
//  * setScanBuffersFor (:462-522): every write of iteration i touches only cells [i-4, i+5] and its conditional
//    EDGE_BROKEN write tests and writes the SAME cell, so the final state of cell c is obtained by replaying, in
//    order, the (pure, precomputed) events of iterations c-5 .. c+4 on that cell alone.
//  * the stable merge sort by curvature (:449, ScanRegistration.h:280-311) is only consumed (a) to pick, in
//    ascending order, up to maxSurfaceFlat unmasked flat points -> iterated block-wide arg-min on (curvature, index)
//    with the +-5 suppression applied between picks, regions chained in order because the suppression leaks into
//    the next region (:269-284, 524-545); (b) to visit every point above the threshold in descending order ->
//    rank by counting on (curvature, index).
//  * pointClassify (:547-666) is a pure function of 11 consecutive points -> evaluated in parallel.
//  * the emission counters of pass 3 (:305-354) become prefix sums over the descending order.
// Compiled with -fmad=false: float operations are IEEE operations in source order (bit parity with the oracle).
#include "cm_host.h"
#include "cm_math.h"
#include <float.h>

namespace cm {

enum { L_CORNER_SHARP = 1, L_SURFACE_FLAT = -1, L_ONESIDE_FLAT = 5, L_MESSY = 9, L_NONE = 0x7f,
       P_SURF_PICKED_NEAR = 3, P_EDGE_BROKEN = -2, P_NEAR_BLOCK = -3, P_BLIND_BLOCK = -4 };

#define SR_THREADS 512
#define SR_MAXR 8          // curvatureRegion upper bound
#define SR_MAXREG 16       // nFeatureRegions upper bound
#define SR_MAXCOLS 8192     // columns per ring the shared-memory layout can hold at most (scanreg_smem_bytes rejects more)

// ------------------------------------------------------------------------------------------------------------
// kernel 1: valid points per ring (OrganizedScanRegistration.cpp:115-123)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool point_valid(const float4& p, float blind_sq) {
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return false;
  return !(p.x * p.x + p.y * p.y + p.z * p.z < blind_sq);
}

__global__ void __launch_bounds__(256) sr_count_kernel(const float4* __restrict__ frames, int rows, int cols, float blind_sq,
                                                       int* __restrict__ ring_count) {
  const int ring = blockIdx.x, s = blockIdx.y;
  const float4* src = frames + ((size_t)s * rows + ring) * cols;
  int c = 0;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) c += point_valid(src[i], blind_sq) ? 1 : 0;
  __shared__ int sm[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; w++) t += sm[w];
    ring_count[s * rows + ring] = t;
  }
}

// ------------------------------------------------------------------------------------------------------------
// block helpers
// ------------------------------------------------------------------------------------------------------------
// exclusive scan of one int per thread; returns the prefix, *total = block sum.  scratch: >= 32 ints, used as two
// alternating halves (`phase` flips on every call, uniformly over the CTA) so that ONE barrier per scan is enough: every
// warp adds up the warp totals itself, and the next scan writes the other half while slow warps still read this one.
__device__ __forceinline__ int block_scan_excl(int v, int* scratch, int* total, int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  int* sc = scratch + 16 * phase;
  phase ^= 1;
  if (lane == 31) sc[warp] = x;
  __syncthreads();
  int w = lane < (SR_THREADS / 32) ? sc[lane] : 0;   // SR_THREADS / 32 = 16 warp totals
  int incl = w;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  const int base = __shfl_sync(0xffffffffu, incl - w, warp);
  *total = __shfl_sync(0xffffffffu, incl, SR_THREADS / 32 - 1);
  return base + x - v;
}

struct ScanRegParamsDev {
  float scan_period, blind_sq, blind_thr, curv_thr, less_flat_leaf;
  int R, nregions, max_sharp, max_flat;
  double cos175, cos5, cos135, cos45;
};

struct ScanRegArgs {
  const float4* frames;      // [S][rows][cols] x, y, z, intensity
  const float* tags;         // optional [S][rows][cols]: the curvature field (ring + relTime) of every slot, precomputed
                             // by the raw-sweep front end; NULL: ring + scanPeriod * col / cols (organised sweeps)
  int rows, cols;
  const int* ring_count;     // [S][rows]
  ScanRegParamsDev prm;
  // per-ring outputs, capacity `cols` each: [S][rows][cols]
  float4* ring_pts[4];       // 0 sharp, 1 lessSharp, 2 flat, 3 lessFlat (after the per-ring voxel filter)
  int* ring_idx[4];          // optional: cloud indices (3 = lessFlatRaw, before the voxel filter)
  int* ring_n;               // [S][rows][5]: counts of the 4 lists + lessFlatRaw
  // full-resolution outputs, [S][rows*cols] (optional)
  float4* cloud; float* cloud_curv;          // _laserCloud: xyz+intensity, curvature field (= ring + relTime)
  signed char* picked; float* curvature; signed char* label;   // debug / parity
  int* scan_range;           // [S][rows][2] inclusive start / end (_scanIndices)
};

// one window of pointClassify (ScanRegistration.cpp:557-602 / 603-649)
__device__ __forceinline__ bool classify_window(const float* px, const float* py, const float* pz, int c, int R, bool forward,
                                                float v[3]) {
  float cx = 0.f, cy = 0.f, cz = 0.f;
  const int first = forward ? c + R : c;   // forward: c+R, ..., c ; backward: c, c-1, ..., c-R
  for (int t = 0; t <= R; t++) { int id = first - t; cx += px[id]; cy += py[id]; cz += pz[id]; }
  float cnt = (float)(R + 1);
  cx /= cnt; cy /= cnt; cz /= cnt;
  float a00 = 0.f, a10 = 0.f, a20 = 0.f, a11 = 0.f, a21 = 0.f, a22 = 0.f;
  for (int t = 0; t <= R; t++) {
    int id = first - t;
    float ax = px[id] - cx, ay = py[id] - cy, az = pz[id] - cz;
    a00 += ax * ax; a10 += ax * ay; a20 += ax * az; a11 += ay * ay; a21 += ay * az; a22 += az * az;
  }
  float A[6] = {a00 / cnt, a10 / cnt, a20 / cnt, a11 / cnt, a21 / cnt, a22 / cnt};
  float w[3], V[9];
  eig3_sym(A, w, V);
  if (w[2] > 100.f * w[1] && w[2] > 10000.f * w[0]) {
    v[0] = V[2]; v[1] = V[5]; v[2] = V[8];
    float vnorm = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int t = 0; t <= R; t++) {
      int id = first - t;
      float ax = px[id] - cx, ay = py[id] - cy, az = pz[id] - cz;
      float kx = ay * v[2] - az * v[1], ky = az * v[0] - ax * v[2], kz = ax * v[1] - ay * v[0];
      float distance = sqrtf(kx * kx + ky * ky + kz * kz) / vnorm;
      if (fabs((double)distance) > 0.08) return false;
    }
    return true;
  }
  return false;
}

__device__ __forceinline__ int point_classify(const float* px, const float* py, const float* pz, int c, const ScanRegParamsDev& prm) {
  float v1[3], v2[3];
  bool line1 = classify_window(px, py, pz, c, prm.R, false, v1);
  bool line2 = classify_window(px, py, pz, c, prm.R, true, v2);
  if (line1 && line2) {
    float ab = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
    float disab = sqrtf(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]) * sqrtf(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
    float diff = ab / disab;
    if ((double)diff < prm.cos175 || (double)diff > prm.cos5) return L_SURFACE_FLAT;
    else if ((double)diff > prm.cos135 && (double)diff < prm.cos45) return L_CORNER_SHARP;
  }
  if (line1 || line2) return L_ONESIDE_FLAT;
  return L_MESSY;
}

__device__ __forceinline__ float cos_angle(const float* px, const float* py, const float* pz, int a, int b) {   // math_utils.h:82-87
  float ab = px[a] * px[b] + py[a] * py[b] + pz[a] * pz[b];
  float disab = sqrtf(px[a] * px[a] + py[a] * py[a] + pz[a] * pz[a]) * sqrtf(px[b] * px[b] + py[b] * py[b] + pz[b] * pz[b]);
  return ab / disab;
}
__device__ __forceinline__ float sq_diff(const float* px, const float* py, const float* pz, int a, int b) {   // math_utils.h:45-51
  float dx = px[a] - px[b], dy = py[a] - py[b], dz = pz[a] - pz[b];
  return dx * dx + dy * dy + dz * dz;
}

// ------------------------------------------------------------------------------------------------------------
// kernel 2: everything for one ring
// ------------------------------------------------------------------------------------------------------------
// dynamic shared memory layout, `cap` = cols rounded up to a multiple of 4, P2 = next power of two >= cols
//   float px[cap], py[cap], pz[cap], pi[cap], curv[cap]
//   u64   key[P2]                        (sort keys: rank sort scratch / voxel sort)
//   u16   col[cap], lst[4][cap], nf[cap], ord[cap]
//   s8    state[cap], snap[cap], ev[cap], lab[cap]
extern __shared__ unsigned char sr_smem[];

__global__ void __launch_bounds__(SR_THREADS) sr_ring_kernel(ScanRegArgs a) {
  const int ring = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const int rows = a.rows, cols = a.cols;
  const ScanRegParamsDev& prm = a.prm;
  const int R = prm.R;
  const int cap = (cols + 3) & ~3;
  int P2 = 1; while (P2 < cols) P2 <<= 1;

  float* px = reinterpret_cast<float*>(sr_smem);
  float* py = px + cap; float* pz = py + cap; float* pin = pz + cap; float* curv = pin + cap;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(curv + cap);
  const int KEYN = (8 * P2 >= 10 * cap) ? P2 : (10 * cap + 7) / 8;   // `key` doubles as five u16 prefix arrays
  unsigned short* colv = reinterpret_cast<unsigned short*>(key + KEYN);
  unsigned short* lst[4] = {colv + cap, colv + 2 * cap, colv + 3 * cap, colv + 4 * cap};
  unsigned short* nfl = colv + 5 * cap;
  unsigned short* ord = colv + 6 * cap;
  signed char* state = reinterpret_cast<signed char*>(colv + 7 * cap);
  signed char* snap = state + cap; signed char* ev = snap + cap; signed char* lab = ev + cap;
  __shared__ int s_scan[32];
  int scan_phase = 0;   // uniform over the CTA: every thread makes the same sequence of block_scan_excl calls
  __shared__ int s_cnt[5];          // list lengths: sharp, lessSharp, flat, lessFlatRaw, (spare)
  __shared__ int s_misc[8];

  // ---- ring ranges (_scanIndices) -------------------------------------------------------------------------
  const int* rc = a.ring_count + s * rows;
  int S0 = 0;
  for (int r = 0; r < ring; r++) S0 += rc[r];
  const int n = rc[ring];
  const int E0 = (S0 + n > 0) ? S0 + n - 1 : 0;   // range.second = cloudSize > 0 ? cloudSize - 1 : 0
  if (tid == 0 && a.scan_range) { a.scan_range[(s * rows + ring) * 2] = S0; a.scan_range[(s * rows + ring) * 2 + 1] = E0; }
  int* ring_n = a.ring_n + (s * rows + ring) * 5;
  if (tid < 5) s_cnt[tid] = 0;

  // ---- ordered compaction of the ring (OrganizedScanRegistration.cpp:102-126) ---------------------------------
  const float4* src = a.frames + ((size_t)s * rows + ring) * cols;
  {
    int base = 0;
    for (int c0 = 0; c0 < cols; c0 += SR_THREADS) {
      int c = c0 + tid;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      bool ok = false;
      if (c < cols) { p = src[c]; ok = point_valid(p, prm.blind_sq); }
      int tot;
      int pos = block_scan_excl(ok ? 1 : 0, s_scan, &tot, scan_phase);
      if (ok) { int d = base + pos; px[d] = p.x; py[d] = p.y; pz[d] = p.z; pin[d] = p.w; colv[d] = (unsigned short)c; }
      base += tot;
    }
  }
  __syncthreads();
  const size_t cbase = (size_t)s * rows * cols + S0;   // this ring's slice of the full-resolution arrays
  for (int i = tid; i < n; i += SR_THREADS) {
    state[i] = 0; snap[i] = 0; ev[i] = 0; lab[i] = L_NONE; curv[i] = -1.f;
    // curvature field of the point = ring + relTime (OrganizedScanRegistration.cpp:109-110); kept in pin[] from here on
    float tag;
    if (a.tags) tag = a.tags[((size_t)s * rows + ring) * cols + colv[i]];
    else {
      const float relTime = (float)((double)prm.scan_period * (double)colv[i] / (double)cols);
      tag = (float)ring + relTime;
    }
    if (a.cloud) {
      a.cloud[cbase + i] = make_float4(px[i], py[i], pz[i], pin[i]);
      a.cloud_curv[cbase + i] = tag;
    }
    pin[i] = tag;
  }
  __syncthreads();
  const bool ring_ok = !(E0 <= S0 + 2 * R) && n > 0;   // "skip empty scans", ScanRegistration.cpp:205
  if (!ring_ok) {
    for (int i = tid; i < n; i += SR_THREADS) {
      if (a.picked) a.picked[cbase + i] = 0;
      if (a.curvature) a.curvature[cbase + i] = -1.f;
      if (a.label) a.label[cbase + i] = L_NONE;
    }
    if (tid < 5) ring_n[tid] = 0;
    return;
  }

  // ---- setScanBuffersFor (ScanRegistration.cpp:462-522), parallel replay ----------------------------------------
  // events of the main loop, i in [R, n-1-R): bit0-1 type (1 blind, 2 jump far->near "A", 3 jump "B"), bit2 ratio test
  // s_evmask: one bit per cell "an event happened here" -- events are rare (range jumps, grazing incidence), so most cells
  // find an empty window and skip the replay loop
  __shared__ unsigned int s_evmask[SR_MAXCOLS / 32 + 2];
  for (int i0 = 0; i0 < n; i0 += SR_THREADS) {
    const int i = i0 + tid;
    int e = 0;
    if (i >= R && i < n - 1 - R) {
      if (cos_angle(px, py, pz, i, i + 1) < prm.blind_thr) e = 1;
      else {
        float diffNext = sq_diff(px, py, pz, i + 1, i);
        if ((double)diffNext > 1.0) {
          float depth1 = sqrtf(px[i] * px[i] + py[i] * py[i] + pz[i] * pz[i]);
          float depth2 = sqrtf(px[i + 1] * px[i + 1] + py[i + 1] * py[i + 1] + pz[i + 1] * pz[i + 1]);
          float diffPrev = sq_diff(px, py, pz, i - 1, i);
          e = (depth1 > depth2) ? 2 : 3;
          if ((double)(diffPrev / diffNext) < 0.2) e |= 4;
        }
      }
    }
    if (i < n) ev[i] = (signed char)e;
    const unsigned int bits = __ballot_sync(0xffffffffu, e != 0);
    if ((tid & 31) == 0) s_evmask[i >> 5] = bits;     // i is a multiple of 32 here; words past n are zero
  }
  if (tid == 0) s_evmask[((n + 31) >> 5)] = 0u;
  if (tid < R) {   // head / tail blind tests (:468-484): R + 1 cells each
    s_misc[tid] = 0;
  }
  __syncthreads();
  if (tid < R) {
    int f = 0;
    if (cos_angle(px, py, pz, tid, tid + 1) < prm.blind_thr) f |= 1;
    if (cos_angle(px, py, pz, n - 1 - tid, n - 2 - tid) < prm.blind_thr) f |= 2;
    s_misc[tid] = f;
  }
  __syncthreads();
  for (int c = tid; c < n; c += SR_THREADS) {
    int st = 0;
    if (c <= 2 * R || c >= n - 1 - 2 * R) {   // only the first / last 2R+1 cells can be hit by the head / tail tests
      for (int i = 0; i < R; i++) {
        int f = s_misc[i];
        if ((f & 1) && c >= i && c <= i + R) st = P_BLIND_BLOCK;
        if ((f & 2) && c >= n - 1 - i - R && c <= n - 1 - i) st = P_BLIND_BLOCK;
      }
    }
    int lo = c - R; if (lo < R) lo = R;
    int hi = c + R - 1; if (hi > n - 2 - R) hi = n - 2 - R;
    bool any = false;
    if (lo <= hi) {   // bits lo..hi of the event bitmap (the window spans at most two words: 2R <= 16)
      const unsigned long long w2 = (unsigned long long)s_evmask[lo >> 5] | ((unsigned long long)s_evmask[(lo >> 5) + 1] << 32);
      any = ((w2 >> (lo & 31)) & ((1ull << (hi - lo + 1)) - 1ull)) != 0ull;
    }
    if (any) {
      for (int i = lo; i <= hi; i++) {
        int e = ev[i];
        int type = e & 3;
        if (type == 1) { if (c >= i - R + 1 && c <= i + R) st = P_BLIND_BLOCK; }
        else if (type == 2) {
          if (c == i + 1 && st > P_NEAR_BLOCK && (e & 4)) st = P_EDGE_BROKEN;
          if (c >= i - R + 1 && c <= i) st = P_NEAR_BLOCK;
        } else if (type == 3) {
          if (c == i && st > P_NEAR_BLOCK && (e & 4)) st = P_EDGE_BROKEN;
          if (c >= i + 1 && c <= i + R) st = P_NEAR_BLOCK;
        }
      }
    }
    state[c] = (signed char)st;
  }

  // ---- curvature (ScanRegistration.cpp:429-445) over the union of the regions ---------------------------------
  const float pointWeight = (float)(-2 * R);
  for (int i = tid; i < n; i += SR_THREADS) {
    if (i >= R && i <= n - 2 - R) {
      float dX = pointWeight * px[i], dY = pointWeight * py[i], dZ = pointWeight * pz[i];
      for (int j = 1; j <= R; j++) {
        dX += px[i + j] + px[i - j];
        dY += py[i + j] + py[i - j];
        dZ += pz[i + j] + pz[i - j];
      }
      curv[i] = dX * dX + dY * dY + dZ * dZ;
    }
  }
  __syncthreads();

  // ---- region bounds (ScanRegistration.cpp:249-262), size_t integer arithmetic on ABSOLUTE indices --------------
  __shared__ int reg_sp[SR_MAXREG], reg_ep[SR_MAXREG];   // relative to the ring; ep < sp: skipped
  if (tid < prm.nregions) {
    unsigned long long s5 = (unsigned long long)S0 + R, e5 = (unsigned long long)E0 - R, NR = prm.nregions, j = tid;
    unsigned long long sp = (s5 * (NR - j) + e5 * j) / NR;
    unsigned long long ep = (s5 * (NR - 1 - j) + e5 * (j + 1)) / NR - 1;
    if (ep <= sp) { reg_sp[tid] = 1; reg_ep[tid] = 0; }
    else { reg_sp[tid] = (int)(sp - S0); reg_ep[tid] = (int)(ep - S0); }
  }
  __syncthreads();

  // From here on the regions of the ring are processed TOGETHER.  Per-region quantities live in small arrays; per-cell
  // work runs over the whole ring so that all 512 threads stay busy.
  __shared__ int nf_begin[SR_MAXREG + 1];      // region j's cells above the threshold: nfl[nf_begin[j] .. nf_begin[j+1])
  __shared__ int p1_begin[SR_MAXREG + 1];      // region j's pass-1 picks: p1buf[p1_begin[j] .. p1_begin[j+1])
  __shared__ unsigned short p1buf[SR_MAXREG * 8];
  __shared__ int cnt2[4][SR_MAXREG], cnt3[4][SR_MAXREG];   // per list (sharp, lessSharp, flat, lessFlatRaw) and region
  __shared__ int start2[2][SR_MAXREG], start3[3][SR_MAXREG];   // exclusive prefixes at region starts
  __shared__ int base[4][SR_MAXREG + 1];
  const int NR = prm.nregions;
  // region of a cell (regions are contiguous and ordered; skipped regions have ep < sp)
  for (int c = tid; c < n; c += SR_THREADS) {   // ev[] is free after the mask replay: region id of every cell (-1: none)
    int rj = -1;
    for (int j = 0; j < NR; j++) if (reg_ep[j] >= reg_sp[j] && c >= reg_sp[j] && c <= reg_ep[j]) rj = j;
    ev[c] = (signed char)rj;
  }
  __syncthreads();
  auto region_of = [&](int c) -> int { return (int)ev[c]; };
  if (tid < 4 * SR_MAXREG) { (&cnt2[0][0])[tid] = 0; (&cnt3[0][0])[tid] = 0; }
  // ---- cells above the curvature threshold, in index order (they are the pass-3 candidates, :305-314) -------------
  {
    int total = 0;
    for (int c0 = 0; c0 < n; c0 += SR_THREADS) {
      const int c = c0 + tid;
      const int rj = c < n ? region_of(c) : -1;
      const bool isnf = rj >= 0 && !(curv[c] < prm.curv_thr);
      int tot;
      const int pos = block_scan_excl(isnf ? 1 : 0, s_scan, &tot, scan_phase);
      if (isnf) nfl[total + pos] = (unsigned short)c;
      if (rj >= 0 && c == reg_sp[rj]) nf_begin[rj] = total + pos;
      total += tot;
    }
    __syncthreads();   // nf_begin[] of the last chunk (the scan itself has a single barrier)
    if (tid == 0) {
      nf_begin[NR] = total;
      for (int j = NR - 1; j >= 0; j--) if (reg_ep[j] < reg_sp[j]) nf_begin[j] = nf_begin[j + 1];   // skipped regions are empty
    }
  }
  __syncthreads();
  const int m_all = nf_begin[NR];
  // ---- warp 0: pass 1, chained over the regions (greedy flat picks with +-R suppression, :267-284, 524-545);
  //      warps 1..15: pointClassify of every pass-3 candidate (:316, 547-666) -- independent of the picks ------------
  if (tid < 32) {
    const int lane = tid;
    int np1 = 0;
    for (int j = 0; j < NR; j++) {
      const int sp = reg_sp[j], ep = reg_ep[j];
      if (lane == 0) p1_begin[j] = np1;
      if (ep >= sp) {
        for (int k = 0; k < prm.max_flat; k++) {
          unsigned long long best = 0xFFFFFFFFFFFFFFFFull;
          for (int c = sp + lane; c <= ep; c += 32) {
            const float cv = curv[c];
            if (state[c] != P_SURF_PICKED_NEAR && cv < prm.curv_thr) {
              const unsigned long long kk = ((unsigned long long)__float_as_uint(cv) << 32) | (unsigned int)c;
              best = kk < best ? kk : best;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o); best = y < best ? y : best; }
          if (best == 0xFFFFFFFFFFFFFFFFull) break;
          const int c = (int)(best & 0xFFFFFFFFu);
          __syncwarp();   // every lane has finished reading state[] (the shuffles order execution, this orders memory)
          if (lane <= 2 * R) state[c - R + lane] = P_SURF_PICKED_NEAR;   // markAsPicked: c-R .. c+R
          if (lane == 0 && np1 < SR_MAXREG * 8) p1buf[np1] = (unsigned short)c;
          np1++;
          __syncwarp();
        }
        for (int c = sp + lane; c <= ep; c += 32) snap[c] = state[c];   // what passes 2 and 3 of region j see
        __syncwarp();
      }
    }
    if (lane == 0) p1_begin[NR] = np1;
  } else {
    // pointClassify(c) looks at the window {c, c-1, .., c-R} and at {c+R, .., c} -- the second one IS the first window of
    // point c+R, summed in the same order, so every window is evaluated once: by its own point when that point is a
    // candidate too, else by the point R cells before it.  Results per window position: line flag + direction.
    float2* wxy = reinterpret_cast<float2*>(key);            // `key` and the four index lists are not in use yet
    float* wz = reinterpret_cast<float*>(lst[0]);            // lst[0..1]
    signed char* wfl = reinterpret_cast<signed char*>(lst[2]);   // window result: 1 line, 3 no line   (first half of lst[2])
    signed char* wcand = wfl + cap;                               // 1: the cell is a candidate itself (second half of lst[2])
    const int NT = SR_THREADS - 32, t = tid - 32;
    for (int c = t; c < n; c += NT) wcand[c] = 0;
    asm volatile("bar.sync 1, %0;" ::"n"(SR_THREADS - 32));
    for (int i = t; i < m_all; i += NT) wcand[nfl[i]] = 1;
    asm volatile("bar.sync 1, %0;" ::"n"(SR_THREADS - 32));
    for (int i = t; i < m_all; i += NT) {
      const int c = nfl[i];
      float v[3];
      const bool own_fwd = wcand[c + R] == 0;                // c+R is not a candidate: this thread owns its window
      const bool l1 = classify_window(px, py, pz, c, R, false, v);
      if (l1) { wxy[c] = make_float2(v[0], v[1]); wz[c] = v[2]; }
      wfl[c] = l1 ? 1 : 3;
      if (own_fwd) {
        const bool l2 = classify_window(px, py, pz, c + R, R, false, v);
        if (l2) { wxy[c + R] = make_float2(v[0], v[1]); wz[c + R] = v[2]; }
        wfl[c + R] = l2 ? 1 : 3;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SR_THREADS - 32));
    for (int i = t; i < m_all; i += NT) {                    // ScanRegistration.cpp:651-665
      const int c = nfl[i];
      const bool line1 = wfl[c] == 1, line2 = wfl[c + R] == 1;
      int l = L_MESSY;
      if (line1 || line2) l = L_ONESIDE_FLAT;
      if (line1 && line2) {
        const float2 a = wxy[c], b = wxy[c + R];
        const float az = wz[c], bz = wz[c + R];
        float ab = a.x * b.x + a.y * b.y + az * bz;
        float disab = sqrtf(a.x * a.x + a.y * a.y + az * az) * sqrtf(b.x * b.x + b.y * b.y + bz * bz);
        float diff = ab / disab;
        if ((double)diff < prm.cos175 || (double)diff > prm.cos5) l = L_SURFACE_FLAT;
        else if ((double)diff > prm.cos135 && (double)diff < prm.cos45) l = L_CORNER_SHARP;
      }
      lab[c] = (signed char)l;
    }
  }
  __syncthreads();
  for (int i = tid; i < m_all; i += SR_THREADS) {
    const int c = nfl[i];
    key[i] = ((unsigned long long)__float_as_uint(curv[c]) << 32) | (unsigned int)c;
  }
  __syncthreads();
  // ---- descending (curvature, index) order inside every region: rank by counting -------------------------------------
  for (int i = tid; i < m_all; i += SR_THREADS) {
    const int rj = region_of(nfl[i]);
    const unsigned long long ki = key[i];
    int rank = 0;
    for (int k = nf_begin[rj]; k < nf_begin[rj + 1]; k++) rank += (key[k] > ki) ? 1 : 0;
    ord[nf_begin[rj] + rank] = nfl[i];
  }
  __syncthreads();
  // `key` is free again: carve four u16 prefix arrays out of it
  unsigned short* pfa = reinterpret_cast<unsigned short*>(key);
  unsigned short* pfb = pfa + cap; unsigned short* pfc = pfb + cap;
  // ---- pass 2 (:286-303), whole ring: c < thr -> lessFlatRaw; EDGE_BROKEN (as seen after the region's picks) -> sharp
  //      and lessSharp.  Ring-wide exclusive prefixes; the value at a region's first cell is that region's offset. ------
  {
    int t1 = 0, t2 = 0;
    for (int c0 = 0; c0 < n; c0 += SR_THREADS) {
      const int c = c0 + tid;
      const int rj = c < n ? region_of(c) : -1;
      const bool isflat = rj >= 0 && (curv[c] < prm.curv_thr);
      const bool isedge = rj >= 0 && (snap[c] == P_EDGE_BROKEN);
      // both counters in one scan: low / high 16 bits (a ring has < 65536 cells)
      int a12;
      const int p12 = block_scan_excl((isflat ? 1 : 0) | (isedge ? 0x10000 : 0), s_scan, &a12, scan_phase);
      const int p1 = p12 & 0xFFFF, p2 = p12 >> 16, a1 = a12 & 0xFFFF, a2 = a12 >> 16;
      if (c < n) { pfa[c] = (unsigned short)(t1 + p1); pfb[c] = (unsigned short)(t2 + p2); }
      if (rj >= 0 && c == reg_sp[rj]) { start2[0][rj] = t1 + p1; start2[1][rj] = t2 + p2; }
      if (isflat) atomicAdd(&cnt2[3][rj], 1);
      if (isedge) { atomicAdd(&cnt2[0][rj], 1); atomicAdd(&cnt2[1][rj], 1); }
      t1 += a1; t2 += a2;
    }
  }
  // ---- pass 3 (:305-354) over the descending order of every region: emission counters become prefix sums -----------
  unsigned short* qfa = pfc; unsigned short* qfb = pfc + cap;   // corner / surf prefixes along `ord`
  {
    int t1 = 0, t2 = 0;
    for (int i0 = 0; i0 < m_all; i0 += SR_THREADS) {
      const int i = i0 + tid;
      const bool in = i < m_all;
      const int c = in ? ord[i] : 0;
      const int l = in ? lab[c] : L_MESSY;
      const int rj = in ? region_of(c) : -1;
      const bool is_corner = in && l == L_CORNER_SHARP && snap[c] > P_EDGE_BROKEN;
      const bool is_surf = in && (l == L_SURFACE_FLAT || l == L_ONESIDE_FLAT);
      int a12;
      const int p12 = block_scan_excl((is_corner ? 1 : 0) | (is_surf ? 0x10000 : 0), s_scan, &a12, scan_phase);
      const int p1 = p12 & 0xFFFF, p2 = p12 >> 16, a1 = a12 & 0xFFFF, a2 = a12 >> 16;
      if (in) { qfa[i] = (unsigned short)(t1 + p1); qfb[i] = (unsigned short)(t2 + p2); }
      if (in && i == nf_begin[rj]) { start3[0][rj] = t1 + p1; start3[1][rj] = t2 + p2; }
      if (is_corner) atomicAdd(&cnt3[1][rj], 1);
      if (is_surf) atomicAdd(&cnt3[3][rj], 1);
      t1 += a1; t2 += a2;
    }
  }
  __syncthreads();
  // flat emissions of pass 3: ONESIDE_FLAT while the shared counter (bumped by SURFACE_FLAT too) is below the cap: at most
  // max_flat per region, found by walking the first surf elements of the region (tiny: one thread per region)
  if (tid < NR) {
    const int j = tid;
    int nflat = 0;
    if (reg_ep[j] >= reg_sp[j]) {
      int seen = 0;
      for (int i = nf_begin[j]; i < nf_begin[j + 1] && seen < prm.max_flat; i++) {
        const int l = lab[ord[i]];
        if (l == L_SURFACE_FLAT || l == L_ONESIDE_FLAT) { if (l == L_ONESIDE_FLAT) nflat++; seen++; }
      }
    }
    cnt3[2][j] = nflat;
    cnt2[2][j] = p1_begin[j + 1] - p1_begin[j];
    cnt3[0][j] = cnt3[1][j] < prm.max_sharp ? cnt3[1][j] : prm.max_sharp;
  }
  __syncthreads();
  if (tid < 4) {   // list bases: regions in order, pass-1/2 entries before pass-3 entries
    int acc = 0;
    for (int j = 0; j < NR; j++) { base[tid][j] = acc; acc += cnt2[tid][j] + cnt3[tid][j]; }
    base[tid][NR] = acc;
    s_cnt[tid] = acc;
  }
  __syncthreads();
  // ---- placement -----------------------------------------------------------------------------------------------------
  for (int c = tid; c < n; c += SR_THREADS) {   // pass 2
    const int rj = region_of(c);
    if (rj < 0) continue;
    if (curv[c] < prm.curv_thr) lst[3][base[3][rj] + (int)pfa[c] - start2[0][rj]] = (unsigned short)c;
    if (snap[c] == P_EDGE_BROKEN) {
      const int o = (int)pfb[c] - start2[1][rj];
      lst[0][base[0][rj] + o] = (unsigned short)c; lst[1][base[1][rj] + o] = (unsigned short)c;
    }
  }
  for (int i = tid; i < m_all; i += SR_THREADS) {   // pass 3
    const int c = ord[i];
    const int l = lab[c];
    const int rj = region_of(c);
    if (l == L_CORNER_SHARP && snap[c] > P_EDGE_BROKEN) {
      const int o = (int)qfa[i] - start3[0][rj];
      lst[1][base[1][rj] + cnt2[1][rj] + o] = (unsigned short)c;
      if (o < prm.max_sharp) lst[0][base[0][rj] + cnt2[0][rj] + o] = (unsigned short)c;
    }
    if (l == L_SURFACE_FLAT || l == L_ONESIDE_FLAT) {
      const int o = (int)qfb[i] - start3[1][rj];
      lst[3][base[3][rj] + cnt2[3][rj] + o] = (unsigned short)c;
    }
  }
  if (tid < NR) {   // flat list: pass-1 picks, then the ONESIDE_FLAT picks of pass 3
    const int j = tid;
    int o = base[2][j];
    for (int i = p1_begin[j]; i < p1_begin[j + 1]; i++) lst[2][o++] = p1buf[i];
    if (reg_ep[j] >= reg_sp[j]) {
      int seen = 0;
      for (int i = nf_begin[j]; i < nf_begin[j + 1] && seen < prm.max_flat; i++) {
        const int c = ord[i];
        const int l = lab[c];
        if (l == L_SURFACE_FLAT || l == L_ONESIDE_FLAT) { if (l == L_ONESIDE_FLAT) lst[2][o++] = (unsigned short)c; seen++; }
      }
    }
  }
  __syncthreads();

  // ---- write the index-based outputs ------------------------------------------------------------------------------
  float4* o_pts[4]; int* o_idx[4];
  const size_t rbase = ((size_t)s * rows + ring) * cols;
  for (int l = 0; l < 4; l++) { o_pts[l] = a.ring_pts[l] + rbase; o_idx[l] = a.ring_idx[l] ? a.ring_idx[l] + rbase : nullptr; }
  for (int l = 0; l < 3; l++) {
    int cnt = s_cnt[l];
    for (int i = tid; i < cnt; i += SR_THREADS) {
      int c = lst[l][i];
      o_pts[l][i] = make_float4(px[c], py[c], pz[c], pin[c]);   // toXYZI: intensity := curvature field
      if (o_idx[l]) o_idx[l][i] = S0 + c;
    }
  }
  const int nlf = s_cnt[3];
  if (o_idx[3]) for (int i = tid; i < nlf; i += SR_THREADS) o_idx[3][i] = S0 + lst[3][i];
  for (int i = tid; i < n; i += SR_THREADS) {
    if (a.picked) a.picked[cbase + i] = state[i];
    if (a.curvature) a.curvature[cbase + i] = curv[i];
    if (a.label) a.label[cbase + i] = lab[i];
  }

  // ---- per-ring voxel filter of the less-flat points (ScanRegistration.cpp:390-399; cm_voxel.cu semantics) -------
  // reuse px..: the member points are addressed through lst[3]; intensity of a member = ring + relTime
  __shared__ int vb_i[8];
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = tid; i < nlf; i += SR_THREADS) {
    int c = lst[3][i];
    mn[0] = fminf(mn[0], px[c]); mn[1] = fminf(mn[1], py[c]); mn[2] = fminf(mn[2], pz[c]);
    mx[0] = fmaxf(mx[0], px[c]); mx[1] = fmaxf(mx[1], py[c]); mx[2] = fmaxf(mx[2], pz[c]);
  }
  {
    float* fs = reinterpret_cast<float*>(key);   // 6 * 16 floats of scratch
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
        mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
      }
    __syncthreads();
    if ((tid & 31) == 0) for (int k = 0; k < 3; k++) { fs[(tid >> 5) * 6 + k] = mn[k]; fs[(tid >> 5) * 6 + 3 + k] = mx[k]; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < SR_THREADS / 32; w++)
        for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], fs[w * 6 + k]); mx[k] = fmaxf(mx[k], fs[w * 6 + 3 + k]); }
      const float inv = 1.0f / prm.less_flat_leaf;
      vb_i[6] = 0;
      if (nlf > 0) {
        long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1,
                  dz = (long long)((mx[2] - mn[2]) * inv) + 1;
        if (dx * dy * dz > 2147483647LL) vb_i[6] = 1;   // passthrough
        int minb[3], maxb[3];
        for (int k = 0; k < 3; k++) { minb[k] = (int)floorf(mn[k] * inv); maxb[k] = (int)floorf(mx[k] * inv); vb_i[k] = minb[k]; }
        int d0 = maxb[0] - minb[0] + 1, d1 = maxb[1] - minb[1] + 1;
        vb_i[3] = d0; vb_i[4] = d0 * d1;
      }
    }
    __syncthreads();
  }
  const bool passthrough = vb_i[6] != 0;
  {
    // PCL sorts the points by voxel index (stable here: emission order inside a voxel).  Consecutive list entries
    // very often share a voxel (the list follows the scan), so the sort is done on RUNS of equal voxel index:
    // key = (voxel, position of the run's first entry); a voxel's runs then come out in emission order and summing
    // run after run reproduces the per-point order exactly, with ~3-5x fewer keys to sort.
    const float inv = 1.0f / prm.less_flat_leaf;
    unsigned int* vidx = reinterpret_cast<unsigned int*>(curv);   // curv[] has been written out: reuse as voxel index per entry
    unsigned short* run_start = nfl;                                // nfl / ord are free as well
    unsigned short* run_end = ord;                                  // indexed by the run's start position
    for (int i = tid; i < nlf; i += SR_THREADS) {
      const int c = lst[3][i];
      unsigned int idx;
      if (passthrough) idx = (unsigned int)i;
      else {
        int ijk0 = (int)(floorf(px[c] * inv) - (float)vb_i[0]);
        int ijk1 = (int)(floorf(py[c] * inv) - (float)vb_i[1]);
        int ijk2 = (int)(floorf(pz[c] * inv) - (float)vb_i[2]);
        idx = (unsigned int)(ijk0 + ijk1 * vb_i[3] + ijk2 * vb_i[4]);
      }
      vidx[i] = idx;
    }
    __syncthreads();
    int nruns = 0;
    for (int i0 = 0; i0 < nlf; i0 += SR_THREADS) {
      const int i = i0 + tid;
      const bool rs = i < nlf && (i == 0 || vidx[i] != vidx[i - 1]);
      int tot;
      const int pos = block_scan_excl(rs ? 1 : 0, s_scan, &tot, scan_phase);
      if (rs) run_start[nruns + pos] = (unsigned short)i;
      nruns += tot;
    }
    __syncthreads();
    int P = 1; while (P < nruns) P <<= 1;
    for (int r = tid; r < P; r += SR_THREADS) {
      unsigned long long k = 0xFFFFFFFFFFFFFFFFull;
      if (r < nruns) {
        const int st = run_start[r];
        run_end[st] = (unsigned short)(r + 1 < nruns ? run_start[r + 1] : nlf);
        k = ((unsigned long long)vidx[st] << 32) | (unsigned int)st;
      }
      key[r] = k;
    }
    __syncthreads();
    // bitonic sort of key[0..P)
    for (int kk = 2; kk <= P; kk <<= 1)
      for (int jj = kk >> 1; jj > 0; jj >>= 1) {
        for (int pr = tid; pr < (P >> 1); pr += SR_THREADS) {   // one compare-exchange per pair: every lane works
          const int i = ((pr & ~(jj - 1)) << 1) | (pr & (jj - 1)), ixj = i | jj;
          const unsigned long long x = key[i], y = key[ixj];
          const bool up = (i & kk) == 0;
          if ((x > y) == up) { key[i] = y; key[ixj] = x; }
        }
        __syncthreads();
      }
    // voxel heads among the sorted runs -> output rank -> centroid (runs in order, entries of a run in order)
    int base = 0;
    for (int r0 = 0; r0 < nruns; r0 += SR_THREADS) {
      const int r = r0 + tid;
      const bool head = r < nruns && (r == 0 || (key[r] >> 32) != (key[r - 1] >> 32));
      int tot;
      const int pos = block_scan_excl(head ? 1 : 0, s_scan, &tot, scan_phase);
      if (head) {
        const unsigned int v = (unsigned int)(key[r] >> 32);
        float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;
        int cntp = 0;
        for (int rr = r; rr < nruns && (unsigned int)(key[rr] >> 32) == v; rr++) {
          const int st = (int)(key[rr] & 0xFFFFFFFFu), en = run_end[st];
          for (int j = st; j < en; j++) {
            const int c = lst[3][j];
            cx += px[c]; cy += py[c]; cz += pz[c]; ci += pin[c];
          }
          cntp += en - st;
        }
        const float cnt = (float)cntp;
        o_pts[3][base + pos] = make_float4(cx / cnt, cy / cnt, cz / cnt, ci / cnt);
      }
      base += tot;
    }
    if (tid == 0) { ring_n[0] = s_cnt[0]; ring_n[1] = s_cnt[1]; ring_n[2] = s_cnt[2]; ring_n[3] = base; ring_n[4] = nlf; }
  }
}

// ------------------------------------------------------------------------------------------------------------
// kernel 3: concatenate the per-ring lists into per-stream clouds (_cornerPointsSharp += ..., :280-399)
// ------------------------------------------------------------------------------------------------------------
struct AssembleArgs {
  const float4* ring_pts[4]; const int* ring_idx[4]; const int* ring_n;
  int rows, cols;
  float4* out_pts[4]; int* out_idx[4];   // [S][cap[l]]
  int cap[4];
  int* out_n;                            // [S][4]
  int* overflow;
};
__global__ void __launch_bounds__(256) sr_assemble_kernel(AssembleArgs a) {
  const int ring = blockIdx.x, s = blockIdx.y;
  for (int l = 0; l < 4; l++) {
    int off = 0;
    for (int r = 0; r < ring; r++) off += a.ring_n[(s * a.rows + r) * 5 + l];
    const int cnt = a.ring_n[(s * a.rows + ring) * 5 + l];
    const size_t rbase = ((size_t)s * a.rows + ring) * a.cols;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      if (off + i < a.cap[l]) {
        a.out_pts[l][(size_t)s * a.cap[l] + off + i] = a.ring_pts[l][rbase + i];
      } else if (a.overflow) atomicExch(a.overflow, 1);
    }
    if (a.out_idx[l] && a.ring_idx[l]) {
      // index list 3 is lessFlatRaw (pre-filter), whose length is ring_n[..][4]
      int ioff = off, icnt = cnt;
      if (l == 3) {
        ioff = 0;
        for (int r = 0; r < ring; r++) ioff += a.ring_n[(s * a.rows + r) * 5 + 4];
        icnt = a.ring_n[(s * a.rows + ring) * 5 + 4];
      }
      for (int i = threadIdx.x; i < icnt; i += blockDim.x) a.out_idx[l][(size_t)s * a.rows * a.cols + ioff + i] = a.ring_idx[l][rbase + i];
    }
    if (ring == a.rows - 1 && threadIdx.x == 0) {
      int tot = off + cnt;
      a.out_n[s * 5 + l] = tot < a.cap[l] ? tot : a.cap[l];
      if (l == 3) {
        int t4 = 0;
        for (int r = 0; r < a.rows; r++) t4 += a.ring_n[(s * a.rows + r) * 5 + 4];
        a.out_n[s * 5 + 4] = t4;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
size_t scanreg_smem_bytes(int cols) {
  int cap = (cols + 3) & ~3;
  int P2 = 1; while (P2 < cols) P2 <<= 1;
  int keyn = (8 * P2 >= 10 * cap) ? P2 : (10 * cap + 7) / 8;
  return (size_t)cap * 4 * 5 + (size_t)keyn * 8 + (size_t)cap * 2 * 7 + (size_t)cap * 4;
}

void ScanRegistrationGpu::run(const ScanRegLaunch& L, cudaStream_t stream) {
  const int S = L.nstreams, rows = L.rows, cols = L.cols;
  const size_t ring_slots = (size_t)S * rows * cols;
  ring_count.reserve(sizeof(int) * S * rows);
  ring_n.reserve(sizeof(int) * S * rows * 5);
  for (int l = 0; l < 4; l++) {
    ring_pts[l].reserve(ring_slots * sizeof(float4));
    if (L.want_idx) ring_idx[l].reserve(ring_slots * sizeof(int));
  }
  ScanRegParamsDev p;
  p.scan_period = L.scan_period; p.blind_sq = L.blind_sq_override >= 0.f ? L.blind_sq_override : L.blind_radius * L.blind_radius; p.blind_thr = L.blind_thr; p.curv_thr = L.curv_thr;
  p.less_flat_leaf = L.less_flat_leaf; p.R = L.R; p.nregions = L.nregions; p.max_sharp = L.max_sharp; p.max_flat = L.max_flat;
  p.cos175 = L.cos175; p.cos5 = L.cos5; p.cos135 = L.cos135; p.cos45 = L.cos45;
  dim3 grid(rows, S);
  CM_LAUNCH(sr_count_kernel, grid, 256, 0, stream, L.frames, rows, cols, p.blind_sq, (int*)ring_count.p);
  ScanRegArgs a;
  a.frames = L.frames; a.tags = L.tags; a.rows = rows; a.cols = cols; a.ring_count = (const int*)ring_count.p; a.prm = p;
  for (int l = 0; l < 4; l++) { a.ring_pts[l] = (float4*)ring_pts[l].p; a.ring_idx[l] = L.want_idx ? (int*)ring_idx[l].p : nullptr; }
  a.ring_n = (int*)ring_n.p;
  a.cloud = L.cloud; a.cloud_curv = L.cloud_curv; a.picked = L.picked; a.curvature = L.curvature; a.label = L.label;
  a.scan_range = L.scan_range;
  size_t smem = scanreg_smem_bytes(cols);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(sr_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  if (L.prof) L.prof->begin(stream);
  CM_LAUNCH(sr_ring_kernel, grid, SR_THREADS, smem, stream, a);
  if (L.prof) L.prof->end(stream);
  AssembleArgs as;
  for (int l = 0; l < 4; l++) {
    as.ring_pts[l] = (const float4*)ring_pts[l].p; as.ring_idx[l] = L.want_idx ? (const int*)ring_idx[l].p : nullptr;
    as.out_pts[l] = L.out_pts[l]; as.out_idx[l] = L.want_idx ? L.out_idx[l] : nullptr; as.cap[l] = L.cap[l];
  }
  as.ring_n = (const int*)ring_n.p; as.rows = rows; as.cols = cols; as.out_n = L.out_n; as.overflow = L.overflow;
  CM_LAUNCH(sr_assemble_kernel, grid, 256, 0, stream, as);
}

}  // namespace cm
