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
#define SR_P1REGS 12       // pass 1 keeps regions of up to 32 * SR_P1REGS cells in registers
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
#define SR_WARPS (SR_THREADS / 32)
#define SR_MAXCH (SR_MAXCOLS / SR_THREADS)   // chunks of SR_THREADS items a ring can have at most

// Ordered compaction offsets for one or two 0/1 flags per item, items e = k * SR_THREADS + tid (chunk-major order).
// Phase 1: vote(k, f0, f1) for every chunk k (a ballot per flag, one store per warp), then ONE __syncthreads(); phase 2:
// pos(k, f0, f1) returns the exclusive prefixes packed as lo | hi << 16 (a ring has < 65536 items) and advances the running
// total.  Compared with a shuffle scan per chunk this is one barrier per scan instead of one per chunk and ~3x fewer
// instructions.  `wt` alternates between two buffers (the caller flips `buf` between scans) so that the next scan's votes
// cannot overtake slow readers of this one.
struct FlagScan {
  int* wt;          // [SR_MAXCH][SR_WARPS]
  int run;          // packed totals of the chunks already passed in phase 2
  int lane, warp;
  __device__ __forceinline__ void begin(int* buffers, int& buf) {
    wt = buffers + buf * (SR_MAXCH * SR_WARPS); buf ^= 1; run = 0;
    lane = threadIdx.x & 31; warp = threadIdx.x >> 5;
  }
  __device__ __forceinline__ void vote(int k, bool f0, bool f1 = false) {
    const unsigned int b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
    if (lane == 0) wt[k * SR_WARPS + warp] = __popc(b0) | (__popc(b1) << 16);
  }
  __device__ __forceinline__ int pos(int k, bool f0, bool f1 = false) {
    const int v = lane < SR_WARPS ? wt[k * SR_WARPS + lane] : 0;
    const int before = __reduce_add_sync(0xffffffffu, lane < warp ? v : 0);
    const int tot = __reduce_add_sync(0xffffffffu, v);
    const unsigned int b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
    const unsigned int lt = (1u << lane) - 1u;
    const int p = run + before + (__popc(b0 & lt) | (__popc(b1 & lt) << 16));
    run += tot;
    return p;
  }
};

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
  const int* ring_count;     // [S][rows] valid points per ring; NULL when no output carries absolute cloud indices
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

// mean and covariance of one window of pointClassify (ScanRegistration.cpp:557-581 / 603-627); the window of position c is
// {c, c-1, .., c-R}, summed in that order
__device__ __forceinline__ void window_cov(const float* px, const float* py, const float* pz, int c, int R, float& cx, float& cy, float& cz,
                                           float A[6]) {
  cx = 0.f; cy = 0.f; cz = 0.f;
  for (int t = 0; t <= R; t++) { int id = c - t; cx += px[id]; cy += py[id]; cz += pz[id]; }
  float cnt = (float)(R + 1);
  cx /= cnt; cy /= cnt; cz /= cnt;
  float a00 = 0.f, a10 = 0.f, a20 = 0.f, a11 = 0.f, a21 = 0.f, a22 = 0.f;
  for (int t = 0; t <= R; t++) {
    int id = c - t;
    float ax = px[id] - cx, ay = py[id] - cy, az = pz[id] - cz;
    a00 += ax * ax; a10 += ax * ay; a20 += ax * az; a11 += ay * ay; a21 += ay * az; a22 += az * az;
  }
  A[0] = a00 / cnt; A[1] = a10 / cnt; A[2] = a20 / cnt; A[3] = a11 / cnt; A[4] = a21 / cnt; A[5] = a22 / cnt;
}

// true: the window provably FAILS the line test  l2 > 100 l1 && l2 > 10000 l0  (ScanRegistration.cpp:587-588 / 633-634), so its
// eigen-solve can be skipped.  With m2 = l0 l1 + l0 l2 + l1 l2 (sum of the principal 2x2 minors) and tr = l0 + l1 + l2: a
// window that passes has l1 < l2 / 100 and l0 < l2 / 10000, hence m2 < 0.010102 l2^2 <= 0.010102 tr^2.  The bound used, 0.0105,
// leaves 4 %: the float rounding of m2 and tr (~1e-6 tr^2) and the backward error of the iterative solver that makes the real
// decision (|l_computed - l| <~ 30 eps l2, i.e. 1.8 % of l2 / 10000) are far inside it.  NaN / zero windows fall through to the solver.
__device__ __forceinline__ bool window_not_line(const float A[6]) {
  const float tr = A[0] + A[3] + A[5];
  const float m2 = (A[0] * A[3] - A[1] * A[1]) + (A[0] * A[5] - A[2] * A[2]) + (A[3] * A[5] - A[4] * A[4]);
  return m2 > 0.0105f * (tr * tr);
}

// the rest of the window test once the covariance is known: eigen-solve, ratio test, 0.08 m line test (:582-602 / 628-649)
__device__ __forceinline__ bool window_line(const float* px, const float* py, const float* pz, int c, int R, float cx, float cy, float cz,
                                            const float A[6], float v[3]) {
  float w[3], V[9];
  eig3_sym(A, w, V);
  if (w[2] > 100.f * w[1] && w[2] > 10000.f * w[0]) {
    v[0] = V[2]; v[1] = V[5]; v[2] = V[8];
    float vnorm = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int t = 0; t <= R; t++) {
      int id = c - t;
      float ax = px[id] - cx, ay = py[id] - cy, az = pz[id] - cz;
      float kx = ay * v[2] - az * v[1], ky = az * v[0] - ax * v[2], kz = ax * v[1] - ay * v[0];
      float distance = sqrtf(kx * kx + ky * ky + kz * kz) / vnorm;
      if (fabs((double)distance) > 0.08) return false;
    }
    return true;
  }
  return false;
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

// ---- in-CTA bitonic sort of P keys (P a power of two, P / E threads hold E consecutive keys each in registers) ------------
// Compare-exchange partners inside a thread cost a few selects, partners in the same warp one shuffle per key, only the
// partners in another warp (strides >= 32 E) go through shared memory with barriers: 10 of the 66 stages at P = 2048.
__device__ __forceinline__ unsigned int shfl_xor_key(unsigned int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ unsigned long long shfl_xor_key(unsigned long long v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

template <typename K, int E>
__device__ __forceinline__ void bitonic_sort_regs(K* key, int P) {
  const int tid = threadIdx.x;
  const int e0 = tid * E;
  const bool act = e0 < P;
  K v[E];
#pragma unroll
  for (int j = 0; j < E; j++) v[j] = act ? key[e0 + j] : (K)~(K)0;
  for (int kk = 2; kk <= P; kk <<= 1) {
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      if (jj >= 32 * E) {                    // partner in another warp
        __syncthreads();
        if (act) {
#pragma unroll
          for (int j = 0; j < E; j++) key[e0 + j] = v[j];
        }
        __syncthreads();
        if (act) {
#pragma unroll
          for (int j = 0; j < E; j++) {
            const int e = e0 + j;
            const K o = key[e ^ jj];
            const bool keepmin = ((e & jj) == 0) == ((e & kk) == 0);
            const bool oless = o < v[j];
            v[j] = (keepmin == oless) ? o : v[j];
          }
        }
      } else if (jj >= E) {                  // partner in another lane of this warp (every lane executes the shuffle)
        const int lm = jj / E;
#pragma unroll
        for (int j = 0; j < E; j++) {
          const int e = e0 + j;
          const K o = shfl_xor_key(v[j], lm);
          const bool keepmin = ((e & jj) == 0) == ((e & kk) == 0);
          const bool oless = o < v[j];
          v[j] = (keepmin == oless) ? o : v[j];
        }
      } else {                               // partner inside this thread: jj in {1, 2} (E <= 4), compile-time register indices
        const bool up = (e0 & kk) == 0;      // e0 .. e0 + E - 1 share every bit >= E, and kk > jj
        if (E >= 2 && jj == 1) {
#pragma unroll
          for (int j = 0; j + 1 < E; j += 2) {
            const bool up_j = (kk == 2 && E == 4) ? (((e0 + j) & kk) == 0) : up;   // kk = 2 < E = 4: the direction alternates inside the thread
            const K a = v[j], b = v[j + 1];
            if ((a > b) == up_j) { v[j] = b; v[j + 1] = a; }
          }
        } else if (E >= 4 && jj == 2) {
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const K a = v[j], b = v[j + 2];
            if ((a > b) == up) { v[j] = b; v[j + 2] = a; }
          }
        }
      }
    }
  }
  __syncthreads();
  if (act) {
#pragma unroll
    for (int j = 0; j < E; j++) key[e0 + j] = v[j];
  }
  __syncthreads();
}

// shared-memory-only version for sizes the register version does not cover (P > 4 * SR_THREADS)
template <typename K>
__device__ __forceinline__ void bitonic_sort_smem(K* key, int P) {
  __syncthreads();
  for (int kk = 2; kk <= P; kk <<= 1)
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      for (int pr = threadIdx.x; pr < (P >> 1); pr += SR_THREADS) {
        const int i = ((pr & ~(jj - 1)) << 1) | (pr & (jj - 1)), ixj = i | jj;
        const K x = key[i], y = key[ixj];
        const bool up = (i & kk) == 0;
        if ((x > y) == up) { key[i] = y; key[ixj] = x; }
      }
      __syncthreads();
    }
}

template <typename K>
__device__ __forceinline__ void bitonic_sort(K* key, int P) {
  if (P <= 1) { __syncthreads(); return; }
  if (P <= SR_THREADS) bitonic_sort_regs<K, 1>(key, P);
  else if (P == 2 * SR_THREADS) bitonic_sort_regs<K, 2>(key, P);
  else if (P == 4 * SR_THREADS) bitonic_sort_regs<K, 4>(key, P);
  else bitonic_sort_smem<K>(key, P);
}

// smallest power of two >= v (v >= 1)
__device__ __forceinline__ int pow2_ge(int v) { return v <= 1 ? 1 : 1 << (32 - __clz(v - 1)); }

// ------------------------------------------------------------------------------------------------------------
// kernel 2: everything for one ring
// ------------------------------------------------------------------------------------------------------------
// dynamic shared memory layout, `cap` = cols rounded up to a multiple of 8, P2 = next power of two >= cols
//   float px[cap], py[cap], pz[cap], pi[cap], curv[cap]
//   u16   col[cap]
//   s8    state[cap], snap[cap], ev[cap], lab[cap]
//   ---- scratch (first the staging area of the raw ring row: float4 row[cols], filled by ONE bulk copy) ----
//   u64   key[KEYN]                      (rank keys / prefix arrays / voxel sort)
//   u16   lst[4][cap], nf[cap], ord[cap]
extern __shared__ __align__(16) unsigned char sr_smem[];

__device__ __forceinline__ unsigned int smem_addr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(SR_THREADS) sr_ring_kernel(ScanRegArgs a) {
  const int ring = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const int rows = a.rows, cols = a.cols;
  const ScanRegParamsDev& prm = a.prm;
  const int R = prm.R;
  const int cap = (cols + 7) & ~7;
  const int P2 = pow2_ge(cols);

  float* px = reinterpret_cast<float*>(sr_smem);
  float* py = px + cap; float* pz = py + cap; float* pin = pz + cap; float* curv = pin + cap;
  unsigned short* colv = reinterpret_cast<unsigned short*>(curv + cap);
  signed char* state = reinterpret_cast<signed char*>(colv + cap);
  signed char* snap = state + cap; signed char* ev = snap + cap; signed char* lab = ev + cap;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(lab + cap);      // 26 * cap bytes in: a multiple of 16
  const int KEYN = (8 * P2 >= 10 * cap) ? P2 : (10 * cap + 7) / 8;   // `key` doubles as five u16 prefix arrays
  unsigned short* lst0 = reinterpret_cast<unsigned short*>(key + KEYN);
  unsigned short* lst[4] = {lst0, lst0 + cap, lst0 + 2 * cap, lst0 + 3 * cap};
  unsigned short* nfl = lst0 + 4 * cap;
  unsigned short* ord = lst0 + 5 * cap;
  const float4* stage = reinterpret_cast<const float4*>(key);         // 16 * cols bytes <= 8 * KEYN + 12 * cap
  __shared__ int s_wt[2 * SR_MAXCH * SR_WARPS];
  __shared__ int s_wt2[SR_MAXCH * SR_WARPS];   // a third flag scanned in the same barrier phase (pass 3)
  __shared__ __align__(8) unsigned long long s_mbar;
  int fs_buf = 0;   // uniform over the CTA: every thread makes the same sequence of scans
  FlagScan fs;
  __shared__ int s_cnt[5];          // list lengths: sharp, lessSharp, flat, lessFlatRaw, (spare)
  __shared__ int s_misc[8];
  const int nch_cols = (cols + SR_THREADS - 1) / SR_THREADS;

  // ---- the ring row -> shared memory with one bulk copy (TMA, 16 * cols bytes), completion on an mbarrier -------------------
  const float4* src = a.frames + ((size_t)s * rows + ring) * cols;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned int bytes = (unsigned int)cols * 16u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(stage)), "l"(src), "r"(bytes), "r"(smem_addr(&s_mbar)) : "memory");
  }
  // ---- ring ranges (_scanIndices): only the optional outputs carry absolute cloud indices ---------------------------------
  int S0 = 0;
  if (a.ring_count) { const int* rc = a.ring_count + s * rows; for (int r = 0; r < ring; r++) S0 += rc[r]; }
  int* ring_n = a.ring_n + (s * rows + ring) * 5;
  if (tid < 5) s_cnt[tid] = 0;
  {
    unsigned int done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_addr(&s_mbar)) : "memory");
    }
  }

  // ---- ordered compaction of the ring (OrganizedScanRegistration.cpp:102-126) ---------------------------------
  int n;
  {
    unsigned int okmask = 0;
    fs.begin(s_wt, fs_buf);
    for (int k = 0; k < nch_cols; k++) {
      const int c = k * SR_THREADS + tid;
      const bool ok = c < cols && point_valid(stage[c], prm.blind_sq);
      okmask |= (ok ? 1u : 0u) << k;
      fs.vote(k, ok);
    }
    __syncthreads();
    for (int k = 0; k < nch_cols; k++) {
      const int c = k * SR_THREADS + tid;
      const bool ok = (okmask >> k) & 1u;
      const int d = fs.pos(k, ok) & 0xFFFF;
      if (ok) { const float4 p = stage[c]; px[d] = p.x; py[d] = p.y; pz[d] = p.z; pin[d] = p.w; colv[d] = (unsigned short)c; }
    }
    n = fs.run & 0xFFFF;
  }
  __syncthreads();   // the staging area is free from here on
  const int E0 = (S0 + n > 0) ? S0 + n - 1 : 0;   // range.second = cloudSize > 0 ? cloudSize - 1 : 0
  if (tid == 0 && a.scan_range) { a.scan_range[(s * rows + ring) * 2] = S0; a.scan_range[(s * rows + ring) * 2 + 1] = E0; }
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
  const int nch = (n + SR_THREADS - 1) / SR_THREADS;

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

  // ---- region bounds (ScanRegistration.cpp:249-262), size_t integer arithmetic on ABSOLUTE indices (the bounds relative to
  //      the ring do not depend on S0: S0 * NR is a multiple of NR and drops out of both floor divisions) ------------------
  __shared__ int reg_sp[SR_MAXREG], reg_ep[SR_MAXREG];   // relative to the ring; skipped regions: sp = INT_MAX, ep = -1
  if (tid < prm.nregions) {
    unsigned long long s5 = (unsigned long long)S0 + R, e5 = (unsigned long long)E0 - R, NR = prm.nregions, j = tid;
    unsigned long long sp = (s5 * (NR - j) + e5 * j) / NR;
    unsigned long long ep = (s5 * (NR - 1 - j) + e5 * (j + 1)) / NR - 1;
    if (ep <= sp) { reg_sp[tid] = 0x7fffffff; reg_ep[tid] = -1; }
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
  __shared__ int s_nw, s_ns;                   // classification work lists: needed windows, windows that need the eigen-solve
  const int NR = prm.nregions;
  // region of a cell (regions are contiguous and ordered): the last region that starts at or before the cell
  for (int c = tid; c < n; c += SR_THREADS) {   // ev[] is free after the mask replay: region id of every cell (-1: none)
    int rj = -1;
    for (int j = 0; j < NR; j++) if (c >= reg_sp[j]) rj = j;
    if (rj >= 0 && c > reg_ep[rj]) rj = -1;
    ev[c] = (signed char)rj;
  }
  if (tid < 4 * SR_MAXREG) { (&cnt2[0][0])[tid] = 0; (&cnt3[0][0])[tid] = 0; }
  if (tid == 0) { s_nw = 0; s_ns = 0; }
  __syncthreads();
  auto region_of = [&](int c) -> int { return (int)ev[c]; };
  // window results of pointClassify live in the scratch area until the index lists are built
  float2* wxy = reinterpret_cast<float2*>(key);            // direction of a line window (x, y)
  float* wz = reinterpret_cast<float*>(lst[0]);            //   and z                                   (lst[0..1])
  signed char* wfl = reinterpret_cast<signed char*>(lst[2]);   // window result: 1 line, 3 no line      (first half of lst[2])
  signed char* wcand = wfl + cap;                               // 1: the cell is a pass-3 candidate    (second half of lst[2])
  unsigned short* wlist = lst[3];                               // windows some candidate needs
  // windows whose eigen-solve cannot be skipped: the tail of `key` behind wxy (KEYN * 8 >= 10 * cap bytes, wxy takes 8 * cap)
  unsigned short* slist = reinterpret_cast<unsigned short*>(key) + 4 * cap;
  unsigned int* ckey = reinterpret_cast<unsigned int*>(key);    // curvature bits of the candidates (rank sort, before wxy is written)
  // ---- cells above the curvature threshold, in index order (they are the pass-3 candidates, :305-314) -------------
  {
    unsigned int m = 0;
    fs.begin(s_wt, fs_buf);
    for (int k = 0; k < nch; k++) {
      const int c = k * SR_THREADS + tid;
      const int rj = c < n ? region_of(c) : -1;
      const bool isnf = rj >= 0 && !(curv[c] < prm.curv_thr);
      m |= (isnf ? 1u : 0u) << k;
      if (c < n) wcand[c] = isnf ? 1 : 0;
      fs.vote(k, isnf);
    }
    __syncthreads();
    for (int k = 0; k < nch; k++) {
      const int c = k * SR_THREADS + tid;
      const bool isnf = (m >> k) & 1u;
      const int pos = fs.pos(k, isnf) & 0xFFFF;
      if (isnf) nfl[pos] = (unsigned short)c;
      const int rj = c < n ? region_of(c) : -1;
      if (rj >= 0 && c == reg_sp[rj]) nf_begin[rj] = pos;
    }
    if (tid == 0) nf_begin[NR] = fs.run & 0xFFFF;
    // the windows pointClassify needs: window w = {w, .., w-R} is the backward window of candidate w and the forward window
    // of candidate w - R (ScanRegistration.cpp:557-649: the second one is summed in the same order), so it is evaluated once
    m = 0;
    fs.begin(s_wt, fs_buf);
    for (int k = 0; k < nch; k++) {
      const int w = k * SR_THREADS + tid;
      const bool need = w < n && (wcand[w] != 0 || (w >= R && wcand[w - R] != 0));
      m |= (need ? 1u : 0u) << k;
      fs.vote(k, need);
    }
    __syncthreads();
    for (int k = 0; k < nch; k++) {
      const int w = k * SR_THREADS + tid;
      const bool need = (m >> k) & 1u;
      const int pos = fs.pos(k, need) & 0xFFFF;
      if (need) wlist[pos] = (unsigned short)w;
    }
    if (tid == 0) {
      s_nw = fs.run & 0xFFFF;
      for (int j = NR - 1; j >= 0; j--) if (reg_ep[j] < 0) nf_begin[j] = nf_begin[j + 1];   // skipped regions are empty
    }
  }
  __syncthreads();
  const int m_all = nf_begin[NR];
  const int n_win = s_nw;
  // ---- warp 0: pass 1, chained over the regions (greedy flat picks with +-R suppression, :267-284, 524-545);
  //      warps 1..15: pointClassify of every pass-3 candidate (:316, 547-666) -- independent of the picks ------------
  if (tid < 32) {
    const int lane = tid;
    int np1 = 0;
    for (int j = 0; j < NR; j++) {
      const int sp = reg_sp[j], ep = reg_ep[j];
      if (lane == 0) p1_begin[j] = np1;
      if (ep >= 0 && ep - sp + 1 <= 32 * SR_P1REGS) {
        // the region's candidates live in registers: cell sp + lane + 32 r in kreg[r], key = curvature bits (curvatures are >= 0, so
        // the bit pattern orders like the value), 0xFFFFFFFF = not a candidate.  One pick = a local arg-min, two warp reductions
        // (smallest bits, then the smallest cell among equal bits: the (curvature, index) order of the stable sort) and the +-R
        // suppression applied to the registers
        unsigned int kreg[SR_P1REGS];
#pragma unroll
        for (int r = 0; r < SR_P1REGS; r++) {
          const int c = sp + lane + 32 * r;
          unsigned int kv = 0xFFFFFFFFu;
          if (c <= ep) { const float cv = curv[c]; if (state[c] != P_SURF_PICKED_NEAR && cv < prm.curv_thr) kv = __float_as_uint(cv); }
          kreg[r] = kv;
        }
        __syncwarp();   // every lane has read state[] before any lane marks a pick (the reductions order execution, this orders memory)
        for (int k = 0; k < prm.max_flat; k++) {
          unsigned int m = kreg[0]; int mr = 0;
#pragma unroll
          for (int r = 1; r < SR_P1REGS; r++) if (kreg[r] < m) { m = kreg[r]; mr = r; }
          const unsigned int mb = __reduce_min_sync(0xffffffffu, m);
          if (mb == 0xFFFFFFFFu) break;
          const int c = (int)__reduce_min_sync(0xffffffffu, m == mb ? (unsigned int)(sp + lane + 32 * mr) : 0x7fffffffu);
#pragma unroll
          for (int r = 0; r < SR_P1REGS; r++) { const int d = sp + lane + 32 * r - c; if (d >= -R && d <= R) kreg[r] = 0xFFFFFFFFu; }
          if (lane <= 2 * R) state[c - R + lane] = P_SURF_PICKED_NEAR;   // markAsPicked: c-R .. c+R
          if (lane == 0 && np1 < SR_MAXREG * 8) p1buf[np1] = (unsigned short)c;
          np1++;
        }
        __syncwarp();
        for (int c = sp + lane; c <= ep; c += 32) snap[c] = state[c];   // what passes 2 and 3 of region j see
        __syncwarp();
      } else if (ep >= 0) {
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
    const int NT = SR_THREADS - 32, t = tid - 32;
    // ---- descending (curvature, index) order inside every region: rank by counting.  The candidate list is in index order, so
    //      an equal curvature later in the list wins the tie: two loops of 32-bit compares on the curvature bits.  Independent
    //      of the picks and of the labels: done here, while warp 0 walks its chain ------------------------------------------------
    for (int i = t; i < m_all; i += NT) ckey[i] = __float_as_uint(curv[nfl[i]]);   // curvatures are >= 0: bit order = value order
    asm volatile("bar.sync 1, %0;" ::"n"(SR_THREADS - 32));
    for (int i = t; i < m_all; i += NT) {
      const int c = nfl[i];
      const int rj = region_of(c);
      const unsigned int ki = ckey[i];
      const int b0 = nf_begin[rj], b1 = nf_begin[rj + 1];
      int rank = 0;
      for (int k = b0; k < i; k++) rank += (ckey[k] > ki) ? 1 : 0;
      for (int k = i + 1; k < b1; k++) rank += (ckey[k] >= ki) ? 1 : 0;
      ord[b0 + rank] = (unsigned short)c;
    }
    // phase A (writes wfl and slist only; wxy, which shares the bytes of ckey, is written after the next barrier):
    // covariance of every needed window + the cheap exact test that rules most of them out (window_not_line)
    for (int i = t; i < n_win; i += NT) {
      const int w = wlist[i];
      float cx, cy, cz, A[6];
      window_cov(px, py, pz, w, R, cx, cy, cz, A);
      if (window_not_line(A)) wfl[w] = 3;
      else slist[atomicAdd(&s_ns, 1)] = (unsigned short)w;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SR_THREADS - 32));
    // phase B: the windows that may be lines, densely packed over the lanes: eigen-solve + 0.08 m line test
    const int n_sv = s_ns;
    for (int i = t; i < n_sv; i += NT) {
      const int w = slist[i];
      float cx, cy, cz, A[6], v[3];
      window_cov(px, py, pz, w, R, cx, cy, cz, A);
      const bool l = window_line(px, py, pz, w, R, cx, cy, cz, A, v);
      if (l) { wxy[w] = make_float2(v[0], v[1]); wz[w] = v[2]; }
      wfl[w] = l ? 1 : 3;
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
  // `key` is free again: carve four u16 prefix arrays out of it
  unsigned short* pfa = reinterpret_cast<unsigned short*>(key);
  unsigned short* pfb = pfa + cap; unsigned short* pfc = pfb + cap;
  // ---- pass 2 (:286-303), whole ring: c < thr -> lessFlatRaw; EDGE_BROKEN (as seen after the region's picks) -> sharp
  //      and lessSharp.  Ring-wide exclusive prefixes; the value at a region's first cell is that region's offset. ------
  {
    unsigned int m = 0;
    fs.begin(s_wt, fs_buf);
    for (int k = 0; k < nch; k++) {
      const int c = k * SR_THREADS + tid;
      const int rj = c < n ? region_of(c) : -1;
      const bool isflat = rj >= 0 && (curv[c] < prm.curv_thr);
      const bool isedge = rj >= 0 && (snap[c] == P_EDGE_BROKEN);
      m |= ((isflat ? 1u : 0u) | (isedge ? 0x10000u : 0u)) << k;
      fs.vote(k, isflat, isedge);
      if (isflat) atomicAdd(&cnt2[3][rj], 1);
      if (isedge) { atomicAdd(&cnt2[0][rj], 1); atomicAdd(&cnt2[1][rj], 1); }
    }
    __syncthreads();
    for (int k = 0; k < nch; k++) {
      const int c = k * SR_THREADS + tid;
      const int p12 = fs.pos(k, (m >> k) & 1u, (m >> (16 + k)) & 1u);
      const int p1 = p12 & 0xFFFF, p2 = p12 >> 16;
      if (c < n) { pfa[c] = (unsigned short)p1; pfb[c] = (unsigned short)p2; }
      const int rj = c < n ? region_of(c) : -1;
      if (rj >= 0 && c == reg_sp[rj]) { start2[0][rj] = p1; start2[1][rj] = p2; }
    }
  }
  // ---- pass 3 (:305-354) over the descending order of every region: emission counters become prefix sums -----------
  // flat emissions: ONESIDE_FLAT while the shared counter (bumped by SURFACE_FLAT too) is below the cap, i.e. the ONESIDE_FLAT
  // elements among the first max_flat surf elements of the region -- decided per element from two prefixes, no serial walk
  unsigned short* qfa = pfc; unsigned short* qfb = pfc + cap; unsigned short* qfc = pfc + 2 * cap;   // corner / surf / one-sided prefixes along `ord`
  {
    const int nchm = (m_all + SR_THREADS - 1) / SR_THREADS;
    unsigned int m = 0, m3 = 0;
    FlagScan fs2; int b2 = 0;
    fs.begin(s_wt, fs_buf);
    fs2.begin(s_wt2, b2);
    for (int k = 0; k < nchm; k++) {
      const int i = k * SR_THREADS + tid;
      const bool in = i < m_all;
      const int c = in ? ord[i] : 0;
      const int l = in ? lab[c] : L_MESSY;
      const int rj = in ? region_of(c) : -1;
      const bool is_corner = in && l == L_CORNER_SHARP && snap[c] > P_EDGE_BROKEN;
      const bool is_surf = in && (l == L_SURFACE_FLAT || l == L_ONESIDE_FLAT);
      const bool is_one = in && l == L_ONESIDE_FLAT;
      m |= ((is_corner ? 1u : 0u) | (is_surf ? 0x10000u : 0u)) << k;
      m3 |= (is_one ? 1u : 0u) << k;
      fs.vote(k, is_corner, is_surf);
      fs2.vote(k, is_one);
      if (is_corner) atomicAdd(&cnt3[1][rj], 1);
      if (is_surf) atomicAdd(&cnt3[3][rj], 1);
    }
    __syncthreads();
    for (int k = 0; k < nchm; k++) {
      const int i = k * SR_THREADS + tid;
      const bool in = i < m_all;
      const int p12 = fs.pos(k, (m >> k) & 1u, (m >> (16 + k)) & 1u);
      const int p3 = fs2.pos(k, (m3 >> k) & 1u) & 0xFFFF;
      const int p1 = p12 & 0xFFFF, p2 = p12 >> 16;
      if (in) {
        qfa[i] = (unsigned short)p1; qfb[i] = (unsigned short)p2; qfc[i] = (unsigned short)p3;
        const int rj = region_of(ord[i]);
        if (i == nf_begin[rj]) { start3[0][rj] = p1; start3[1][rj] = p2; start3[2][rj] = p3; }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < m_all; i += SR_THREADS) {   // ONESIDE_FLAT among the first max_flat surf elements of its region
    const int c = ord[i];
    if (lab[c] == L_ONESIDE_FLAT) {
      const int rj = region_of(c);
      if ((int)qfb[i] - start3[1][rj] < prm.max_flat) atomicAdd(&cnt3[2][rj], 1);
    }
  }
  if (tid < NR) {
    const int j = tid;
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
  for (int i = tid; i < m_all; i += SR_THREADS) {   // flat list, pass-3 part: after the region's pass-1 picks, in `ord` order
    const int c = ord[i];
    if (lab[c] == L_ONESIDE_FLAT) {
      const int rj = region_of(c);
      if ((int)qfb[i] - start3[1][rj] < prm.max_flat) lst[2][base[2][rj] + cnt2[2][rj] + (int)qfc[i] - start3[2][rj]] = (unsigned short)c;
    }
  }
  if (tid < NR) {   // flat list, pass-1 picks (at most max_flat per region)
    const int j = tid;
    int o = base[2][j];
    for (int i = p1_begin[j]; i < p1_begin[j + 1]; i++) lst[2][o++] = p1buf[i];
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
  if (a.picked || a.curvature || a.label)
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
    float* fs_ = reinterpret_cast<float*>(key);   // 6 * 16 floats of scratch
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
        mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
      }
    __syncthreads();
    if ((tid & 31) == 0) for (int k = 0; k < 3; k++) { fs_[(tid >> 5) * 6 + k] = mn[k]; fs_[(tid >> 5) * 6 + 3 + k] = mx[k]; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < SR_THREADS / 32; w++)
        for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], fs_[w * 6 + k]); mx[k] = fmaxf(mx[k], fs_[w * 6 + 3 + k]); }
      const float inv = 1.0f / prm.less_flat_leaf;
      vb_i[6] = 0; vb_i[7] = 32;
      if (nlf > 0) {
        long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1,
                  dz = (long long)((mx[2] - mn[2]) * inv) + 1;
        if (dx * dy * dz > 2147483647LL) vb_i[6] = 1;   // passthrough
        int minb[3], maxb[3];
        for (int k = 0; k < 3; k++) { minb[k] = (int)floorf(mn[k] * inv); maxb[k] = (int)floorf(mx[k] * inv); vb_i[k] = minb[k]; }
        int d0 = maxb[0] - minb[0] + 1, d1 = maxb[1] - minb[1] + 1;
        vb_i[3] = d0; vb_i[4] = d0 * d1;
        // bits of the largest voxel index (passthrough: of the largest list position)
        const long long cells = vb_i[6] ? (long long)nlf : (long long)d0 * d1 * (maxb[2] - minb[2] + 1);
        int b = 1; while (b < 32 && (1LL << b) < cells) b++;
        vb_i[7] = (cells > 0 && cells <= 0x7fffffffLL) ? b : 32;
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
    const int nchl = (nlf + SR_THREADS - 1) / SR_THREADS;
    int nruns;
    {
      unsigned int m = 0;
      fs.begin(s_wt, fs_buf);
      for (int k = 0; k < nchl; k++) {
        const int i = k * SR_THREADS + tid;
        const bool rs = i < nlf && (i == 0 || vidx[i] != vidx[i - 1]);
        m |= (rs ? 1u : 0u) << k;
        fs.vote(k, rs);
      }
      __syncthreads();
      for (int k = 0; k < nchl; k++) {
        const int i = k * SR_THREADS + tid;
        const bool rs = (m >> k) & 1u;
        const int pos = fs.pos(k, rs) & 0xFFFF;
        if (rs) run_start[pos] = (unsigned short)i;
      }
      nruns = fs.run & 0xFFFF;
    }
    __syncthreads();
    const int P = pow2_ge(nruns);
    // 32-bit sort keys (voxel << posbits | start) whenever the voxel index and the list position fit together
    const int posbits = 32 - __clz(nlf > 1 ? nlf - 1 : 1);
    const bool narrow = vb_i[7] + posbits <= 31;   // a real key never has bit 31 set, so it cannot equal the padding key
    unsigned int* key32 = reinterpret_cast<unsigned int*>(key) + P;   // upper half of key[0..P): expanded to 64 bits afterwards
    for (int r = tid; r < P; r += SR_THREADS) {
      unsigned long long k = 0xFFFFFFFFFFFFFFFFull;
      unsigned int k32 = 0xFFFFFFFFu;
      if (r < nruns) {
        const int st = run_start[r];
        run_end[st] = (unsigned short)(r + 1 < nruns ? run_start[r + 1] : nlf);
        k = ((unsigned long long)vidx[st] << 32) | (unsigned int)st;
        k32 = (vidx[st] << posbits) | (unsigned int)st;
      }
      if (narrow) key32[r] = k32; else key[r] = k;
    }
    __syncthreads();
    if (narrow) {
      bitonic_sort<unsigned int>(key32, P);
      // expand in place: every thread reads its keys, barrier, writes the 64-bit form the rest of the filter uses
      unsigned int mine[SR_MAXCH];
      const unsigned int posmask = (1u << posbits) - 1u;
#pragma unroll
      for (int q = 0; q < SR_MAXCH; q++) { const int r = q * SR_THREADS + tid; mine[q] = r < P ? key32[r] : 0xFFFFFFFFu; }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < SR_MAXCH; q++) {
        const int r = q * SR_THREADS + tid;
        if (r < P) key[r] = mine[q] == 0xFFFFFFFFu ? 0xFFFFFFFFFFFFFFFFull : (((unsigned long long)(mine[q] >> posbits) << 32) | (mine[q] & posmask));
      }
      __syncthreads();
    } else {
      bitonic_sort<unsigned long long>(key, P);
    }
    // voxel heads among the sorted runs -> output rank -> centroid (runs in order, entries of a run in order)
    const int nchr = (nruns + SR_THREADS - 1) / SR_THREADS;
    int nvox;
    {
      unsigned int m = 0;
      fs.begin(s_wt, fs_buf);
      for (int k = 0; k < nchr; k++) {
        const int r = k * SR_THREADS + tid;
        const bool head = r < nruns && (r == 0 || (key[r] >> 32) != (key[r - 1] >> 32));
        m |= (head ? 1u : 0u) << k;
        fs.vote(k, head);
      }
      __syncthreads();
      for (int k = 0; k < nchr; k++) {
        const int r = k * SR_THREADS + tid;
        const bool head = (m >> k) & 1u;
        const int pos = fs.pos(k, head) & 0xFFFF;
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
          o_pts[3][pos] = make_float4(cx / cnt, cy / cnt, cz / cnt, ci / cnt);
        }
      }
      nvox = fs.run & 0xFFFF;
    }
    if (tid == 0) { ring_n[0] = s_cnt[0]; ring_n[1] = s_cnt[1]; ring_n[2] = s_cnt[2]; ring_n[3] = nvox; ring_n[4] = nlf; }
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
  int cap = (cols + 7) & ~7;
  int P2 = 1; while (P2 < cols) P2 <<= 1;
  int keyn = (8 * P2 >= 10 * cap) ? P2 : (10 * cap + 7) / 8;
  return (size_t)cap * 4 * 5 + (size_t)cap * 2 + (size_t)cap * 4 + (size_t)keyn * 8 + (size_t)cap * 2 * 6;
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
  // ring offsets inside the concatenated cloud are only needed by outputs that carry absolute cloud indices (diagnostics / parity):
  // the feature clouds themselves do not depend on them, so the pipeline skips this pass over the sweep
  const bool want_abs = L.want_idx || L.cloud || L.picked || L.curvature || L.label || L.scan_range;
  if (want_abs) CM_LAUNCH(sr_count_kernel, grid, 256, 0, stream, L.frames, rows, cols, p.blind_sq, (int*)ring_count.p);
  ScanRegArgs a;
  a.frames = L.frames; a.tags = L.tags; a.rows = rows; a.cols = cols; a.ring_count = want_abs ? (const int*)ring_count.p : nullptr; a.prm = p;
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
