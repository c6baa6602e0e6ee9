// cm_odom.inl -- K8: scan-to-scan odometry correspondences (included by cm_match.cu; shares its helpers).
//
// Replaces the data-parallel inner loops of LaserOdometry::scanMatch (L_SLAM/src/odometry/LaserOdometry.cpp:355-497,
// 524-577), transformToStart / transformToEnd (:135-190) and the 4-argument coefficient overloads
// (util/feature_utils.h:28-61, 77-95).  The reduction and the 6x6 step are the mapping solver's kernels with the
// odometry's constants (b = -0.05 d, eigenvalue threshold 10, "< 10 rows -> skip the iteration", NaN guards).

// bounding box + ring range of 32 consecutive points of a last cloud (points [32 b, 32 b + 32)): lets the ring walks below skip the
// chunks that cannot hold a correspondent (all farther than the 5 m gate) without changing what the walk finds
struct ChunkBox { float mnx, mny, mnz; int ring_min; float mxx, mxy, mxz; int ring_max; };

struct OdomArgs {
  const ChunkBox* box_corner; const ChunkBox* box_surf;   // optional (NULL: every point of the walk is visited)
  const float4* sharp; const float4* flat; int n_sharp, n_flat;
  const float4* last_corner; const float4* last_surf; int bound_corner, bound_surf;   // scan bounds (clamped)
  GridView grid_corner, grid_surf;      // over the last clouds, pts[].w = original index
  const MatchState* state;
  int* ind;                             // [2 * n_sharp + 3 * n_flat]: corner {closest, second}, surf {closest, second, third}
  RowOut* rows;                         // [n_sharp + n_flat]
  int iter;
  int spread;                           // a warp takes 32 >> spread query slots (its first lanes): see launch_odom_corr_batch
};

// transformToStart (LaserOdometry.cpp:135-142): s = 10 * frac(intensity); po = T(_transform * s) * pi
__device__ __forceinline__ void odom_to_start(const float tf[6], const float4& p, float* x, float* y, float* z) {
  const float s = 10 * (p.w - (float)(int)p.w);
  float t[6];
#pragma unroll
  for (int k = 0; k < 6; k++) t[k] = tf[k] * s;   // Twist::operator*(scale), Twist.h:28-35
  float R[9];
  pose_to_matrix(t, R);
  transform_point(R, t + 3, p.x, p.y, p.z, x, y, z);
}

__device__ __forceinline__ float odom_sqdiff(const float4& a, float bx, float by, float bz) {   // calcSquaredDiff(a, pointSel)
  float dx = a.x - bx, dy = a.y - by, dz = a.z - bz;
  return dx * dx + dy * dy + dz * dz;
}

// one stream's correspondences and rows for one Gauss-Newton iteration (every thread of the CTA calls; uniform early exits only)
__device__ __forceinline__ void odom_corr_body(const OdomArgs& a, uint4* rng, PoseCoef& kc, float* tf) {
  const MatchState& st = *a.state;
  if (st.done) return;
  if (threadIdx.x == 0) make_pose_coef(st, kc);
  if (threadIdx.x < 6) tf[threadIdx.x] = st.pose[threadIdx.x];
  __syncthreads();
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int nT = ((a.n_sharp + 31) & ~31) + a.n_flat;
  const int per = 32 >> a.spread;
  if ((gt >> 5) * per >= nT) return;
  const int t = (gt & 31) < per ? (gt >> 5) * per + (gt & 31) : nT + 32;   // idle lanes of a thinned warp: an invalid slot
  bool isCorner; int src, row;
  const bool valid = decode_query(t, a.n_sharp, a.n_flat, &isCorner, &src, &row);
  if ((gt & 31) >= per) isCorner = (gt >> 5) * per < ((a.n_sharp + 31) & ~31);   // (the class of the warp's real slots)
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  float sx = 0.f, sy = 0.f, sz = 0.f;
  if (valid) { p = isCorner ? a.sharp[src] : a.flat[src]; odom_to_start(tf, p, &sx, &sy, &sz); }
  int* ind = isCorner ? a.ind + 2 * src : a.ind + 2 * a.n_sharp + 3 * src;
  if (a.iter % 5 == 0) {   // LaserOdometry.cpp:358,424: correspondences are refreshed every 5th iteration
    Top5 best;
    knn5_search<true, 0>(isCorner ? a.grid_corner : a.grid_surf, valid, sx, sy, sz, 25.0f, rng, best);   // only the nearest point is used
    // The second (and third) correspondent: the reference walks the last cloud from the closest point forwards, then backwards, until
    // the ring index leaves +-2.5 rings (LaserOdometry.cpp:363-398, 427-476) -- hundreds to thousands of points per query, one after
    // the other.  Here the WARP walks them for one query at a time, 32 consecutive points per round (one coalesced load): a ballot
    // finds the first point past the ring limit, every lane keeps the first minimum of the points it saw (strict <, in visit order),
    // and the lanes are merged by (distance, visit order) -- the point the sequential walk would have kept.
    int closest = -1, min2 = -1, min3 = -1;
    if (valid && best.d(0) < 25.f && best.slot[0] >= 0) closest = best.idx(0);
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned int work = __ballot_sync(FULL, closest >= 0);
    while (work) {
      const int h = __ffs(work) - 1;
      work &= work - 1;
      const int c0 = __shfl_sync(FULL, closest, h);
      const float bx = __shfl_sync(FULL, sx, h), by = __shfl_sync(FULL, sy, h), bz = __shfl_sync(FULL, sz, h);
      const float4* cloud = isCorner ? a.last_corner : a.last_surf;        // (a warp never mixes the two classes: decode_query)
      const int bound = isCorner ? a.bound_corner : a.bound_surf;
      const int scan = (int)cloud[c0].w;
      float d2 = 25.f, d3 = 25.f;
      unsigned int v2 = 0xFFFFFFFFu, v3 = 0xFFFFFFFFu;   // visit rank of the lane's candidate: forward 1, 2, ..; backward 0x40000000 + 1, 2, ..
      const ChunkBox* boxes = isCorner ? a.box_corner : a.box_surf;
      // one chunk of 32 consecutive points, lane order = visit order; returns the ballot of the points past the ring limit (the walk
      // ends at the first of them), the points before it are candidates
      auto visit = [&](int jbase, bool forward) -> unsigned int {
        const int j = forward ? jbase + lane : jbase - lane;
        const bool in = forward ? (j > c0 && j < bound) : (j >= 0 && j < c0);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in) q = cloud[j];
        const int ring = (int)q.w;
        const unsigned int brk = __ballot_sync(FULL, in && (forward ? (double)ring > (double)scan + 2.5 : (double)ring < (double)scan - 2.5));
        if (in && (brk == 0 || lane < __ffs(brk) - 1)) {
          const float d = odom_sqdiff(q, bx, by, bz);
          const unsigned int v = forward ? (unsigned int)(j - c0) : 0x40000000u + (unsigned int)(c0 - j);
          const bool same_side = forward ? ring <= scan : ring >= scan;   // surf: the second point, else the third (LaserOdometry.cpp:441-474)
          if (isCorner) { if ((forward ? ring > scan : ring < scan) && d < d2) { d2 = d; v2 = v; } }
          else if (same_side) { if (d < d2) { d2 = d; v2 = v; } }
          else { if (d < d3) { d3 = d; v3 = v; } }
        }
        return brk;
      };
      if (!boxes) {
        for (int j0 = c0 + 1; j0 < bound; j0 += 32) if (visit(j0, true)) break;
        for (int j0 = c0 - 1; j0 >= 0; j0 -= 32) if (visit(j0, false)) break;
      } else {
        // 32 chunk boxes per round: a chunk is visited when it is within the gate OR holds a point past the ring limit (the walk may
        // end inside it); a far chunk without such a point contributes nothing and is skipped.  (0.9999: the box distance is a float
        // lower bound of the point distances it stands for.)
        auto wanted = [&](int cb, bool have, bool forward) -> bool {
          if (!have) return false;
          const ChunkBox b = boxes[cb];
          const float ex = fmaxf(fmaxf(b.mnx - bx, bx - b.mxx), 0.f), ey = fmaxf(fmaxf(b.mny - by, by - b.mxy), 0.f), ez = fmaxf(fmaxf(b.mnz - bz, bz - b.mxz), 0.f);
          const bool near = (ex * ex + ey * ey + ez * ez) * 0.9999f < 25.f;
          const bool limit = forward ? b.ring_max > scan + 2 : b.ring_min < scan - 2;
          return near || limit;
        };
        bool stop = false;
        if (c0 + 1 < bound) {
          const int cbl = (bound - 1) >> 5;
          for (int cb0 = (c0 + 1) >> 5; cb0 <= cbl && !stop; cb0 += 32) {
            unsigned int todo = __ballot_sync(FULL, wanted(cb0 + lane, cb0 + lane <= cbl, true));
            while (todo) {
              const int kq = __ffs(todo) - 1;
              todo &= todo - 1;
              if (visit((cb0 + kq) << 5, true)) { stop = true; break; }
            }
          }
        }
        stop = false;
        if (c0 > 0) {
          for (int cb0 = (c0 - 1) >> 5; cb0 >= 0 && !stop; cb0 -= 32) {
            unsigned int todo = __ballot_sync(FULL, wanted(cb0 - lane, cb0 - lane >= 0, false));
            while (todo) {
              const int kq = __ffs(todo) - 1;
              todo &= todo - 1;
              if (visit(((cb0 - kq) << 5) + 31, false)) { stop = true; break; }
            }
          }
        }
      }
      // merge: smallest distance, earliest visit among equals (squared distances are >= 0: bit order = value order)
      int w2 = -1, w3 = -1;
      {
        const unsigned int db = v2 != 0xFFFFFFFFu ? __float_as_uint(d2) : 0xFFFFFFFFu;
        const unsigned int mind = __reduce_min_sync(FULL, db);
        const unsigned int minv = __reduce_min_sync(FULL, (db == mind) ? v2 : 0xFFFFFFFFu);
        if (mind != 0xFFFFFFFFu && minv != 0xFFFFFFFFu) w2 = minv >= 0x40000000u ? c0 - (int)(minv - 0x40000000u) : c0 + (int)minv;
      }
      if (!isCorner) {
        const unsigned int db = v3 != 0xFFFFFFFFu ? __float_as_uint(d3) : 0xFFFFFFFFu;
        const unsigned int mind = __reduce_min_sync(FULL, db);
        const unsigned int minv = __reduce_min_sync(FULL, (db == mind) ? v3 : 0xFFFFFFFFu);
        if (mind != 0xFFFFFFFFu && minv != 0xFFFFFFFFu) w3 = minv >= 0x40000000u ? c0 - (int)(minv - 0x40000000u) : c0 + (int)minv;
      }
      if (lane == h) { min2 = w2; min3 = w3; }
    }
    if (valid) {
      ind[0] = closest; ind[1] = min2;
      if (!isCorner) ind[2] = min3;
    }
  }
  if (!valid) return;
  RowOut rowv;
#pragma unroll
  for (int k = 0; k < 6; k++) rowv.a[k] = 0.f;
  rowv.b = 0.f; rowv.flag = 0;
  float co[4];
  bool keep = false;
  if (isCorner) {
    if (ind[1] >= 0) {   // getLinePointDistance + getCornerFeatureCoefficients(A, B, X, iter), feature_utils.h:17-26, 42-61
      const float4 A = a.last_corner[ind[0]], B = a.last_corner[ind[1]];
      float bx = sx - B.x, by = sy - B.y, bz = sz - B.z;
      float ax = sx - A.x, ay = sy - A.y, az = sz - A.z;
      float kx = by * az - bz * ay, ky = bz * ax - bx * az, kz = bx * ay - by * ax;
      float knorm = norm3f(kx, ky, kz);
      float lengthAB = norm3f(A.x - B.x, A.y - B.y, A.z - B.z);
      float ex = B.x - A.x, ey = B.y - A.y, ez = B.z - A.z;
      float ux = ky * ez - kz * ey, uy = kz * ex - kx * ez, uz = kx * ey - ky * ex;
      float den = knorm * lengthAB;
      float dirx = -ux / den, diry = -uy / den, dirz = -uz / den;
      float distance = knorm / lengthAB;
      float weight = 1.0f;
      if (a.iter >= 5) weight = (float)(1 - 1.8 * fabs((double)distance));
      co[0] = dirx * weight; co[1] = diry * weight; co[2] = dirz * weight; co[3] = distance * weight;
      keep = ((double)weight > 0.1 && distance != 0);
    }
  } else {
    if (ind[1] >= 0 && ind[2] >= 0) {   // getSurfacePointDistance + getSurfaceFeatureCoefficients(A, B, C, X, iter), :28-40, 77-95
      const float4 A = a.last_surf[ind[0]], B = a.last_surf[ind[1]], C = a.last_surf[ind[2]];
      float b0 = B.x - A.x, b1 = B.y - A.y, b2 = B.z - A.z, c0 = C.x - A.x, c1 = C.y - A.y, c2 = C.z - A.z;
      float nx = b1 * c2 - b2 * c1, ny = b2 * c0 - b0 * c2, nz = b0 * c1 - b1 * c0;
      float nn = norm3f(nx, ny, nz);
      if (nn > 0.f) { nx /= nn; ny /= nn; nz /= nn; }
      float dsigned = ((sx - A.x) * nx + (sy - A.y) * ny) + (sz - A.z) * nz;
      float cosv = dsigned / norm3f(nx, ny, nz) / norm3f(A.x - sx, A.y - sy, A.z - sz);
      if (cosv < 0) { nx *= -1.0f; ny *= -1.0f; nz *= -1.0f; }
      float distance = (float)fabs((double)dsigned);
      float weight = 1.f;
      if (a.iter >= 5) weight = (float)(1 - 1.8 * fabs((double)distance) / sqrt((double)norm3f(sx, sy, sz)));
      co[0] = weight * nx; co[1] = weight * ny; co[2] = weight * nz; co[3] = weight * distance;
      keep = ((double)weight > 0.1 && distance != 0);
    }
  }
  if (keep) {
    const float x = p.x, y = p.y, z = p.z;   // LaserOdometry.cpp:556-577 (same Jacobian text as the mapping solver)
    float arx = (kc.x1 * y + kc.x2 * z) * co[0] + (kc.x3 * y - kc.x4 * z) * co[1] + (kc.x5 * y - kc.x6 * z) * co[2];
    float ary = (kc.y1 * x + kc.y2 * y + kc.y3 * z) * co[0] + (kc.y4 * x + kc.y5 * y + kc.y6 * z) * co[1] +
                (kc.y7 * x - kc.y8 * y - kc.y9 * z) * co[2];
    float arz = (kc.z1 * x - kc.x4 * y + kc.z3 * z) * co[0] + (kc.z4 * x + kc.z5 * y + kc.z6 + kc.z7 * z) * co[1] + 0 * co[2];
    rowv.a[0] = arx; rowv.a[1] = ary; rowv.a[2] = arz; rowv.a[3] = co[0]; rowv.a[4] = co[1]; rowv.a[5] = co[2];
    rowv.b = (float)(-0.05 * (double)co[3]);
    rowv.flag = 3;
  }
  float4* dst = reinterpret_cast<float4*>(a.rows + row);
  dst[0] = make_float4(rowv.a[0], rowv.a[1], rowv.a[2], rowv.a[3]);
  dst[1] = make_float4(rowv.a[4], rowv.a[5], rowv.b, __int_as_float(rowv.flag));
}

__global__ void __launch_bounds__(128) odom_corr_kernel(OdomArgs a) {
  __shared__ uint4 rng[8 * 128];
  __shared__ PoseCoef kc;
  __shared__ float tf[6];
  odom_corr_body(a, rng, kc, tf);
}

// the same for a batch of independent streams (blockIdx.y): clouds [S][cap], per-stream counts, grids and states
struct OdomBatchArgs {
  const float4* sharp; const float4* flat; int cap_sharp, cap_flat; const int* n_sharp; const int* n_flat;
  const float4* last_corner; const float4* last_surf; int cap_last_corner, cap_last_surf; const int* bound_corner; const int* bound_surf;
  const GridView* grid_corner; const GridView* grid_surf;
  const MatchState* state; int* ind; RowOut* rows; int iter;
  const int* iter_dev;   // optional: overrides iter (graph WHILE loop)
  int spread;
  const ChunkBox* box_corner; const ChunkBox* box_surf; int box_cap_corner, box_cap_surf;   // [S][box_cap_*] (optional)
};
__global__ void __launch_bounds__(128) odom_corr_batch_kernel(OdomBatchArgs b) {
  __shared__ uint4 rng[8 * 128];
  __shared__ PoseCoef kc;
  __shared__ float tf[6];
  const int s = blockIdx.y;
  OdomArgs a;
  a.sharp = b.sharp + (size_t)s * b.cap_sharp; a.flat = b.flat + (size_t)s * b.cap_flat;
  a.n_sharp = b.n_sharp[s]; a.n_flat = b.n_flat[s];
  a.last_corner = b.last_corner + (size_t)s * b.cap_last_corner; a.last_surf = b.last_surf + (size_t)s * b.cap_last_surf;
  a.bound_corner = b.bound_corner[s]; a.bound_surf = b.bound_surf[s];
  a.grid_corner = b.grid_corner[s]; a.grid_surf = b.grid_surf[s];
  a.state = b.state + s;
  a.ind = b.ind + (size_t)s * (2 * b.cap_sharp + 3 * b.cap_flat);
  a.rows = b.rows + (size_t)s * (b.cap_sharp + b.cap_flat);
  a.iter = b.iter_dev ? *b.iter_dev : b.iter;
  a.spread = b.spread;
  a.box_corner = b.box_corner ? b.box_corner + (size_t)s * b.box_cap_corner : nullptr;
  a.box_surf = b.box_surf ? b.box_surf + (size_t)s * b.box_cap_surf : nullptr;
  odom_corr_body(a, rng, kc, tf);
}

// chunk boxes of a batch of last clouds: one warp per chunk of 32 points
__global__ void __launch_bounds__(128) odom_boxes_batch_kernel(const float4* __restrict__ cloud, int cap, const int* __restrict__ n, ChunkBox* __restrict__ boxes,
                                                               int box_cap) {
  const int s = blockIdx.y, lane = threadIdx.x & 31;
  const int cb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (cb >= box_cap) return;
  const int j = cb * 32 + lane;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int rmin = 0x7fffffff, rmax = -0x7fffffff;
  if (j < n[s]) {
    const float4 q = cloud[(size_t)s * cap + j];
    mn[0] = mx[0] = q.x; mn[1] = mx[1] = q.y; mn[2] = mx[2] = q.z;
    rmin = rmax = (int)q.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
    rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
  }
  if (lane == 0) {
    ChunkBox b;
    b.mnx = mn[0]; b.mny = mn[1]; b.mnz = mn[2]; b.ring_min = rmin; b.mxx = mx[0]; b.mxy = mx[1]; b.mxz = mx[2]; b.ring_max = rmax;
    boxes[(size_t)s * box_cap + cb] = b;
  }
}

// transformToEnd for a batch: cloud [S][cap], tf6 [S][6], inv12 [S][12]; streams with on[s] == 0 keep their cloud as it is
__global__ void odom_to_end_batch_kernel(float4* __restrict__ cloud, int cap, const int* __restrict__ n, const float* __restrict__ tf6,
                                         const float* __restrict__ inv12, const int* __restrict__ on) {
  const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (!on[s] || i >= n[s]) return;
  float tf[6], R[9], t[3];
#pragma unroll
  for (int k = 0; k < 6; k++) tf[k] = tf6[6 * s + k];
#pragma unroll
  for (int k = 0; k < 9; k++) R[k] = inv12[12 * s + k];
#pragma unroll
  for (int k = 0; k < 3; k++) t[k] = inv12[12 * s + 9 + k];
  float4 p = cloud[(size_t)s * cap + i];
  float sx, sy, sz, x, y, z;
  odom_to_start(tf, p, &sx, &sy, &sz);
  transform_point(R, t, sx, sy, sz, &x, &y, &z);
  cloud[(size_t)s * cap + i] = make_float4(x, y, z, p.w);
}
// streams that do not run scanMatch this frame (first frame, or too few points in the last clouds, LaserOdometry.cpp:338)
__global__ void odom_gate_kernel(MatchState* state, const int* __restrict__ active, int nstreams) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nstreams && !active[s]) state[s].done = 1;
}

// transformToEnd (LaserOdometry.cpp:156-168): every point to the sweep start, then through the inverse of the full transform
__global__ void odom_to_end_kernel(float4* __restrict__ cloud, int n, const float* __restrict__ tf6, const float* __restrict__ inv12) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float tf[6], R[9], t[3];
#pragma unroll
  for (int k = 0; k < 6; k++) tf[k] = tf6[k];
#pragma unroll
  for (int k = 0; k < 9; k++) R[k] = inv12[k];
#pragma unroll
  for (int k = 0; k < 3; k++) t[k] = inv12[9 + k];
  float4 p = cloud[i];
  float sx, sy, sz, x, y, z;
  odom_to_start(tf, p, &sx, &sy, &sz);
  transform_point(R, t, sx, sy, sz, &x, &y, &z);
  cloud[i] = make_float4(x, y, z, p.w);
}
