// cm_mapping.cu -- the scan-to-map STAGE on top of the kernels: host mirror of LaserMapping::process
// (L_SLAM/src/odometry/LaserMapping.cpp:39-59) = transformMerge (LaserMatcher.cpp:333-340), prepareFeatureFrame
// (:288-301), prepareFeatureSurround (:303-325 -> FeatureMap::update FeatureMap.h:232-254, 308-352),
// optimizeTransform (:327-331 -> ScanMatch.cpp:349-360), transformUpdate (:342-347), featureMapUpdate (:349-355),
// batched over independent streams; plus the full pipeline entry (scan registration -> mapping).
// The host part is a few dozen float operations per stream and frame (pose chaining, cube window); everything that
// touches points runs on the device.
#include "cm_ctx.h"
#include "cm_math.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdint.h>
#include <algorithm>

namespace cm {

// ---- Eigen::Isometry3f algebra (same operation order as Transform * Transform / inverse(Isometry)) ---------------------
static HostIso iso_identity() { HostIso i; for (int k = 0; k < 9; k++) i.R[k] = (k % 4 == 0) ? 1.f : 0.f; i.t[0] = i.t[1] = i.t[2] = 0.f; return i; }
static HostIso iso_mul(const HostIso& a, const HostIso& b) {
  HostIso r;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) r.R[i * 3 + j] = (a.R[i * 3 + 0] * b.R[0 * 3 + j] + a.R[i * 3 + 1] * b.R[1 * 3 + j]) + a.R[i * 3 + 2] * b.R[2 * 3 + j];
    r.t[i] = ((a.R[i * 3 + 0] * b.t[0] + a.R[i * 3 + 1] * b.t[1]) + a.R[i * 3 + 2] * b.t[2]) + a.t[i];
  }
  return r;
}
static HostIso iso_inverse(const HostIso& a) {
  HostIso r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; i++) r.t[i] = -((r.R[i * 3 + 0] * a.t[0] + r.R[i * 3 + 1] * a.t[1]) + r.R[i * 3 + 2] * a.t[2]);
  return r;
}
// convertTransform(Isometry3f&, Twist&), transform_utils.h:313-331 (getEulerAngles :54-60)
static void iso_to_twist(const HostIso& it, float pose[6]) {
  pose[3] = it.t[0]; pose[4] = it.t[1]; pose[5] = it.t[2];
  pose[0] = atan2f(it.R[7], it.R[8]);
  pose[1] = asinf(-it.R[6]);
  pose[2] = atan2f(it.R[3], it.R[0]);
}
static void twist_to_iso(const float pose[6], HostIso& it) { pose_to_matrix(pose, it.R); it.t[0] = pose[3]; it.t[1] = pose[4]; it.t[2] = pose[5]; }

// FeatureMap::update + computeActiveAera (FeatureMap.h:232-254, 308-352).  shift_d returns the argument of the reference's
// shift(newGrid - grid) (all zero: the cubes stay); the origin in `st` is already moved by it.
static void update_window(const cm_config& cfg, MappingStream& st, const float sensor[3], CubeWindow& w, int shift_d[3]) {
  const int dims[3] = {cfg.cube_w, cfg.cube_h, cfg.cube_d};
  int g[3];
  for (int k = 0; k < 3; k++) g[k] = (int)(roundf(sensor[k] / cfg.cube_size) + (float)st.origin[k]);   // worldToCube :479-481
  const int PAD = 3;
  for (int k = 0; k < 3; k++) {
    int ng = std::min(std::max(g[k], PAD), dims[k] - PAD - 1);
    shift_d[k] = ng - g[k];         // shift(newGrid - grid), then _cubeOrigin += newGrid - grid (:243-249)
    st.origin[k] += shift_d[k];
    st.cur[k] = ng;
  }
  memset(&w, 0, sizeof(w));
  for (int k = 0; k < 3; k++) { w.origin[k] = st.origin[k]; w.dims[k] = dims[k]; w.w0[k] = st.cur[k] - 3; }
  w.cube_size = cfg.cube_size;
  const int window = (int)ceilf(cfg.valid_distance / cfg.cube_size);   // 3 with the reference's 150 / 50
  for (int i = st.cur[0] - window; i <= st.cur[0] + window; i++)
    for (int j = st.cur[1] - window; j <= st.cur[1] + window; j++)
      for (int k = st.cur[2] - window; k <= st.cur[2] + window; k++) {
        if (!(0 <= i && i < dims[0] && 0 <= j && j < dims[1] && 0 <= k && k < dims[2])) continue;
        float centerX = cfg.cube_size * (i - st.origin[0]);
        float centerY = cfg.cube_size * (j - st.origin[1]);
        float centerZ = cfg.cube_size * (k - st.origin[2]);
        bool inFov = false;
        for (int ii = -1; ii <= 1 && !inFov; ii += 2)
          for (int jj = -1; jj <= 1 && !inFov; jj += 2)
            for (int kk = -1; kk <= 1 && !inFov; kk += 2) {
              float cx = (float)(centerX + cfg.cube_size / 2.0 * ii);
              float cy = (float)(centerY + cfg.cube_size / 2.0 * jj);
              float cz = (float)(centerZ + cfg.cube_size / 2.0 * kk);
              float dx = sensor[0] - cx, dy = sensor[1] - cy, dz = sensor[2] - cz;
              float sq = dx * dx + dy * dy + dz * dz;
              if (sqrt((double)sq) < cfg.valid_distance) inFov = true;
            }
        int wi = i - w.w0[0], wj = j - w.w0[1], wk = k - w.w0[2];
        if (inFov && wi >= 0 && wi < 7 && wj >= 0 && wj < 7 && wk >= 0 && wk < 7) w.active[(wi * 7 + wj) * 7 + wk] = 1;
      }
  for (int a = 0; a < 343; a++) {
    int i = a / 49, j = (a / 7) % 7, k = a % 7;
    bool all = true;
    for (int di = -1; di <= 1 && all; di++)
      for (int dj = -1; dj <= 1 && all; dj++)
        for (int dk = -1; dk <= 1 && all; dk++) {
          int x = i + di, y = j + dj, z = k + dk;
          if (x < 0 || x > 6 || y < 0 || y > 6 || z < 0 || z > 6 || !w.active[(x * 7 + y) * 7 + z]) all = false;
        }
    w.interior[a] = all ? 1 : 0;
  }
}

__global__ void gather_counts_kernel(const int* __restrict__ n5, int* __restrict__ n2, int S) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S) { n2[s] = n5[s * 5 + 1]; n2[S + s] = n5[s * 5 + 3]; }   // lessSharp -> corner, lessFlat -> surf
}

// transformPointCloud (transform_utils.h:601-614) of a filtered frame cloud into a window frame (LaserMappingLocal.cpp:67-70)
struct IsoArg { float R[9]; float t[3]; };
__global__ void local_transform_kernel(const float4* __restrict__ in, int n, IsoArg tf, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  float4 q; q.w = p.w;
  transform_point(tf.R, tf.t, p.x, p.y, p.z, &q.x, &q.y, &q.z);
  out[i] = q;
}

static int default_kdiv(float cell, float leaf, int mult) {
  if (cell > 0.f) { int k = (int)floorf(cell / leaf + 0.5f); return k < 1 ? 1 : k; }
  return mult;
}

}  // namespace cm

using namespace cm;
static int fail(cm_ctx* ctx, int code, const std::string& msg) { return ctx_fail(ctx, code, msg); }

extern "C" {

int cm_mapping_create(cm_ctx* ctx, int nstreams, size_t max_corner_points, size_t max_surf_points) {
  if (!ctx || nstreams <= 0 || nstreams > 256 || max_corner_points == 0 || max_surf_points == 0) return fail(ctx, CM_ERR_ARG, "bad argument");
  const cm_config& c = ctx->cfg;
  if (ceilf(c.valid_distance / c.cube_size) > 3.f) return fail(ctx, CM_ERR_UNSUPPORTED, "valid_distance / cube_size > 3");
  try {
    cudaSetDevice(c.device);
    MapConfig mc;
    mc.max_corner = max_corner_points; mc.max_surf = max_surf_points;
    mc.leaf_corner = c.map_filter_corner; mc.leaf_surf = c.map_filter_surf;
    mc.kdiv_corner = default_kdiv(c.cell_corner, c.map_filter_corner, 8);
    mc.kdiv_surf = default_kdiv(c.cell_surf, c.map_filter_surf, 4);
    mc.cube_size = c.cube_size;
    mc.dims[0] = c.cube_w; mc.dims[1] = c.cube_h; mc.dims[2] = c.cube_d;
    for (int k = 0; k < 3; k++) mc.origin[k] = (int)round((mc.dims[k] - 1) / 2.0);   // FeatureMap.h:63-65
    ctx->map.create(nstreams, mc, ctx->stream);
    ctx->wins_shadow.clear();
    ctx->map_streams = nstreams;
    ctx->mstreams.assign(nstreams, MappingStream());
    for (auto& st : ctx->mstreams) {
      st.mappedLast = st.mappedNew = st.odomLast = iso_identity();   // LaserMatcher.cpp:31-32
      for (int k = 0; k < 3; k++) { st.origin[k] = mc.origin[k]; st.cur[k] = 0; }
    }
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

static int pipeline_prefetch(cm_ctx* ctx, const void* frames, int rows, int cols, bool is_host, const void* const* clouds = nullptr, size_t stride = 0);
// the prefetch registered with cm_pipeline_prefetch_deferred_* (if any): submitted while the device works on the current step
static void issue_deferred_prefetch(cm_ctx* ctx) {
  if (!ctx->defer_frames) return;
  const void* f = ctx->defer_frames;
  ctx->defer_frames = nullptr;
  ctx->defer_rc = pipeline_prefetch(ctx, f, ctx->defer_rows, ctx->defer_cols, ctx->defer_is_host);
}

// Core of the stage on DEVICE clouds.  d_corner/d_surf: [S][cap] with device counts d_n (corner counts then surf counts,
// [2][S]).  h_odom: S odometry poses (host).  Outputs on the host: mapped poses and stats.
// max_in_c / max_in_s: host-known upper bounds of the input counts (sizes the sorts).
// localise: LaserLocalization::process (LaserLocalization.cpp:163-188) instead of LaserMapping::process -- the matcher is
// FeatureMap::scanMatchScan (FeatureMap.h:490-690: neighbours from the query's own cube only, no reference-size gate,
// 10 iterations, 0.05 deg / 0.05 cm) and the map is not updated.
static int mapping_process_dev(cm_ctx* ctx, const float4* d_corner, int cap_c, const float4* d_surf, int cap_s, const int* d_n,
                               int max_in_c, int max_in_s, const cm_iso* h_odom, cm_iso* h_mapped, cm_match_stats* h_stats,
                               bool localise = false) {
  const cm_config& cfg = ctx->cfg;
  const int S = ctx->map_streams;
  cudaStream_t st = ctx->stream;
  MatchParamsDev prm = dev_params(cfg);
  if (localise) {
    prm.own_cube_only = 1; prm.max_iterations = 10; prm.delta_t_abort = 0.05f; prm.delta_r_abort = 0.05f;
    prm.min_ref_corner = 0; prm.min_ref_surf = 0;
  }
  // transformMerge
  // poses and cube windows go to the device through pinned staging + a copy kernel (COOPERMAP_STAGED_UPLOAD=0: plain memcpys)
  static const bool staged = !(getenv("COOPERMAP_STAGED_UPLOAD") && atoi(getenv("COOPERMAP_STAGED_UPLOAD")) == 0);
  const size_t pose_bytes = (sizeof(float) * 6 * S + 255) & ~(size_t)255, stage_bytes = pose_bytes + sizeof(CubeWindow) * S;
  if (ctx->h_stage_cap < stage_bytes) {
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
    CM_CUDA_CHECK(ctx, cudaHostAlloc(&ctx->h_stage, stage_bytes, cudaHostAllocMapped | cudaHostAllocPortable));
    ctx->h_stage_cap = stage_bytes;
  }
  float* pose_in = (float*)ctx->h_stage;
  CubeWindow* wins = (CubeWindow*)((char*)ctx->h_stage + pose_bytes);
  {
    // every stream's window is validated on a copy first: a failure leaves no stream half-updated
    std::vector<MappingStream> next(ctx->mstreams.begin(), ctx->mstreams.end());
    std::vector<int> shifts;   // (stream, d[3]) of the streams whose cube grid is re-centred this frame
    for (int s = 0; s < S; s++) {
      MappingStream& ms = next[s];
      HostIso odomNew; memcpy(odomNew.R, h_odom[s].R, 36); memcpy(odomNew.t, h_odom[s].t, 12);
      HostIso L2W = iso_mul(ms.mappedLast, iso_inverse(ms.odomLast));   // transformAssociate, transform_utils.h:502-507
      ms.mappedNew = iso_mul(L2W, odomNew);
      int d[3];
      update_window(cfg, ms, ms.mappedNew.t, wins[s], d);
      if (d[0] || d[1] || d[2]) {
        if (ctx->dist.on) return fail(ctx, CM_ERR_UNSUPPORTED, "FeatureMap::shift on a sharded map");
        const bool wrong_way = d[0] > 0 || (d[0] == 0 && (d[1] > 0 || (d[1] == 0 && d[2] > 0)));
        if (wrong_way && ctx->map.cur_epoch[s] >= 255) return fail(ctx, CM_ERR_UNSUPPORTED, "more than 255 wrong-way FeatureMap::shift calls");
        shifts.push_back(s); shifts.push_back(d[0]); shifts.push_back(d[1]); shifts.push_back(d[2]);
      }
      iso_to_twist(ms.mappedNew, &pose_in[6 * s]);
    }
    ctx->mstreams.swap(next);
    for (size_t q = 0; q < shifts.size(); q += 4) {   // FeatureMap::shift: relabel / drop the stored points of that stream
      const int s = shifts[q];
      ctx->map.shift(s, &shifts[q + 1], ctx->mstreams[s].origin, st);
      ctx->n_shifts++;
    }
  }
  // prepareFeatureFrame: voxel filters
  ctx->m_corner_ds.reserve((size_t)S * cap_c * sizeof(float4));
  ctx->m_surf_ds.reserve((size_t)S * cap_s * sizeof(float4));
  ctx->m_n_ds.reserve(sizeof(int) * 2 * S);
  ctx->d_flag.reserve(sizeof(int));
  CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_flag.p, 0, sizeof(int), st));
  int* d_nds = (int*)ctx->m_n_ds.p;
  if (!ctx->aux_stream) {
    CM_CUDA_CHECK(ctx, cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, ctx->prio_high));
    CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
    CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming));
  }
  cudaStream_t aux = ctx->aux_stream;
  // the two filter chains (and, below, the two insert chains) are fixed launch sequences: replayed from CUDA graphs keyed by
  // their arguments, with the element bounds rounded up so that the keys repeat from step to step
  static const bool no_graph = getenv("COOPERMAP_NO_GRAPH") != nullptr;
  const bool use_graphs = !no_graph && !g_timeline.on;
  auto P = [](const void* p) { return (unsigned long long)(uintptr_t)p; };
  // both classes of every stream in one launch (one CTA per cloud); the counts never leave the device
  ctx->voxel.run2(S, d_corner, d_n, cap_c, cfg.filter_corner, (float4*)ctx->m_corner_ds.p, d_nds, cap_c,
                  d_surf, d_n + S, cap_s, cfg.filter_surf, (float4*)ctx->m_surf_ds.p, d_nds + S, cap_s,
                  std::max(max_in_c, max_in_s), (int*)ctx->d_flag.p, st);
  // The filtered counts stay on the device: no host round trip in the middle of a step.  Everything downstream is launched
  // for an ESTIMATE -- the previous step's largest filtered clouds + 25 % (a LiDAR's feature counts move by a few per cent
  // from sweep to sweep), never more than the unfiltered counts, which bound them.  The correspondence kernels loop when a
  // stream exceeds the estimate, the map insertion is skipped as a whole and repeated below (flags[4], [5]).
  const int bound_c = std::min(cap_c, std::max(max_in_c, 1)), bound_s = std::min(cap_s, std::max(max_in_s, 1));
  auto estimate = [](int prev, int bound, int round) {
    if (prev <= 0) return bound;
    const long long e = ((long long)prev * 5 / 4 + round + round - 1) / round * round;
    return (int)std::min<long long>(e, bound);
  };
  int max_c = estimate(ctx->est_c, bound_c, 256), max_s = estimate(ctx->est_s, bound_s, 1024);
  if (getenv("COOPERMAP_TEST_UNDERESTIMATE")) { max_c = std::min(max_c, 64); max_s = std::min(max_s, 256); }   // tests: force the overflow paths
  const int max_q = std::min(max_c + max_s, bound_c + bound_s);
  launch_step_guard(d_nds, S, max_c, max_s, (int*)ctx->map.flags.p + 6, st);
  // prepareFeatureSurround: cube window -> searchable views
  {
    // the window only changes when a sensor crosses a cube face: skip the upload when the device copy is current
    const size_t wb = sizeof(CubeWindow) * S;
    const bool same = ctx->wins_shadow.size() == wb && memcmp(ctx->wins_shadow.data(), wins, wb) == 0;
    if (!same) ctx->wins_shadow.assign((const unsigned char*)wins, (const unsigned char*)wins + wb);
    ctx->map.set_windows(same ? nullptr : wins, prm.knn_gate, st, staged);
  }
  // optimizeTransform
  ctx->m_pose.reserve(sizeof(float) * 6 * S);
  ctx->m_state.reserve(sizeof(MatchState) * S);
  ctx->m_rows.reserve((size_t)S * (cap_c + cap_s) * sizeof(RowOut));
  ctx->m_slots.reserve((size_t)S * (cap_c + cap_s) * 5 * sizeof(int));
  ctx->m_sums.reserve(sizeof(double) * 32 * S);
  if (staged) staged_upload(ctx->m_pose.p, pose_in, sizeof(float) * 6 * S, st);
  else CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_pose.p, pose_in, sizeof(float) * 6 * S, cudaMemcpyHostToDevice, st));
  MatchLaunch m;
  m.nstreams = S;
  m.corner = (const float4*)ctx->m_corner_ds.p; m.surf = (const float4*)ctx->m_surf_ds.p;
  m.n_corner = d_nds; m.n_surf = d_nds + S; m.cap_corner = cap_c; m.cap_surf = cap_s;
  m.grid_corner = (const GridView*)ctx->map.views[0].p; m.grid_surf = (const GridView*)ctx->map.views[1].p;
  m.pose_in = (const float*)ctx->m_pose.p; m.state = (MatchState*)ctx->m_state.p; m.rows = (RowOut*)ctx->m_rows.p; m.nn_slot = (int*)ctx->m_slots.p;
  m.sums = (double*)ctx->m_sums.p; m.trace = nullptr; m.nn = nullptr; m.orig_idx = 0; m.prm = prm;
  m.max_queries = max_q; m.bound_queries = bound_c + bound_s;
  m.skip = (const int*)ctx->map.flags.p + 6;   // set by launch_step_guard below when an estimate was too small
  const bool sharded = ctx->dist.on && ctx->dist.dev.nranks > 1;
  if (sharded) { m.dist_rank = ctx->dist.dev.rank; m.dist_nranks = ctx->dist.dev.nranks; }
  // One map over several ranks: every rank runs the same loop on its own queries; between the per-stream sums and the 6x6 step
  // the library's exchange kernel adds the sums of all ranks in rank order (cm_dist.cu), so every rank takes the same step.
  // Launch by launch: every rank must issue the same max_iterations exchanges whatever its streams converge to.
  auto run_gn_sharded = [&](const MatchLaunch& ml) -> int {
    ctx->dist.scratch.reserve(sizeof(double) * 2 * (size_t)S);
    ctx->map.pack_npts((double*)ctx->dist.scratch.p, st);          // the reference-size gate (ScanMatch.cpp:57-61) looks at the WHOLE map
    int rc = dist_allreduce(ctx, (double*)ctx->dist.scratch.p, 2 * S, st);
    if (rc < 0) return rc;
    ctx->map.unpack_npts((const double*)ctx->dist.scratch.p, st);
    launch_match_init(ml, st);
    for (int it = 0; it < ml.prm.max_iterations; it++) {
      launch_match_partial(ml, it, st, &ctx->prof, true, true);
      rc = dist_allreduce(ctx, ml.sums, 32 * S, st);
      if (rc < 0) return rc;
      launch_match_solve_warp(ml, it, st);
    }
    return CM_OK;
  };
  // capacities from the bound, in whole 256-query tiles: the Gauss-Newton graph key then repeats from frame to frame
  ctx->hardq.attach(m, (size_t)S * (size_t)(((m.bound_queries + 32 + 255) / 256) * 256));
  if (ctx->dbg_on) {
    ctx->dbg_words = (size_t)((max_q + 32 + 255) / 256) * S * 8 * 4;
    ctx->dbg_trace.reserve(ctx->dbg_words * sizeof(unsigned long long));
    m.dbg = (unsigned long long*)ctx->dbg_trace.p; m.dbg_iter = ctx->dbg_iter;
  }
  {
    int G = cfg.gn_groups > 0 ? cfg.gn_groups : 1;
    if (const char* e = getenv("COOPERMAP_GN_GROUPS")) G = atoi(e);   // development override
    if (G > CM_MAX_GN_GROUPS) G = CM_MAX_GN_GROUPS;
    if (G > S) G = S;
    if (G > 1) {
      if (!ctx->gn_fork) CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->gn_fork, cudaEventDisableTiming));
      for (int g = 0; g < G; g++) {
        if (!ctx->gn_stream[g]) CM_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->gn_stream[g], cudaStreamNonBlocking));
        if (!ctx->gn_join[g]) CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->gn_join[g], cudaEventDisableTiming));
      }
      launch_match_groups(m, st, G, ctx->gn_stream, ctx->gn_fork, ctx->gn_join, &ctx->prof);
    } else {
      // replay from a CUDA graph unless per-launch events are wanted (bench.py's kernel timing, the timeline, the search trace)
      const bool want_events = ctx->prof.enabled || g_timeline.on || ctx->dbg_on || no_graph;
      if (sharded) { const int rc = run_gn_sharded(m); if (rc < 0) return rc; }
      else if (want_events || !ctx->match_graphs.launch(m, st)) launch_match(m, st, &ctx->prof);
    }
  }
  issue_deferred_prefetch(ctx);   // the Gauss-Newton loop is on its way: submit the next sweep's upload + scan registration now
  // results: states, filtered counts and flags come back together, into pinned memory, with an event behind them -- the step's
  // only host synchronisation.  The map insertion is enqueued BEHIND that event: the caller has its poses while the insertion
  // still runs (a synchronous caller's next sweep is uploaded and scan-registered on the side stream meanwhile).  What the
  // insertion reports (capacity, voxel range) stays in the device flags and comes back with the next step's results, or with
  // cm_mapping_sync / the map read-outs.
  const size_t res_bytes = sizeof(MatchState) * S + sizeof(int) * 2 * S + sizeof(int) * 8;
  if (ctx->h_result_cap < res_bytes) {
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    ctx->h_result = nullptr; ctx->h_result_cap = 0;
    CM_CUDA_CHECK(ctx, cudaHostAlloc(&ctx->h_result, res_bytes, cudaHostAllocDefault));
    ctx->h_result_cap = res_bytes;
  }
  if (!ctx->result_ready) CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->result_ready, cudaEventDisableTiming));
  MatchState* hs = (MatchState*)ctx->h_result;
  int* nds = (int*)(hs + S);
  int* flags = nds + 2 * S;
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(hs, ctx->m_state.p, sizeof(MatchState) * S, cudaMemcpyDeviceToHost, st));
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(nds, d_nds, sizeof(int) * 2 * S, cudaMemcpyDeviceToHost, st));
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(flags, ctx->map.flags.p, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
  CM_CUDA_CHECK(ctx, cudaEventRecord(ctx->result_ready, st));
  // featureMapUpdate (commented out in LaserLocalization::process, LaserLocalization.cpp:186)
  if (!localise) {
    CM_CUDA_CHECK(ctx, cudaEventRecord(ctx->aux_fork, st));
    CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(aux, ctx->aux_fork, 0));
    const int ic_r = max_c, is_r = max_s;
    const int* step_skip = (const int*)ctx->map.flags.p + 6;
    auto ins_c = [&]() { ctx->map.insert(0, (const float4*)ctx->m_corner_ds.p, d_nds, cap_c, ic_r, (const MatchState*)ctx->m_state.p, nullptr, aux, step_skip); };
    auto ins_s = [&]() { ctx->map.insert(1, (const float4*)ctx->m_surf_ds.p, d_nds + S, cap_s, is_r, (const MatchState*)ctx->m_state.p, nullptr, st, step_skip); };
    if (use_graphs) {
      ctx->stage_graphs.run({3, (unsigned long long)S, P(ctx->m_corner_ds.p), P(d_nds), (unsigned long long)cap_c, (unsigned long long)ic_r, P(ctx->m_state.p), P(aux)}, aux, ins_c);
      CM_CUDA_CHECK(ctx, cudaEventRecord(ctx->aux_join, aux));
      ctx->stage_graphs.run({4, (unsigned long long)S, P(ctx->m_surf_ds.p), P(d_nds), (unsigned long long)cap_s, (unsigned long long)is_r, P(ctx->m_state.p), P(st)}, st, ins_s);
    } else {
      ins_c();
      CM_CUDA_CHECK(ctx, cudaEventRecord(ctx->aux_join, aux));
      ins_s();
    }
    CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(st, ctx->aux_join, 0));
  }
  CM_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->result_ready));
  CM_CUDA_CHECK(ctx, cudaGetLastError());
  if (sharded) {
    int derr = 0;
    CM_CUDA_CHECK(ctx, cudaMemcpy(&derr, ctx->dist.err.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (derr) return fail(ctx, CM_ERR_CUDA, "sharded map: the exchange timed out waiting for a peer rank");
  }
  {
    int act_c = 1, act_s = 1;
    for (int s = 0; s < S; s++) { act_c = std::max(act_c, nds[s]); act_s = std::max(act_s, nds[S + s]); }
    ctx->est_c = act_c; ctx->est_s = act_s;
    if (flags[6]) {
      // a stream had more filtered points than this step's launches were sized for: the Gauss-Newton kernels and the map
      // insertion saw the flag and did nothing.  Repeat both with exact sizes (launch by launch: this is the rare path).
      ++ctx->insert_redos;
      CM_CUDA_CHECK(ctx, cudaMemsetAsync((int*)ctx->map.flags.p + 4, 0, sizeof(int) * 3, st));
      m.max_queries = std::min(act_c, cap_c) + std::min(act_s, cap_s);
      m.skip = nullptr;
      if (sharded) { const int rc = run_gn_sharded(m); if (rc < 0) return rc; }
      else launch_match(m, st, &ctx->prof);
      if (!localise) {
        ctx->map.insert(0, (const float4*)ctx->m_corner_ds.p, d_nds, cap_c, std::min(cap_c, act_c), (const MatchState*)ctx->m_state.p, nullptr, st);
        ctx->map.insert(1, (const float4*)ctx->m_surf_ds.p, d_nds + S, cap_s, std::min(cap_s, act_s), (const MatchState*)ctx->m_state.p, nullptr, st);
      }
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(hs, ctx->m_state.p, sizeof(MatchState) * S, cudaMemcpyDeviceToHost, st));
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(flags, ctx->map.flags.p, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
      CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
      CM_CUDA_CHECK(ctx, cudaGetLastError());
    }
  }
  if (flags[0] || flags[2] || flags[3]) {
    // reported once: the flags are cleared so that the context stays usable (the dropped appends of THIS frame are lost, the
    // poses below are not advanced -- the caller may retry the frame after making room)
    cudaMemsetAsync(ctx->map.flags.p, 0, sizeof(int) * 8, st);
    cudaStreamSynchronize(st);
    if (flags[0]) return fail(ctx, CM_ERR_UNSUPPORTED, "map point outside the supported voxel range (+-65536 voxels)");
    return fail(ctx, CM_ERR_CAPACITY, "map capacity exhausted (raise max_*_points in cm_mapping_create)");
  }
  ctx->last_query_iters = 0; ctx->last_queries = 0; ctx->last_inserted = 0;
  for (int s = 0; s < S; s++) {
    const unsigned long long q = (unsigned long long)(nds[s] + nds[S + s]);
    const int evals = hs[s].iterations + ((hs[s].flags & CM_F_TOO_FEW_MATCHES) ? 1 : 0);   // correspondence passes actually run
    ctx->last_query_iters += q * (unsigned long long)evals; ctx->last_queries += q; ctx->last_inserted += localise ? 0 : q;
  }
  for (int s = 0; s < S; s++) {
    MappingStream& ms = ctx->mstreams[s];
    twist_to_iso(hs[s].pose, ms.mappedNew);   // ScanMatch.cpp:358 (always, also when the map was too small)
    ms.mappedLast = ms.mappedNew;             // transformUpdate
    memcpy(ms.odomLast.R, h_odom[s].R, 36); memcpy(ms.odomLast.t, h_odom[s].t, 12);
    if (h_mapped) { memcpy(h_mapped[s].R, ms.mappedNew.R, 36); memcpy(h_mapped[s].t, ms.mappedNew.t, 12); }
    if (h_stats) fill_match_stats(cfg, hs[s], (size_t)(nds[s] + nds[S + s]), &h_stats[s]);
  }
  return CM_OK;
}

static int mapping_process_host(cm_ctx* ctx, const cm_iso* odom, const cm_point* corner, const int* n_corner, int cap_corner,
                                const cm_point* surf, const int* n_surf, int cap_surf, cm_iso* mapped, cm_match_stats* stats, bool localise) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!odom || !corner || !surf || !n_corner || !n_surf || cap_corner <= 0 || cap_surf <= 0) return fail(ctx, CM_ERR_ARG, "bad argument");
  const int S = ctx->map_streams;
  for (int s = 0; s < S; s++)
    if (n_corner[s] < 0 || n_corner[s] > cap_corner || n_surf[s] < 0 || n_surf[s] > cap_surf) return fail(ctx, CM_ERR_ARG, "count out of range");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    ctx->m_corner_in.reserve((size_t)S * cap_corner * sizeof(cm_point));
    ctx->m_surf_in.reserve((size_t)S * cap_surf * sizeof(cm_point));
    ctx->m_n_in.reserve(sizeof(int) * 2 * S);
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_corner_in.p, corner, (size_t)S * cap_corner * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_surf_in.p, surf, (size_t)S * cap_surf * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_n_in.p, n_corner, sizeof(int) * S, cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync((int*)ctx->m_n_in.p + S, n_surf, sizeof(int) * S, cudaMemcpyHostToDevice, st));
    int mc_ = 1, ms_ = 1;
    for (int s = 0; s < S; s++) { mc_ = std::max(mc_, n_corner[s]); ms_ = std::max(ms_, n_surf[s]); }
    return mapping_process_dev(ctx, (const float4*)ctx->m_corner_in.p, cap_corner, (const float4*)ctx->m_surf_in.p, cap_surf,
                               (const int*)ctx->m_n_in.p, mc_, ms_, odom, mapped, stats, localise);
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
}

int cm_mapping_process_host(cm_ctx* ctx, const cm_iso* odom, const cm_point* corner, const int* n_corner, int cap_corner,
                            const cm_point* surf, const int* n_surf, int cap_surf, cm_iso* mapped, cm_match_stats* stats) {
  return mapping_process_host(ctx, odom, corner, n_corner, cap_corner, surf, n_surf, cap_surf, mapped, stats, false);
}
int cm_localization_process_host(cm_ctx* ctx, const cm_iso* odom, const cm_point* corner, const int* n_corner, int cap_corner,
                                 const cm_point* surf, const int* n_surf, int cap_surf, cm_iso* mapped, cm_match_stats* stats) {
  return mapping_process_host(ctx, odom, corner, n_corner, cap_corner, surf, n_surf, cap_surf, mapped, stats, true);
}

// ---- LaserMappingLocal: the mapping stage over a sliding window of frames (LocalFeatureMap) ---------------------------------
int cm_mapping_local_create(cm_ctx* ctx, int use_mapped_pose) {
  if (!ctx) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  LocalWindow& w = ctx->local;
  w.frames.clear();
  w.created = true; w.use_mapped = use_mapped_pose != 0;
  w.accum = 0.0; w.first = true;
  w.mappedLast = w.mappedNew = w.odomLast = iso_identity();   // LaserMatcher.cpp:31-32
  w.n_surround[0] = w.n_surround[1] = 0;
  return CM_OK;
}

int cm_mapping_local_process_host(cm_ctx* ctx, const cm_iso* odom, const cm_point* corner, size_t nc, const cm_point* surf, size_t ns,
                                  cm_iso* mapped, cm_match_stats* stats) {
  if (!ctx || !ctx->local.created) return fail(ctx, CM_ERR_ARG, "cm_mapping_local_create has not been called");
  if (!odom || (!corner && nc) || (!surf && ns) || nc > 0x7fffffffu || ns > 0x7fffffffu) return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    const cm_config& cfg = ctx->cfg;
    cudaStream_t st = ctx->stream;
    LocalWindow& w = ctx->local;
    // transformMerge, LaserMatcher.cpp:333-340
    HostIso odomNew; memcpy(odomNew.R, odom->R, 36); memcpy(odomNew.t, odom->t, 12);
    w.mappedNew = iso_mul(iso_mul(w.mappedLast, iso_inverse(w.odomLast)), odomNew);
    // window totals; the concatenation (LocalFeatureMap.h:88-92) is a device-to-device gather in queue order
    size_t wc = 0, wsn = 0;
    for (auto& f : w.frames) { wc += (size_t)f->nc; wsn += (size_t)f->ns; }
    if (wc > 0x7fffffffu || wsn > 0x7fffffffu) return fail(ctx, CM_ERR_CAPACITY, "window too large");
    // slots: 0 window corner, 1 window surf (leaf 0.2 / 0.4, LocalFeatureMap.h:29-31), 2 frame corner, 3 frame surf (prepareFeatureFrame)
    const size_t n[4] = {wc, wsn, nc, ns};
    const float leaf[4] = {0.2f, 0.4f, cfg.filter_corner, cfg.filter_surf};
    DeviceBuffer* raw[4] = {&ctx->d_ref_corner, &ctx->d_ref_surf, &ctx->d_corner, &ctx->d_surf};
    DeviceBuffer* ds[4] = {&ctx->d_l_ds[0], &ctx->d_l_ds[1], &ctx->d_l_ds[2], &ctx->d_l_ds[3]};
    ctx->d_vn_in.reserve(4 * sizeof(int)); ctx->d_vn_out.reserve(4 * sizeof(int)); ctx->d_flag.reserve(sizeof(int));
    int nin[4] = {(int)wc, (int)wsn, (int)nc, (int)ns};
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_vn_in.p, nin, sizeof(nin), cudaMemcpyHostToDevice, st));
    for (int k = 0; k < 4; k++) { const size_t cap = n[k] ? n[k] : 1; raw[k]->reserve(cap * sizeof(cm_point)); ds[k]->reserve(cap * sizeof(cm_point)); }
    {
      size_t oc = 0, os = 0;
      for (auto& f : w.frames) {
        if (f->nc) CM_CUDA_CHECK(ctx, cudaMemcpyAsync((float4*)raw[0]->p + oc, f->corner.p, (size_t)f->nc * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        if (f->ns) CM_CUDA_CHECK(ctx, cudaMemcpyAsync((float4*)raw[1]->p + os, f->surf.p, (size_t)f->ns * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        oc += (size_t)f->nc; os += (size_t)f->ns;
      }
    }
    if (nc) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(raw[2]->p, corner, nc * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    if (ns) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(raw[3]->p, surf, ns * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    for (int k = 0; k < 4; k++) {
      const size_t cap = n[k] ? n[k] : 1;
      ctx->voxel.run(1, (const float4*)raw[k]->p, (const int*)ctx->d_vn_in.p + k, (int)cap, (int)n[k], leaf[k], (float4*)ds[k]->p,
                     (int*)ctx->d_vn_out.p + k, (int)cap, (int*)ctx->d_flag.p, st);
    }
    int nout[4];
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(nout, ctx->d_vn_out.p, sizeof(nout), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    w.n_surround[0] = nout[0]; w.n_surround[1] = nout[1];
    // optimizeTransform, LaserMatcher.cpp:327-331
    float pose6[6];
    iso_to_twist(w.mappedNew, pose6);
    cm_pose tw; tw.rx = pose6[0]; tw.ry = pose6[1]; tw.rz = pose6[2]; tw.tx = pose6[3]; tw.ty = pose6[4]; tw.tz = pose6[5];
    const int rc = cm_match_stateless_dev(ctx, (const float4*)ds[0]->p, (size_t)nout[0], (const float4*)ds[1]->p, (size_t)nout[1],
                                          (const float4*)ds[2]->p, (size_t)nout[2], (const float4*)ds[3]->p, (size_t)nout[3], &tw, stats,
                                          nullptr, nullptr, nullptr);
    if (rc < 0) return rc;
    pose6[0] = tw.rx; pose6[1] = tw.ry; pose6[2] = tw.rz; pose6[3] = tw.tx; pose6[4] = tw.ty; pose6[5] = tw.tz;
    twist_to_iso(pose6, w.mappedNew);   // ScanMatch.cpp:358
    // transformUpdate, LaserMatcher.cpp:342-347
    w.mappedLast = w.mappedNew;
    w.odomLast = odomNew;
    if (mapped) { memcpy(mapped->R, w.mappedNew.R, 36); memcpy(mapped->t, w.mappedNew.t, 12); }
    // featureMapUpdate, LaserMappingLocal.cpp:62-76
    HostIso tf;
    if (w.use_mapped) tf = w.mappedNew;
    else { const float zero[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; twist_to_iso(zero, tf); }   // convertTransform of the never-assigned Twist
    std::unique_ptr<LocalFrame> fr(new LocalFrame());
    fr->nc = nout[2]; fr->ns = nout[3];
    fr->corner.reserve((size_t)(fr->nc ? fr->nc : 1) * sizeof(float4));
    fr->surf.reserve((size_t)(fr->ns ? fr->ns : 1) * sizeof(float4));
    IsoArg ta; memcpy(ta.R, tf.R, 36); memcpy(ta.t, tf.t, 12);
    if (fr->nc) CM_LAUNCH(local_transform_kernel, (fr->nc + 255) / 256, 256, 0, st, (const float4*)ds[2]->p, fr->nc, ta, (float4*)fr->corner.p);
    if (fr->ns) CM_LAUNCH(local_transform_kernel, (fr->ns + 255) / 256, 256, 0, st, (const float4*)ds[3]->p, fr->ns, ta, (float4*)fr->surf.p);
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    // LocalFeatureMap::addDataFrame (:62-69) -> FrameUpdater::update (FrameUpdater.hpp:17-42) on the Isometry3d of the float pose
    double R[9], t[3];
    for (int k = 0; k < 9; k++) R[k] = (double)tf.R[k];
    for (int k = 0; k < 3; k++) t[k] = (double)tf.t[k];
    if (w.first) {
      w.first = false;
    } else {
      // delta = prev_keypose.inverse() * pose; its translation = Rp^T * t + (-(Rp^T * tp))
      double v[3];
      for (int r = 0; r < 3; r++) {
        const double a = (w.prevR[0 + r] * t[0] + w.prevR[3 + r] * t[1]) + w.prevR[6 + r] * t[2];
        const double b = -((w.prevR[0 + r] * w.prevT[0] + w.prevR[3 + r] * w.prevT[1]) + w.prevR[6 + r] * w.prevT[2]);
        v[r] = a + b;
      }
      w.accum += sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    }
    memcpy(w.prevR, R, sizeof(R)); memcpy(w.prevT, t, sizeof(t));
    fr->accum = w.accum;
    w.frames.push_back(std::move(fr));
    // LocalFeatureMap::clean (:70-82): counts the leading frames more than queue_distance_threshold (30 m) behind and erases
    // one more than it counted
    int del = 0;
    for (auto& f : w.frames) {
      if (f->accum > (w.accum - 30.0)) break;
      ++del;
    }
    if (del > 0) w.frames.erase(w.frames.begin(), w.frames.begin() + del + 1);
    return rc;
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
}

int cm_mapping_local_window_host(cm_ctx* ctx, int* n_frames, size_t* n_corner, size_t* n_surf, int* n_surround2, double* accum_distance,
                                 cm_point* corner_out, size_t cap_corner, cm_point* surf_out, size_t cap_surf) {
  if (!ctx || !ctx->local.created) return fail(ctx, CM_ERR_ARG, "cm_mapping_local_create has not been called");
  cudaSetDevice(ctx->cfg.device);
  LocalWindow& w = ctx->local;
  size_t wc = 0, wsn = 0;
  for (auto& f : w.frames) { wc += (size_t)f->nc; wsn += (size_t)f->ns; }
  if (n_frames) *n_frames = (int)w.frames.size();
  if (n_corner) *n_corner = wc;
  if (n_surf) *n_surf = wsn;
  if (n_surround2) { n_surround2[0] = w.n_surround[0]; n_surround2[1] = w.n_surround[1]; }
  if (accum_distance) *accum_distance = w.accum;
  if ((corner_out && cap_corner < wc) || (surf_out && cap_surf < wsn)) return fail(ctx, CM_ERR_CAPACITY, "output buffer too small");
  size_t oc = 0, os = 0;
  for (auto& f : w.frames) {
    if (corner_out && f->nc) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(corner_out + oc, f->corner.p, (size_t)f->nc * sizeof(cm_point), cudaMemcpyDeviceToHost, ctx->stream));
    if (surf_out && f->ns) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(surf_out + os, f->surf.p, (size_t)f->ns * sizeof(cm_point), cudaMemcpyDeviceToHost, ctx->stream));
    oc += (size_t)f->nc; os += (size_t)f->ns;
  }
  CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return CM_OK;
}

// Full pipeline: OrganisedScanRegistration::process -> (/laser_cloud_less_sharp, /laser_cloud_less_flat) ->
// LaserMapping::process, for the S streams of the context.  (The reference routes the clouds through LaserOdometry,
// which re-projects them to the sweep end; with instantaneous synthetic sweeps that projection is the identity.)
// Stage 1 (scan registration into a slot's buffers) only depends on the sweep, so it can be issued ahead of time on the
// side stream; stage 2 (mapping) consumes the slot on the context's main stream.
static void pipeline_scanreg(cm_ctx* ctx, cm_ctx::PipeSlot& slot, const float4* d_frames, int rows, int cols, cudaStream_t st, const float* d_tags = nullptr) {
  const cm_config& cfg = ctx->cfg;
  const int S = ctx->map_streams;
  const int cap = rows * cols;
  for (int k = 0; k < 4; k++) slot.pts[k].reserve((size_t)S * cap * sizeof(float4));
  slot.n.reserve(sizeof(int) * 7 * S);
  ScanRegLaunch L;
  memset(&L, 0, sizeof(L));
  L.nstreams = S; L.rows = rows; L.cols = cols; L.frames = d_frames;
  L.tags = d_tags;   // ring + relTime of every point (raw sweeps: computed by the front end); NULL: from the column index
  fill_scanreg_params(cfg, L);
  L.blind_sq_override = d_tags ? 0.f : -1.f;   // raw sweeps: MultiScanRegistration has no blind radius (its front end dropped r^2 < 1e-4 already)
  L.prof = &ctx->prof_sr;
  for (int k = 0; k < 4; k++) { L.out_pts[k] = (float4*)slot.pts[k].p; L.cap[k] = cap; }
  L.out_n = (int*)slot.n.p + 2 * S;   // [S][5], after the [2][S] count rows the mapping stage reads
  slot.scanreg.run(L, st);
  CM_LAUNCH(gather_counts_kernel, (S + 63) / 64, 64, 0, st, (const int*)slot.n.p + 2 * S, (int*)slot.n.p, S);
}

static int pipeline_mapping(cm_ctx* ctx, cm_ctx::PipeSlot& slot, int rows, int cols, const cm_iso* odom, cm_iso* mapped, cm_match_stats* stats) {
  const int S = ctx->map_streams;
  const int cap = rows * cols;
  cudaStream_t st = ctx->stream;
  // feature-cloud sizes: upper bounds for the scratch of the frame voxel filters (and the byte accounting)
  std::vector<int> n5v;
  const int* n5;
  if (slot.counts_ready && (size_t)S * cap <= ((size_t)4 << 20) && cudaEventQuery(slot.done) != cudaSuccess) {
    // A synchronous caller with a small batch (the latency case): scan registration is still running.  The feature counts only
    // size scratch memory and bound the launch estimates, so the step is enqueued behind it with the capacity as the bound instead
    // of waiting for the counts on the host (~15 us of idle GPU per sweep); the counters are filled in once the step is through.
    slot.counts_ready = false;
    const int rc = mapping_process_dev(ctx, (const float4*)slot.pts[1].p, cap, (const float4*)slot.pts[3].p, cap, (const int*)slot.n.p, cap,
                                       cap, odom, mapped, stats, false);
    CM_CUDA_CHECK(ctx, cudaEventSynchronize(slot.done));   // (complete: the step waited for it on the device)
    ctx->last_features = 0;
    for (int s = 0; s < S; s++)
      for (int k = 0; k < 4; k++) ctx->last_features += (unsigned long long)slot.h_n5[s * 5 + k];
    return rc;
  }
  if (slot.counts_ready) {
    // read back by the prefetch on the side stream: normally complete long before the step starts
    CM_CUDA_CHECK(ctx, cudaEventSynchronize(slot.done));
    n5 = slot.h_n5;
    slot.counts_ready = false;
  } else {
    n5v.resize(5 * S);
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(n5v.data(), (const int*)slot.n.p + 2 * S, sizeof(int) * 5 * S, cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    n5 = n5v.data();
  }
  int max_c = 1, max_s = 1;
  ctx->last_features = 0;
  for (int s = 0; s < S; s++) {
    max_c = std::max(max_c, n5[s * 5 + 1]); max_s = std::max(max_s, n5[s * 5 + 3]);
    for (int k = 0; k < 4; k++) ctx->last_features += (unsigned long long)n5[s * 5 + k];
  }
  return mapping_process_dev(ctx, (const float4*)slot.pts[1].p, cap, (const float4*)slot.pts[3].p, cap, (const int*)slot.n.p, max_c,
                             max_s, odom, mapped, stats, false);
}

// clouds != NULL: one strided host cloud per stream (cm_pipeline_prefetch_strided_host); `frames` is then the slot key, clouds[0]
static int pipeline_prefetch(cm_ctx* ctx, const void* frames, int rows, int cols, bool is_host, const void* const* clouds, size_t stride) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!frames || rows <= 0 || cols <= 0 || cols > 65535) return fail(ctx, CM_ERR_ARG, "bad argument");
  if (scanreg_smem_bytes(cols) > 220 * 1024) return fail(ctx, CM_ERR_UNSUPPORTED, "cols too large for one CTA per ring");
  try {
    cudaSetDevice(ctx->cfg.device);
    if (!ctx->side_stream) CM_CUDA_CHECK(ctx, cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, ctx->prio_low));
    if (!ctx->copy_stream) CM_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->copy_stream2) CM_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking));
    int si = -1;
    for (int i = 0; i < CM_PIPE_SLOTS; i++) if (!ctx->pipe[i].src) { si = i; break; }
    if (si < 0) return fail(ctx, CM_ERR_ARG, "four sweeps are already pending: run cm_pipeline_step on one of them first");
    cm_ctx::PipeSlot& slot = ctx->pipe[si];
    if (!slot.done) CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&slot.done, cudaEventDisableTiming));
    if (!slot.copied) CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&slot.copied, cudaEventDisableTiming));
    if (!slot.copied2) CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&slot.copied2, cudaEventDisableTiming));
    // a free slot fed a step that has returned (the pipeline entries synchronise): nothing reads its buffers any more
    const float4* d_frames = (const float4*)frames;
    if (clouds) {
      const int rc = stage_upload_strided(ctx, slot, clouds, stride, rows, cols, ctx->side_stream);
      if (rc != CM_OK) return rc;
      slot.frames_valid = 0;
      d_frames = (const float4*)slot.frames.p;
    } else if (is_host) {
      const size_t bytes = (size_t)ctx->map_streams * rows * cols * sizeof(cm_point);
      slot.frames.reserve(bytes);
      // upload on its own streams: the copy of sweep k+2 runs while sweep k+1 is in scan registration
      // (in NC parts on NC copy streams: two concurrent transfers move ~1.5x the bytes per second of one)
      static const int NC = []() { const char* e = getenv("COOPERMAP_COPY_STREAMS"); int n = e ? atoi(e) : 2; return n < 1 ? 1 : (n > 4 ? 4 : n); }();
      for (int c = 2; c < NC; c++) {
        if (!ctx->copy_stream_x[c - 2]) CM_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream_x[c - 2], cudaStreamNonBlocking));
        if (!slot.copied_x[c - 2]) CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&slot.copied_x[c - 2], cudaEventDisableTiming));
      }
      cudaStream_t cs[4] = {ctx->copy_stream, ctx->copy_stream2, ctx->copy_stream_x[0], ctx->copy_stream_x[1]};
      cudaEvent_t ce[4] = {slot.copied, slot.copied2, slot.copied_x[0], slot.copied_x[1]};
      const size_t part = (bytes / NC) & ~(size_t)255;
      static const bool skip_upload = getenv("COOPERMAP_SKIP_UPLOAD") != nullptr;   // development aid: time the pipeline without PCIe
      if (!skip_upload || slot.frames_valid != bytes) {
        for (int c = 0; c < NC; c++) {
          const size_t off = (size_t)c * part, len = (c == NC - 1) ? bytes - off : part;
          CM_TIMED("h2d_upload(part)", cs[c],
                   CM_CUDA_CHECK(ctx, cudaMemcpyAsync((char*)slot.frames.p + off, (const char*)frames + off, len, cudaMemcpyHostToDevice, cs[c])));
        }
        slot.frames_valid = bytes;
      }
      for (int c = 0; c < NC; c++) {
        CM_CUDA_CHECK(ctx, cudaEventRecord(ce[c], cs[c]));
        CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->side_stream, ce[c], 0));
      }
      d_frames = (const float4*)slot.frames.p;
    }
    pipeline_scanreg(ctx, slot, d_frames, rows, cols, ctx->side_stream);
    {
      const int S = ctx->map_streams;
      if (slot.h_streams < S) {
        if (slot.h_n5) cudaFreeHost(slot.h_n5);
        slot.h_n5 = nullptr; slot.h_streams = 0;
        CM_CUDA_CHECK(ctx, cudaHostAlloc((void**)&slot.h_n5, sizeof(int) * 5 * S, cudaHostAllocDefault));
        slot.h_streams = S;
      }
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(slot.h_n5, (const int*)slot.n.p + 2 * S, sizeof(int) * 5 * S, cudaMemcpyDeviceToHost, ctx->side_stream));
      slot.counts_ready = true;
    }
    CM_CUDA_CHECK(ctx, cudaEventRecord(slot.done, ctx->side_stream));
    slot.src = frames; slot.rows = rows; slot.cols = cols; slot.is_host = is_host;
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

static int pipeline_step(cm_ctx* ctx, const void* frames, int rows, int cols, bool is_host, const cm_iso* odom, cm_iso* mapped,
                         cm_match_stats* stats, const void* const* clouds = nullptr, size_t stride = 0) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!frames || !odom || rows <= 0 || cols <= 0 || cols > 65535) return fail(ctx, CM_ERR_ARG, "bad argument");
  if (scanreg_smem_bytes(cols) > 220 * 1024) return fail(ctx, CM_ERR_UNSUPPORTED, "cols too large for one CTA per ring");
  try {
    cudaSetDevice(ctx->cfg.device);
    for (int i = 0; i < CM_PIPE_SLOTS; i++) {
      cm_ctx::PipeSlot& slot = ctx->pipe[i];
      if (slot.src == frames && slot.rows == rows && slot.cols == cols && slot.is_host == is_host) {
        // prefetched: its upload and scan registration were issued on the side stream; wait for them on the device
        CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, slot.done, 0));
        ctx->defer_rc = 0;
        const int rc = pipeline_mapping(ctx, slot, rows, cols, odom, mapped, stats);
        slot.src = nullptr;
        issue_deferred_prefetch(ctx);   // (only still pending when the step failed before its Gauss-Newton loop)
        return (rc >= 0 && ctx->defer_rc < 0) ? ctx->defer_rc : rc;
      }
    }
    // not prefetched: when a slot is free, the sweep takes the prefetch route all the same -- its upload and scan registration go
    // to the copy / side streams, where they do not queue behind the previous step's map insertion (still running on
    // ctx->stream when a synchronous caller comes back with the next sweep)
    for (int i = 0; i < CM_PIPE_SLOTS; i++)
      if (!ctx->pipe[i].src) {
        const int prc = pipeline_prefetch(ctx, frames, rows, cols, is_host, clouds, stride);
        if (prc != CM_OK) return prc;
        cm_ctx::PipeSlot& ps = ctx->pipe[i];
        CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ps.done, 0));
        ctx->defer_rc = 0;
        const int rc = pipeline_mapping(ctx, ps, rows, cols, odom, mapped, stats);
        ps.src = nullptr;
        issue_deferred_prefetch(ctx);
        return (rc >= 0 && ctx->defer_rc < 0) ? ctx->defer_rc : rc;
      }
    cm_ctx::PipeSlot& slot = ctx->pipe[CM_PIPE_SLOTS];
    const float4* d_frames = (const float4*)frames;
    if (clouds) {
      if (!ctx->copy_stream) CM_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      if (!ctx->copy_stream2) CM_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking));
      if (!slot.copied) CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&slot.copied, cudaEventDisableTiming));
      if (!slot.copied2) CM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&slot.copied2, cudaEventDisableTiming));
      const int rc = stage_upload_strided(ctx, slot, clouds, stride, rows, cols, ctx->stream);
      if (rc != CM_OK) return rc;
      d_frames = (const float4*)slot.frames.p;
    } else if (is_host) {
      const size_t bytes = (size_t)ctx->map_streams * rows * cols * sizeof(cm_point);
      slot.frames.reserve(bytes);
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(slot.frames.p, frames, bytes, cudaMemcpyHostToDevice, ctx->stream));
      d_frames = (const float4*)slot.frames.p;
    }
    pipeline_scanreg(ctx, slot, d_frames, rows, cols, ctx->stream);
    slot.counts_ready = false;
    ctx->defer_rc = 0;
    const int rc = pipeline_mapping(ctx, slot, rows, cols, odom, mapped, stats);
    issue_deferred_prefetch(ctx);
    return (rc >= 0 && ctx->defer_rc < 0) ? ctx->defer_rc : rc;
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
}

// ---- the whole LOAM chain for one sweep per stream: scan registration -> laserOdometry -> laserMapping, clouds never leave the device --
// What the three nodelets exchange over ROS topics in the reference (/laser_cloud_sharp, _less_sharp, _flat, _less_flat into
// LaserOdometry; /laser_cloud_corner_last, /laser_cloud_surf_last and /laser_odom_to_init into LaserMapping) stays in device memory:
// the odometry batch reads the four clouds scan registration left in the slot, the mapping stage reads the clouds the odometry
// projected to the sweep end.  The host sees the feature counts (they size the launches) and the two poses per stream.
static int pipeline_chain_step(cm_ctx* ctx, const void* frames, int rows, int cols, bool is_host, cm_iso* odom, cm_iso* mapped,
                               cm_odom_stats* ostats, cm_match_stats* mstats) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  OdomBatch& b = ctx->obatch;
  if (b.S != ctx->map_streams || ctx->chain_rows != rows || ctx->chain_cols != cols)
    return fail(ctx, CM_ERR_ARG, "cm_pipeline_chain_create has not been called for this sweep shape");
  if (!frames || rows <= 0 || cols <= 0) return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    const int S = ctx->map_streams, cap = rows * cols;
    int si = -1;
    for (int i = 0; i < CM_PIPE_SLOTS; i++) {
      const cm_ctx::PipeSlot& ps = ctx->pipe[i];
      if (ps.src == frames && ps.rows == rows && ps.cols == cols && ps.is_host == is_host) { si = i; break; }
    }
    if (si < 0) {
      const int prc = pipeline_prefetch(ctx, frames, rows, cols, is_host, nullptr, 0);
      if (prc != CM_OK) return prc;
      for (int i = 0; i < CM_PIPE_SLOTS; i++) if (ctx->pipe[i].src == frames) { si = i; break; }
    }
    cm_ctx::PipeSlot& slot = ctx->pipe[si];
    CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, slot.done, 0));
    CM_CUDA_CHECK(ctx, cudaEventSynchronize(slot.done));   // the feature counts (read back by the prefetch)
    slot.counts_ready = false;
    std::vector<int> n4((size_t)4 * S);
    ctx->last_features = 0;
    for (int s = 0; s < S; s++)
      for (int k = 0; k < 4; k++) { n4[(size_t)k * S + s] = slot.h_n5[s * 5 + k]; ctx->last_features += (unsigned long long)slot.h_n5[s * 5 + k]; }
    for (int s = 0; s < S; s++)
      if (n4[s] > b.cap_sharp || n4[S + s] > b.cap_less_sharp || n4[2 * S + s] > b.cap_flat || n4[3 * S + s] > b.cap_less_flat)
        return fail(ctx, CM_ERR_CAPACITY, "a feature cloud exceeds the chain's capacity");
    const size_t pitch = (size_t)cap * sizeof(cm_point);
    std::vector<cm_iso> od(S);
    int rc = odometry_batch_core(ctx, slot.pts[0].p, pitch, &n4[0], slot.pts[1].p, pitch, &n4[S], slot.pts[2].p, pitch, &n4[2 * S], slot.pts[3].p,
                                 pitch, &n4[3 * S], od.data(), nullptr, nullptr, nullptr, ostats);
    slot.src = nullptr;   // the odometry stage has copied what it keeps: the slot is free for the next sweep
    if (rc != CM_OK) return rc;
    if (odom) memcpy(odom, od.data(), sizeof(cm_iso) * S);
    int max_ls = 1, max_lf = 1;
    for (int s = 0; s < S; s++) { max_ls = std::max(max_ls, n4[S + s]); max_lf = std::max(max_lf, n4[3 * S + s]); }
    // the counts of the projected clouds sit in the odometry batch's integer block: rows 5 and 6 = [2][S], the layout the mapping stage reads
    return mapping_process_dev(ctx, (const float4*)b.last_c.p, b.cap_less_sharp, (const float4*)b.last_s.p, b.cap_less_flat,
                               (const int*)b.ints.p + 5 * S, max_ls, max_lf, od.data(), mapped, mstats, false);
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
}

// The chain for ONE stream fed with RAW spinning-LiDAR sweeps (what the Velodyne driver publishes: an unorganised, azimuth-major
// cloud): MultiScanRegistration::process (front end on the device, cm_frontend.cu) -> feature extraction -> LaserOdometry ->
// LaserMapping, nothing but the sweep going up and the two poses coming down.  scan_time >= 0: de-skew with the IMU states pushed
// with cm_imu_push_host (cm_scanreg_sweep_imu_host); < 0: no IMU.
int cm_pipeline_chain_sweep_create(cm_ctx* ctx, size_t max_points) {
  if (!ctx || ctx->map_streams != 1) return fail(ctx, CM_ERR_ARG, "cm_mapping_create(ctx, 1, ...) first: the raw-sweep chain drives one stream");
  if (max_points == 0 || max_points > 0x7fffffffu) return fail(ctx, CM_ERR_ARG, "bad argument");
  const int cap = (int)max_points;
  const int rc = cm_odometry_batch_create(ctx, 1, cap, cap, cap, cap);
  if (rc != CM_OK) return rc;
  ctx->chain_rows = -1; ctx->chain_cols = cap;   // rows < 0: raw-sweep mode, chain_cols = point capacity
  return CM_OK;
}
int cm_pipeline_chain_step_sweep_host(cm_ctx* ctx, const cm_point* sweep, size_t n, int lidar, double scan_time, cm_iso* odom, cm_iso* mapped,
                                      cm_odom_stats* ostats, cm_match_stats* mstats) {
  float lo, up; int nr;
  if (!ctx || ctx->map_streams != 1 || ctx->chain_rows != -1 || ctx->obatch.S != 1)
    return fail(ctx, CM_ERR_ARG, "cm_pipeline_chain_sweep_create has not been called");
  if ((!sweep && n) || !frontend_mapper(lidar, &lo, &up, &nr) || n > (size_t)ctx->chain_cols) return fail(ctx, CM_ERR_ARG, "bad argument / sweep larger than max_points");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    OdomBatch& b = ctx->obatch;
    int rows = nr, cols = 1;
    ctx->d_sweep.reserve((n ? n : 1) * sizeof(cm_point));
    if (n) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_sweep.p, sweep, n * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 first = n ? make_float4(sweep[0].x, sweep[0].y, sweep[0].z, 0.f) : zero;
    const float4 last = n ? make_float4(sweep[n - 1].x, sweep[n - 1].y, sweep[n - 1].z, 0.f) : zero;
    ctx->frontend.run((const float4*)ctx->d_sweep.p, (int)n, first, last, lidar, ctx->cfg.scan_period, st, &rows, &cols,
                      (scan_time >= 0.0 && !ctx->imu.stamp.empty()) ? &ctx->imu : nullptr, scan_time, nullptr);
    if (scanreg_smem_bytes(cols) > 220 * 1024) return fail(ctx, CM_ERR_UNSUPPORTED, "a ring of this sweep is too long for one CTA");
    cm_ctx::PipeSlot& slot = ctx->pipe[CM_PIPE_SLOTS];
    pipeline_scanreg(ctx, slot, (const float4*)ctx->frontend.frame.p, rows, cols, st, (const float*)ctx->frontend.tags.p);
    int n5[5];
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(n5, (const int*)slot.n.p + 2, sizeof(n5), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    ctx->last_features = 0;
    for (int k = 0; k < 4; k++) ctx->last_features += (unsigned long long)n5[k];
    if (n5[0] > b.cap_sharp || n5[1] > b.cap_less_sharp || n5[2] > b.cap_flat || n5[3] > b.cap_less_flat)
      return fail(ctx, CM_ERR_CAPACITY, "a feature cloud exceeds the chain's capacity");
    const size_t pitch = (size_t)rows * cols * sizeof(cm_point);
    cm_iso od;
    int rc = odometry_batch_core(ctx, slot.pts[0].p, pitch, &n5[0], slot.pts[1].p, pitch, &n5[1], slot.pts[2].p, pitch, &n5[2], slot.pts[3].p, pitch, &n5[3],
                                 &od, nullptr, nullptr, nullptr, ostats);
    if (rc != CM_OK) return rc;
    if (odom) *odom = od;
    return mapping_process_dev(ctx, (const float4*)b.last_c.p, b.cap_less_sharp, (const float4*)b.last_s.p, b.cap_less_flat, (const int*)b.ints.p + 5,
                               std::max(n5[1], 1), std::max(n5[3], 1), &od, mapped, mstats, false);
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
}

int cm_pipeline_chain_create(cm_ctx* ctx, int rows, int cols) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (rows <= 0 || cols <= 0 || cols > 65535) return fail(ctx, CM_ERR_ARG, "bad argument");
  // every feature class can take most of a sweep (EDGE_BROKEN points all join sharp / less-sharp, every point below the curvature
  // threshold is flat, ScanRegistration.cpp:286-303): the odometry stage is sized for whole sweeps
  const long long cap = (long long)rows * cols;
  const int cs = (int)cap, cls = (int)cap, cf = (int)cap;
  const int rc = cm_odometry_batch_create(ctx, ctx->map_streams, std::max(cs, 1), std::max(cls, 1), std::max(cf, 1), (int)cap);
  if (rc != CM_OK) return rc;
  ctx->chain_rows = rows; ctx->chain_cols = cols;
  return CM_OK;
}
int cm_pipeline_chain_step_host(cm_ctx* ctx, const cm_point* frames, int rows, int cols, cm_iso* odom, cm_iso* mapped, cm_odom_stats* ostats,
                                cm_match_stats* mstats) {
  return pipeline_chain_step(ctx, frames, rows, cols, true, odom, mapped, ostats, mstats);
}
int cm_pipeline_chain_step_dev(cm_ctx* ctx, const void* d_frames, int rows, int cols, cm_iso* odom, cm_iso* mapped, cm_odom_stats* ostats,
                               cm_match_stats* mstats) {
  return pipeline_chain_step(ctx, d_frames, rows, cols, false, odom, mapped, ostats, mstats);
}

int cm_pipeline_prefetch_host(cm_ctx* ctx, const cm_point* frames, int rows, int cols) { return pipeline_prefetch(ctx, frames, rows, cols, true); }
int cm_pipeline_prefetch_dev(cm_ctx* ctx, const void* d_frames, int rows, int cols) { return pipeline_prefetch(ctx, d_frames, rows, cols, false); }
static bool strided_args_ok(cm_ctx* ctx, const void* const* clouds, size_t stride) {
  if (!clouds || stride < 12) return false;   // any stride and alignment: a PointCloud2 point_step of 22 leaves the floats unaligned
  for (int s = 0; s < ctx->map_streams; s++) if (!clouds[s]) return false;
  return true;
}
int cm_pipeline_prefetch_strided_host(cm_ctx* ctx, const void* const* clouds, size_t stride, int rows, int cols) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!strided_args_ok(ctx, clouds, stride)) return fail(ctx, CM_ERR_ARG, "bad argument");
  return pipeline_prefetch(ctx, clouds[0], rows, cols, true, clouds, stride);
}
int cm_pipeline_step_strided_host(cm_ctx* ctx, const void* const* clouds, size_t stride, int rows, int cols, const cm_iso* odom,
                                  cm_iso* mapped, cm_match_stats* stats) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!strided_args_ok(ctx, clouds, stride)) return fail(ctx, CM_ERR_ARG, "bad argument");
  return pipeline_step(ctx, clouds[0], rows, cols, true, odom, mapped, stats, clouds, stride);
}
static int pipeline_prefetch_deferred(cm_ctx* ctx, const void* frames, int rows, int cols, bool is_host) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!frames || rows <= 0 || cols <= 0 || cols > 65535) return fail(ctx, CM_ERR_ARG, "bad argument");
  if (ctx->defer_frames) return fail(ctx, CM_ERR_ARG, "a deferred prefetch is already registered: run cm_pipeline_step first");
  ctx->defer_frames = frames; ctx->defer_rows = rows; ctx->defer_cols = cols; ctx->defer_is_host = is_host;
  return CM_OK;
}
int cm_pipeline_prefetch_deferred_host(cm_ctx* ctx, const cm_point* frames, int rows, int cols) { return pipeline_prefetch_deferred(ctx, frames, rows, cols, true); }
int cm_pipeline_prefetch_deferred_dev(cm_ctx* ctx, const void* d_frames, int rows, int cols) { return pipeline_prefetch_deferred(ctx, d_frames, rows, cols, false); }
int cm_pipeline_step_host(cm_ctx* ctx, const cm_point* frames, int rows, int cols, const cm_iso* odom, cm_iso* mapped,
                          cm_match_stats* stats) {
  return pipeline_step(ctx, frames, rows, cols, true, odom, mapped, stats);
}
int cm_pipeline_step_dev(cm_ctx* ctx, const void* d_frames, int rows, int cols, const cm_iso* odom, cm_iso* mapped,
                         cm_match_stats* stats) {
  return pipeline_step(ctx, d_frames, rows, cols, false, odom, mapped, stats);
}

// FeatureMap::update(sensorPose) (FeatureMap.h:232-254) of one stream outside a stage step: shift if needed, new valid-cube window
int cm_map_update_host(cm_ctx* ctx, int stream_index, const float* sensor) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!sensor || stream_index < 0 || stream_index >= ctx->map_streams) return fail(ctx, CM_ERR_ARG, "bad argument");
  const int S = ctx->map_streams;
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    const size_t wb = sizeof(CubeWindow) * (size_t)S;
    std::vector<CubeWindow> wins(S);
    if (ctx->wins_shadow.size() == wb) memcpy(wins.data(), ctx->wins_shadow.data(), wb);
    else {   // no stream has a window yet: the others get an empty one around their current cube
      memset(wins.data(), 0, wb);
      for (int s = 0; s < S; s++) {
        const MappingStream& o = ctx->mstreams[s];
        for (int k = 0; k < 3; k++) { wins[s].origin[k] = o.origin[k]; wins[s].w0[k] = o.cur[k] - 3; }
        wins[s].dims[0] = ctx->cfg.cube_w; wins[s].dims[1] = ctx->cfg.cube_h; wins[s].dims[2] = ctx->cfg.cube_d; wins[s].cube_size = ctx->cfg.cube_size;
      }
    }
    MappingStream ms = ctx->mstreams[stream_index];
    int d[3];
    update_window(ctx->cfg, ms, sensor, wins[stream_index], d);
    if (d[0] || d[1] || d[2]) {
      if (ctx->dist.on) return fail(ctx, CM_ERR_UNSUPPORTED, "FeatureMap::shift on a sharded map");
      const bool wrong_way = d[0] > 0 || (d[0] == 0 && (d[1] > 0 || (d[1] == 0 && d[2] > 0)));
      if (wrong_way && ctx->map.cur_epoch[stream_index] >= 255) return fail(ctx, CM_ERR_UNSUPPORTED, "more than 255 wrong-way FeatureMap::shift calls");
      ctx->map.shift(stream_index, d, ms.origin, st);
      ctx->n_shifts++;
    }
    ctx->mstreams[stream_index] = ms;
    ctx->wins_shadow.assign((const unsigned char*)wins.data(), (const unsigned char*)wins.data() + wb);
    ctx->map.set_windows(wins.data(), dev_params(ctx->cfg).knn_gate, st, false);
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));   // `wins` goes out of scope
    CM_CUDA_CHECK(ctx, cudaGetLastError());
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_map_insert_host(cm_ctx* ctx, const cm_point* corner, const int* n_corner, int cap_corner, const cm_point* surf,
                       const int* n_surf, int cap_surf, const cm_iso* tf) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!corner || !surf || !n_corner || !n_surf || !tf || cap_corner <= 0 || cap_surf <= 0) return fail(ctx, CM_ERR_ARG, "bad argument");
  const int S = ctx->map_streams;
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    ctx->m_corner_in.reserve((size_t)S * cap_corner * sizeof(cm_point));
    ctx->m_surf_in.reserve((size_t)S * cap_surf * sizeof(cm_point));
    ctx->m_n_in.reserve(sizeof(int) * 2 * S);
    ctx->m_tf.reserve(sizeof(float) * 12 * S);
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_corner_in.p, corner, (size_t)S * cap_corner * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_surf_in.p, surf, (size_t)S * cap_surf * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_n_in.p, n_corner, sizeof(int) * S, cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync((int*)ctx->m_n_in.p + S, n_surf, sizeof(int) * S, cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_tf.p, tf, sizeof(float) * 12 * S, cudaMemcpyHostToDevice, st));
    ctx->map.insert(0, (const float4*)ctx->m_corner_in.p, (const int*)ctx->m_n_in.p, cap_corner, 0, nullptr, (const float*)ctx->m_tf.p, st);
    ctx->map.insert(1, (const float4*)ctx->m_surf_in.p, (const int*)ctx->m_n_in.p + S, cap_surf, 0, nullptr, (const float*)ctx->m_tf.p, st);
    int flags[8];
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(flags, ctx->map.flags.p, sizeof(flags), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    if (flags[0] || flags[2] || flags[3]) {
      cudaMemsetAsync(ctx->map.flags.p, 0, sizeof(int) * 8, st);   // reported once (see mapping_process_dev)
      cudaStreamSynchronize(st);
      if (flags[0]) return fail(ctx, CM_ERR_UNSUPPORTED, "map point outside the supported voxel range (+-65536 voxels)");
      return fail(ctx, CM_ERR_CAPACITY, "map capacity exhausted (raise max_*_points in cm_mapping_create)");
    }
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_map_export_host(cm_ctx* ctx, int stream_index, int cls, cm_point* out, int* cube_index, size_t cap, size_t* n_out) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (stream_index < 0 || stream_index >= ctx->map_streams || cls < 0 || cls > 1 || !n_out) return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    const size_t c = cap ? cap : 1;
    ctx->m_exp_pts.reserve(c * sizeof(float4)); ctx->m_exp_cube.reserve(c * sizeof(int)); ctx->m_exp_n.reserve(sizeof(unsigned int));
    ctx->map.export_points(cls, stream_index, (float4*)ctx->m_exp_pts.p, (int*)ctx->m_exp_cube.p, (unsigned int*)ctx->m_exp_n.p,
                           (unsigned int)cap, st);
    unsigned int n = 0;
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(&n, ctx->m_exp_n.p, sizeof(n), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    *n_out = n;
    size_t m = n < cap ? n : cap;
    if (m && out) CM_CUDA_CHECK(ctx, cudaMemcpy(out, ctx->m_exp_pts.p, m * sizeof(cm_point), cudaMemcpyDeviceToHost));
    if (m && cube_index) CM_CUDA_CHECK(ctx, cudaMemcpy(cube_index, ctx->m_exp_cube.p, m * sizeof(int), cudaMemcpyDeviceToHost));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

/* ---- "map cloud out": FeatureMap::getSurroundFeature (FeatureMap.h:256-265) and getFullMap (:267-287) ------------------------------
 * The map lives in a hash of cells; the reference's clouds are ordered: valid cubes in (i, j, k) loop order (:308-352), inside a
 * cube the output order of pcl::VoxelGrid (by voxel index: z, then y, then x).  The export is ordered on the host: this is the
 * publishing path (every _mapFrameNum-th frame / a service call, LaserMatcher.cpp:164-188, 357-394), not the per-sweep path. */
namespace {
struct ExportedCloud { std::vector<float4> pts; std::vector<int> cube; };
int export_class(cm_ctx* ctx, int stream_index, int cls, ExportedCloud& out) {
  cudaStream_t st = ctx->stream;
  ctx->m_exp_n.reserve(sizeof(unsigned int));
  ctx->m_exp_pts.reserve(sizeof(float4)); ctx->m_exp_cube.reserve(sizeof(int));
  unsigned int n = 0;
  ctx->map.export_points(cls, stream_index, (float4*)ctx->m_exp_pts.p, (int*)ctx->m_exp_cube.p, (unsigned int*)ctx->m_exp_n.p, 0u, st);   // count
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(&n, ctx->m_exp_n.p, sizeof(n), cudaMemcpyDeviceToHost, st));
  CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  out.pts.resize(n); out.cube.resize(n);
  if (!n) return CM_OK;
  ctx->m_exp_pts.reserve((size_t)n * sizeof(float4)); ctx->m_exp_cube.reserve((size_t)n * sizeof(int));
  ctx->map.export_points(cls, stream_index, (float4*)ctx->m_exp_pts.p, (int*)ctx->m_exp_cube.p, (unsigned int*)ctx->m_exp_n.p, n, st);
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out.pts.data(), ctx->m_exp_pts.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, st));
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out.cube.data(), ctx->m_exp_cube.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  CM_CUDA_CHECK(ctx, cudaGetLastError());
  return CM_OK;
}
// order of a cube's cloud after downsizeValidCloud: pcl::VoxelGrid output order = ascending voxel index (z, y, x)
struct VoxOrder {
  float inv;
  bool operator()(const float4& a, const float4& b) const {
    const float za = floorf(a.z * inv), zb = floorf(b.z * inv); if (za != zb) return za < zb;
    const float ya = floorf(a.y * inv), yb = floorf(b.y * inv); if (ya != yb) return ya < yb;
    return floorf(a.x * inv) < floorf(b.x * inv);
  }
};
}  // namespace

int cm_map_surround_host(cm_ctx* ctx, int stream_index, cm_point* out_corner, size_t cap_corner, cm_point* out_surf, size_t cap_surf, size_t* n_out2) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (stream_index < 0 || stream_index >= ctx->map_streams || !n_out2) return fail(ctx, CM_ERR_ARG, "bad argument");
  const size_t wb = sizeof(CubeWindow) * (size_t)ctx->map_streams;
  n_out2[0] = n_out2[1] = 0;
  if (ctx->wins_shadow.size() != wb) return CM_OK;   // no FeatureMap::update yet: _cubeValidInd is empty
  try {
    cudaSetDevice(ctx->cfg.device);
    const CubeWindow& w = ((const CubeWindow*)ctx->wins_shadow.data())[stream_index];
    const int W = w.dims[0], H = w.dims[1];
    for (int cls = 0; cls < 2; cls++) {
      ExportedCloud e;
      const int rc = export_class(ctx, stream_index, cls, e);
      if (rc < 0) return rc;
      const float leaf = cls == 0 ? ctx->cfg.map_filter_corner : ctx->cfg.map_filter_surf;
      VoxOrder vo{1.0f / leaf};
      // valid cubes in the order computeActiveAera pushes them: i outermost, k innermost (FeatureMap.h:311-349)
      struct Item { long long cube_rank; float4 p; };
      std::vector<Item> items;
      items.reserve(e.pts.size());
      for (size_t q = 0; q < e.pts.size(); q++) {
        const int c = e.cube[q];
        const int i = c % W, j = (c / W) % H, k = c / (W * H);
        const int wi = i - w.w0[0], wj = j - w.w0[1], wk = k - w.w0[2];
        if (wi < 0 || wi > 6 || wj < 0 || wj > 6 || wk < 0 || wk > 6 || !w.active[(wi * 7 + wj) * 7 + wk]) continue;
        items.push_back(Item{((long long)wi * 7 + wj) * 7 + wk, e.pts[q]});
      }
      std::stable_sort(items.begin(), items.end(), [&](const Item& a, const Item& b) {
        if (a.cube_rank != b.cube_rank) return a.cube_rank < b.cube_rank;
        return vo(a.p, b.p);
      });
      n_out2[cls] = items.size();
      cm_point* out = cls == 0 ? out_corner : out_surf;
      const size_t cap = cls == 0 ? cap_corner : cap_surf;
      if (out) for (size_t q = 0; q < items.size() && q < cap; q++) out[q] = cm_point{items[q].p.x, items[q].p.y, items[q].p.z, items[q].p.w};
    }
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_map_full_host(cm_ctx* ctx, int stream_index, float leaf, cm_point* out, size_t cap, size_t* n_out) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (stream_index < 0 || stream_index >= ctx->map_streams || !n_out || !(leaf > 0.f)) return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    // getFullMap: for every cube in index order: VoxelGrid(leaf) of its corner cloud, then of its surf cloud
    ExportedCloud e[2];
    for (int cls = 0; cls < 2; cls++) { const int rc = export_class(ctx, stream_index, cls, e[cls]); if (rc < 0) return rc; }
    struct Seg { int cube, cls; size_t first, count; };
    std::vector<float4> sorted[2];
    std::vector<Seg> segs;
    for (int cls = 0; cls < 2; cls++) {
      const float mleaf = cls == 0 ? ctx->cfg.map_filter_corner : ctx->cfg.map_filter_surf;
      VoxOrder vo{1.0f / mleaf};
      std::vector<size_t> order(e[cls].pts.size());
      for (size_t q = 0; q < order.size(); q++) order[q] = q;
      std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        if (e[cls].cube[a] != e[cls].cube[b]) return e[cls].cube[a] < e[cls].cube[b];
        return vo(e[cls].pts[a], e[cls].pts[b]);
      });
      sorted[cls].resize(order.size());
      for (size_t q = 0; q < order.size(); q++) {
        sorted[cls][q] = e[cls].pts[order[q]];
        const int c = e[cls].cube[order[q]];
        if (q == 0 || e[cls].cube[order[q - 1]] != c) segs.push_back(Seg{c, cls, q, 0});
        segs.back().count++;
      }
    }
    std::stable_sort(segs.begin(), segs.end(), [](const Seg& a, const Seg& b) { return a.cube != b.cube ? a.cube < b.cube : a.cls < b.cls; });
    *n_out = 0;
    if (segs.empty()) return CM_OK;
    size_t cap_in = 1;
    for (const Seg& sg : segs) cap_in = std::max(cap_in, sg.count);
    const int nseg = (int)segs.size();
    std::vector<float4> packed((size_t)nseg * cap_in);
    std::vector<int> n_in(nseg);
    for (int q = 0; q < nseg; q++) {
      n_in[q] = (int)segs[q].count;
      memcpy(&packed[(size_t)q * cap_in], &sorted[segs[q].cls][segs[q].first], segs[q].count * sizeof(float4));
    }
    ctx->d_vin.reserve(packed.size() * sizeof(float4)); ctx->d_vout.reserve(packed.size() * sizeof(float4));
    ctx->d_vn_in.reserve(sizeof(int) * nseg); ctx->d_vn_out.reserve(sizeof(int) * nseg); ctx->d_flag.reserve(sizeof(int));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_vin.p, packed.data(), packed.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_vn_in.p, n_in.data(), sizeof(int) * nseg, cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_flag.p, 0, sizeof(int), st));
    ctx->voxel.run(nseg, (const float4*)ctx->d_vin.p, (const int*)ctx->d_vn_in.p, (int)cap_in, (int)cap_in, leaf, (float4*)ctx->d_vout.p,
                   (int*)ctx->d_vn_out.p, (int)cap_in, (int*)ctx->d_flag.p, st);
    std::vector<int> n_o(nseg);
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(n_o.data(), ctx->d_vn_out.p, sizeof(int) * nseg, cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(packed.data(), ctx->d_vout.p, packed.size() * sizeof(float4), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    size_t o = 0;
    for (int q = 0; q < nseg; q++)
      for (int r = 0; r < n_o[q]; r++, o++)
        if (out && o < cap) { const float4 p = packed[(size_t)q * cap_in + r]; out[o] = cm_point{p.x, p.y, p.z, p.w}; }
    *n_out = o;
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

/* ---- measurement helpers (bench.py) -------------------------------------------------------------------------------- */
int cm_timer_record(cm_ctx* ctx, int which) {
  if (!ctx || which < 0 || which > 1) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  if (!ctx->timer[which]) CM_CUDA_CHECK(ctx, cudaEventCreate(&ctx->timer[which]));
  CM_CUDA_CHECK(ctx, cudaEventRecord(ctx->timer[which], ctx->stream));
  return CM_OK;
}
/* stage-alone measurements (bench.py): events on the SIDE stream, where prefetched scan registration runs; cm_pipeline_wait blocks
 * until every prefetched sweep is uploaded and registered; cm_pipeline_discard drops a prefetched sweep without running its step */
int cm_timer_record_side(cm_ctx* ctx, int which) {
  if (!ctx || which < 0 || which > 1) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  if (!ctx->side_stream) CM_CUDA_CHECK(ctx, cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, ctx->prio_low));
  if (!ctx->timer[which]) CM_CUDA_CHECK(ctx, cudaEventCreate(&ctx->timer[which]));
  CM_CUDA_CHECK(ctx, cudaEventRecord(ctx->timer[which], ctx->side_stream));
  return CM_OK;
}
int cm_pipeline_wait(cm_ctx* ctx) {
  if (!ctx) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  for (int i = 0; i < CM_PIPE_SLOTS; i++)
    if (ctx->pipe[i].src && ctx->pipe[i].done) CM_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->pipe[i].done));
  return CM_OK;
}
/* cm_mapping_process_* / cm_pipeline_step_* return when the poses are on the host; the map insertion they enqueued may still be
 * running.  This waits for it and reports what it hit (capacity, voxel range) -- otherwise the next step reports it. */
int cm_mapping_sync(cm_ctx* ctx) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  try {
    cudaSetDevice(ctx->cfg.device);
    int flags[8];
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(flags, ctx->map.flags.p, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (flags[0] || flags[2] || flags[3]) {
      cudaMemsetAsync(ctx->map.flags.p, 0, sizeof(int) * 8, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      if (flags[0]) return fail(ctx, CM_ERR_UNSUPPORTED, "map point outside the supported voxel range (+-65536 voxels)");
      return fail(ctx, CM_ERR_CAPACITY, "map capacity exhausted (raise max_*_points in cm_mapping_create)");
    }
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}
int cm_pipeline_discard(cm_ctx* ctx, const void* frames) {
  if (!ctx || !frames) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  for (int i = 0; i < CM_PIPE_SLOTS; i++)
    if (ctx->pipe[i].src == frames) {
      if (ctx->pipe[i].done) CM_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->pipe[i].done));
      ctx->pipe[i].src = nullptr; ctx->pipe[i].counts_ready = false;
      return CM_OK;
    }
  return fail(ctx, CM_ERR_ARG, "no prefetched sweep with this address");
}
int cm_timer_elapsed_ms(cm_ctx* ctx, float* ms) {
  if (!ctx || !ms || !ctx->timer[0] || !ctx->timer[1]) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  CM_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->timer[1]));
  CM_CUDA_CHECK(ctx, cudaEventElapsedTime(ms, ctx->timer[0], ctx->timer[1]));
  return CM_OK;
}
/* development aid: per-warp trace of search_kernel (4 words per warp: t0 ns, t1 ns, max<<32|sum candidates, hard<<32|smid<<16|corner)
 * of Gauss-Newton evaluation `iter` of the following cm_pipeline_step / cm_mapping_process calls; iter < 0 switches it off. */
int cm_debug_search_trace_enable(cm_ctx* ctx, int iter) { if (!ctx) return CM_ERR_ARG; ctx->dbg_on = iter >= 0; ctx->dbg_iter = iter; return CM_OK; }
int cm_debug_search_trace_read(cm_ctx* ctx, unsigned long long* out, size_t cap_words, size_t* n_words) {
  if (!ctx || !n_words) return CM_ERR_ARG;
  *n_words = ctx->dbg_words;
  if (out && ctx->dbg_words) {
    cudaSetDevice(ctx->cfg.device);
    CM_CUDA_CHECK(ctx, cudaMemcpy(out, ctx->dbg_trace.p, sizeof(unsigned long long) * (cap_words < ctx->dbg_words ? cap_words : ctx->dbg_words), cudaMemcpyDeviceToHost));
  }
  return CM_OK;
}
/* development aid: how the mapping stage's Gauss-Newton loop is being submitted: n_graphs cached CUDA graphs, while_loop = 1 when
 * they are conditional WHILE graphs (0: unrolled max_iterations graphs, the fallback) */
int cm_debug_graph_info(cm_ctx* ctx, int* n_graphs, int* while_loop) {
  if (!ctx) return CM_ERR_ARG;
  if (n_graphs) *n_graphs = (int)ctx->match_graphs.entries.size();
  ctx->dbg_graph_builds = ctx->match_graphs.builds; ctx->dbg_stage_captures = ctx->stage_graphs.captures;
  if (while_loop) *while_loop = ctx->match_graphs.use_while ? 1 : 0;
  return CM_OK;
}
/* [0] Gauss-Newton loop graphs built so far, [1] map-insert chain captures (a steady-state pipeline stops building), [2] steps whose
 * map insertion was repeated with exact sizes (the estimate of the filtered counts was too small), [3] reserved */
int cm_debug_graph_builds(cm_ctx* ctx, unsigned long long* out4) {
  if (!ctx || !out4) return CM_ERR_ARG;
  out4[0] = ctx->match_graphs.builds; out4[1] = ctx->stage_graphs.captures; out4[2] = ctx->insert_redos; out4[3] = 0;
  return CM_OK;
}
/* development aids: the neighbour slots (5 per query row, [stream][cap_corner + cap_surf][5]) the last mapping step's final search
 * left behind, the filtered query clouds, and map points by pool slot */
int cm_debug_read_slots(cm_ctx* ctx, int* out, size_t n_ints) {
  if (!ctx || !out) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  CM_CUDA_CHECK(ctx, cudaMemcpy(out, ctx->m_slots.p, std::min(n_ints * sizeof(int), ctx->m_slots.cap), cudaMemcpyDeviceToHost));
  return CM_OK;
}
int cm_debug_read_queries(cm_ctx* ctx, int cls, float* out, size_t n_floats, int* counts) {
  if (!ctx || !out || cls < 0 || cls > 1) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  cm::DeviceBuffer& b = cls == 0 ? ctx->m_corner_ds : ctx->m_surf_ds;
  CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  CM_CUDA_CHECK(ctx, cudaMemcpy(out, b.p, std::min(n_floats * sizeof(float), b.cap), cudaMemcpyDeviceToHost));
  if (counts) CM_CUDA_CHECK(ctx, cudaMemcpy(counts, (const int*)ctx->m_n_ds.p + cls * ctx->map_streams, sizeof(int) * ctx->map_streams, cudaMemcpyDeviceToHost));
  return CM_OK;
}
int cm_debug_read_map_points(cm_ctx* ctx, int stream_index, int cls, const int* slots, int n, float* out4) {
  if (!ctx || !slots || !out4 || cls < 0 || cls > 1 || stream_index < 0 || stream_index >= ctx->map_streams) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  const float4* pool = (const float4*)ctx->map.pts[cls].p + (size_t)stream_index * ctx->map.pool_cap[cls];
  CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; i++) {
    if (slots[i] < 0 || (unsigned int)slots[i] >= ctx->map.pool_cap[cls]) { out4[4 * i] = out4[4 * i + 1] = out4[4 * i + 2] = out4[4 * i + 3] = nanf(""); continue; }
    CM_CUDA_CHECK(ctx, cudaMemcpy(out4 + 4 * i, pool + slots[i], sizeof(float4), cudaMemcpyDeviceToHost));
  }
  return CM_OK;
}
int cm_prof_enable(cm_ctx* ctx, int on) { if (!ctx) return CM_ERR_ARG; ctx->prof.enabled = ctx->prof_sr.enabled = on != 0; return CM_OK; }
int cm_prof_drain_scanreg(cm_ctx* ctx, double* kernel_ms, int* launches) {
  if (!ctx || !kernel_ms) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->side_stream) CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->side_stream));
  *kernel_ms = ctx->prof_sr.drain_ms(launches);
  return CM_OK;
}
int cm_prof_drain(cm_ctx* ctx, double* kernel_ms, int* launches) {
  if (!ctx || !kernel_ms) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  *kernel_ms = ctx->prof.drain_ms(launches);
  return CM_OK;
}
int cm_last_step_counters(cm_ctx* ctx, unsigned long long* out4) {
  if (!ctx || !out4) return CM_ERR_ARG;
  out4[0] = ctx->last_query_iters; out4[1] = ctx->last_queries; out4[2] = ctx->last_inserted; out4[3] = ctx->last_features;
  return CM_OK;
}

}  // extern "C"
