// cm_odometry.cu -- the scan-to-scan odometry STAGE: host mirror of LaserOdometry::process
// (L_SLAM/src/odometry/LaserOdometry.cpp:288-326) around the K8 kernels (cm_odom.inl) and the shared reduce / solve kernels.
#include "cm_ctx.h"
#include "cm_math.h"
#include <math.h>
#include <string.h>
#include <algorithm>

using namespace cm;

// cell size of the voxel-cell hash over the last clouds (the odometry's 1-NN search with a 5 m gate): development override
static float odom_cell(bool corner) {
  static const float c = getenv("COOPERMAP_ODOM_CELL_C") ? (float)atof(getenv("COOPERMAP_ODOM_CELL_C")) : 2.5f;
  static const float f = getenv("COOPERMAP_ODOM_CELL_S") ? (float)atof(getenv("COOPERMAP_ODOM_CELL_S")) : 2.5f;
  return corner ? c : f;
}
static int fail(cm_ctx* ctx, int code, const std::string& msg) { return ctx_fail(ctx, code, msg); }

static HostIso h_identity() { HostIso i; for (int k = 0; k < 9; k++) i.R[k] = (k % 4 == 0) ? 1.f : 0.f; i.t[0] = i.t[1] = i.t[2] = 0.f; return i; }
static HostIso h_mul(const HostIso& a, const HostIso& b) {
  HostIso r;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) r.R[i * 3 + j] = (a.R[i * 3 + 0] * b.R[0 * 3 + j] + a.R[i * 3 + 1] * b.R[1 * 3 + j]) + a.R[i * 3 + 2] * b.R[2 * 3 + j];
    r.t[i] = ((a.R[i * 3 + 0] * b.t[0] + a.R[i * 3 + 1] * b.t[1]) + a.R[i * 3 + 2] * b.t[2]) + a.t[i];
  }
  return r;
}
static HostIso h_inverse(const HostIso& a) {
  HostIso r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; i++) r.t[i] = -((r.R[i * 3 + 0] * a.t[0] + r.R[i * 3 + 1] * a.t[1]) + r.R[i * 3 + 2] * a.t[2]);
  return r;
}

extern "C" {

int cm_odometry_reset(cm_ctx* ctx) {
  if (!ctx) return CM_ERR_ARG;
  ctx->odom_inited = false;
  for (int k = 0; k < 6; k++) ctx->odom_tf[k] = 0.f;
  ctx->odom_Tsum = h_identity();
  ctx->o_n_last_c = ctx->o_n_last_s = 0;
  return CM_OK;
}

int cm_odometry_process_host(cm_ctx* ctx, const cm_point* sharp, int n_sharp, const cm_point* less_sharp, int n_less_sharp,
                             const cm_point* flat, int n_flat, const cm_point* less_flat, int n_less_flat, cm_iso* odom,
                             cm_pose* transform, cm_point* corner_last, cm_point* surf_last, cm_odom_stats* stats,
                             cm_iter_trace* trace) {
  if (!ctx || n_sharp < 0 || n_less_sharp < 0 || n_flat < 0 || n_less_flat < 0 || (!sharp && n_sharp) || (!less_sharp && n_less_sharp) ||
      (!flat && n_flat) || (!less_flat && n_less_flat))
    return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    if (!ctx->odom_inited && ctx->o_n_last_c == 0 && ctx->o_n_last_s == 0) ctx->odom_Tsum = h_identity();
    cm_odom_stats local;
    memset(&local, 0, sizeof(local));
    const int MAXIT = 25;   // LaserOdometry.cpp:24
    bool first = !ctx->odom_inited;
    if (!first && ctx->o_n_last_c > 10 && ctx->o_n_last_s > 100) {   // LaserOdometry.cpp:338
      // ---- scanMatch ----
      MatchParamsDev prm;
      prm.max_iterations = MAXIT; prm.delta_t_abort = 0.1f; prm.delta_r_abort = 0.1f; prm.knn_gate = 25.f; prm.plane_max_dist = 0.f;
      prm.min_ref_corner = 0; prm.min_ref_surf = 0; prm.min_rows = 10; prm.eig_threshold = 10.f; prm.few_rows_continue = 1; prm.nan_guard = 1; prm.own_cube_only = 0;
      const int capC = std::max(n_sharp, 1), capS = std::max(n_flat, 1), capQ = capC + capS;
      ctx->o_sharp.reserve(capC * sizeof(cm_point)); ctx->o_flat.reserve(capS * sizeof(cm_point));
      ctx->o_ind.reserve((size_t)(2 * capC + 3 * capS) * sizeof(int));
      ctx->o_rows.reserve((size_t)capQ * sizeof(RowOut)); ctx->o_state.reserve(sizeof(MatchState)); ctx->o_sums.reserve(32 * sizeof(double));
      ctx->o_pose.reserve(6 * sizeof(float)); ctx->o_counts.reserve(2 * sizeof(int)); ctx->o_views.reserve(2 * sizeof(GridView));
      ctx->o_trace.reserve(sizeof(IterTrace) * MAXIT);
      if (n_sharp) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->o_sharp.p, sharp, n_sharp * sizeof(cm_point), cudaMemcpyHostToDevice, st));
      if (n_flat) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->o_flat.p, flat, n_flat * sizeof(cm_point), cudaMemcpyHostToDevice, st));
      CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->o_ind.p, 0xFF, (size_t)(2 * capC + 3 * capS) * sizeof(int), st));
      CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->o_rows.p, 0, (size_t)capQ * sizeof(RowOut), st));
      CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->o_trace.p, 0, sizeof(IterTrace) * MAXIT, st));
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->o_pose.p, ctx->odom_tf, 6 * sizeof(float), cudaMemcpyHostToDevice, st));
      int counts[2] = {n_sharp, n_flat};
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->o_counts.p, counts, sizeof(counts), cudaMemcpyHostToDevice, st));
      GridView views[2] = {ctx->o_grid_c.view, ctx->o_grid_s.view};
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->o_views.p, views, sizeof(views), cudaMemcpyHostToDevice, st));
      MatchLaunch m;
      m.nstreams = 1;
      m.corner = (const float4*)ctx->o_sharp.p; m.surf = (const float4*)ctx->o_flat.p;
      m.n_corner = (const int*)ctx->o_counts.p; m.n_surf = (const int*)ctx->o_counts.p + 1;
      m.cap_corner = capC; m.cap_surf = capS;
      m.grid_corner = (const GridView*)ctx->o_views.p; m.grid_surf = (const GridView*)ctx->o_views.p + 1;
      m.pose_in = (const float*)ctx->o_pose.p; m.state = (MatchState*)ctx->o_state.p; m.rows = (RowOut*)ctx->o_rows.p;
      m.nn_slot = nullptr; m.sums = (double*)ctx->o_sums.p; m.trace = (IterTrace*)ctx->o_trace.p; m.nn = nullptr;
      m.orig_idx = 1; m.max_queries = n_sharp + n_flat; m.prm = prm;
      launch_match_init(m, st);
      OdomLaunch o;
      o.sharp = m.corner; o.flat = m.surf; o.n_sharp = n_sharp; o.n_flat = n_flat;
      o.last_corner = (const float4*)ctx->o_last_c.p; o.last_surf = (const float4*)ctx->o_last_s.p;
      o.bound_corner = std::min(n_sharp, ctx->o_n_last_c); o.bound_surf = std::min(n_flat, ctx->o_n_last_s);   // LaserOdometry.cpp:370,434 (clamped)
      o.grid_corner = ctx->o_grid_c.view; o.grid_surf = ctx->o_grid_s.view;
      o.state = m.state; o.ind = (int*)ctx->o_ind.p; o.rows = m.rows;
      for (int it = 0; it < MAXIT; it++) {
        launch_odom_corr(o, it, st);
        launch_match_reduce(m, it, st);
        launch_match_solve(m, it, (const double*)m.sums, st);
      }
      MatchState hs;
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(&hs, ctx->o_state.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
      if (trace) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(trace, ctx->o_trace.p, sizeof(IterTrace) * MAXIT, cudaMemcpyDeviceToHost, st));
      CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
      CM_CUDA_CHECK(ctx, cudaGetLastError());
      for (int k = 0; k < 6; k++) ctx->odom_tf[k] = hs.pose[k];
      local.iterations = hs.iterations; local.rows = hs.rows; local.converged = (hs.flags & CM_F_CONVERGED) ? 1 : 0;
      local.degenerate = (hs.flags & CM_F_DEGENERATE) ? 1 : 0; local.matched = 1;
    }
    if (!first) {
      // transformUpdate, LaserOdometry.cpp:649-653
      HostIso update; pose_to_matrix(ctx->odom_tf, update.R); update.t[0] = ctx->odom_tf[3]; update.t[1] = ctx->odom_tf[4]; update.t[2] = ctx->odom_tf[5];
      ctx->odom_Tsum = h_mul(ctx->odom_Tsum, update);
    }
    // the less-sharp / less-flat clouds become the "last" clouds (first frame: as they are, :295-303; later: projected to
    // the sweep end, :311-315)
    ctx->o_last_c.reserve(std::max(n_less_sharp, 1) * sizeof(cm_point)); ctx->o_last_s.reserve(std::max(n_less_flat, 1) * sizeof(cm_point));
    if (n_less_sharp) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->o_last_c.p, less_sharp, n_less_sharp * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    if (n_less_flat) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->o_last_s.p, less_flat, n_less_flat * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    if (!first) {
      HostIso it; pose_to_matrix(ctx->odom_tf, it.R); it.t[0] = ctx->odom_tf[3]; it.t[1] = ctx->odom_tf[4]; it.t[2] = ctx->odom_tf[5];
      HostIso inv = h_inverse(it);
      float inv12[12]; memcpy(inv12, inv.R, 36); memcpy(inv12 + 9, inv.t, 12);
      ctx->o_tf.reserve(6 * sizeof(float)); ctx->o_inv.reserve(12 * sizeof(float));
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->o_tf.p, ctx->odom_tf, 6 * sizeof(float), cudaMemcpyHostToDevice, st));
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->o_inv.p, inv12, sizeof(inv12), cudaMemcpyHostToDevice, st));
      launch_odom_to_end((float4*)ctx->o_last_c.p, n_less_sharp, (const float*)ctx->o_tf.p, (const float*)ctx->o_inv.p, st);
      launch_odom_to_end((float4*)ctx->o_last_s.p, n_less_flat, (const float*)ctx->o_tf.p, (const float*)ctx->o_inv.p, st);
    }
    ctx->o_n_last_c = n_less_sharp; ctx->o_n_last_s = n_less_flat;
    if (first || (n_less_sharp > 10 && n_less_flat > 100)) {   // KD-trees rebuilt, :299-300, 320-323
      ctx->o_grid_c.build((const float4*)ctx->o_last_c.p, n_less_sharp, odom_cell(true), 25.f, 0, st);
      ctx->o_grid_s.build((const float4*)ctx->o_last_s.p, n_less_flat, odom_cell(false), 25.f, 0, st);
    }
    if (corner_last && n_less_sharp) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(corner_last, ctx->o_last_c.p, n_less_sharp * sizeof(cm_point), cudaMemcpyDeviceToHost, st));
    if (surf_last && n_less_flat) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(surf_last, ctx->o_last_s.p, n_less_flat * sizeof(cm_point), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->odom_inited = true;
    if (odom) { memcpy(odom->R, ctx->odom_Tsum.R, 36); memcpy(odom->t, ctx->odom_Tsum.t, 12); }
    if (transform) { transform->rx = ctx->odom_tf[0]; transform->ry = ctx->odom_tf[1]; transform->rz = ctx->odom_tf[2];
                     transform->tx = ctx->odom_tf[3]; transform->ty = ctx->odom_tf[4]; transform->tz = ctx->odom_tf[5]; }
    local.initialising = first ? 1 : 0;
    if (stats) *stats = local;
    return CM_OK;
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
}

// ---- the same stage for every stream of a context in ONE set of launches ------------------------------------------------------
// Clouds are [nstreams][cap_*] with per-stream counts.  Per stream this is exactly cm_odometry_process_host (LaserOdometry::process,
// LaserOdometry.cpp:288-326): streams on their first frame only store their clouds, streams whose last clouds are too small
// (<= 10 corner or <= 100 surf points, :338) skip scanMatch, the others iterate together -- 25 x (correspondences + rows, reduction,
// 6x6 step) = 75 launches for the whole batch instead of 75 per stream.
int cm_odometry_batch_create(cm_ctx* ctx, int nstreams, int cap_sharp, int cap_less_sharp, int cap_flat, int cap_less_flat) {
  if (!ctx || nstreams <= 0 || cap_sharp <= 0 || cap_less_sharp <= 0 || cap_flat <= 0 || cap_less_flat <= 0) return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    OdomBatch& b = ctx->obatch;
    b.S = nstreams; b.cap_sharp = cap_sharp; b.cap_less_sharp = cap_less_sharp; b.cap_flat = cap_flat; b.cap_less_flat = cap_less_flat;
    b.inited.assign(nstreams, 0); b.n_last_c.assign(nstreams, 0); b.n_last_s.assign(nstreams, 0);
    b.tf.assign((size_t)6 * nstreams, 0.f); b.Tsum.assign(nstreams, h_identity());
    const size_t S = nstreams;
    b.sharp.reserve(S * cap_sharp * sizeof(cm_point)); b.flat.reserve(S * cap_flat * sizeof(cm_point));
    b.last_c.reserve(S * cap_less_sharp * sizeof(cm_point)); b.last_s.reserve(S * cap_less_flat * sizeof(cm_point));
    b.ints.reserve(S * 10 * sizeof(int));
    b.ind.reserve(S * (2 * (size_t)cap_sharp + 3 * (size_t)cap_flat) * sizeof(int));
    b.rows.reserve(S * ((size_t)cap_sharp + cap_flat) * sizeof(RowOut));
    b.state.reserve(S * sizeof(MatchState)); b.sums.reserve(S * 32 * sizeof(double)); b.pose.reserve(S * 6 * sizeof(float));
    b.tfinv.reserve(S * 18 * sizeof(float));
    b.box_c.reserve(S * (size_t)((cap_less_sharp + 31) / 32) * 32); b.box_s.reserve(S * (size_t)((cap_less_flat + 31) / 32) * 32);   // 32 bytes per chunk of 32 points
    b.grid_c.create(nstreams, cap_less_sharp, ctx->stream); b.grid_s.create(nstreams, cap_less_flat, ctx->stream);
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

}  // extern "C"

namespace cm {
// The batch stage with its four input clouds anywhere (host or device: the copies are cudaMemcpyDefault), stream s of a cloud
// starting at src + s * pitch bytes.  cm_odometry_batch_process_host passes its packed host arrays; the pipeline chain
// (cm_pipeline_chain_step_host, cm_mapping.cu) passes the device clouds scan registration left behind.
int odometry_batch_core(cm_ctx* ctx, const void* sharp, size_t pitch_sharp, const int* n_sharp, const void* less_sharp, size_t pitch_less_sharp,
                        const int* n_less_sharp, const void* flat, size_t pitch_flat, const int* n_flat, const void* less_flat,
                        size_t pitch_less_flat, const int* n_less_flat, cm_iso* odom, cm_pose* transform, cm_point* corner_last,
                        cm_point* surf_last, cm_odom_stats* stats) {
  if (!ctx || ctx->obatch.S <= 0) return fail(ctx, CM_ERR_ARG, "cm_odometry_batch_create has not been called");
  if (!sharp || !less_sharp || !flat || !less_flat || !n_sharp || !n_less_sharp || !n_flat || !n_less_flat) return fail(ctx, CM_ERR_ARG, "bad argument");
  OdomBatch& b = ctx->obatch;
  const int S = b.S;
  for (int s = 0; s < S; s++)
    if (n_sharp[s] < 0 || n_sharp[s] > b.cap_sharp || n_less_sharp[s] < 0 || n_less_sharp[s] > b.cap_less_sharp || n_flat[s] < 0 ||
        n_flat[s] > b.cap_flat || n_less_flat[s] < 0 || n_less_flat[s] > b.cap_less_flat)
      return fail(ctx, CM_ERR_CAPACITY, "a feature cloud exceeds the capacity given to cm_odometry_batch_create");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    const int MAXIT = 25;   // LaserOdometry.cpp:24
    const int capQ = b.cap_sharp + b.cap_flat;
    // per-stream integers: [0] n_sharp [1] n_flat [2] bound_corner [3] bound_surf [4] active [5] n_less_sharp [6] n_less_flat [7] to_end [8] rebuild
    std::vector<int> hi((size_t)10 * S, 0);
    int any_active = 0, max_sharp = 1, max_flat = 1, max_ls = 0, max_lf = 0;
    for (int s = 0; s < S; s++) {
      hi[0 * S + s] = n_sharp[s]; hi[1 * S + s] = n_flat[s];
      hi[2 * S + s] = std::min(n_sharp[s], b.n_last_c[s]); hi[3 * S + s] = std::min(n_flat[s], b.n_last_s[s]);   // LaserOdometry.cpp:370,434 (clamped)
      const int active = b.inited[s] && b.n_last_c[s] > 10 && b.n_last_s[s] > 100;                                // :338
      hi[4 * S + s] = active; any_active |= active;
      hi[5 * S + s] = n_less_sharp[s]; hi[6 * S + s] = n_less_flat[s];
      hi[7 * S + s] = b.inited[s];
      hi[8 * S + s] = (!b.inited[s] || (n_less_sharp[s] > 10 && n_less_flat[s] > 100)) ? 1 : 0;                  // :299-300, 320-323
      if (active) { max_sharp = std::max(max_sharp, n_sharp[s]); max_flat = std::max(max_flat, n_flat[s]); }
      max_ls = std::max(max_ls, n_less_sharp[s]); max_lf = std::max(max_lf, n_less_flat[s]);
    }
    int* d_i = (int*)b.ints.p;
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(d_i, hi.data(), hi.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    std::vector<cm_odom_stats> local(S);
    memset(local.data(), 0, sizeof(cm_odom_stats) * S);
    std::vector<MatchState> hs(S);
    if (any_active) {
      CM_CUDA_CHECK(ctx, cudaMemcpy2DAsync(b.sharp.p, b.cap_sharp * sizeof(cm_point), sharp, pitch_sharp, (size_t)max_sharp * sizeof(cm_point), S, cudaMemcpyDefault, st));
      CM_CUDA_CHECK(ctx, cudaMemcpy2DAsync(b.flat.p, b.cap_flat * sizeof(cm_point), flat, pitch_flat, (size_t)max_flat * sizeof(cm_point), S, cudaMemcpyDefault, st));
      CM_CUDA_CHECK(ctx, cudaMemsetAsync(b.ind.p, 0xFF, (size_t)S * (2 * (size_t)b.cap_sharp + 3 * (size_t)b.cap_flat) * sizeof(int), st));
      CM_CUDA_CHECK(ctx, cudaMemsetAsync(b.rows.p, 0, (size_t)S * capQ * sizeof(RowOut), st));
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(b.pose.p, b.tf.data(), (size_t)6 * S * sizeof(float), cudaMemcpyHostToDevice, st));
      MatchParamsDev prm;
      prm.max_iterations = MAXIT; prm.delta_t_abort = 0.1f; prm.delta_r_abort = 0.1f; prm.knn_gate = 25.f; prm.plane_max_dist = 0.f;
      prm.min_ref_corner = 0; prm.min_ref_surf = 0; prm.min_rows = 10; prm.eig_threshold = 10.f; prm.few_rows_continue = 1; prm.nan_guard = 1; prm.own_cube_only = 0;
      MatchLaunch m;
      m.nstreams = S;
      m.corner = (const float4*)b.sharp.p; m.surf = (const float4*)b.flat.p;
      m.n_corner = d_i; m.n_surf = d_i + S; m.cap_corner = b.cap_sharp; m.cap_surf = b.cap_flat;
      m.grid_corner = (const GridView*)b.grid_c.views.p; m.grid_surf = (const GridView*)b.grid_s.views.p;
      m.pose_in = (const float*)b.pose.p; m.state = (MatchState*)b.state.p; m.rows = (RowOut*)b.rows.p;
      m.nn_slot = nullptr; m.sums = (double*)b.sums.p; m.trace = nullptr; m.nn = nullptr;
      m.orig_idx = 1; m.max_queries = max_sharp + max_flat; m.prm = prm;
      OdomBatchLaunch o;
      o.nstreams = S; o.sharp = m.corner; o.flat = m.surf; o.cap_sharp = b.cap_sharp; o.cap_flat = b.cap_flat; o.n_sharp = d_i; o.n_flat = d_i + S;
      o.max_sharp = max_sharp; o.max_flat = max_flat;
      o.last_corner = (const float4*)b.last_c.p; o.last_surf = (const float4*)b.last_s.p; o.cap_last_corner = b.cap_less_sharp; o.cap_last_surf = b.cap_less_flat;
      o.bound_corner = d_i + 2 * S; o.bound_surf = d_i + 3 * S;
      static const bool no_boxes = getenv("COOPERMAP_ODOM_NO_BOXES") != nullptr;   // development: the plain ring walks
      if (!no_boxes) { o.box_corner = b.box_c.p; o.box_surf = b.box_s.p; o.box_cap_corner = (b.cap_less_sharp + 31) / 32; o.box_cap_surf = (b.cap_less_flat + 31) / 32; }
      o.grid_corner = m.grid_corner; o.grid_surf = m.grid_surf; o.state = m.state; o.ind = (int*)b.ind.p; o.rows = m.rows;
      // launch sizes in whole tiles: the loop's graph is keyed by them and then repeats from frame to frame
      o.max_sharp = (max_sharp + 1023) & ~1023; o.max_flat = (max_flat + 1023) & ~1023;
      static const bool no_graph = getenv("COOPERMAP_NO_GRAPH") != nullptr;
      if (no_graph || !b.graphs.launch(m, o, d_i + 4 * S, st)) {
        launch_match_init(m, st);
        launch_odom_gate(m.state, d_i + 4 * S, S, st);
        for (int it = 0; it < MAXIT; it++) {
          launch_odom_corr_batch(o, it, st);
          launch_match_reduce_solve(m, it, st);
        }
      }
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(hs.data(), b.state.p, sizeof(MatchState) * S, cudaMemcpyDeviceToHost, st));
      CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
      CM_CUDA_CHECK(ctx, cudaGetLastError());
      for (int s = 0; s < S; s++) {
        if (!hi[4 * S + s]) continue;
        for (int k = 0; k < 6; k++) b.tf[6 * s + k] = hs[s].pose[k];
        local[s].iterations = hs[s].iterations; local[s].rows = hs[s].rows; local[s].converged = (hs[s].flags & CM_F_CONVERGED) ? 1 : 0;
        local[s].degenerate = (hs[s].flags & CM_F_DEGENERATE) ? 1 : 0; local[s].matched = 1;
      }
    }
    // transformUpdate (:649-653), then the new clouds become the last clouds, projected to the sweep end (:311-315)
    std::vector<float> tfinv((size_t)18 * S, 0.f);
    for (int s = 0; s < S; s++) {
      if (!b.inited[s]) continue;
      HostIso it; pose_to_matrix(&b.tf[6 * s], it.R); it.t[0] = b.tf[6 * s + 3]; it.t[1] = b.tf[6 * s + 4]; it.t[2] = b.tf[6 * s + 5];
      b.Tsum[s] = h_mul(b.Tsum[s], it);
      const HostIso inv = h_inverse(it);
      memcpy(&tfinv[(size_t)6 * s], &b.tf[6 * s], 24);
      memcpy(&tfinv[(size_t)6 * S + 12 * s], inv.R, 36); memcpy(&tfinv[(size_t)6 * S + 12 * s + 9], inv.t, 12);
    }
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(b.tfinv.p, tfinv.data(), tfinv.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    if (max_ls) CM_CUDA_CHECK(ctx, cudaMemcpy2DAsync(b.last_c.p, b.cap_less_sharp * sizeof(cm_point), less_sharp, pitch_less_sharp, (size_t)max_ls * sizeof(cm_point), S, cudaMemcpyDefault, st));
    if (max_lf) CM_CUDA_CHECK(ctx, cudaMemcpy2DAsync(b.last_s.p, b.cap_less_flat * sizeof(cm_point), less_flat, pitch_less_flat, (size_t)max_lf * sizeof(cm_point), S, cudaMemcpyDefault, st));
    const float* d_tf6 = (const float*)b.tfinv.p; const float* d_inv = d_tf6 + (size_t)6 * S;
    launch_odom_to_end_batch((float4*)b.last_c.p, b.cap_less_sharp, d_i + 5 * S, max_ls, S, d_tf6, d_inv, d_i + 7 * S, st);
    launch_odom_to_end_batch((float4*)b.last_s.p, b.cap_less_flat, d_i + 6 * S, max_lf, S, d_tf6, d_inv, d_i + 7 * S, st);
    b.grid_c.build((const float4*)b.last_c.p, d_i + 5 * S, std::max(max_ls, 1), d_i + 8 * S, odom_cell(true), 25.f, st);
    b.grid_s.build((const float4*)b.last_s.p, d_i + 6 * S, std::max(max_lf, 1), d_i + 8 * S, odom_cell(false), 25.f, st);
    launch_odom_boxes_batch((const float4*)b.last_c.p, b.cap_less_sharp, d_i + 5 * S, max_ls, S, b.box_c.p, (b.cap_less_sharp + 31) / 32, st);
    launch_odom_boxes_batch((const float4*)b.last_s.p, b.cap_less_flat, d_i + 6 * S, max_lf, S, b.box_s.p, (b.cap_less_flat + 31) / 32, st);
    if (corner_last) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(corner_last, b.last_c.p, (size_t)S * b.cap_less_sharp * sizeof(cm_point), cudaMemcpyDeviceToHost, st));
    if (surf_last) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(surf_last, b.last_s.p, (size_t)S * b.cap_less_flat * sizeof(cm_point), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    for (int s = 0; s < S; s++) {
      local[s].initialising = b.inited[s] ? 0 : 1;
      b.n_last_c[s] = n_less_sharp[s]; b.n_last_s[s] = n_less_flat[s];
      b.inited[s] = 1;
      if (odom) { memcpy(odom[s].R, b.Tsum[s].R, 36); memcpy(odom[s].t, b.Tsum[s].t, 12); }
      if (transform) { transform[s].rx = b.tf[6 * s]; transform[s].ry = b.tf[6 * s + 1]; transform[s].rz = b.tf[6 * s + 2];
                       transform[s].tx = b.tf[6 * s + 3]; transform[s].ty = b.tf[6 * s + 4]; transform[s].tz = b.tf[6 * s + 5]; }
      if (stats) stats[s] = local[s];
    }
    return CM_OK;
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
}
}  // namespace cm

extern "C" {
int cm_odometry_batch_process_host(cm_ctx* ctx, const cm_point* sharp, const int* n_sharp, const cm_point* less_sharp, const int* n_less_sharp,
                                   const cm_point* flat, const int* n_flat, const cm_point* less_flat, const int* n_less_flat, cm_iso* odom,
                                   cm_pose* transform, cm_point* corner_last, cm_point* surf_last, cm_odom_stats* stats) {
  if (!ctx || ctx->obatch.S <= 0) return fail(ctx, CM_ERR_ARG, "cm_odometry_batch_create has not been called");
  const OdomBatch& b = ctx->obatch;
  return odometry_batch_core(ctx, sharp, b.cap_sharp * sizeof(cm_point), n_sharp, less_sharp, b.cap_less_sharp * sizeof(cm_point), n_less_sharp, flat,
                             b.cap_flat * sizeof(cm_point), n_flat, less_flat, b.cap_less_flat * sizeof(cm_point), n_less_flat, odom, transform,
                             corner_last, surf_last, stats);
}
}  // extern "C"
