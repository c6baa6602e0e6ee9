// cm_odometry.cu -- the scan-to-scan odometry STAGE: host mirror of LaserOdometry::process
// (L_SLAM/src/odometry/LaserOdometry.cpp:288-326) around the K8 kernels (cm_odom.inl) and the shared reduce / solve kernels.
#include "cm_ctx.h"
#include "cm_math.h"
#include <math.h>
#include <string.h>
#include <algorithm>

using namespace cm;
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
      ctx->o_grid_c.build((const float4*)ctx->o_last_c.p, n_less_sharp, 2.5f, 25.f, 0, st);
      ctx->o_grid_s.build((const float4*)ctx->o_last_s.p, n_less_flat, 2.5f, 25.f, 0, st);
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

}  // extern "C"
