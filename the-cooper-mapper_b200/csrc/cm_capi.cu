// cm_capi.cu -- the extern "C" boundary declared in include/coopermap.h.
#include "cm_ctx.h"
#include "cm_math.h"
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace cm {
unsigned long long g_launch_count = 0;
unsigned long long g_alloc_generation = 0;
Timeline g_timeline;

MatchParamsDev dev_params(const cm_config& c) {
  MatchParamsDev p;
  p.max_iterations = c.max_iterations;
  p.delta_t_abort = c.delta_t_abort; p.delta_r_abort = c.delta_r_abort;
  p.knn_gate = 5.0f; p.plane_max_dist = 0.2f;
  p.min_ref_corner = 50; p.min_ref_surf = 100; p.min_rows = 50; p.eig_threshold = 100.f;
  p.few_rows_continue = 0; p.nan_guard = 0; p.own_cube_only = 0;
  return p;
}
int ctx_fail(cm_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}
}  // namespace cm

using namespace cm;
static int fail(cm_ctx* ctx, int code, const std::string& msg) { return ctx_fail(ctx, code, msg); }

extern "C" {

void cm_config_default(cm_config* c) {
  memset(c, 0, sizeof(*c));
  c->device = 0;
  c->scan_period = 0.1f; c->n_feature_regions = 6; c->curvature_region = 5; c->max_corner_sharp = 2; c->max_surface_flat = 4;
  c->less_flat_filter_size = 0.2f; c->surface_curvature_threshold = 0.02f; c->blind_degree_threshold = 0.5f; c->blind_radius = 2.5f;
  c->max_iterations = 10; c->delta_t_abort = 0.1f; c->delta_r_abort = 0.1f; c->use_score = 0; c->score_threshold = 800.0;
  c->match_percentage_threshold = 0.4f;
  c->filter_corner = 1.0f; c->filter_surf = 1.0f; c->map_filter_corner = 1.0f; c->map_filter_surf = 1.0f;
  c->cube_w = 121; c->cube_h = 121; c->cube_d = 11; c->cube_size = 50.f; c->valid_distance = 150.f;
  c->cell_corner = 0.f; c->cell_surf = 0.f;
  c->gn_groups = 0;
}

int cm_ctx_create(const cm_config* cfg, cm_ctx** out) {
  if (!cfg || !out) return CM_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
    fprintf(stderr, "coopermap: no usable CUDA device (%s); there is no CPU fallback\n", cudaGetErrorString(e));
    return CM_ERR_CUDA;
  }
  if (cudaSetDevice(cfg->device) != cudaSuccess) return CM_ERR_CUDA;
  cm_ctx* ctx = new cm_ctx();
  ctx->cfg = *cfg;
  // the main stream (matching, map kernels: chains of short, latency-bound launches) gets the highest priority: its CTAs are
  // placed first whenever an SM frees resources, so it interleaves with the long, issue-bound scan registration of the NEXT
  // sweep on the (default-priority) side stream instead of queueing behind its 4096 CTAs  (COOPERMAP_NO_PRIORITY=1: off)
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  if (getenv("COOPERMAP_NO_PRIORITY")) prio_greatest = prio_least;
  ctx->prio_high = prio_greatest; ctx->prio_low = prio_least;
  if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) { delete ctx; return CM_ERR_CUDA; }
  *out = ctx;
  return CM_OK;
}

void cm_ctx_destroy(cm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
  if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
  if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->copy_stream2) { cudaStreamSynchronize(ctx->copy_stream2); cudaStreamDestroy(ctx->copy_stream2); }
  for (int i = 0; i < 2; i++) if (ctx->copy_stream_x[i]) { cudaStreamSynchronize(ctx->copy_stream_x[i]); cudaStreamDestroy(ctx->copy_stream_x[i]); }
  if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
  for (int g = 0; g < CM_MAX_GN_GROUPS; g++) {
    if (ctx->gn_stream[g]) { cudaStreamSynchronize(ctx->gn_stream[g]); cudaStreamDestroy(ctx->gn_stream[g]); }
    if (ctx->gn_join[g]) cudaEventDestroy(ctx->gn_join[g]);
  }
  if (ctx->gn_fork) cudaEventDestroy(ctx->gn_fork);
  for (int i = 0; i <= CM_PIPE_SLOTS; i++) {
    if (ctx->pipe[i].done) cudaEventDestroy(ctx->pipe[i].done);
    if (ctx->pipe[i].copied) cudaEventDestroy(ctx->pipe[i].copied);
    if (ctx->pipe[i].copied2) cudaEventDestroy(ctx->pipe[i].copied2);
    for (int j = 0; j < 2; j++) if (ctx->pipe[i].copied_x[j]) cudaEventDestroy(ctx->pipe[i].copied_x[j]);
    if (ctx->pipe[i].h_n5) cudaFreeHost(ctx->pipe[i].h_n5);
    if (ctx->pipe[i].h_xyz) cudaFreeHost(ctx->pipe[i].h_xyz);
  }
  cm::dist_destroy(ctx);
  cm::stage_pool_destroy(ctx);
  if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  if (ctx->result_ready) cudaEventDestroy(ctx->result_ready);
  delete ctx;
}

const char* cm_last_error(const cm_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
unsigned long long cm_launch_count(const cm_ctx*) { return cm::g_launch_count; }

static float cell_or_default(float cell, float leaf, float mult) { return cell > 0.f ? cell : mult * leaf; }

int cm_knn5_host(cm_ctx* ctx, const cm_point* map, size_t n_map, float cell, float gate, const float* q, size_t nq,
                 int* idx_out, float* d2_out) {
  if (!ctx || (!map && n_map) || (!q && nq) || !idx_out || !d2_out || !(gate > 0.f)) return fail(ctx, CM_ERR_ARG, "bad argument");
  if (nq == 0) return CM_OK;
  try {
    cudaSetDevice(ctx->cfg.device);
    if (!(cell > 0.f)) cell = 1.2f;
    ctx->d_ref_surf.reserve((n_map ? n_map : 1) * sizeof(cm_point));
    ctx->d_q.reserve(nq * 3 * sizeof(float));
    ctx->d_idx.reserve(nq * 5 * sizeof(int));
    ctx->d_d2.reserve(nq * 5 * sizeof(float));
    if (n_map) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_ref_surf.p, map, n_map * sizeof(cm_point), cudaMemcpyHostToDevice, ctx->stream));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_q.p, q, nq * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ctx->grid_a.build((const float4*)ctx->d_ref_surf.p, (int)n_map, cell, gate, 0, ctx->stream);
    launch_knn5(ctx->grid_a.view, (const float*)ctx->d_q.p, (int)nq, gate, (int*)ctx->d_idx.p, (float*)ctx->d_d2.p, ctx->stream);
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(idx_out, ctx->d_idx.p, nq * 5 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(d2_out, ctx->d_d2.p, nq * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

}  // extern "C"
namespace cm {
void fill_match_stats(const cm_config& cfg, const MatchState& st, size_t nq, cm_match_stats* out) {
  out->converged = (st.flags & CM_F_CONVERGED) ? 1 : 0;
  out->degenerate = (st.flags & CM_F_DEGENERATE) ? 1 : 0;
  out->iterations = st.iterations;
  out->rows = st.rows; out->line_matches = st.line; out->plane_matches = st.plane;
  out->score = st.score;
  out->ret = 0;
  if (st.flags & CM_F_TOO_FEW_REF) out->status = CM_TOO_FEW_REF;
  else if (st.flags & CM_F_TOO_FEW_MATCHES) out->status = CM_TOO_FEW_MATCHES;
  else if (!(st.flags & CM_F_CONVERGED)) out->status = CM_NOT_CONVERGED;
  else {
    out->status = CM_OK;
    if (cfg.use_score) {   // ScanMatch.cpp:263-341
      double match_count = (double)(st.line + st.plane);
      float percent = (float)(match_count / (double)nq);
      if (st.score < cfg.score_threshold || (double)percent < (double)cfg.match_percentage_threshold) out->status = CM_LOW_SCORE;
      else out->ret = 1;
    }
  }
}
}  // namespace cm
extern "C" {

// ScanMatch::scanMatchScan on clouds that are already in device memory (host counts).  Internal (declared in cm_ctx.h).
int cm_match_stateless_dev(cm_ctx* ctx, const float4* d_rc, size_t nrc, const float4* d_rs, size_t nrs, const float4* d_c, size_t nc,
                               const float4* d_s, size_t ns, cm_pose* pose, cm_match_stats* stats, cm_iter_trace* trace,
                               int* nn_corner, int* nn_surf) {
  const cm_config& cfg = ctx->cfg;
  MatchParamsDev prm = dev_params(cfg);
  cudaStream_t st = ctx->stream;
  const int capC = (int)(nc ? nc : 1), capS = (int)(ns ? ns : 1), capQ = capC + capS;
  ctx->d_counts.reserve(2 * sizeof(int));
  ctx->d_views.reserve(2 * sizeof(GridView));
  ctx->d_pose.reserve(6 * sizeof(float));
  ctx->d_state.reserve(sizeof(MatchState));
  ctx->d_rows.reserve((size_t)capQ * sizeof(RowOut));
  ctx->d_slots.reserve((size_t)capQ * 5 * sizeof(int));
  ctx->d_sums.reserve(32 * sizeof(double));
  const bool want_nn = nn_corner || nn_surf;
  if (trace) ctx->d_trace.reserve(sizeof(IterTrace) * prm.max_iterations);
  if (want_nn) ctx->d_nn.reserve((size_t)prm.max_iterations * capQ * 5 * sizeof(int));
  int counts[2] = {(int)nc, (int)ns};
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_counts.p, counts, sizeof(counts), cudaMemcpyHostToDevice, st));
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_pose.p, pose, 6 * sizeof(float), cudaMemcpyHostToDevice, st));
  ctx->grid_a.build(d_rc, (int)nrc, cell_or_default(cfg.cell_corner, cfg.map_filter_corner, 8.f), prm.knn_gate, 0, st);
  ctx->grid_b.build(d_rs, (int)nrs, cell_or_default(cfg.cell_surf, cfg.map_filter_surf, 4.f), prm.knn_gate, 0, st);
  GridView views[2] = {ctx->grid_a.view, ctx->grid_b.view};
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_views.p, views, sizeof(views), cudaMemcpyHostToDevice, st));
  if (trace) CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_trace.p, 0, sizeof(IterTrace) * prm.max_iterations, st));
  if (want_nn) CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_nn.p, 0xFF, (size_t)prm.max_iterations * capQ * 5 * sizeof(int), st));
  MatchLaunch m;
  m.nstreams = 1;
  m.corner = d_c; m.surf = d_s;
  m.n_corner = (const int*)ctx->d_counts.p; m.n_surf = (const int*)ctx->d_counts.p + 1;
  m.cap_corner = capC; m.cap_surf = capS;
  m.grid_corner = (const GridView*)ctx->d_views.p; m.grid_surf = (const GridView*)ctx->d_views.p + 1;
  m.pose_in = (const float*)ctx->d_pose.p;
  m.state = (MatchState*)ctx->d_state.p;
  m.rows = (RowOut*)ctx->d_rows.p;
  m.nn_slot = (int*)ctx->d_slots.p;
  m.sums = (double*)ctx->d_sums.p;
  m.trace = trace ? (IterTrace*)ctx->d_trace.p : nullptr;
  m.nn = want_nn ? (int*)ctx->d_nn.p : nullptr;
  m.orig_idx = 1;
  m.max_queries = (int)(nc + ns);
  m.prm = prm;
  ctx->hardq.attach(m, nc + ns + 32);
  launch_match(m, st);
  MatchState hs;
  CM_CUDA_CHECK(ctx, cudaMemcpyAsync(&hs, ctx->d_state.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
  std::vector<int> hnn;
  if (want_nn) {
    hnn.resize((size_t)prm.max_iterations * capQ * 5);
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(hnn.data(), ctx->d_nn.p, hnn.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  if (trace) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(trace, ctx->d_trace.p, sizeof(IterTrace) * prm.max_iterations, cudaMemcpyDeviceToHost, st));
  CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  CM_CUDA_CHECK(ctx, cudaGetLastError());
  if (want_nn) {
    for (int it = 0; it < prm.max_iterations; it++) {
      const int* src = hnn.data() + (size_t)it * capQ * 5;
      if (nn_corner && nc) memcpy(nn_corner + (size_t)it * nc * 5, src, nc * 5 * sizeof(int));
      if (nn_surf && ns) memcpy(nn_surf + (size_t)it * ns * 5, src + nc * 5, ns * 5 * sizeof(int));
    }
  }
  pose->rx = hs.pose[0]; pose->ry = hs.pose[1]; pose->rz = hs.pose[2];
  pose->tx = hs.pose[3]; pose->ty = hs.pose[4]; pose->tz = hs.pose[5];
  cm_match_stats local;
  fill_match_stats(cfg, hs, nc + ns, &local);
  if (stats) *stats = local;
  return local.status;
}

int cm_match_stateless_host(cm_ctx* ctx, const cm_point* ref_corner, size_t nrc, const cm_point* ref_surf, size_t nrs,
                            const cm_point* corner, size_t nc, const cm_point* surf, size_t ns, cm_pose* pose,
                            cm_match_stats* stats, cm_iter_trace* trace, int* nn_corner, int* nn_surf) {
  if (!ctx || !pose || (!ref_corner && nrc) || (!ref_surf && nrs) || (!corner && nc) || (!surf && ns))
    return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    ctx->d_ref_corner.reserve((nrc ? nrc : 1) * sizeof(cm_point));
    ctx->d_ref_surf.reserve((nrs ? nrs : 1) * sizeof(cm_point));
    ctx->d_corner.reserve((nc ? nc : 1) * sizeof(cm_point));
    ctx->d_surf.reserve((ns ? ns : 1) * sizeof(cm_point));
    if (nrc) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_ref_corner.p, ref_corner, nrc * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    if (nrs) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_ref_surf.p, ref_surf, nrs * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    if (nc) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_corner.p, corner, nc * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    if (ns) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_surf.p, surf, ns * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    return cm_match_stateless_dev(ctx, (const float4*)ctx->d_ref_corner.p, nrc, (const float4*)ctx->d_ref_surf.p, nrs,
                               (const float4*)ctx->d_corner.p, nc, (const float4*)ctx->d_surf.p, ns, pose, stats, trace, nn_corner, nn_surf);
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
}

// ScanMatch::scanMatchLocal (ScanMatch.cpp:375-398): voxel-filter all four clouds (corner 0.2, surf 0.4, ScanMatch.cpp:29-30)
// on the device, then scanMatchScan.
int cm_match_local_host(cm_ctx* ctx, const cm_point* ref_corner, size_t nrc, const cm_point* ref_surf, size_t nrs,
                        const cm_point* corner, size_t nc, const cm_point* surf, size_t ns, cm_pose* pose, cm_match_stats* stats) {
  if (!ctx || !pose || (!ref_corner && nrc) || (!ref_surf && nrs) || (!corner && nc) || (!surf && ns))
    return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    const cm_point* src[4] = {ref_corner, ref_surf, corner, surf};
    const size_t n[4] = {nrc, nrs, nc, ns};
    const float leaf[4] = {0.2f, 0.4f, 0.2f, 0.4f};
    DeviceBuffer* raw[4] = {&ctx->d_ref_corner, &ctx->d_ref_surf, &ctx->d_corner, &ctx->d_surf};
    DeviceBuffer* ds[4] = {&ctx->d_l_ds[0], &ctx->d_l_ds[1], &ctx->d_l_ds[2], &ctx->d_l_ds[3]};
    ctx->d_vn_in.reserve(4 * sizeof(int)); ctx->d_vn_out.reserve(4 * sizeof(int)); ctx->d_flag.reserve(sizeof(int));
    int nin[4] = {(int)nrc, (int)nrs, (int)nc, (int)ns};
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_vn_in.p, nin, sizeof(nin), cudaMemcpyHostToDevice, st));
    for (int k = 0; k < 4; k++) {
      const size_t cap = n[k] ? n[k] : 1;
      raw[k]->reserve(cap * sizeof(cm_point)); ds[k]->reserve(cap * sizeof(cm_point));
      if (n[k]) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(raw[k]->p, src[k], n[k] * sizeof(cm_point), cudaMemcpyHostToDevice, st));
      ctx->voxel.run(1, (const float4*)raw[k]->p, (const int*)ctx->d_vn_in.p + k, (int)cap, (int)n[k], leaf[k], (float4*)ds[k]->p,
                     (int*)ctx->d_vn_out.p + k, (int)cap, (int*)ctx->d_flag.p, st);
    }
    int nout[4];
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(nout, ctx->d_vn_out.p, sizeof(nout), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    return cm_match_stateless_dev(ctx, (const float4*)ds[0]->p, nout[0], (const float4*)ds[1]->p, nout[1], (const float4*)ds[2]->p, nout[2],
                               (const float4*)ds[3]->p, nout[3], pose, stats, nullptr, nullptr, nullptr);
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
}

// host-side Isometry <-> Twist, transform_utils.h:54-60, 288-299, 308-331
static void iso_to_twist_host(const cm_iso* it, cm_pose* p) {
  p->tx = it->t[0]; p->ty = it->t[1]; p->tz = it->t[2];
  p->rx = atan2f(it->R[7], it->R[8]);
  p->ry = asinf(-it->R[6]);
  p->rz = atan2f(it->R[3], it->R[0]);
}
static void twist_to_iso_host(const cm_pose* p, cm_iso* it) {
  float pose[6] = {p->rx, p->ry, p->rz, p->tx, p->ty, p->tz};
  pose_to_matrix(pose, it->R);
  it->t[0] = p->tx; it->t[1] = p->ty; it->t[2] = p->tz;
}

int cm_match_stateless_iso_host(cm_ctx* ctx, const cm_point* ref_corner, size_t nrc, const cm_point* ref_surf, size_t nrs,
                                const cm_point* corner, size_t nc, const cm_point* surf, size_t ns, cm_iso* pose,
                                cm_match_stats* stats) {
  if (!pose) return fail(ctx, CM_ERR_ARG, "bad argument");
  cm_pose tw;
  iso_to_twist_host(pose, &tw);
  int rc = cm_match_stateless_host(ctx, ref_corner, nrc, ref_surf, nrs, corner, nc, surf, ns, &tw, stats, nullptr, nullptr, nullptr);
  if (rc < 0) return rc;
  twist_to_iso_host(&tw, pose);
  return rc;
}

}  // extern "C"
namespace cm {
void fill_scanreg_params(const cm_config& c, ScanRegLaunch& L) {
  L.scan_period = c.scan_period; L.blind_radius = c.blind_radius;
  // ScanRegistration.cpp:27,46: blindThreshold = cos(deg2rad(blindDegreeThreshold)), deg2rad(float) math_utils.h:38
  float rad = (float)(c.blind_degree_threshold * M_PI / 180.0);
  L.blind_thr = (float)cos((double)rad);
  L.curv_thr = c.surface_curvature_threshold; L.less_flat_leaf = c.less_flat_filter_size;
  L.R = c.curvature_region; L.nregions = c.n_feature_regions; L.max_sharp = c.max_corner_sharp; L.max_flat = c.max_surface_flat;
  // ScanRegistration.cpp:653-655: cos(deg2rad(175.0)) ... in double
  L.cos175 = cos(175.0 * M_PI / 180.0); L.cos5 = cos(5.0 * M_PI / 180.0);
  L.cos135 = cos(135.0 * M_PI / 180.0); L.cos45 = cos(45.0 * M_PI / 180.0);
}
}  // namespace cm
extern "C" {

// shared by the organised and the raw-sweep entries: frames (and optional tags) are HOST arrays [S][rows][cols]
// (d_frames_in / d_tags_in non-NULL: the rows are already on the device -- the raw-sweep front end produced them there)
static int scanreg_run_host(cm_ctx* ctx, const cm_point* frames, const float* tags, float blind_sq_override, int nstreams, int rows,
                            int cols, cm_scanreg_out* out, size_t full_res_entries, const float4* d_frames_in = nullptr,
                            const float* d_tags_in = nullptr) {
  const cm_config& cfg = ctx->cfg;
  if (cols > 65535 || cfg.curvature_region < 1 || cfg.curvature_region > 8 || cfg.n_feature_regions < 1 || cfg.n_feature_regions > 16 ||
      cfg.max_surface_flat < 0 || cfg.max_surface_flat > 8 || cfg.max_corner_sharp < 0)
    return fail(ctx, CM_ERR_UNSUPPORTED, "scan registration parameters outside the supported range");
  if (scanreg_smem_bytes(cols) > 220 * 1024) return fail(ctx, CM_ERR_UNSUPPORTED, "cols too large for one CTA per ring");
  try {
    cudaSetDevice(cfg.device);
    cudaStream_t st = ctx->stream;
    const size_t npts = (size_t)nstreams * rows * cols;
    if (!d_frames_in) {
      ctx->d_frames.reserve(npts * sizeof(cm_point));
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_frames.p, frames, npts * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    }
    ScanRegLaunch L;
    memset(&L, 0, sizeof(L));
    L.nstreams = nstreams; L.rows = rows; L.cols = cols; L.frames = d_frames_in ? d_frames_in : (const float4*)ctx->d_frames.p;
    fill_scanreg_params(cfg, L);
    L.blind_sq_override = blind_sq_override;
    if (d_tags_in) L.tags = d_tags_in;
    else if (tags) {
      ctx->d_tags.reserve(npts * sizeof(float));
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_tags.p, tags, npts * sizeof(float), cudaMemcpyHostToDevice, st));
      L.tags = (const float*)ctx->d_tags.p;
    }
    for (int k = 0; k < 4; k++) {
      ctx->d_sr_pts[k].reserve((size_t)nstreams * out->cap[k] * sizeof(cm_point));
      L.out_pts[k] = (float4*)ctx->d_sr_pts[k].p; L.cap[k] = out->cap[k];
    }
    ctx->d_sr_n.reserve(sizeof(int) * 5 * nstreams); L.out_n = (int*)ctx->d_sr_n.p;
    ctx->d_flag.reserve(sizeof(int)); L.overflow = (int*)ctx->d_flag.p;
    CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_flag.p, 0, sizeof(int), st));
    L.want_idx = (out->idx[0] || out->idx[1] || out->idx[2] || out->idx[3]) ? 1 : 0;
    if (L.want_idx) for (int k = 0; k < 4; k++) { ctx->d_sr_idx[k].reserve(npts * sizeof(int)); L.out_idx[k] = (int*)ctx->d_sr_idx[k].p; }
    if (out->cloud) { ctx->d_sr_cloud.reserve(npts * sizeof(cm_point)); ctx->d_sr_ccurv.reserve(npts * sizeof(float));
                      L.cloud = (float4*)ctx->d_sr_cloud.p; L.cloud_curv = (float*)ctx->d_sr_ccurv.p; }
    if (out->picked) { ctx->d_sr_picked.reserve(npts); L.picked = (signed char*)ctx->d_sr_picked.p; }
    if (out->curvature) { ctx->d_sr_curv.reserve(npts * sizeof(float)); L.curvature = (float*)ctx->d_sr_curv.p; }
    if (out->label) { ctx->d_sr_label.reserve(npts); L.label = (signed char*)ctx->d_sr_label.p; }
    if (out->scan_ranges) { ctx->d_sr_range.reserve(sizeof(int) * 2 * nstreams * rows); L.scan_range = (int*)ctx->d_sr_range.p; }
    ctx->scanreg.run(L, st);
    const size_t fr = full_res_entries < npts ? full_res_entries : npts;   // full-resolution outputs: entries the caller sized
    for (int k = 0; k < 4; k++)
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out->pts[k], L.out_pts[k], (size_t)nstreams * out->cap[k] * sizeof(cm_point), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out->n, L.out_n, sizeof(int) * 5 * nstreams, cudaMemcpyDeviceToHost, st));
    if (L.want_idx) for (int k = 0; k < 4; k++) if (out->idx[k])
      CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out->idx[k], L.out_idx[k], fr * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (out->cloud) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out->cloud, L.cloud, fr * sizeof(cm_point), cudaMemcpyDeviceToHost, st));
    if (out->cloud && out->cloud_curvature) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out->cloud_curvature, L.cloud_curv, fr * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out->picked) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out->picked, L.picked, fr, cudaMemcpyDeviceToHost, st));
    if (out->curvature) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out->curvature, L.curvature, fr * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out->label) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out->label, L.label, fr, cudaMemcpyDeviceToHost, st));
    if (out->scan_ranges) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out->scan_ranges, L.scan_range, sizeof(int) * 2 * nstreams * rows, cudaMemcpyDeviceToHost, st));
    int ovf = 0;
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(&ovf, ctx->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    if (ovf) return fail(ctx, CM_ERR_CAPACITY, "a feature cloud exceeds its output capacity");
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_scanreg_organised_host(cm_ctx* ctx, const cm_point* frames, int nstreams, int rows, int cols, cm_scanreg_out* out) {
  if (!ctx || !frames || !out || nstreams <= 0 || rows <= 0 || cols <= 0 || !out->n) return fail(ctx, CM_ERR_ARG, "bad argument");
  for (int k = 0; k < 4; k++) if (!out->pts[k] || out->cap[k] <= 0) return fail(ctx, CM_ERR_ARG, "bad output buffers");
  return scanreg_run_host(ctx, frames, nullptr, -1.f, nstreams, rows, cols, out, (size_t)nstreams * rows * cols);
}

// Raw-sweep front end, MultiScanRegistration::process (MultiScanRegistration.cpp:95-190): per point axis swap, validity, ring from
// the elevation angle, azimuth unwrap with the half-sweep flag, relTime, stable per-ring append -- on the device (cm_frontend.cu);
// the rings go to the same extraction kernels as ring-major rows with their precomputed curvature field.  The sweep crosses PCIe
// once, unsorted; the only host work is two atan2 for the sweep's start / end orientation.
int cm_imu_push_host(cm_ctx* ctx, const cm_imu_sample* m) {
  if (!ctx || !m) return fail(ctx, CM_ERR_ARG, "bad argument");
  ctx->imu.push(m->stamp, m->roll, m->pitch, m->yaw, m->ax, m->ay, m->az);
  return CM_OK;
}
int cm_imu_clear(cm_ctx* ctx) {
  if (!ctx) return CM_ERR_ARG;
  ctx->imu.clear();
  return CM_OK;
}
static int scanreg_sweep_common(cm_ctx* ctx, const cm_point* sweep, size_t n, int lidar, cm_scanreg_out* out, int* rows_out, int* cols_out,
                                bool use_imu, double scan_time, float* imu_trans);
int cm_scanreg_sweep_host(cm_ctx* ctx, const cm_point* sweep, size_t n, int lidar, cm_scanreg_out* out, int* rows_out, int* cols_out) {
  return scanreg_sweep_common(ctx, sweep, n, lidar, out, rows_out, cols_out, false, 0.0, nullptr);
}
int cm_scanreg_sweep_imu_host(cm_ctx* ctx, const cm_point* sweep, size_t n, int lidar, double scan_time, cm_scanreg_out* out, int* rows_out,
                              int* cols_out, float* imu_trans) {
  return scanreg_sweep_common(ctx, sweep, n, lidar, out, rows_out, cols_out, true, scan_time, imu_trans);
}
static int scanreg_sweep_common(cm_ctx* ctx, const cm_point* sweep, size_t n, int lidar, cm_scanreg_out* out, int* rows_out, int* cols_out,
                                bool use_imu, double scan_time, float* imu_trans) {
  float lo, up; int nr;
  if (!ctx || (!sweep && n) || !out || !out->n || !frontend_mapper(lidar, &lo, &up, &nr) || n > 0x7fffffffu) return fail(ctx, CM_ERR_ARG, "bad argument");
  for (int k = 0; k < 4; k++) if (!out->pts[k] || out->cap[k] <= 0) return fail(ctx, CM_ERR_ARG, "bad output buffers");
  int rows = nr, cols = 1;
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    ctx->d_sweep.reserve((n ? n : 1) * sizeof(cm_point));
    if (n) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_sweep.p, sweep, n * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 first = n ? make_float4(sweep[0].x, sweep[0].y, sweep[0].z, 0.f) : zero;
    const float4 last = n ? make_float4(sweep[n - 1].x, sweep[n - 1].y, sweep[n - 1].z, 0.f) : zero;
    ctx->frontend.run((const float4*)ctx->d_sweep.p, (int)n, first, last, lidar, ctx->cfg.scan_period, st, &rows, &cols,
                      (use_imu && !ctx->imu.stamp.empty()) ? &ctx->imu : nullptr, scan_time, imu_trans);
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  if (rows_out) *rows_out = rows;
  if (cols_out) *cols_out = cols;
  // full-resolution outputs come back in the reference's concatenated _laserCloud order: at most n entries
  return scanreg_run_host(ctx, nullptr, nullptr, 0.f, 1, rows, cols, out, n, (const float4*)ctx->frontend.frame.p, (const float*)ctx->frontend.tags.p);
}

/* ---- sharded-map matching: one rank's part of ScanMatch::scanMatchScan when the reference map is split over ranks ---- */
int cm_shard_set_map_host(cm_ctx* ctx, const cm_point* corner, size_t nc, const cm_point* surf, size_t ns) {
  if (!ctx || (!corner && nc) || (!surf && ns)) return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    const cm_config& cfg = ctx->cfg;
    cudaStream_t st = ctx->stream;
    ctx->d_ref_corner.reserve((nc ? nc : 1) * sizeof(cm_point));
    ctx->d_ref_surf.reserve((ns ? ns : 1) * sizeof(cm_point));
    if (nc) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_ref_corner.p, corner, nc * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    if (ns) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_ref_surf.p, surf, ns * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    ctx->grid_a.build((const float4*)ctx->d_ref_corner.p, (int)nc, cell_or_default(cfg.cell_corner, cfg.map_filter_corner, 8.f), 5.0f, 0, st);
    ctx->grid_b.build((const float4*)ctx->d_ref_surf.p, (int)ns, cell_or_default(cfg.cell_surf, cfg.map_filter_surf, 4.f), 5.0f, 0, st);
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->shard_ready = true;
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_shard_begin_host(cm_ctx* ctx, const cm_point* corner, size_t nc, const cm_point* surf, size_t ns, const cm_pose* init,
                        size_t total_ref_corner, size_t total_ref_surf) {
  if (!ctx || !ctx->shard_ready || !init || (!corner && nc) || (!surf && ns)) return fail(ctx, CM_ERR_ARG, "bad argument / no shard map");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    MatchParamsDev prm = dev_params(ctx->cfg);
    const int capC = (int)(nc ? nc : 1), capS = (int)(ns ? ns : 1), capQ = capC + capS;
    ctx->d_corner.reserve(capC * sizeof(cm_point)); ctx->d_surf.reserve(capS * sizeof(cm_point));
    ctx->d_counts.reserve(2 * sizeof(int)); ctx->d_views.reserve(2 * sizeof(GridView)); ctx->d_pose.reserve(6 * sizeof(float));
    ctx->d_state.reserve(sizeof(MatchState)); ctx->d_rows.reserve((size_t)capQ * sizeof(RowOut));
    ctx->d_slots.reserve((size_t)capQ * 5 * sizeof(int)); ctx->d_sums.reserve(32 * sizeof(double)); ctx->d_box.reserve(6 * sizeof(float));
    if (nc) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_corner.p, corner, nc * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    if (ns) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_surf.p, surf, ns * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    int counts[2] = {(int)nc, (int)ns};
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_counts.p, counts, sizeof(counts), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_pose.p, init, 6 * sizeof(float), cudaMemcpyHostToDevice, st));
    GridView views[2] = {ctx->grid_a.view, ctx->grid_b.view};
    views[0].npts = (int)total_ref_corner; views[1].npts = (int)total_ref_surf;   // the 50 / 100 gate looks at the WHOLE map
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_views.p, views, sizeof(views), cudaMemcpyHostToDevice, st));
    MatchLaunch& m = ctx->shard;
    m = MatchLaunch();
    m.nstreams = 1;
    m.corner = (const float4*)ctx->d_corner.p; m.surf = (const float4*)ctx->d_surf.p;
    m.n_corner = (const int*)ctx->d_counts.p; m.n_surf = (const int*)ctx->d_counts.p + 1;
    m.cap_corner = capC; m.cap_surf = capS;
    m.grid_corner = (const GridView*)ctx->d_views.p; m.grid_surf = (const GridView*)ctx->d_views.p + 1;
    m.pose_in = (const float*)ctx->d_pose.p; m.state = (MatchState*)ctx->d_state.p; m.rows = (RowOut*)ctx->d_rows.p;
    m.nn_slot = (int*)ctx->d_slots.p; m.sums = (double*)ctx->d_sums.p; m.trace = nullptr; m.nn = nullptr;
    m.orig_idx = 1; m.max_queries = (int)(nc + ns); m.own_box = (const float*)ctx->d_box.p; m.prm = prm;
    ctx->shard_nq = nc + ns;
    ctx->hardq.attach(m, nc + ns + 32);
    launch_match_init(m, st);
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_shard_partial_host(cm_ctx* ctx, int iter, const float* own_lo, const float* own_hi, double* sums32) {
  if (!ctx || !ctx->shard_ready || !own_lo || !own_hi || !sums32 || iter < 0) return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    float box[6] = {own_lo[0], own_lo[1], own_lo[2], own_hi[0], own_hi[1], own_hi[2]};
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_box.p, box, sizeof(box), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_sums.p, 0, 32 * sizeof(double), st));
    launch_match_partial(ctx->shard, iter, st, nullptr);
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(sums32, ctx->d_sums.p, 32 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_shard_solve_host(cm_ctx* ctx, int iter, const double* sums32, cm_pose* pose, int* done, cm_match_stats* stats) {
  if (!ctx || !ctx->shard_ready || !sums32 || !pose || !done || iter < 0) return fail(ctx, CM_ERR_ARG, "bad argument");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_sums.p, sums32, 32 * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_match_solve(ctx->shard, iter, (const double*)ctx->d_sums.p, st);
    MatchState hs;
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(&hs, ctx->d_state.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    pose->rx = hs.pose[0]; pose->ry = hs.pose[1]; pose->rz = hs.pose[2]; pose->tx = hs.pose[3]; pose->ty = hs.pose[4]; pose->tz = hs.pose[5];
    *done = hs.done;
    if (stats) fill_match_stats(ctx->cfg, hs, ctx->shard_nq, stats);
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_timeline_enable(cm_ctx* ctx, int on) { (void)ctx; g_timeline.on = on != 0; return CM_OK; }
// writes "name total_us launches" lines, sorted by time, into buf; resets the timeline
int cm_timeline_report(cm_ctx* ctx, char* buf, size_t cap) {
  if (!ctx || !buf || cap == 0) return CM_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<double, int>> agg;
  for (auto& r : g_timeline.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { auto& e = agg[r.name]; e.first += ms * 1e3; e.second++; }
    g_timeline.pool.push_back(r.a); g_timeline.pool.push_back(r.b);
  }
  g_timeline.recs.clear();
  std::vector<std::pair<double, std::string>> v;
  for (auto& kv : agg) v.push_back({kv.second.first, kv.first});
  std::sort(v.begin(), v.end(), [](const std::pair<double, std::string>& a, const std::pair<double, std::string>& b) { return a.first > b.first; });
  std::string out;
  for (auto& e : v) { char line[256]; snprintf(line, sizeof(line), "%s %.1f %d\n", e.second.c_str(), e.first, agg[e.second].second); out += line; }
  snprintf(buf, cap, "%s", out.c_str());
  return CM_OK;
}

int cm_voxel_filter_host(cm_ctx* ctx, const cm_point* in, int nseg, const int* n_in, int cap_in, float leaf, cm_point* out,
                         int* n_out, int cap_out) {
  if (!ctx || nseg <= 0 || !n_in || cap_in <= 0 || !(leaf > 0.f) || !out || !n_out || cap_out <= 0 || !in)
    return fail(ctx, CM_ERR_ARG, "bad argument");
  for (int s = 0; s < nseg; s++) if (n_in[s] < 0 || n_in[s] > cap_in) return fail(ctx, CM_ERR_ARG, "n_in[s] out of range");
  try {
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    ctx->d_vin.reserve((size_t)nseg * cap_in * sizeof(cm_point));
    ctx->d_vout.reserve((size_t)nseg * cap_out * sizeof(cm_point));
    ctx->d_vn_in.reserve(nseg * sizeof(int)); ctx->d_vn_out.reserve(nseg * sizeof(int)); ctx->d_flag.reserve(sizeof(int));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_vin.p, in, (size_t)nseg * cap_in * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_vn_in.p, n_in, nseg * sizeof(int), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_flag.p, 0, sizeof(int), st));
    ctx->voxel.run(nseg, (const float4*)ctx->d_vin.p, (const int*)ctx->d_vn_in.p, cap_in, 0, leaf, (float4*)ctx->d_vout.p,
                   (int*)ctx->d_vn_out.p, cap_out, (int*)ctx->d_flag.p, st);
    int ovf = 0;
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out, ctx->d_vout.p, (size_t)nseg * cap_out * sizeof(cm_point), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(n_out, ctx->d_vn_out.p, nseg * sizeof(int), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(&ovf, ctx->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    if (ovf) return fail(ctx, CM_ERR_CAPACITY, "voxel filter output exceeds cap_out");
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_debug_math_host(cm_ctx* ctx, int op, const float* in, size_t n, float* out) {
  int nin, nout;
  if (!ctx || !in || !out || debug_math_dims(op, &nin, &nout) != 0) return fail(ctx, CM_ERR_ARG, "bad argument");
  if (n == 0) return CM_OK;
  try {
    cudaSetDevice(ctx->cfg.device);
    ctx->d_q.reserve(n * nin * sizeof(float));
    ctx->d_d2.reserve(n * nout * sizeof(float));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_q.p, in, n * nin * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    launch_debug_math(op, (const float*)ctx->d_q.p, nin, (float*)ctx->d_d2.p, nout, (int)n, ctx->stream);
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out, ctx->d_d2.p, n * nout * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

}  // extern "C"
