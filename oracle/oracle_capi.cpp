// oracle_capi.cpp -- extern "C" surface of the CPU oracle for ctypes (TEST INFRASTRUCTURE, see cm_oracle.h).
#include "cm_oracle.h"
#include "../the-cooper-mapper_b200/csrc/cm_math.h"
#include <dlfcn.h>
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstring>
#include <string>

using namespace cmo;

namespace {
// ---- nanoflann backend through oracle/_ref/libcm_ref_nanoflann.so -------------------------------------------
typedef void* (*ref_build_t)(const float*, size_t);
typedef void (*ref_free_t)(void*);
typedef int (*ref_query_t)(void*, const float*, int, int*, float*);
ref_build_t g_ref_build = nullptr; ref_free_t g_ref_free = nullptr; ref_query_t g_ref_query = nullptr;
void nf_build(const PointI* pts, size_t n, void** h) { *h = g_ref_build((const float*)pts, n); }
void nf_free(void* h) { g_ref_free(h); }
void nf_query(void* h, const float q[3], int k, int* idx, float* d2) {
  int found = g_ref_query(h, q, k, idx, d2);
  for (int i = found; i < k; i++) { idx[i] = -1; d2[i] = FLT_MAX; }
  // canonical interior order: (d2, index)
  for (int i = 1; i < found; i++) {
    int ii = idx[i]; float dd = d2[i]; int j = i;
    while (j > 0 && (d2[j - 1] > dd || (d2[j - 1] == dd && idx[j - 1] > ii))) { d2[j] = d2[j - 1]; idx[j] = idx[j - 1]; j--; }
    d2[j] = dd; idx[j] = ii;
  }
}
KnnBackend pick_backend(int useNanoflann) {
  if (useNanoflann && g_ref_build) return KnnBackend{nf_build, nf_query, nf_free};
  return brute_force_backend();
}
template <class T> void copy_out(const std::vector<T>& v, void* dst) { if (!v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T)); }
}  // namespace

extern "C" {

int cmo_load_nanoflann(const char* path) {
  void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!lib) return 0;
  g_ref_build = (ref_build_t)dlsym(lib, "cmref_kd_build");
  g_ref_free = (ref_free_t)dlsym(lib, "cmref_kd_free");
  g_ref_query = (ref_query_t)dlsym(lib, "cmref_kd_query");
  return (g_ref_build && g_ref_free && g_ref_query) ? 1 : 0;
}

// ---- math known-answer hooks ----------------------------------------------------------------------------------
void cmo_sincosf(const float* x, int n, float* s, float* c) { for (int i = 0; i < n; i++) cm::cm_sincosf(x[i], s + i, c + i); }
void cmo_eig3(const float* A6, float* w, float* V) { cm::eig3_sym(A6, w, V); }
void cmo_eig6(const float* A36, float* w, float* V) { cm::eig_sym<6>(A36, w, V); }
void cmo_qr_solve_5x3(const float* A, const float* b, float* x) { float a[15], bb[5]; std::memcpy(a, A, sizeof(a)); std::memcpy(bb, b, sizeof(bb)); cm::colpiv_qr_solve<5, 3>(a, bb, x); }
void cmo_qr_solve_6x6(const float* A, const float* b, float* x) { float a[36], bb[6]; std::memcpy(a, A, sizeof(a)); std::memcpy(bb, b, sizeof(bb)); cm::colpiv_qr_solve<6, 6>(a, bb, x); }
int cmo_inverse6(const float* A, float* inv) { return cm::inverse_lu<6>(A, inv) ? 1 : 0; }
void cmo_pose_to_matrix(const float* pose, float* R) { cm::pose_to_matrix(pose, R); }
void cmo_iso_to_twist(const float* R, const float* t, float* pose) { Iso i; std::memcpy(i.R, R, 36); std::memcpy(i.t, t, 12); iso_to_twist(i, pose); }

// same op codes as cm_debug_math_host (include/coopermap.h): the shared header compiled for the HOST
void cmo_debug_math(int op, const float* in, size_t n, float* out) {
  static const int ni[8] = {42, 20, 6, 36, 36, 36, 6, 2}, no[8] = {6, 3, 12, 42, 6, 36, 15, 2};
  for (size_t t = 0; t < n; t++) {
    const float* a = in + t * ni[op]; float* o = out + t * no[op];
    if (op == 0) { float A[36], b[6]; std::memcpy(A, a, 144); std::memcpy(b, a + 36, 24); cm::colpiv_qr_solve<6, 6>(A, b, o); }
    else if (op == 1) { float A[15], b[5]; std::memcpy(A, a, 60); std::memcpy(b, a + 15, 20); cm::colpiv_qr_solve<5, 3>(A, b, o); }
    else if (op == 2) cm::eig3_sym(a, o, o + 3);
    else if (op == 3) cm::eig_sym<6>(a, o, o + 6);
    else if (op == 4) cm::eig_sym<6>(a, o, (float*)nullptr);
    else if (op == 5) { float inv[36]; bool ok = cm::inverse_lu<6>(a, inv); for (int i = 0; i < 36; i++) o[i] = ok ? inv[i] : 0.f; }
    else if (op == 7) { o[0] = cm::cm_atan2f(a[0], a[1]); o[1] = cm::cm_atanf(a[0] / a[1]); }
    else if (op == 6) { cm::pose_to_matrix(a, o); for (int i = 0; i < 3; i++) cm::cm_sincosf(a[i], o + 9 + i, o + 12 + i); }
  }
}

// ---- scan registration ------------------------------------------------------------------------------------------
struct ScanRegHandle { ScanRegResult r; };
static ScanRegParams make_prm(const float* f, const int* iv) {
  ScanRegParams p;
  if (f) { p.scanPeriod = f[0]; p.lessFlatFilterSize = f[1]; p.surfaceCurvatureThreshold = f[2]; p.blindDegreeThreshold = f[3]; p.blindRadius = f[4]; }
  if (iv) { p.nFeatureRegions = iv[0]; p.curvatureRegion = iv[1]; p.maxCornerSharp = iv[2]; p.maxSurfaceFlat = iv[3]; }
  return p;
}
// fparams: scanPeriod, lessFlatFilterSize, surfaceCurvatureThreshold, blindDegreeThreshold, blindRadius
// iparams: nFeatureRegions, curvatureRegion, maxCornerSharp, maxSurfaceFlat       (NULL = reference defaults)
void* cmo_scanreg_organised(const float* fparams, const int* iparams, const float* xyzi, int rows, int cols) {
  ScanRegHandle* h = new ScanRegHandle();
  scanreg_organised(make_prm(fparams, iparams), xyzi, rows, cols, h->r);
  return h;
}
void* cmo_scanreg_sweep(const float* fparams, const int* iparams, const float* xyzi, int n, int lidar) {
  ScanRegHandle* h = new ScanRegHandle();
  scanreg_sweep(make_prm(fparams, iparams), xyzi, n, lidar, h->r);
  return h;
}
// imu: [nimu][7] doubles (stamp, roll, pitch, yaw, ax, ay, az) pushed in order through ImuHistory::push; imuTrans: 12 floats out
void* cmo_scanreg_sweep_imu(const float* fparams, const int* iparams, const float* xyzi, int n, int lidar, double scanTime, const double* imu,
                            int nimu, float* imuTrans) {
  ScanRegHandle* h = new ScanRegHandle();
  ImuHistory hist;
  for (int k = 0; k < nimu; k++) hist.push(imu[7 * k], imu[7 * k + 1], imu[7 * k + 2], imu[7 * k + 3], imu[7 * k + 4], imu[7 * k + 5], imu[7 * k + 6]);
  scanreg_sweep_imu(make_prm(fparams, iparams), xyzi, n, lidar, scanTime, hist, h->r, imuTrans);
  return h;
}
void cmo_scanreg_free(void* h) { delete (ScanRegHandle*)h; }
// field ids: 0 cloud(5f) 1 scanStart 2 scanEnd 3 sharp(4f) 4 lessSharp 5 flat 6 lessFlat 7 sharpIdx 8 lessSharpIdx
// 9 flatIdx 10 lessFlatRawIdx 11 lessFlatRawRing 12 picked 13 curvature 14 classLabel 15..18 dbg blind/block/slop/curv
size_t cmo_scanreg_size(void* hh, int field) {
  ScanRegResult& r = ((ScanRegHandle*)hh)->r;
  switch (field) {
    case 0: return r.cloud.size(); case 1: return r.scanStart.size(); case 2: return r.scanEnd.size();
    case 3: return r.sharp.size(); case 4: return r.lessSharp.size(); case 5: return r.flat.size(); case 6: return r.lessFlat.size();
    case 7: return r.sharpIdx.size(); case 8: return r.lessSharpIdx.size(); case 9: return r.flatIdx.size();
    case 10: return r.lessFlatRawIdx.size(); case 11: return r.lessFlatRawRing.size(); case 12: return r.picked.size();
    case 13: return r.curvature.size(); case 14: return r.classLabel.size();
    case 15: return r.dbgBlind.size(); case 16: return r.dbgBlock.size(); case 17: return r.dbgSlop.size(); case 18: return r.dbgCurv.size();
  }
  return 0;
}
void cmo_scanreg_copy(void* hh, int field, void* dst) {
  ScanRegResult& r = ((ScanRegHandle*)hh)->r;
  switch (field) {
    case 0: copy_out(r.cloud, dst); break; case 1: copy_out(r.scanStart, dst); break; case 2: copy_out(r.scanEnd, dst); break;
    case 3: copy_out(r.sharp, dst); break; case 4: copy_out(r.lessSharp, dst); break; case 5: copy_out(r.flat, dst); break;
    case 6: copy_out(r.lessFlat, dst); break; case 7: copy_out(r.sharpIdx, dst); break; case 8: copy_out(r.lessSharpIdx, dst); break;
    case 9: copy_out(r.flatIdx, dst); break; case 10: copy_out(r.lessFlatRawIdx, dst); break; case 11: copy_out(r.lessFlatRawRing, dst); break;
    case 12: copy_out(r.picked, dst); break; case 13: copy_out(r.curvature, dst); break; case 14: copy_out(r.classLabel, dst); break;
    case 15: copy_out(r.dbgBlind, dst); break; case 16: copy_out(r.dbgBlock, dst); break; case 17: copy_out(r.dbgSlop, dst); break;
    case 18: copy_out(r.dbgCurv, dst); break;
  }
}

// ---- voxel filter -----------------------------------------------------------------------------------------------
size_t cmo_voxel_filter(const float* xyzi, size_t n, float leaf, float* out, size_t cap) {
  std::vector<PointI> o;
  voxel_filter((const PointI*)xyzi, n, leaf, o);
  if (out) std::memcpy(out, o.data(), std::min(cap, o.size()) * sizeof(PointI));
  return o.size();
}

// ---- KNN ----------------------------------------------------------------------------------------------------------
void cmo_knn(int useNanoflann, const float* pts, size_t n, const float* q, size_t nq, int k, int* idx, float* d2) {
  KnnBackend b = pick_backend(useNanoflann);
  void* h = nullptr;
  b.build((const PointI*)pts, n, &h);
  for (size_t i = 0; i < nq; i++) b.query(h, q + 3 * i, k, idx + (size_t)k * i, d2 + (size_t)k * i);
  b.free(h);
}

// ---- scan-to-map solver -----------------------------------------------------------------------------------------
struct MatchHandle { MatchResult res; size_t nc, ns; };
// fparams: deltaTAbort, deltaRAbort, knnGate, planeMaxDistance ; iparams: maxIterations, useScore
void* cmo_scan_match(const float* fparams, const int* iparams, int useNanoflann, const float* refCorner, size_t nrc,
                     const float* refSurf, size_t nrs, const float* corner, size_t nc, const float* surf, size_t ns,
                     float* pose, int keepLog) {
  MatchParams p;
  if (fparams) { p.deltaTAbort = fparams[0]; p.deltaRAbort = fparams[1]; p.knnGate = fparams[2]; p.planeMaxDistance = fparams[3]; }
  if (iparams) { p.maxIterations = iparams[0]; p.useScore = iparams[1] != 0; }
  MatchHandle* h = new MatchHandle();
  h->nc = nc; h->ns = ns;
  scan_match(p, pick_backend(useNanoflann), (const PointI*)refCorner, nrc, (const PointI*)refSurf, nrs,
             (const PointI*)corner, nc, (const PointI*)surf, ns, pose, h->res, keepLog != 0);
  return h;
}
void cmo_match_free(void* h) { delete (MatchHandle*)h; }
// out[0..9]: ok, converged, tooFewRef, tooFewMatches, degenerate, iterations, lastRows, lastLine, lastPlane, nLog
void cmo_match_stats(void* hh, int* out, double* score) {
  MatchResult& r = ((MatchHandle*)hh)->res;
  out[0] = r.ok; out[1] = r.converged; out[2] = r.tooFewRef; out[3] = r.tooFewMatches; out[4] = r.degenerate;
  out[5] = r.iterations; out[6] = r.lastRows; out[7] = r.lastLine; out[8] = r.lastPlane; out[9] = (int)r.log.size();
  if (score) *score = r.score;
}
// per-iteration log: pose_in[6], AtA[36], AtB[6], x[6], counts[4] = rows, line, plane, degenerate; nn arrays 5*nc, 5*ns
void cmo_match_log(void* hh, int it, float* pose_in, float* AtA, float* AtB, float* x, int* counts, int* nnCorner, int* nnSurf) {
  MatchHandle* h = (MatchHandle*)hh;
  const IterLog& l = h->res.log[it];
  std::memcpy(pose_in, l.pose_in, 24); std::memcpy(AtA, l.AtA, 144); std::memcpy(AtB, l.AtB, 24); std::memcpy(x, l.x, 24);
  counts[0] = l.rows; counts[1] = l.lineMatches; counts[2] = l.planeMatches; counts[3] = l.degenerate;
  if (nnCorner) copy_out(l.nnCorner, nnCorner);
  if (nnSurf) copy_out(l.nnSurf, nnSurf);
}

// ---- mapping loop -----------------------------------------------------------------------------------------------
struct MappingHandle { LaserMapping* m; };
// mfparams: cubeSize, validDistance, mapFilterCorner, mapFilterSurf, filterCorner, filterSurf ; miparams: cubeW, cubeH, cubeD
void* cmo_mapping_create(const float* mfparams, const int* miparams, const float* sfparams, const int* siparams, int useNanoflann) {
  MapParams mp; MatchParams sp;
  if (mfparams) { mp.cubeSize = mfparams[0]; mp.validDistance = mfparams[1]; mp.mapFilterCorner = mfparams[2]; mp.mapFilterSurf = mfparams[3]; mp.filterCorner = mfparams[4]; mp.filterSurf = mfparams[5]; }
  if (miparams) { mp.cubeW = miparams[0]; mp.cubeH = miparams[1]; mp.cubeD = miparams[2]; }
  if (sfparams) { sp.deltaTAbort = sfparams[0]; sp.deltaRAbort = sfparams[1]; sp.knnGate = sfparams[2]; sp.planeMaxDistance = sfparams[3]; }
  if (siparams) { sp.maxIterations = siparams[0]; sp.useScore = siparams[1] != 0; }
  MappingHandle* h = new MappingHandle();
  h->m = new LaserMapping(mp, sp, pick_backend(useNanoflann));
  return h;
}
void cmo_mapping_free(void* hh) { MappingHandle* h = (MappingHandle*)hh; delete h->m; delete h; }
// odom / out pose: R[9] row-major + t[3]
void cmo_mapping_process(void* hh, const float* odomR, const float* odomT, const float* corner, size_t nc, const float* surf,
                         size_t ns, float* outR, float* outT, int* stats) {
  LaserMapping* m = ((MappingHandle*)hh)->m;
  Iso od; std::memcpy(od.R, odomR, 36); std::memcpy(od.t, odomT, 12);
  std::vector<PointI> c((const PointI*)corner, (const PointI*)corner + nc), s((const PointI*)surf, (const PointI*)surf + ns);
  Iso r = m->process(od, c, s);
  std::memcpy(outR, r.R, 36); std::memcpy(outT, r.t, 12);
  if (stats) {
    const MatchResult& q = m->lastMatch;
    stats[0] = q.ok; stats[1] = q.converged; stats[2] = q.tooFewRef; stats[3] = q.tooFewMatches; stats[4] = q.degenerate;
    stats[5] = q.iterations; stats[6] = q.lastRows; stats[7] = q.lastLine; stats[8] = q.lastPlane;
    stats[9] = (int)m->cornerDS.size(); stats[10] = (int)m->surfDS.size();
    stats[11] = (int)m->surroundCorner.size(); stats[12] = (int)m->surroundSurf.size();
  }
}
// LaserLocalization::process
void cmo_mapping_localize(void* hh, const float* odomR, const float* odomT, const float* corner, size_t nc, const float* surf,
                          size_t ns, float* outR, float* outT, int* stats) {
  LaserMapping* m = ((MappingHandle*)hh)->m;
  Iso od; std::memcpy(od.R, odomR, 36); std::memcpy(od.t, odomT, 12);
  std::vector<PointI> c((const PointI*)corner, (const PointI*)corner + nc), s((const PointI*)surf, (const PointI*)surf + ns);
  Iso r = m->localize(od, c, s);
  std::memcpy(outR, r.R, 36); std::memcpy(outT, r.t, 12);
  if (stats) {
    const MatchResult& q = m->lastMatch;
    stats[0] = q.ok; stats[1] = q.converged; stats[2] = q.tooFewRef; stats[3] = q.tooFewMatches; stats[4] = q.degenerate;
    stats[5] = q.iterations; stats[6] = q.lastRows; stats[7] = q.lastLine; stats[8] = q.lastPlane;
    stats[9] = (int)m->cornerDS.size(); stats[10] = (int)m->surfDS.size(); stats[11] = stats[12] = 0;
  }
}
// ---- LaserMappingLocal (sliding-window map) ----------------------------------------------------------------------
struct MappingLocalHandle { LaserMappingLocal* m; };
void* cmo_mapping_local_create(const float* mfparams, const float* sfparams, const int* siparams, int useNanoflann, int useMappedPose) {
  MapParams mp; MatchParams sp;
  if (mfparams) { mp.filterCorner = mfparams[4]; mp.filterSurf = mfparams[5]; }
  if (sfparams) { sp.deltaTAbort = sfparams[0]; sp.deltaRAbort = sfparams[1]; sp.knnGate = sfparams[2]; sp.planeMaxDistance = sfparams[3]; }
  if (siparams) { sp.maxIterations = siparams[0]; sp.useScore = siparams[1] != 0; }
  MappingLocalHandle* h = new MappingLocalHandle();
  h->m = new LaserMappingLocal(mp, sp, pick_backend(useNanoflann), useMappedPose != 0);
  return h;
}
void cmo_mapping_local_free(void* hh) { MappingLocalHandle* h = (MappingLocalHandle*)hh; delete h->m; delete h; }
// stats as cmo_mapping_process, then [13] frames in the window, [14] corner points, [15] surf points of the window; accum: travelled distance
void cmo_mapping_local_process(void* hh, const float* odomR, const float* odomT, const float* corner, size_t nc, const float* surf,
                               size_t ns, float* outR, float* outT, int* stats, double* accum) {
  LaserMappingLocal* m = ((MappingLocalHandle*)hh)->m;
  Iso od; std::memcpy(od.R, odomR, 36); std::memcpy(od.t, odomT, 12);
  std::vector<PointI> c((const PointI*)corner, (const PointI*)corner + nc), s((const PointI*)surf, (const PointI*)surf + ns);
  Iso r = m->process(od, c, s);
  std::memcpy(outR, r.R, 36); std::memcpy(outT, r.t, 12);
  if (stats) {
    const MatchResult& q = m->lastMatch;
    stats[0] = q.ok; stats[1] = q.converged; stats[2] = q.tooFewRef; stats[3] = q.tooFewMatches; stats[4] = q.degenerate;
    stats[5] = q.iterations; stats[6] = q.lastRows; stats[7] = q.lastLine; stats[8] = q.lastPlane;
    stats[9] = (int)m->cornerDS.size(); stats[10] = (int)m->surfDS.size();
    stats[11] = (int)m->surroundCorner.size(); stats[12] = (int)m->surroundSurf.size();
    size_t wc = 0, ws = 0;
    for (const LocalFrame& f : m->window) { wc += f.corner.size(); ws += f.surf.size(); }
    stats[13] = (int)m->window.size(); stats[14] = (int)wc; stats[15] = (int)ws;
  }
  if (accum) *accum = m->accumDistance;
}
// window contents in queue order (which: 0 corner, 1 surf); NULL out: count only
size_t cmo_mapping_local_window(void* hh, int which, float* out, size_t cap) {
  LaserMappingLocal* m = ((MappingLocalHandle*)hh)->m;
  size_t n = 0;
  for (const LocalFrame& f : m->window) {
    const std::vector<PointI>& v = which == 0 ? f.corner : f.surf;
    if (out && n + v.size() <= cap) std::memcpy(out + 4 * n, v.data(), v.size() * sizeof(PointI));
    n += v.size();
  }
  return n;
}
// saveCloudToFiles enumeration: fills type / i / j / k per file (NULL: count only); cube clouds through cmo_mapping_cube
size_t cmo_mapping_file_order(void* hh, int* type, int* ci, int* cj, int* ck, size_t cap) {
  std::vector<int> t, a, b, c;
  ((MappingHandle*)hh)->m->map.fileOrder(t, a, b, c);
  for (size_t n = 0; n < t.size() && n < cap && type; n++) { type[n] = t[n]; ci[n] = a[n]; cj[n] = b[n]; ck[n] = c[n]; }
  return t.size();
}
size_t cmo_mapping_cube(void* hh, int type, int i, int j, int k, const int* dims, float* out, size_t cap) {
  LaserMapping* m = ((MappingHandle*)hh)->m;
  const std::vector<PointI>& v = (type == 0 ? m->map.cornerCube : m->map.surfCube)[i + j * dims[0] + k * dims[0] * dims[1]];
  if (out) std::memcpy(out, v.data(), std::min(cap, v.size()) * sizeof(PointI));
  return v.size();
}
void cmo_mapping_load_cube(void* hh, int type, int i, int j, int k, const float* cloud, size_t n) {
  std::vector<PointI> c((const PointI*)cloud, (const PointI*)cloud + n);
  ((MappingHandle*)hh)->m->map.loadCube(type, i, j, k, c);
}
// which: 0 surround corner, 1 surround surf (as used by the last process()), 2 cornerDS, 3 surfDS,
// 4 all corner cubes concatenated in cube-index order, 5 all surf cubes
size_t cmo_mapping_cloud(void* hh, int which, float* out, size_t cap) {
  LaserMapping* m = ((MappingHandle*)hh)->m;
  std::vector<PointI> tmp;
  const std::vector<PointI>* v = nullptr;
  if (which == 0) v = &m->surroundCorner; else if (which == 1) v = &m->surroundSurf;
  else if (which == 2) v = &m->cornerDS; else if (which == 3) v = &m->surfDS;
  else {
    auto& cubes = which == 4 ? m->map.cornerCube : m->map.surfCube;
    for (auto& c : cubes) tmp.insert(tmp.end(), c.begin(), c.end());
    v = &tmp;
  }
  if (out) std::memcpy(out, v->data(), std::min(cap, v->size()) * sizeof(PointI));
  return v->size();
}
void cmo_mapping_origin(void* hh, int* out3) {   // _cubeOriginWidth / Height / Depth (moved by FeatureMap::update's shift)
  LaserMapping* m = ((MappingHandle*)hh)->m;
  out3[0] = m->map.originW; out3[1] = m->map.originH; out3[2] = m->map.originD;
}
// direct map access for tests: update + surround, add
void cmo_mapping_map_update(void* hh, const float* sensor) { ((MappingHandle*)hh)->m->map.update(sensor); }
void cmo_mapping_map_add(void* hh, const float* corner, size_t nc, const float* surf, size_t ns, const float* R, const float* t) {
  LaserMapping* m = ((MappingHandle*)hh)->m;
  Iso tf; std::memcpy(tf.R, R, 36); std::memcpy(tf.t, t, 12);
  std::vector<PointI> c((const PointI*)corner, (const PointI*)corner + nc), s((const PointI*)surf, (const PointI*)surf + ns);
  m->map.addFeatureCloud(c, s, tf);
}
size_t cmo_mapping_map_surround(void* hh, int which, float* out, size_t cap) {
  LaserMapping* m = ((MappingHandle*)hh)->m;
  std::vector<PointI> c, s;
  m->map.getSurroundFeature(c, s);
  const std::vector<PointI>& v = which == 0 ? c : s;
  if (out) std::memcpy(out, v.data(), std::min(cap, v.size()) * sizeof(PointI));
  return v.size();
}

// ---- odometry ---------------------------------------------------------------------------------------------------
struct OdomHandle { LaserOdometry o; KnnBackend knn; };
void* cmo_odom_create(int useNanoflann) { OdomHandle* h = new OdomHandle(); h->knn = pick_backend(useNanoflann); return h; }
void cmo_odom_free(void* h) { delete (OdomHandle*)h; }
// out: transform[6], Tsum R[9] t[3]; counts: iterations, lastRows, nLastCorner, nLastSurf, nLog
void cmo_odom_process(void* hh, const float* sharp, size_t n0, const float* lessSharp, size_t n1, const float* flat, size_t n2,
                      const float* lessFlat, size_t n3, float* transform, float* R, float* t, int* counts) {
  OdomHandle* h = (OdomHandle*)hh;
  auto vec = [](const float* p, size_t n) { return std::vector<PointI>((const PointI*)p, (const PointI*)p + n); };
  h->o.process(vec(sharp, n0), vec(lessSharp, n1), vec(flat, n2), vec(lessFlat, n3), h->knn);
  std::memcpy(transform, h->o.transform, 24); std::memcpy(R, h->o.Tsum.R, 36); std::memcpy(t, h->o.Tsum.t, 12);
  counts[0] = h->o.iterations; counts[1] = h->o.lastRows; counts[2] = (int)h->o.lastCorner.size(); counts[3] = (int)h->o.lastSurf.size();
  counts[4] = (int)h->o.log.size();
}
void cmo_odom_last_clouds(void* hh, float* corner, float* surf) {
  OdomHandle* h = (OdomHandle*)hh;
  copy_out(h->o.lastCorner, corner); copy_out(h->o.lastSurf, surf);
}
void cmo_odom_log(void* hh, int it, float* pose_in, float* x, int* rows) {
  const OdomIterLog& l = ((OdomHandle*)hh)->o.log[it];
  std::memcpy(pose_in, l.pose_in, 24); std::memcpy(x, l.x, 24); *rows = l.rows;
}
size_t cmo_odom_indices(void* hh, int* out, size_t cap) {
  OdomHandle* h = (OdomHandle*)hh;
  if (out) std::memcpy(out, h->o.ind.data(), std::min(cap, h->o.ind.size()) * sizeof(int));
  return h->o.ind.size();
}

}  // extern "C"
