// oracle_match.cpp -- CPU restatement of the scan-to-map Gauss-Newton solver (TEST INFRASTRUCTURE, see
// cm_oracle.h).  Follows L_SLAM/src/scan_to_scan_match/ScanMatch.cpp:51-347, util/feature_utils.h:17-26,63-75,
// 97-204, util/transform_utils.h:288-331,476-482, util/Angle.h:17-29, util/Twist.h.
// Canonical choices where the reference's arithmetic is compiler/library defined (documented in DESIGN.md):
//  * AtA / AtB (Eigen dynamic float GEMM, ScanMatch.cpp:206-208; its blocked/vectorised summation order is
//    unknowable) are defined as the CORRECTLY ROUNDED float product: float products are exact in double, they
//    are accumulated in double and rounded to float once.  Any summation order gives the same float (up to
//    ~1e-13 relative before rounding), which is what lets the GPU's tree reduction match bit for bit.
//  * the 5 neighbours of a query are ordered by (d2, index) (nanoflann orders ties by traversal).
//  * unqualified fabs()/sqrt() in feature_utils.h resolve to the C double overloads.
#include "cm_oracle.h"
#include "../the-cooper-mapper_b200/csrc/cm_math.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace cmo {

static inline float norm3(float x, float y, float z) { return std::sqrt(x * x + y * y + z * z); }

// ---- brute-force KNN, the canonical (d2, index) order ------------------------------------------------------
void knn_brute(const PointI* pts, size_t n, const float q[3], int k, int* idx, float* d2) {
  int cnt = 0;
  for (size_t p = 0; p < n; p++) {
    float dx = q[0] - pts[p].x, dy = q[1] - pts[p].y, dz = q[2] - pts[p].z;
    float d = 0.f;             // L2_Simple_Adaptor::evalMetric nanoflann.hpp:364-372
    d += dx * dx; d += dy * dy; d += dz * dz;
    if (cnt == k && !(d < d2[k - 1])) continue;   // equal distance: the lower index (seen first) stays
    int pos = cnt < k ? cnt : k - 1;
    while (pos > 0 && d2[pos - 1] > d) { d2[pos] = d2[pos - 1]; idx[pos] = idx[pos - 1]; pos--; }
    d2[pos] = d; idx[pos] = (int)p;
    if (cnt < k) cnt++;
  }
  for (int i = cnt; i < k; i++) { idx[i] = -1; d2[i] = FLT_MAX; }
}
namespace {
struct BruteHandle { const PointI* pts; size_t n; };
void bf_build(const PointI* pts, size_t n, void** h) { *h = new BruteHandle{pts, n}; }
void bf_query(void* h, const float q[3], int k, int* idx, float* d2) {
  BruteHandle* b = (BruteHandle*)h; knn_brute(b->pts, b->n, q, k, idx, d2);
}
void bf_free(void* h) { delete (BruteHandle*)h; }
}  // namespace
KnnBackend brute_force_backend() { return KnnBackend{bf_build, bf_query, bf_free}; }

// ---- feature_utils.h ---------------------------------------------------------------------------------------
// findLine, feature_utils.h:108-154
bool find_line(const PointI* cloud, const int* ind, float lineA[3], float lineB[3]) {
  float cx = 0, cy = 0, cz = 0;
  for (int j = 0; j < 5; j++) { cx += cloud[ind[j]].x; cy += cloud[ind[j]].y; cz += cloud[ind[j]].z; }
  cx /= 5.0f; cy /= 5.0f; cz /= 5.0f;
  float a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
  for (int j = 0; j < 5; j++) {
    float ax = cloud[ind[j]].x - cx, ay = cloud[ind[j]].y - cy, az = cloud[ind[j]].z - cz;
    a00 += ax * ax; a10 += ax * ay; a20 += ax * az; a11 += ay * ay; a21 += ay * az; a22 += az * az;
  }
  float A[6] = {a00 / 5.0f, a10 / 5.0f, a20 / 5.0f, a11 / 5.0f, a21 / 5.0f, a22 / 5.0f};
  float w[3], V[9];
  cm::eig3_sym(A, w, V);
  if (w[2] > 5 * w[1]) {
    float vx = V[2], vy = V[5], vz = V[8];
    lineA[0] = cx - vx * 0.1f; lineA[1] = cy - vy * 0.1f; lineA[2] = cz - vz * 0.1f;
    lineB[0] = cx + vx * 0.1f; lineB[1] = cy + vy * 0.1f; lineB[2] = cz + vz * 0.1f;
    return true;
  }
  return false;
}

// findPlane, feature_utils.h:157-204
bool find_plane(const PointI* cloud, const int* ind, float maxDistance, float plane[4]) {
  float A[15], B[5], X[3];
  float cx = 0, cy = 0, cz = 0;
  for (int j = 0; j < 5; j++) {
    cx += cloud[ind[j]].x; cy += cloud[ind[j]].y; cz += cloud[ind[j]].z;
    A[j * 3 + 0] = cloud[ind[j]].x; A[j * 3 + 1] = cloud[ind[j]].y; A[j * 3 + 2] = cloud[ind[j]].z;
    B[j] = -1.f;
  }
  cx /= 5.0f; cy /= 5.0f; cz /= 5.0f;
  cm::colpiv_qr_solve<5, 3>(A, B, X);
  plane[0] = X[0]; plane[1] = X[1]; plane[2] = X[2]; plane[3] = 0;
  float norm = std::sqrt(plane[0] * plane[0] + plane[1] * plane[1] + plane[2] * plane[2] + plane[3] * plane[3]);
  plane[0] /= norm; plane[1] /= norm; plane[2] /= norm; plane[3] /= norm;
  plane[3] = -(plane[0] * cx + plane[1] * cy + plane[2] * cz);
  for (int j = 0; j < 5; j++) {
    float distance = (plane[0] * cloud[ind[j]].x + plane[1] * cloud[ind[j]].y + plane[2] * cloud[ind[j]].z) + plane[3];
    if (std::fabs((double)distance) > maxDistance) return false;
  }
  return true;
}

// getLinePointDistance + getCornerFeatureCoefficients, feature_utils.h:17-26, 63-75
bool corner_coefficients(const float A[3], const float B[3], const float X[3], float coeff[4]) {
  float bx = X[0] - B[0], by = X[1] - B[1], bz = X[2] - B[2];
  float ax = X[0] - A[0], ay = X[1] - A[1], az = X[2] - A[2];
  float kx = by * az - bz * ay, ky = bz * ax - bx * az, kz = bx * ay - by * ax;   // (X-B) x (X-A)
  float knorm = norm3(kx, ky, kz);
  float lengthAB = norm3(A[0] - B[0], A[1] - B[1], A[2] - B[2]);
  float ex = B[0] - A[0], ey = B[1] - A[1], ez = B[2] - A[2];
  float ux = ky * ez - kz * ey, uy = kz * ex - kx * ez, uz = kx * ey - ky * ex;   // k x (B-A)
  float den = knorm * lengthAB;
  float dirx = -ux / den, diry = -uy / den, dirz = -uz / den;
  float distance = knorm / lengthAB;
  float weight = (float)(1 - 0.9f * std::fabs((double)distance));
  coeff[0] = dirx * weight; coeff[1] = diry * weight; coeff[2] = dirz * weight;
  coeff[3] = distance * weight;
  return ((double)weight > 0.1);
}

// getSurfaceFeatureCoefficients(planeCoef, X, coefficients), feature_utils.h:97-106
bool surface_coefficients(const float plane[4], const float X[3], float coeff[4]) {
  float distance = ((plane[0] * X[0] + plane[1] * X[1]) + plane[2] * X[2]) + plane[3];
  float xn = norm3(X[0], X[1], X[2]);
  float weight = (float)(1 - 0.9 * std::fabs((double)distance) / std::sqrt((double)xn));
  coeff[0] = plane[0] * weight; coeff[1] = plane[1] * weight; coeff[2] = plane[2] * weight;
  coeff[3] = distance * weight;
  return ((double)weight > 0.1);
}

// Angle(float) : Angle.h:19-20 (sin/cos through cm_sincosf, see cm_math.h header comment)
struct AngleO {
  float rad = 0.f, c = 1.f, s = 0.f;
  void set(float r) { rad = r; cm::cm_sincosf(r, &s, &c); }
};

static inline float rad2degf(float r) { return (float)(r * 180.0 / M_PI); }   // math_utils.h:24

// The Gauss-Newton loop shared by ScanMatch::scanMatchScan (ScanMatch.cpp:51-347) and FeatureMap::scanMatchScan
// (FeatureMap.h:490-690): the two differ only in where the 5 neighbours of a query come from (`lookup`: returns the cloud
// the indices refer to, or NULL when the reference `continue`s before searching).
void scan_match_impl(const MatchParams& prm, const NeighbourLookup& lookup, const PointI* corner, size_t CornerNum,
                     const PointI* surf, size_t SurfNum, float pose[6], MatchResult& res, bool keepLog) {
  AngleO rot_x, rot_y, rot_z;
  rot_x.set(pose[0]); rot_y.set(pose[1]); rot_z.set(pose[2]);
  float pos[3] = {pose[3], pose[4], pose[5]};
  bool converge = false, isDegenerate = false;
  float matP[36];
  std::vector<PointI> laserCloudOri, coeffSel;
  int line_match_count = 0, plane_match_count = 0;
  int ind[5]; float sq[5];
  for (int iterCount = 0; iterCount < prm.maxIterations; iterCount++) {
    laserCloudOri.clear(); coeffSel.clear();
    line_match_count = 0; plane_match_count = 0;
    IterLog lg;
    float cur[6] = {rot_x.rad, rot_y.rad, rot_z.rad, pos[0], pos[1], pos[2]};
    if (keepLog) { std::memcpy(lg.pose_in, cur, sizeof(cur)); lg.nnCorner.assign(5 * CornerNum, -1); lg.nnSurf.assign(5 * SurfNum, -1); }
    float Rm[9];
    cm::pose_to_matrix(cur, Rm);   // pointAssociateToMap -> convertTransform(Twist, Isometry)
    for (size_t i = 0; i < CornerNum; i++) {
      const PointI& pointOri = corner[i];
      float sel[3];
      cm::transform_point(Rm, pos, pointOri.x, pointOri.y, pointOri.z, &sel[0], &sel[1], &sel[2]);
      const PointI* refCorner = lookup(true, sel, ind, sq);
      if (!refCorner) continue;
      if (sq[4] < prm.knnGate) {
        if (keepLog) for (int t = 0; t < 5; t++) lg.nnCorner[5 * i + t] = ind[t];
        float lineA[3], lineB[3];
        if (find_line(refCorner, ind, lineA, lineB)) {
          float co[4];
          if (corner_coefficients(lineA, lineB, sel, co)) {
            laserCloudOri.push_back(pointOri);
            coeffSel.push_back(PointI{co[0], co[1], co[2], co[3]});
          }
          line_match_count++;
        }
      }
    }
    for (size_t i = 0; i < SurfNum; i++) {
      const PointI& pointOri = surf[i];
      float sel[3];
      cm::transform_point(Rm, pos, pointOri.x, pointOri.y, pointOri.z, &sel[0], &sel[1], &sel[2]);
      const PointI* refSurf = lookup(false, sel, ind, sq);
      if (!refSurf) continue;
      if (sq[4] < prm.knnGate) {
        if (keepLog) for (int t = 0; t < 5; t++) lg.nnSurf[5 * i + t] = ind[t];
        float plane[4];
        if (find_plane(refSurf, ind, prm.planeMaxDistance, plane)) {
          float co[4];
          if (surface_coefficients(plane, sel, co)) {
            laserCloudOri.push_back(pointOri);
            coeffSel.push_back(PointI{co[0], co[1], co[2], co[3]});
          }
          plane_match_count++;
        }
      }
    }
    float srx = rot_x.s, crx = rot_x.c, sry = rot_y.s, cry = rot_y.c, srz = rot_z.s, crz = rot_z.c;
    size_t laserCloudSelNum = laserCloudOri.size();
    res.lastRows = (int)laserCloudSelNum; res.lastLine = line_match_count; res.lastPlane = plane_match_count;
    if (laserCloudSelNum < 50) {   // :142-145
      res.tooFewMatches = true;
      if (keepLog) { lg.rows = (int)laserCloudSelNum; lg.lineMatches = line_match_count; lg.planeMatches = plane_match_count;
        std::memset(lg.AtA, 0, sizeof(lg.AtA)); std::memset(lg.AtB, 0, sizeof(lg.AtB)); std::memset(lg.x, 0, sizeof(lg.x));
        lg.degenerate = isDegenerate; res.log.push_back(lg); }
      break;
    }
    float AtA[36], AtB[6], matX[6];
    double AtAd[36], AtBd[6];
    for (int t = 0; t < 36; t++) AtAd[t] = 0.0;
    for (int t = 0; t < 6; t++) AtBd[t] = 0.0;
    for (size_t i = 0; i < laserCloudSelNum; i++) {
      const PointI& pointOri = laserCloudOri[i];
      const PointI& coeff = coeffSel[i];
      // ScanMatch.cpp:185-195, literal (including the unparenthesised arz middle term and 0*coeff.z)
      float arx = ((crz*sry*crx + srz*srx)* pointOri.y +(srz*crx-crz*sry*srx)* pointOri.z)*coeff.x +
        ((srz*sry*crx-crz*srx)*pointOri.y -(srz*sry*srx+crz*crx)*pointOri.z)*coeff.y +
        (cry*crx*pointOri.y-cry*srx*pointOri.z)*coeff.z;
      float ary = (-crz*sry*pointOri.x+crz*cry*srx*pointOri.y+crz*cry*crx*pointOri.z)*coeff.x +
        (-srz*sry*pointOri.x+srz*cry*srx*pointOri.y +srz*cry*crx*pointOri.z)*coeff.y +
        (-cry*pointOri.x-sry*srx*pointOri.y-sry*crx*pointOri.z)*coeff.z;
      float arz = (-srz*cry*pointOri.x -(srz*sry*srx+crz*crx)*pointOri.y+(crz*srx-srz*sry*crx)*pointOri.z)*coeff.x+
        (crz*cry*pointOri.x+ (crz*sry*srx-srz*crx)*pointOri.y+crz*sry*crx+srz*srx*pointOri.z)*coeff.y+
        0*coeff.z;
      float row[6] = {arx, ary, arz, coeff.x, coeff.y, coeff.z};
      float b = -coeff.intensity;
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) AtAd[r * 6 + c] += (double)row[r] * (double)row[c];
        AtBd[r] += (double)row[r] * (double)b;
      }
    }
    for (int t = 0; t < 36; t++) AtA[t] = (float)AtAd[t];
    for (int t = 0; t < 6; t++) AtB[t] = (float)AtBd[t];
    {
      float Aw[36], bw[6];
      std::memcpy(Aw, AtA, sizeof(Aw)); std::memcpy(bw, AtB, sizeof(bw));
      cm::colpiv_qr_solve<6, 6>(Aw, bw, matX);   // :209
    }
    if (iterCount == 0) {   // :211-235
      float matE[6], matV[36], matV2[36];
      cm::eig_sym<6>(AtA, matE, matV);
      std::memcpy(matV2, matV, sizeof(matV));
      isDegenerate = false;
      for (int i = 0; i < 6; i++) {
        if (matE[i] < 100.f) { for (int j = 0; j < 6; j++) matV2[i * 6 + j] = 0; isDegenerate = true; }
        else break;
      }
      float Vinv[36];
      if (!cm::inverse_lu<6>(matV, Vinv)) { for (int t = 0; t < 36; t++) Vinv[t] = NAN; }
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
          float s = 0.f;
          for (int k = 0; k < 6; k++) s += Vinv[r * 6 + k] * matV2[k * 6 + c];
          matP[r * 6 + c] = s;
        }
    }
    if (isDegenerate) {   // :237-240
      float x2[6];
      std::memcpy(x2, matX, sizeof(x2));
      for (int r = 0; r < 6; r++) {
        float s = 0.f;
        for (int k = 0; k < 6; k++) s += matP[r * 6 + k] * x2[k];
        matX[r] = s;
      }
    }
    rot_x.set(rot_x.rad + matX[0]); rot_y.set(rot_y.rad + matX[1]); rot_z.set(rot_z.rad + matX[2]);
    pos[0] += matX[3]; pos[1] += matX[4]; pos[2] += matX[5];
    res.iterations = iterCount + 1;
    float deltaR = (float)std::sqrt(std::pow((double)rad2degf(matX[0]), 2) + std::pow((double)rad2degf(matX[1]), 2) +
                                    std::pow((double)rad2degf(matX[2]), 2));
    float deltaT = (float)std::sqrt(std::pow((double)(matX[3] * 100), 2) + std::pow((double)(matX[4] * 100), 2) +
                                    std::pow((double)(matX[5] * 100), 2));
    if (keepLog) {
      std::memcpy(lg.AtA, AtA, sizeof(AtA)); std::memcpy(lg.AtB, AtB, sizeof(AtB)); std::memcpy(lg.x, matX, sizeof(matX));
      lg.rows = (int)laserCloudSelNum; lg.lineMatches = line_match_count; lg.planeMatches = plane_match_count;
      lg.degenerate = isDegenerate; res.log.push_back(lg);
    }
    if (deltaR < prm.deltaRAbort && deltaT < prm.deltaTAbort) { converge = true; break; }
  }
  res.converged = converge; res.degenerate = isDegenerate;
  pose[0] = rot_x.rad; pose[1] = rot_y.rad; pose[2] = rot_z.rad; pose[3] = pos[0]; pose[4] = pos[1]; pose[5] = pos[2];
  if (converge && prm.useScore) {   // :263-341 (fineScore is always false, ScanMatch.cpp:32)
    double score = 0;
    for (size_t i = 0; i < coeffSel.size(); i++) score += std::exp(-std::fabs((double)coeffSel[i].intensity));
    res.score = score;
    double match_count = line_match_count + plane_match_count;
    float percent = (float)(match_count / (double)(CornerNum + SurfNum));
    if (score < prm.scoreThreshold) { res.ok = false; return; }
    if ((double)percent < (double)prm.matchPercentageThreshold) { res.ok = false; return; }
    res.ok = true;
    return;
  }
  res.ok = false;   // :342-346 (pose is still written back)
}

// ScanMatch::scanMatchScan(..., Twist&), ScanMatch.cpp:51-347
void scan_match(const MatchParams& prm, const KnnBackend& knn, const PointI* refCorner, size_t nRefCorner,
                const PointI* refSurf, size_t nRefSurf, const PointI* corner, size_t CornerNum,
                const PointI* surf, size_t SurfNum, float pose[6], MatchResult& res, bool keepLog) {
  res = MatchResult();
  if (nRefCorner < 50 || nRefSurf < 100) { res.tooFewRef = true; return; }   // :57-61
  void *kdCorner = nullptr, *kdSurf = nullptr;
  knn.build(refCorner, nRefCorner, &kdCorner);   // :75-76
  knn.build(refSurf, nRefSurf, &kdSurf);
  NeighbourLookup lookup = [&](bool isCorner, const float sel[3], int* ind, float* sq) -> const PointI* {
    knn.query(isCorner ? kdCorner : kdSurf, sel, 5, ind, sq);   // :100-101, 119
    return isCorner ? refCorner : refSurf;
  };
  scan_match_impl(prm, lookup, corner, CornerNum, surf, SurfNum, pose, res, keepLog);
  knn.free(kdCorner); knn.free(kdSurf);
}

// ---- Isometry helpers --------------------------------------------------------------------------------------
Iso iso_identity() { Iso i; for (int k = 0; k < 9; k++) i.R[k] = (k % 4 == 0) ? 1.f : 0.f; i.t[0] = i.t[1] = i.t[2] = 0.f; return i; }
Iso iso_mul(const Iso& a, const Iso& b) {   // 4x4 product restricted to the affine part
  Iso r;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) r.R[i * 3 + j] = (a.R[i * 3 + 0] * b.R[0 * 3 + j] + a.R[i * 3 + 1] * b.R[1 * 3 + j]) + a.R[i * 3 + 2] * b.R[2 * 3 + j];
    r.t[i] = ((a.R[i * 3 + 0] * b.t[0] + a.R[i * 3 + 1] * b.t[1]) + a.R[i * 3 + 2] * b.t[2]) + a.t[i];
  }
  return r;
}
Iso iso_inverse(const Iso& a) {   // Transform::inverse(Isometry): R^T, -R^T t
  Iso r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; i++) r.t[i] = -((r.R[i * 3 + 0] * a.t[0] + r.R[i * 3 + 1] * a.t[1]) + r.R[i * 3 + 2] * a.t[2]);
  return r;
}
void twist_to_iso(const float pose[6], Iso& it) { cm::pose_to_matrix(pose, it.R); it.t[0] = pose[3]; it.t[1] = pose[4]; it.t[2] = pose[5]; }
void iso_to_twist(const Iso& it, float pose[6]) {   // getEulerAngles transform_utils.h:54-60
  pose[3] = it.t[0]; pose[4] = it.t[1]; pose[5] = it.t[2];
  pose[0] = std::atan2(it.R[7], it.R[8]);
  pose[1] = std::asin(-it.R[6]);
  pose[2] = std::atan2(it.R[3], it.R[0]);
}

}  // namespace cmo
