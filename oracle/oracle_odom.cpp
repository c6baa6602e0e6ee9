// oracle_odom.cpp -- CPU restatement of the scan-to-scan odometry stage (TEST INFRASTRUCTURE, see cm_oracle.h).
// Follows L_SLAM/src/odometry/LaserOdometry.cpp:135-190 (transformToStart / transformToEnd), :288-326 (process),
// :328-647 (scanMatch), :649-653 (transformUpdate) and the 4-argument coefficient overloads util/feature_utils.h:28-61,77-95.
// Canonical choices beyond those listed in oracle_match.cpp:
//  * the reference bounds the ring-neighbour scans by the CURRENT feature count (LaserOdometry.cpp:370,434),
//    which reads past the end of the last cloud when that count is larger; the bound used here is
//    min(feature count, last cloud size) -- identical whenever the reference stays in bounds.
//  * the nearest neighbour of the 1-NN search is the (d2, index) minimum.
#include "cm_oracle.h"
#include "../the-cooper-mapper_b200/csrc/cm_math.h"
#include <cfloat>
#include <cmath>
#include <cstring>

namespace cmo {

static inline float sqdiff(const PointI& a, const float b[3]) {   // calcSquaredDiff(a, b), math_utils.h:45-51
  float dx = a.x - b[0], dy = a.y - b[1], dz = a.z - b[2];
  return dx * dx + dy * dy + dz * dz;
}
static inline float norm3(float x, float y, float z) { return std::sqrt(x * x + y * y + z * z); }
static inline float rad2degf(float r) { return (float)(r * 180.0 / M_PI); }

// getLinePointDistance + 4-argument getCornerFeatureCoefficients, feature_utils.h:17-26, 42-61
static bool corner_coeff_iter(const PointI& A, const PointI& B, const float X[3], int iter, float co[4]) {
  float bx = X[0] - B.x, by = X[1] - B.y, bz = X[2] - B.z;
  float ax = X[0] - A.x, ay = X[1] - A.y, az = X[2] - A.z;
  float kx = by * az - bz * ay, ky = bz * ax - bx * az, kz = bx * ay - by * ax;
  float knorm = norm3(kx, ky, kz);
  float lengthAB = norm3(A.x - B.x, A.y - B.y, A.z - B.z);
  float ex = B.x - A.x, ey = B.y - A.y, ez = B.z - A.z;
  float ux = ky * ez - kz * ey, uy = kz * ex - kx * ez, uz = kx * ey - ky * ex;
  float den = knorm * lengthAB;
  float dirx = -ux / den, diry = -uy / den, dirz = -uz / den;
  float distance = knorm / lengthAB;
  float weight = 1.0;
  if (iter >= 5) weight = (float)(1 - 1.8 * std::fabs((double)distance));
  co[0] = dirx * weight; co[1] = diry * weight; co[2] = dirz * weight; co[3] = distance * weight;
  return ((double)weight > 0.1 && distance != 0);
}

// getSurfacePointDistance + 4-argument getSurfaceFeatureCoefficients, feature_utils.h:28-40, 77-95
static bool surf_coeff_iter(const PointI& A, const PointI& B, const PointI& C, const float X[3], int iter, float co[4]) {
  float b0 = B.x - A.x, b1 = B.y - A.y, b2 = B.z - A.z, c0 = C.x - A.x, c1 = C.y - A.y, c2 = C.z - A.z;
  float nx = b1 * c2 - b2 * c1, ny = b2 * c0 - b0 * c2, nz = b0 * c1 - b1 * c0;
  float nn = norm3(nx, ny, nz);
  if (nn > 0.f) { nx /= nn; ny /= nn; nz /= nn; }   // MatrixBase::normalize()
  float dsigned = ((X[0] - A.x) * nx + (X[1] - A.y) * ny) + (X[2] - A.z) * nz;
  float cosv = dsigned / norm3(nx, ny, nz) / norm3(A.x - X[0], A.y - X[1], A.z - X[2]);
  if (cosv < 0) { nx *= -1.0f; ny *= -1.0f; nz *= -1.0f; }
  float distance = (float)std::fabs((double)dsigned);
  float weight = 1;
  if (iter >= 5) weight = (float)(1 - 1.8 * std::fabs((double)distance) / std::sqrt((double)norm3(X[0], X[1], X[2])));
  co[0] = weight * nx; co[1] = weight * ny; co[2] = weight * nz; co[3] = weight * distance;
  return ((double)weight > 0.1 && distance != 0);
}

// transformToStart, LaserOdometry.cpp:135-142: s = 10 * frac(intensity); Twist t = _transform * s; po = T(t) * pi
static void to_start(const float tf[6], const PointI& pi, float out[3]) {
  float s = 10 * (pi.intensity - int(pi.intensity));
  float t[6];
  for (int k = 0; k < 6; k++) t[k] = tf[k] * s;   // Twist::operator*(scale), Twist.h:28-35
  float R[9];
  cm::pose_to_matrix(t, R);
  cm::transform_point(R, t + 3, pi.x, pi.y, pi.z, &out[0], &out[1], &out[2]);
}

LaserOdometry::LaserOdometry() { Tsum = iso_identity(); for (int k = 0; k < 6; k++) transform[k] = 0.f; }

// LaserOdometry::scanMatch, LaserOdometry.cpp:328-647
void LaserOdometry::scanMatch(const std::vector<PointI>& sharp, const std::vector<PointI>& flat, const KnnBackend& knn) {
  iterations = 0; lastRows = 0;
  log.clear();
  bool isDegenerate = false;
  float matP[36];
  const size_t lastC = lastCorner.size(), lastS = lastSurf.size();
  if (!(lastC > 10 && lastS > 100)) return;
  const size_t nSharp = sharp.size(), nFlat = flat.size();
  ind.assign(2 * nSharp + 3 * nFlat, -1);
  int* c1 = ind.data(); int* c2 = c1 + nSharp; int* s1 = c2 + nSharp; int* s2 = s1 + nFlat; int* s3 = s2 + nFlat;
  void *kdC = nullptr, *kdS = nullptr;
  knn.build(lastCorner.data(), lastC, &kdC);
  knn.build(lastSurf.data(), lastS, &kdS);
  const int boundC = (int)std::min(nSharp, lastC), boundS = (int)std::min(nFlat, lastS);
  std::vector<PointI> ori, coeffs;
  for (int iterCount = 0; iterCount < maxIterations; iterCount++) {
    ori.clear(); coeffs.clear();
    for (size_t i = 0; i < nSharp; i++) {
      float sel[3];
      to_start(transform, sharp[i], sel);
      if (iterCount % 5 == 0) {
        int id; float d2;
        knn.query(kdC, sel, 1, &id, &d2);
        int closest = -1, min2 = -1;
        if (d2 < 25) {
          closest = id;
          int scan = int(lastCorner[closest].intensity);
          float minD2 = 25;
          for (int j = closest + 1; j < boundC; j++) {
            if (int(lastCorner[j].intensity) > scan + 2.5) break;
            float d = sqdiff(lastCorner[j], sel);
            if (int(lastCorner[j].intensity) > scan) { if (d < minD2) { minD2 = d; min2 = j; } }
          }
          for (int j = closest - 1; j >= 0; j--) {
            if (int(lastCorner[j].intensity) < scan - 2.5) break;
            float d = sqdiff(lastCorner[j], sel);
            if (int(lastCorner[j].intensity) < scan) { if (d < minD2) { minD2 = d; min2 = j; } }
          }
        }
        c1[i] = closest; c2[i] = min2;
      }
      if (c2[i] >= 0) {
        float co[4];
        if (corner_coeff_iter(lastCorner[c1[i]], lastCorner[c2[i]], sel, iterCount, co)) {
          ori.push_back(sharp[i]); coeffs.push_back(PointI{co[0], co[1], co[2], co[3]});
        }
      }
    }
    for (size_t i = 0; i < nFlat; i++) {
      float sel[3];
      to_start(transform, flat[i], sel);
      if (iterCount % 5 == 0) {
        int id; float d2;
        knn.query(kdS, sel, 1, &id, &d2);
        int closest = -1, min2 = -1, min3 = -1;
        if (d2 < 25) {
          closest = id;
          int scan = int(lastSurf[closest].intensity);
          float minD2 = 25, minD3 = 25;
          for (int j = closest + 1; j < boundS; j++) {
            if (int(lastSurf[j].intensity) > scan + 2.5) break;
            float d = sqdiff(lastSurf[j], sel);
            if (int(lastSurf[j].intensity) <= scan) { if (d < minD2) { minD2 = d; min2 = j; } }
            else { if (d < minD3) { minD3 = d; min3 = j; } }
          }
          for (int j = closest - 1; j >= 0; j--) {
            if (int(lastSurf[j].intensity) < scan - 2.5) break;
            float d = sqdiff(lastSurf[j], sel);
            if (int(lastSurf[j].intensity) >= scan) { if (d < minD2) { minD2 = d; min2 = j; } }
            else { if (d < minD3) { minD3 = d; min3 = j; } }
          }
        }
        s1[i] = closest; s2[i] = min2; s3[i] = min3;
      }
      if (s2[i] >= 0 && s3[i] >= 0) {
        float co[4];
        if (surf_coeff_iter(lastSurf[s1[i]], lastSurf[s2[i]], lastSurf[s3[i]], sel, iterCount, co)) {
          ori.push_back(flat[i]); coeffs.push_back(PointI{co[0], co[1], co[2], co[3]});
        }
      }
    }
    const int pointSelNum = (int)ori.size();
    lastRows = pointSelNum;
    OdomIterLog lg;
    std::memcpy(lg.pose_in, transform, sizeof(transform)); lg.rows = pointSelNum;
    std::memset(lg.x, 0, sizeof(lg.x));
    if (pointSelNum < 10) { log.push_back(lg); continue; }   // :501-503
    float srx, crx, sry, cry, srz, crz;
    cm::cm_sincosf(transform[0], &srx, &crx); cm::cm_sincosf(transform[1], &sry, &cry); cm::cm_sincosf(transform[2], &srz, &crz);
    double AtAd[36], AtBd[6];
    for (int t = 0; t < 36; t++) AtAd[t] = 0.0;
    for (int t = 0; t < 6; t++) AtBd[t] = 0.0;
    for (int i = 0; i < pointSelNum; i++) {
      const PointI& pointOri = ori[i];
      const PointI& coeff = coeffs[i];
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
      float b = (float)(-0.05 * coeff.intensity);   // :575
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) AtAd[r * 6 + c] += (double)row[r] * (double)row[c];
        AtBd[r] += (double)row[r] * (double)b;
      }
    }
    float AtA[36], AtB[6], matX[6];
    for (int t = 0; t < 36; t++) AtA[t] = (float)AtAd[t];
    for (int t = 0; t < 6; t++) AtB[t] = (float)AtBd[t];
    { float Aw[36], bw[6]; std::memcpy(Aw, AtA, sizeof(Aw)); std::memcpy(bw, AtB, sizeof(bw)); cm::colpiv_qr_solve<6, 6>(Aw, bw, matX); }
    if (iterCount == 0) {
      float matE[6], matV[36], matV2[36];
      cm::eig_sym<6>(AtA, matE, matV);
      std::memcpy(matV2, matV, sizeof(matV));
      isDegenerate = false;
      for (int i = 0; i < 6; i++) {
        if (matE[i] < 10.f) { for (int j = 0; j < 6; j++) matV2[i * 6 + j] = 0; isDegenerate = true; }
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
    if (isDegenerate) {
      float x2[6]; std::memcpy(x2, matX, sizeof(x2));
      for (int r = 0; r < 6; r++) { float s = 0.f; for (int k = 0; k < 6; k++) s += matP[r * 6 + k] * x2[k]; matX[r] = s; }
    }
    for (int k = 0; k < 6; k++) transform[k] += matX[k];
    for (int k = 0; k < 6; k++) if (!std::isfinite(transform[k])) transform[k] = 0.f;   // :622-634
    iterations = iterCount + 1;
    std::memcpy(lg.x, matX, sizeof(matX));
    log.push_back(lg);
    float deltaR = (float)std::sqrt(std::pow((double)rad2degf(matX[0]), 2) + std::pow((double)rad2degf(matX[1]), 2) + std::pow((double)rad2degf(matX[2]), 2));
    float deltaT = (float)std::sqrt(std::pow((double)(matX[3] * 100), 2) + std::pow((double)(matX[4] * 100), 2) + std::pow((double)(matX[5] * 100), 2));
    if (deltaR < deltaRAbort && deltaT < deltaTAbort) break;
  }
  knn.free(kdC); knn.free(kdS);
}

// transformToEnd, LaserOdometry.cpp:156-168
static void to_end(const float tf[6], std::vector<PointI>& cloud) {
  Iso it; twist_to_iso(tf, it);
  Iso inv = iso_inverse(it);
  for (PointI& p : cloud) {
    float st[3];
    to_start(tf, p, st);
    float x, y, z;
    cm::transform_point(inv.R, inv.t, st[0], st[1], st[2], &x, &y, &z);
    p.x = x; p.y = y; p.z = z;
  }
}

// LaserOdometry::process, LaserOdometry.cpp:288-326
void LaserOdometry::process(const std::vector<PointI>& sharp, const std::vector<PointI>& lessSharp, const std::vector<PointI>& flat,
                            const std::vector<PointI>& lessFlat, const KnnBackend& knn) {
  if (!systemInited) {
    lastCorner = lessSharp; lastSurf = lessFlat;
    systemInited = true;
    return;
  }
  scanMatch(sharp, flat, knn);
  Iso update; twist_to_iso(transform, update);   // transformUpdate :649-653
  Tsum = iso_mul(Tsum, update);
  std::vector<PointI> c = lessSharp, s = lessFlat;
  to_end(transform, c); to_end(transform, s);
  lastCorner.swap(c); lastSurf.swap(s);
}

}  // namespace cmo
