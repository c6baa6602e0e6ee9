// oracle_scanreg.cpp -- CPU restatement of L_SLAM scan registration (TEST INFRASTRUCTURE, see cm_oracle.h).
// Follows L_SLAM/src/odometry/ScanRegistration.cpp:190-666, OrganizedScanRegistration.cpp:82-150,
// MultiScanRegistration.cpp:95-200, MultiScanRegistration.h:57-102, ScanRegistration.h:280-311,
// util/math_utils.h:16-99, util/pcl_util.h:30-37.
#include "cm_oracle.h"
#include "../the-cooper-mapper_b200/csrc/cm_math.h"
#include <algorithm>
#include <cmath>

namespace cmo {

// math_utils.h:45-51
static inline float calcSquaredDiff(const PointIN& a, const PointIN& b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}
// math_utils.h:72-74
static inline float calcPointDistance(const PointIN& p) { return std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z); }
// math_utils.h:82-87
static inline float calcCosAngleDiff(const PointIN& a, const PointIN& b) {
  float ab = a.x * b.x + a.y * b.y + a.z * b.z;
  float disab = calcPointDistance(a) * calcPointDistance(b);
  return ab / disab;
}
// pcl_util.h:30-37
static inline PointI toXYZI(const PointIN& p) { return PointI{p.x, p.y, p.z, p.curvature}; }
static inline float deg2radf(float d) { return (float)(d * M_PI / 180.0); }   // math_utils.h:38
static inline double deg2radd(double d) { return d * M_PI / 180.0; }          // math_utils.h:31

// One 6-point window of pointClassify (ScanRegistration.cpp:557-602 / 603-649): mean, lower-triangle
// covariance /6, eigen-solve, "line" test and the 0.08 m inlier check.  `sign` = -1 walks cloudIdx-j for
// j = 0..R (first block), +1 walks cloudIdx-j for j = -R..0 (second block: cloudIdx+R first ... cloudIdx last).
static bool classify_window(const std::vector<PointIN>& cloud, size_t cloudIdx, int R, bool forward, float v[3]) {
  float cx = 0, cy = 0, cz = 0;
  size_t ids[16];
  int n = 0;
  if (!forward) { for (int j = 0; j <= R; j++) ids[n++] = cloudIdx - j; }
  else { for (int j = -R; j <= 0; j++) ids[n++] = cloudIdx - j; }
  for (int t = 0; t < n; t++) { cx += cloud[ids[t]].x; cy += cloud[ids[t]].y; cz += cloud[ids[t]].z; }
  float cnt = (float)(R + 1);
  cx /= cnt; cy /= cnt; cz /= cnt;
  float a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
  for (int t = 0; t < n; t++) {
    float ax = cloud[ids[t]].x - cx, ay = cloud[ids[t]].y - cy, az = cloud[ids[t]].z - cz;
    a00 += ax * ax; a10 += ax * ay; a20 += ax * az; a11 += ay * ay; a21 += ay * az; a22 += az * az;
  }
  float A[6] = {a00 / cnt, a10 / cnt, a20 / cnt, a11 / cnt, a21 / cnt, a22 / cnt};
  float w[3], V[9];
  cm::eig3_sym(A, w, V);
  if (w[2] > 100 * w[1] && w[2] > 10000 * w[0]) {
    v[0] = V[2]; v[1] = V[5]; v[2] = V[8];
    float vnorm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int t = 0; t < n; t++) {
      float ax = cloud[ids[t]].x - cx, ay = cloud[ids[t]].y - cy, az = cloud[ids[t]].z - cz;
      float kx = ay * v[2] - az * v[1], ky = az * v[0] - ax * v[2], kz = ax * v[1] - ay * v[0];
      float distance = std::sqrt(kx * kx + ky * ky + kz * kz) / vnorm;
      if (std::fabs((double)distance) > 0.08) return false;
    }
    return true;
  }
  return false;
}

// ScanRegistration.cpp:547-666
int point_classify(const std::vector<PointIN>& cloud, size_t cloudIdx, int R) {
  float v1[3], v2[3];
  bool line1 = classify_window(cloud, cloudIdx, R, false, v1);
  bool line2 = classify_window(cloud, cloudIdx, R, true, v2);
  if (line1 && line2) {
    float ab = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
    float disab = std::sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]) *
                  std::sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
    float diff = ab / disab;   // calcCosAngleDiff(Vector3f, Vector3f) math_utils.h:89-93
    if ((double)diff < std::cos(deg2radd(175.0)) || (double)diff > std::cos(deg2radd(5.0))) return SURFACE_FLAT;
    else if ((double)diff > std::cos(deg2radd(135.0)) && (double)diff < std::cos(deg2radd(45.0))) return CORNER_SHARP;
  }
  if (line1 || line2) return ONESIDE_FLAT;
  return MESSY;
}

namespace {
struct Ctx {
  const ScanRegParams& prm;
  ScanRegResult& r;
  std::vector<float> regionCurvature;
  std::vector<size_t> regionSortIndices, swapIdx;
  std::vector<int> picked;   // _scanNeighborPicked of the current ring
  float blindThreshold;

  // ScanRegistration.h:280-311
  void mergeArray(std::vector<size_t>& a, int first, int mid, int last, std::vector<size_t>& tmp) {
    int i = first, j = mid + 1, m = mid, n = last, k = 0;
    while (i <= m && j <= n) {
      if (regionCurvature[a[i]] <= regionCurvature[a[j]]) tmp[k++] = a[i++];
      else tmp[k++] = a[j++];
    }
    while (i <= m) tmp[k++] = a[i++];
    while (j <= n) tmp[k++] = a[j++];
    for (int t = 0; t < k; t++) a[first + t] = tmp[t];
  }
  void mergeSort(std::vector<size_t>& a, int first, int last, std::vector<size_t>& tmp) {
    if (first < last) {
      int mid = (first + last) / 2;
      mergeSort(a, first, mid, tmp);
      mergeSort(a, mid + 1, last, tmp);
      mergeArray(a, first, mid, last, tmp);
    }
  }
  // ScanRegistration.cpp:420-460
  void setRegionBuffersFor(size_t startIdx, size_t endIdx) {
    size_t regionSize = endIdx - startIdx + 1;
    regionCurvature.resize(regionSize);
    regionSortIndices.resize(regionSize);
    swapIdx.resize(regionSize);
    float pointWeight = -2 * prm.curvatureRegion;
    const std::vector<PointIN>& c = r.cloud;
    for (size_t i = startIdx, ri = 0; i <= endIdx; i++, ri++) {
      float diffX = pointWeight * c[i].x, diffY = pointWeight * c[i].y, diffZ = pointWeight * c[i].z;
      for (int j = 1; j <= prm.curvatureRegion; j++) {
        diffX += c[i + j].x + c[i - j].x;
        diffY += c[i + j].y + c[i - j].y;
        diffZ += c[i + j].z + c[i - j].z;
      }
      regionCurvature[ri] = diffX * diffX + diffY * diffY + diffZ * diffZ;
      regionSortIndices[ri] = i - startIdx;
      r.curvature[i] = regionCurvature[ri];
    }
    mergeSort(regionSortIndices, 0, (int)regionSize - 1, swapIdx);
    for (size_t i = 0; i < regionSize; i++) regionSortIndices[i] += startIdx;
  }
  // ScanRegistration.cpp:462-522
  void setScanBuffersFor(size_t startIdx, size_t endIdx) {
    const int R = prm.curvatureRegion;
    size_t scanSize = endIdx - startIdx + 1;
    picked.assign(scanSize, 0);
    const std::vector<PointIN>& c = r.cloud;
    for (int i = 0; i < R; ++i)
      if (calcCosAngleDiff(c[startIdx + i], c[startIdx + i + 1]) < blindThreshold)
        std::fill_n(&picked[i], R + 1, (int)BLIND_BLOCK);
    for (int i = 0; i < R; ++i)
      if (calcCosAngleDiff(c[endIdx - i], c[endIdx - i - 1]) < blindThreshold)
        std::fill_n(&picked[endIdx - i - startIdx - R], R + 1, (int)BLIND_BLOCK);
    for (size_t i = startIdx + R; i < endIdx - R; i++) {
      const PointIN& previousPoint = c[i - 1];
      const PointIN& point = c[i];
      const PointIN& nextPoint = c[i + 1];
      float diffNext = calcSquaredDiff(nextPoint, point);
      if (calcCosAngleDiff(point, nextPoint) < blindThreshold) {
        std::fill_n(&picked[i - startIdx - R + 1], R * 2, (int)BLIND_BLOCK);
        continue;
      }
      if ((double)diffNext > 1.0) {
        float depth1 = calcPointDistance(point);
        float depth2 = calcPointDistance(nextPoint);
        float diffPrev = calcSquaredDiff(previousPoint, point);
        if (depth1 > depth2) {
          if (picked[i - startIdx + 1] > NEAR_BLOCK && (double)(diffPrev / diffNext) < 0.2)
            picked[i - startIdx + 1] = EDGE_BROKEN;
          std::fill_n(&picked[i - startIdx - R + 1], R, (int)NEAR_BLOCK);
        } else {
          if (picked[i - startIdx] > NEAR_BLOCK && (double)(diffPrev / diffNext) < 0.2)
            picked[i - startIdx] = EDGE_BROKEN;
          std::fill_n(&picked[i - startIdx + 1], R, (int)NEAR_BLOCK);
        }
      }
    }
  }
  // ScanRegistration.cpp:524-545
  void markAsPicked(size_t scanIdx, int label) {
    picked[scanIdx] = label;
    for (int i = 1; i <= prm.curvatureRegion; i++) picked[scanIdx + i] = label;
    for (int i = 1; i <= prm.curvatureRegion; i++) picked[scanIdx - i] = label;
  }
};
}  // namespace

// ScanRegistration.cpp:190-418
void extract_features(const ScanRegParams& prm, ScanRegResult& r) {
  Ctx cx{prm, r, {}, {}, {}, {}, 0.f};
  // ScanRegistration.cpp:27,46: cos() is unqualified there -> the C double overload, narrowed into the float member
  cx.blindThreshold = (float)std::cos((double)deg2radf(prm.blindDegreeThreshold));
  const int R = prm.curvatureRegion;
  r.picked.assign(r.cloud.size(), 0);
  r.curvature.assign(r.cloud.size(), -1.f);
  r.classLabel.assign(r.cloud.size(), 0x7f);
  size_t nScans = r.scanStart.size();
  for (size_t i = 0; i < nScans; i++) {
    std::vector<PointI> lessFlatScan;
    size_t scanStartIdx = r.scanStart[i], scanEndIdx = r.scanEnd[i];
    if (scanEndIdx <= scanStartIdx + 2 * R) continue;   // :205
    cx.setScanBuffersFor(scanStartIdx, scanEndIdx);
    for (int j = 0; j < prm.nFeatureRegions; j++) {
      size_t sp = ((scanStartIdx + R) * (prm.nFeatureRegions - j) + (scanEndIdx - R) * j) / prm.nFeatureRegions;
      size_t ep = ((scanStartIdx + R) * (prm.nFeatureRegions - 1 - j) + (scanEndIdx - R) * (j + 1)) /
                      prm.nFeatureRegions - 1;
      if (ep <= sp) continue;
      size_t regionSize = ep - sp + 1;
      cx.setRegionBuffersFor(sp, ep);
      std::vector<int> regionLabel(regionSize, (int)UNKNOW);
      // pass 1: flat surface features (:267-284)
      int surfPickedNum = 0;
      for (size_t k = 0; k < regionSize && surfPickedNum < prm.maxSurfaceFlat; k++) {
        size_t idx = cx.regionSortIndices[k];
        size_t scanIdx = idx - scanStartIdx, regionIdx = idx - sp;
        if (cx.picked[scanIdx] != SURF_PICKED_NEAR && cx.regionCurvature[regionIdx] < prm.surfaceCurvatureThreshold) {
          surfPickedNum++;
          regionLabel[regionIdx] = SURFACE_FLAT;
          r.flat.push_back(toXYZI(r.cloud[idx])); r.flatIdx.push_back((int)idx);
          cx.markAsPicked(scanIdx, SURF_PICKED_NEAR);
        }
      }
      // pass 2: less flat + edge_broken (:286-303)
      for (size_t k = 0; k < regionSize; k++) {
        size_t idx = sp + k, scanIdx = idx - scanStartIdx;
        if (cx.regionCurvature[k] < prm.surfaceCurvatureThreshold) {
          lessFlatScan.push_back(toXYZI(r.cloud[idx]));
          r.lessFlatRawIdx.push_back((int)idx); r.lessFlatRawRing.push_back((int)i);
          if (regionLabel[k] != SURFACE_FLAT) regionLabel[k] = SURFACE_LESS_FLAT;
        }
        if (cx.picked[scanIdx] == EDGE_BROKEN) {
          r.sharp.push_back(toXYZI(r.cloud[idx])); r.sharpIdx.push_back((int)idx);
          r.lessSharp.push_back(toXYZI(r.cloud[idx])); r.lessSharpIdx.push_back((int)idx);
          regionLabel[k] = CORNER_SHARP;
          r.dbgBlock.push_back(toXYZI(r.cloud[idx]));
        }
      }
      // pass 3: classify everything above the curvature threshold, most curved first (:305-354)
      int cornerPickedNum = 0;
      surfPickedNum = 0;
      for (size_t k = regionSize; k > 0;) {
        size_t idx = cx.regionSortIndices[--k];
        size_t scanIdx = idx - scanStartIdx, regionIdx = idx - sp;
        if (cx.regionCurvature[regionIdx] < prm.surfaceCurvatureThreshold) break;
        int label = point_classify(r.cloud, idx, R);
        r.classLabel[idx] = label;
        switch (label) {
          case MESSY: break;
          case SURFACE_FLAT:
            if (surfPickedNum < prm.maxSurfaceFlat) surfPickedNum++;
            lessFlatScan.push_back(toXYZI(r.cloud[idx]));
            r.lessFlatRawIdx.push_back((int)idx); r.lessFlatRawRing.push_back((int)i);
            r.dbgBlind.push_back(toXYZI(r.cloud[idx]));
            break;
          case CORNER_SHARP:
            if (cx.picked[scanIdx] > EDGE_BROKEN) {
              if (cornerPickedNum < prm.maxCornerSharp) {
                cornerPickedNum++;
                r.sharp.push_back(toXYZI(r.cloud[idx])); r.sharpIdx.push_back((int)idx);
              }
              r.lessSharp.push_back(toXYZI(r.cloud[idx])); r.lessSharpIdx.push_back((int)idx);
              r.dbgSlop.push_back(toXYZI(r.cloud[idx]));
            }
            break;
          case ONESIDE_FLAT:
            if (surfPickedNum < prm.maxSurfaceFlat) {
              surfPickedNum++;
              r.flat.push_back(toXYZI(r.cloud[idx])); r.flatIdx.push_back((int)idx);
            }
            lessFlatScan.push_back(toXYZI(r.cloud[idx]));
            r.lessFlatRawIdx.push_back((int)idx); r.lessFlatRawRing.push_back((int)i);
            r.dbgCurv.push_back(toXYZI(r.cloud[idx]));
            break;
        }
      }
    }
    for (size_t t = 0; t < cx.picked.size(); t++) r.picked[scanStartIdx + t] = cx.picked[t];
    // per-ring voxel filter of the less-flat points (:390-399)
    std::vector<PointI> ds;
    voxel_filter(lessFlatScan.data(), lessFlatScan.size(), prm.lessFlatFilterSize, ds);
    r.lessFlat.insert(r.lessFlat.end(), ds.begin(), ds.end());
  }
}

static void finish_rings(std::vector<std::vector<PointIN>>& rings, ScanRegResult& r) {
  // construct sorted full resolution cloud (OrganizedScanRegistration.cpp:128-139, MultiScanRegistration.cpp:179-190)
  size_t cloudSize = 0;
  for (size_t i = 0; i < rings.size(); i++) {
    r.cloud.insert(r.cloud.end(), rings[i].begin(), rings[i].end());
    size_t first = cloudSize;
    cloudSize += rings[i].size();
    r.scanStart.push_back((int)first);
    r.scanEnd.push_back(cloudSize > 0 ? (int)(cloudSize - 1) : 0);
  }
}

// OrganizedScanRegistration.cpp:82-150 (ring = row, PointXYZIT.ring = row)
void scanreg_organised(const ScanRegParams& prm, const float* xyzi, int rows, int cols, ScanRegResult& r) {
  r = ScanRegResult();
  std::vector<std::vector<PointIN>> rings(rows);
  for (int row = 0; row < rows; row++)
    for (int col = 0; col < cols; col++) {
      const float* p = xyzi + 4 * ((size_t)row * cols + col);
      PointIN point;
      point.x = p[0]; point.y = p[1]; point.z = p[2]; point.intensity = p[3];
      float relTime = prm.scanPeriod * static_cast<double>(col) / cols;
      point.curvature = (int)(uint16_t)row + relTime;
      if (!std::isfinite(point.x) || !std::isfinite(point.y) || !std::isfinite(point.z)) continue;
      if (point.x * point.x + point.y * point.y + point.z * point.z < prm.blindRadius * prm.blindRadius) continue;
      rings[row].push_back(point);
    }
  finish_rings(rings, r);
  extract_features(prm, r);
}

// ring of a vertical angle in degrees for the Pandar40 (MultiScanMapperP::getRingForAngle = scanID_pandar(rad2deg(angle)),
// MultiScanRegistration.h:37, lidar_type.h:78-104): piecewise-linear bins; angles no branch covers (exactly -15, -5.8, or >= 7.5
// degrees; the reference also prints "ERROR" for the last) fall to ring 0
static int scan_id_pandar(float angle) {
  int scanID = 0;
  if (angle < -15.0) scanID = 0;
  else if (angle > -15.0 && angle < -5.8) scanID = (int)(angle + 16.0 + 0.5);
  else if (angle > -5.8 && angle < 2.8) scanID = (int)((angle + 5.667) / 0.33 + 0.5) + 10;
  else if (angle > 1.8 && angle < 7.5) scanID = (int)(angle - 2.0 + 0.5) + 34;
  return scanID;
}

// MultiScanRegistration.cpp:95-200.  lidar: 0 VLP-16, 1 HDL-32, 2 HDL-64E (linear mappers, MultiScanRegistration.h:90-102),
// 3 Pandar40 (MultiScanMapperP, :24-42).  atan / atan2 through cm_atanf / cm_atan2f (canonical, see cm_math.h).
void scanreg_sweep(const ScanRegParams& prm, const float* xyzi, int n, int lidar, ScanRegResult& r) {
  r = ScanRegResult();
  float lower, upper; int nRings;
  if (lidar == 0) { lower = -15; upper = 15; nRings = 16; }
  else if (lidar == 1) { lower = -30.67f; upper = 10.67f; nRings = 32; }
  else if (lidar == 3) { lower = -15.444f; upper = 6.96f; nRings = 40; }
  else { lower = -24.9f; upper = 2; nRings = 64; }
  float factor = (nRings - 1) / (upper - lower);   // MultiScanRegistration.h:63
  std::vector<std::vector<PointIN>> rings(nRings);
  if (n > 0) {
    const float* in = xyzi;
    float startOri = -cm::cm_atan2f(in[1], in[0]);
    float endOri = -cm::cm_atan2f(in[4 * (n - 1) + 1], in[4 * (n - 1) + 0]) + 2 * float(M_PI);
    if (endOri - startOri > 3 * M_PI) endOri -= 2 * M_PI;
    else if (endOri - startOri < M_PI) endOri += 2 * M_PI;
    bool halfPassed = false;
    for (int i = 0; i < n; i++) {
      PointIN point;
      point.x = in[4 * i + 1]; point.y = in[4 * i + 2]; point.z = in[4 * i + 0]; point.intensity = in[4 * i + 3];
      if (!std::isfinite(point.x) || !std::isfinite(point.y) || !std::isfinite(point.z)) continue;
      if (point.x * point.x + point.y * point.y + point.z * point.z < 0.0001) continue;
      float angle = cm::cm_atanf(point.y / std::sqrt(point.x * point.x + point.z * point.z));
      int scanID = lidar == 3 ? scan_id_pandar((float)(angle * 180.0 / M_PI))   // rad2deg(float), math_utils.h:24
                              : int(((angle * 180 / M_PI) - lower) * factor + 0.5);   // MultiScanRegistration.h:85-87
      if (scanID >= nRings || scanID < 0) continue;
      float ori = -cm::cm_atan2f(point.x, point.z);
      if (!halfPassed) {
        if (ori < startOri - M_PI / 2) ori += 2 * M_PI;
        else if (ori > startOri + M_PI * 3 / 2) ori -= 2 * M_PI;
        if (ori - startOri > M_PI) halfPassed = true;
      } else {
        ori += 2 * M_PI;
        if (ori < endOri - M_PI * 3 / 2) ori += 2 * M_PI;
        else if (ori > endOri + M_PI / 2) ori -= 2 * M_PI;
      }
      float relTime = prm.scanPeriod * (ori - startOri) / (endOri - startOri);
      point.curvature = scanID + relTime;
      rings[scanID].push_back(point);
    }
  }
  finish_rings(rings, r);
  extract_features(prm, r);
}

// ---- IMU de-skew ------------------------------------------------------------------------------------------------------------
namespace {
struct Ang { float rad, c, s; };                                  // Angle.h: radian + buffered cos / sin (cm_sincosf, see cm_math.h)
inline Ang ang(float r) { Ang a; a.rad = r; cm::cm_sincosf(r, &a.s, &a.c); return a; }
inline Ang neg(const Ang& a) { Ang o; o.rad = -a.rad; o.c = a.c; o.s = -a.s; return o; }   // Angle::operator-()
inline void rotX(float v[3], const Ang& a) { float y = v[1]; v[1] = a.c * y - a.s * v[2]; v[2] = a.s * y + a.c * v[2]; }   // math_utils.h:115-145
inline void rotY(float v[3], const Ang& a) { float x = v[0]; v[0] = a.c * x + a.s * v[2]; v[2] = a.c * v[2] - a.s * x; }
inline void rotZ(float v[3], const Ang& a) { float x = v[0]; v[0] = a.c * x - a.s * v[1]; v[1] = a.s * x + a.c * v[1]; }
inline void rotateZXY(float v[3], const Ang& z, const Ang& x, const Ang& y) { rotZ(v, z); rotX(v, x); rotY(v, y); }   // :184-205
inline void rotateYXZ(float v[3], const Ang& y, const Ang& x, const Ang& z) { rotY(v, y); rotX(v, x); rotZ(v, z); }   // :215-236
// IMUState::interpolate(start, end, ratio, result), ScanRegistration.h:151-169
ImuState imu_interpolate(const ImuState& start, const ImuState& end, float ratio) {
  ImuState r;
  float invRatio = 1 - ratio;
  r.roll = start.roll * invRatio + end.roll * ratio;
  r.pitch = start.pitch * invRatio + end.pitch * ratio;
  if (start.yaw - end.yaw > M_PI) r.yaw = (float)(start.yaw * invRatio + (end.yaw + 2 * M_PI) * ratio);
  else if (start.yaw - end.yaw < -M_PI) r.yaw = (float)(start.yaw * invRatio + (end.yaw - 2 * M_PI) * ratio);
  else r.yaw = start.yaw * invRatio + end.yaw * ratio;
  for (int k = 0; k < 3; k++) { r.vel[k] = start.vel[k] * invRatio + end.vel[k] * ratio; r.pos[k] = start.pos[k] * invRatio + end.pos[k] * ratio; }
  return r;
}
// ScanRegistration::interpolateIMUStateFor (ScanRegistration.cpp:168-186); imuIdx is the member that only moves forward
ImuState imu_state_for(const ImuHistory& imu, double scanTime, float relTime, size_t& imuIdx) {
  double timeDiff = (scanTime - imu.h[imuIdx].stamp) + relTime;
  while (imuIdx < imu.h.size() - 1 && timeDiff > 0) { imuIdx++; timeDiff = (scanTime - imu.h[imuIdx].stamp) + relTime; }
  if (imuIdx == 0 || timeDiff > 0) return imu.h[imuIdx];
  float ratio = (float)(-timeDiff / (imu.h[imuIdx].stamp - imu.h[imuIdx - 1].stamp));
  return imu_interpolate(imu.h[imuIdx], imu.h[imuIdx - 1], ratio);
}
}  // namespace

void ImuHistory::push(double stamp, double roll, double pitch, double yaw, double ax, double ay, double az) {
  ImuState n;
  n.acc[0] = float(ay - std::sin(roll) * std::cos(pitch) * 9.81);     // ScanRegistration.cpp:96-99
  n.acc[1] = float(az - std::cos(roll) * std::cos(pitch) * 9.81);
  n.acc[2] = float(ax + std::sin(pitch) * 9.81);
  n.stamp = stamp; n.roll = (float)roll; n.pitch = (float)pitch; n.yaw = (float)yaw;
  if (!h.empty()) {                                                   // :108-118
    float acc[3] = {n.acc[0], n.acc[1], n.acc[2]};
    rotateZXY(acc, ang(n.roll), ang(n.pitch), ang(n.yaw));
    const ImuState& prev = h.back();
    float timeDiff = float(n.stamp - prev.stamp);
    for (int k = 0; k < 3; k++) {
      n.pos[k] = (prev.pos[k] + (prev.vel[k] * timeDiff)) + (((0.5f * acc[k]) * timeDiff) * timeDiff);
      n.vel[k] = prev.vel[k] + acc[k] * timeDiff;
    }
  }
  if (h.size() < capacity) h.push_back(n);
  else { h.erase(h.begin()); h.push_back(n); }                        // CircularBuffer::push: the oldest state goes
}

void scanreg_sweep_imu(const ScanRegParams& prm, const float* xyzi, int n, int lidar, double scanTime, const ImuHistory& imu,
                       ScanRegResult& r, float imuTrans[12]) {
  r = ScanRegResult();
  for (int k = 0; k < 12; k++) imuTrans[k] = 0.f;
  float lower, upper; int nRings;
  if (lidar == 0) { lower = -15; upper = 15; nRings = 16; }
  else if (lidar == 1) { lower = -30.67f; upper = 10.67f; nRings = 32; }
  else if (lidar == 3) { lower = -15.444f; upper = 6.96f; nRings = 40; }
  else { lower = -24.9f; upper = 2; nRings = 64; }
  float factor = (nRings - 1) / (upper - lower);
  std::vector<std::vector<PointIN>> rings(nRings);
  const bool hasImu = !imu.h.empty();
  size_t imuIdx = 0;                                                  // reset(scanTime): _imuIdx = 0, _imuStart = state at relTime 0
  ImuState imuStart, imuCur;
  float shift[3] = {0, 0, 0};
  if (hasImu) imuStart = imu_state_for(imu, scanTime, 0.f, imuIdx);
  if (n > 0) {
    const float* in = xyzi;
    float startOri = -cm::cm_atan2f(in[1], in[0]);
    float endOri = -cm::cm_atan2f(in[4 * (n - 1) + 1], in[4 * (n - 1) + 0]) + 2 * float(M_PI);
    if (endOri - startOri > 3 * M_PI) endOri -= 2 * M_PI;
    else if (endOri - startOri < M_PI) endOri += 2 * M_PI;
    bool halfPassed = false;
    for (int i = 0; i < n; i++) {
      PointIN point;
      point.x = in[4 * i + 1]; point.y = in[4 * i + 2]; point.z = in[4 * i + 0]; point.intensity = in[4 * i + 3];
      if (!std::isfinite(point.x) || !std::isfinite(point.y) || !std::isfinite(point.z)) continue;
      if (point.x * point.x + point.y * point.y + point.z * point.z < 0.0001) continue;
      float angle = cm::cm_atanf(point.y / std::sqrt(point.x * point.x + point.z * point.z));
      int scanID = lidar == 3 ? scan_id_pandar((float)(angle * 180.0 / M_PI)) : int(((angle * 180 / M_PI) - lower) * factor + 0.5);
      if (scanID >= nRings || scanID < 0) continue;
      float ori = -cm::cm_atan2f(point.x, point.z);
      if (!halfPassed) {
        if (ori < startOri - M_PI / 2) ori += 2 * M_PI;
        else if (ori > startOri + M_PI * 3 / 2) ori -= 2 * M_PI;
        if (ori - startOri > M_PI) halfPassed = true;
      } else {
        ori += 2 * M_PI;
        if (ori < endOri - M_PI * 3 / 2) ori += 2 * M_PI;
        else if (ori > endOri + M_PI / 2) ori -= 2 * M_PI;
      }
      float relTime = prm.scanPeriod * (ori - startOri) / (endOri - startOri);
      point.curvature = scanID + relTime;
      if (hasImu) {                                                   // setIMUTransformFor + transformToStartIMU, :145-166
        imuCur = imu_state_for(imu, scanTime, relTime, imuIdx);
        float relSweepTime = (float)(0.0 + relTime);                  // (_scanTime - _sweepStart).toSec() + relTime, a new sweep
        for (int k = 0; k < 3; k++) shift[k] = (imuCur.pos[k] - imuStart.pos[k]) - imuStart.vel[k] * relSweepTime;
        float v[3] = {point.x, point.y, point.z};
        rotateZXY(v, ang(imuCur.roll), ang(imuCur.pitch), ang(imuCur.yaw));
        v[0] += shift[0]; v[1] += shift[1]; v[2] += shift[2];
        rotateYXZ(v, neg(ang(imuStart.yaw)), neg(ang(imuStart.pitch)), neg(ang(imuStart.roll)));
        point.x = v[0]; point.y = v[1]; point.z = v[2];
      }
      rings[scanID].push_back(point);
    }
  }
  if (hasImu) {                                                       // publishResult, :681-708
    imuTrans[0] = imuStart.pitch; imuTrans[1] = imuStart.yaw; imuTrans[2] = imuStart.roll;
    imuTrans[3] = imuCur.pitch; imuTrans[4] = imuCur.yaw; imuTrans[5] = imuCur.roll;
    float s3[3] = {shift[0], shift[1], shift[2]};
    rotateYXZ(s3, neg(ang(imuStart.yaw)), neg(ang(imuStart.pitch)), neg(ang(imuStart.roll)));
    imuTrans[6] = s3[0]; imuTrans[7] = s3[1]; imuTrans[8] = s3[2];
    float v3[3] = {imuCur.vel[0] - imuStart.vel[0], imuCur.vel[1] - imuStart.vel[1], imuCur.vel[2] - imuStart.vel[2]};
    rotateYXZ(v3, neg(ang(imuStart.yaw)), neg(ang(imuStart.pitch)), neg(ang(imuStart.roll)));
    imuTrans[9] = v3[0]; imuTrans[10] = v3[1]; imuTrans[11] = v3[2];
  }
  finish_rings(rings, r);
  extract_features(prm, r);
}

}  // namespace cmo
