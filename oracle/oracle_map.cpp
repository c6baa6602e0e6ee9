// oracle_map.cpp -- CPU restatement of the cube-grid local map and the LaserMapping frame loop
// (TEST INFRASTRUCTURE, see cm_oracle.h).  Follows L_SLAM/src/util/FeatureMap.h:59-74,102-108,146-148,189-376,
// 464-487, odometry/LaserMatcher.cpp:18-33,288-355, odometry/LaserMapping.cpp:39-59,
// util/transform_utils.h:502-507,601-614, scan_to_scan_match/ScanMatch.cpp:349-360.
#include "cm_oracle.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace cmo {

FeatureMap::FeatureMap(const MapParams& p) : _p(p) {
  int n = p.cubeW * p.cubeH * p.cubeD;
  cornerCube.resize(n); surfCube.resize(n);
  originW = (int)std::round((p.cubeW - 1) / 2.0);   // FeatureMap.h:63-65
  originH = (int)std::round((p.cubeH - 1) / 2.0);
  originD = (int)std::round((p.cubeD - 1) / 2.0);
}
bool FeatureMap::isIndexValid(int i, int j, int k) const {
  return 0 <= i && i < _p.cubeW && 0 <= j && j < _p.cubeH && 0 <= k && k < _p.cubeD;
}
// FeatureMap.h:475-487
bool FeatureMap::worldToCube(float x, float y, float z, int& i, int& j, int& k) const {
  i = (int)(std::round(x / _p.cubeSize) + originW);
  j = (int)(std::round(y / _p.cubeSize) + originH);
  k = (int)(std::round(z / _p.cubeSize) + originD);
  return isIndexValid(i, j, k);
}
// FeatureMap.h:354-376 (literal, including the in-place swap order)
void FeatureMap::shift(int di, int dj, int dk) {
  if (di != 0 || dj != 0 || dk != 0) {
    for (int i = 0; i < _p.cubeW; i++)
      for (int j = 0; j < _p.cubeH; j++)
        for (int k = 0; k < _p.cubeD; k++) {
          int oi = i - di, oj = j - dj, ok = k - dk;
          if (isIndexValid(oi, oj, ok)) {
            std::swap(cornerCube[toIndex(i, j, k)], cornerCube[toIndex(oi, oj, ok)]);
            std::swap(surfCube[toIndex(i, j, k)], surfCube[toIndex(oi, oj, ok)]);
          } else {
            cornerCube[toIndex(i, j, k)].clear();
            surfCube[toIndex(i, j, k)].clear();
          }
        }
  }
}
// FeatureMap.h:308-352
void FeatureMap::computeActiveArea(const float s[3]) {
  _cubeValidInd.clear();
  int window = (int)std::ceil(_p.validDistance / _p.cubeSize);
  for (int i = _curW - window; i <= _curW + window; i++)
    for (int j = _curH - window; j <= _curH + window; j++)
      for (int k = _curD - window; k <= _curD + window; k++) {
        if (!isIndexValid(i, j, k)) continue;
        float centerX = _p.cubeSize * (i - originW);
        float centerY = _p.cubeSize * (j - originH);
        float centerZ = _p.cubeSize * (k - originD);
        bool inFov = false;
        for (int ii = -1; ii <= 1 && !inFov; ii += 2)
          for (int jj = -1; jj <= 1 && !inFov; jj += 2)
            for (int kk = -1; kk <= 1 && !inFov; kk += 2) {
              float cx = (float)(centerX + _p.cubeSize / 2.0 * ii);
              float cy = (float)(centerY + _p.cubeSize / 2.0 * jj);
              float cz = (float)(centerZ + _p.cubeSize / 2.0 * kk);
              float dx = s[0] - cx, dy = s[1] - cy, dz = s[2] - cz;
              float sq = dx * dx + dy * dy + dz * dz;
              if (std::sqrt((double)sq) < _p.validDistance) inFov = true;
            }
        if (inFov) _cubeValidInd.push_back(toIndex(i, j, k));
      }
}
// FeatureMap.h:232-254
void FeatureMap::update(const float s[3]) {
  int gi, gj, gk;
  worldToCube(s[0], s[1], s[2], gi, gj, gk);
  const int PAD = 3;
  int ni = std::min(std::max(gi, PAD), _p.cubeW - PAD - 1);
  int nj = std::min(std::max(gj, PAD), _p.cubeH - PAD - 1);
  int nk = std::min(std::max(gk, PAD), _p.cubeD - PAD - 1);
  shift(ni - gi, nj - gj, nk - gk);
  originW += ni - gi; originH += nj - gj; originD += nk - gk;
  _curW = ni; _curH = nj; _curD = nk;
  computeActiveArea(s);
}
void FeatureMap::getSurroundFeature(std::vector<PointI>& corner, std::vector<PointI>& surf) const {
  corner.clear(); surf.clear();
  for (size_t v : _cubeValidInd) {
    corner.insert(corner.end(), cornerCube[v].begin(), cornerCube[v].end());
    surf.insert(surf.end(), surfCube[v].begin(), surfCube[v].end());
  }
}
// FeatureMap.h:289-306
void FeatureMap::downsizeValidCloud() {
  std::vector<PointI> tmp;
  for (size_t v : _cubeValidInd) {
    voxel_filter(cornerCube[v].data(), cornerCube[v].size(), _p.mapFilterCorner, tmp);
    cornerCube[v].swap(tmp);
    voxel_filter(surfCube[v].data(), surfCube[v].size(), _p.mapFilterSurf, tmp);
    surfCube[v].swap(tmp);
  }
}
// FeatureMap.h:189-230 + transformPointCloud transform_utils.h:601-614
void FeatureMap::addFeatureCloud(const std::vector<PointI>& corner, const std::vector<PointI>& surf, const Iso& tf) {
  auto push = [&](const std::vector<PointI>& in, std::vector<std::vector<PointI>>& cubes) {
    for (const PointI& p : in) {
      PointI q = p;
      q.x = ((tf.R[0] * p.x + tf.R[1] * p.y) + tf.R[2] * p.z) + tf.t[0];
      q.y = ((tf.R[3] * p.x + tf.R[4] * p.y) + tf.R[5] * p.z) + tf.t[1];
      q.z = ((tf.R[6] * p.x + tf.R[7] * p.y) + tf.R[8] * p.z) + tf.t[2];
      int i, j, k;
      if (worldToCube(q.x, q.y, q.z, i, j, k)) cubes[toIndex(i, j, k)].push_back(q);
    }
  };
  push(corner, cornerCube);
  push(surf, surfCube);
  downsizeValidCloud();
}
int FeatureMap::worldToIndex(float x, float y, float z) const {   // FeatureMap.h:464-473
  int i, j, k;
  return worldToCube(x, y, z, i, j, k) ? toIndex(i, j, k) : -1;
}
// FeatureMap.h:490-690.  The reference builds the per-cube KD-trees when the map is loaded (:437,451); here they are built
// on first use inside the call (same trees, the map does not change during a match).
void FeatureMap::scanMatchScan(const KnnBackend& knn, const std::vector<PointI>& corner, const std::vector<PointI>& surf, float pose[6],
                               MatchResult& res, bool keepLog) {
  res = MatchResult();
  MatchParams prm;
  prm.maxIterations = 10; prm.deltaRAbort = 0.05f; prm.deltaTAbort = 0.05f; prm.useScore = false;   // :514, 676
  std::map<int, void*> kdCorner, kdSurf;
  NeighbourLookup lookup = [&](bool isCorner, const float sel[3], int* ind, float* sq) -> const PointI* {
    const int idx = worldToIndex(sel[0], sel[1], sel[2]);
    if (idx < 0) return nullptr;                                    // :521-522
    const std::vector<PointI>& cube = isCorner ? cornerCube[idx] : surfCube[idx];
    if (cube.size() < 5) return nullptr;                            // :523
    std::map<int, void*>& trees = isCorner ? kdCorner : kdSurf;
    auto it = trees.find(idx);
    if (it == trees.end()) { void* h = nullptr; knn.build(cube.data(), cube.size(), &h); it = trees.insert(std::make_pair(idx, h)).first; }
    knn.query(it->second, sel, 5, ind, sq);                         // :524
    return cube.data();
  };
  scan_match_impl(prm, lookup, corner.data(), corner.size(), surf.data(), surf.size(), pose, res, keepLog);
  for (auto& kv : kdCorner) knn.free(kv.second);
  for (auto& kv : kdSurf) knn.free(kv.second);
}
void FeatureMap::fileOrder(std::vector<int>& type, std::vector<int>& ci, std::vector<int>& cj, std::vector<int>& ck) const {
  type.clear(); ci.clear(); cj.clear(); ck.clear();
  for (int i = 0; i < _p.cubeW; i++)
    for (int j = 0; j < _p.cubeH; j++)
      for (int k = 0; k < _p.cubeD; k++) {
        if (!cornerCube[toIndex(i, j, k)].empty()) { type.push_back(0); ci.push_back(i); cj.push_back(j); ck.push_back(k); }
        if (!surfCube[toIndex(i, j, k)].empty()) { type.push_back(1); ci.push_back(i); cj.push_back(j); ck.push_back(k); }
      }
}
void FeatureMap::loadCube(int type, int i, int j, int k, const std::vector<PointI>& cloud) {
  if (!isIndexValid(i, j, k)) return;
  std::vector<PointI> ds;
  voxel_filter(cloud.data(), cloud.size(), type == 0 ? _p.mapFilterCorner : _p.mapFilterSurf, ds);
  (type == 0 ? cornerCube : surfCube)[toIndex(i, j, k)].swap(ds);
}
size_t FeatureMap::totalPoints() const {
  size_t n = 0;
  for (auto& c : cornerCube) n += c.size();
  for (auto& c : surfCube) n += c.size();
  return n;
}

LaserMapping::LaserMapping(const MapParams& mp, const MatchParams& sp, const KnnBackend& knn)
    : map(mp), _mp(mp), _sp(sp), _knn(knn) {
  mappedLast = mappedNew = odomLast = iso_identity();   // LaserMatcher.cpp:31-32
}

// LaserMapping.cpp:39-59 (hasNewData / publishResult are ROS plumbing and are not restated)
Iso LaserMapping::process(const Iso& odomNew, const std::vector<PointI>& corner, const std::vector<PointI>& surf) {
  // transformMerge, LaserMatcher.cpp:333-340 -> transformAssociate transform_utils.h:502-507
  Iso L2W = iso_mul(mappedLast, iso_inverse(odomLast));
  mappedNew = iso_mul(L2W, odomNew);
  // prepareFeatureFrame, LaserMatcher.cpp:288-301
  voxel_filter(corner.data(), corner.size(), _mp.filterCorner, cornerDS);
  voxel_filter(surf.data(), surf.size(), _mp.filterSurf, surfDS);
  // prepareFeatureSurround, LaserMatcher.cpp:303-325
  map.update(mappedNew.t);
  map.getSurroundFeature(surroundCorner, surroundSurf);
  // optimizeTransform, LaserMatcher.cpp:327-331 -> ScanMatch.cpp:349-360
  float pose[6];
  iso_to_twist(mappedNew, pose);
  scan_match(_sp, _knn, surroundCorner.data(), surroundCorner.size(), surroundSurf.data(), surroundSurf.size(),
             cornerDS.data(), cornerDS.size(), surfDS.data(), surfDS.size(), pose, lastMatch, keepLog);
  twist_to_iso(pose, mappedNew);
  // transformUpdate, LaserMatcher.cpp:342-347
  mappedLast = mappedNew;
  odomLast = odomNew;
  // featureMapUpdate, LaserMatcher.cpp:349-355
  map.addFeatureCloud(cornerDS, surfDS, mappedNew);
  return mappedNew;
}

// LaserLocalization::process, LaserLocalization.cpp:163-188 (IMU blending of transformUpdate not restated)
Iso LaserMapping::localize(const Iso& odomNew, const std::vector<PointI>& corner, const std::vector<PointI>& surf) {
  Iso L2W = iso_mul(mappedLast, iso_inverse(odomLast));
  mappedNew = iso_mul(L2W, odomNew);
  voxel_filter(corner.data(), corner.size(), _mp.filterCorner, cornerDS);
  voxel_filter(surf.data(), surf.size(), _mp.filterSurf, surfDS);
  map.update(mappedNew.t);
  float pose[6];
  iso_to_twist(mappedNew, pose);
  map.scanMatchScan(_knn, cornerDS, surfDS, pose, lastMatch, keepLog);   // optimizeTransform, :124-138 -> FeatureMap.h:692-700
  twist_to_iso(pose, mappedNew);
  mappedLast = mappedNew;
  odomLast = odomNew;
  return mappedNew;
}

// ---- LaserMappingLocal (LaserMappingLocal.cpp:33-78) over LocalFeatureMap (io_module/LocalFeatureMap.h) -------------------
LaserMappingLocal::LaserMappingLocal(const MapParams& mp, const MatchParams& sp, const KnnBackend& knn, bool useMappedPose)
    : _mp(mp), _sp(sp), _knn(knn), _useMapped(useMappedPose) {
  mappedLast = mappedNew = odomLast = iso_identity();   // LaserMatcher.cpp:31-32
}

Iso LaserMappingLocal::process(const Iso& odomNew, const std::vector<PointI>& corner, const std::vector<PointI>& surf) {
  // transformMerge, LaserMatcher.cpp:333-340
  Iso L2W = iso_mul(mappedLast, iso_inverse(odomLast));
  mappedNew = iso_mul(L2W, odomNew);
  // prepareFeatureFrame, LaserMatcher.cpp:288-301
  voxel_filter(corner.data(), corner.size(), _mp.filterCorner, cornerDS);
  voxel_filter(surf.data(), surf.size(), _mp.filterSurf, surfDS);
  // prepareFeatureSurround, LaserMappingLocal.cpp:55-60 -> LocalFeatureMap::getSurroundFeature, LocalFeatureMap.h:84-99:
  // concatenation of the window in queue order, then VoxelGrid 0.2 (corner) / 0.4 (surf), :29-31
  std::vector<PointI> catC, catS;
  for (const LocalFrame& f : window) {
    catC.insert(catC.end(), f.corner.begin(), f.corner.end());
    catS.insert(catS.end(), f.surf.begin(), f.surf.end());
  }
  voxel_filter(catC.data(), catC.size(), 0.2f, surroundCorner);
  voxel_filter(catS.data(), catS.size(), 0.4f, surroundSurf);
  // optimizeTransform, LaserMatcher.cpp:327-331
  float pose[6];
  iso_to_twist(mappedNew, pose);
  scan_match(_sp, _knn, surroundCorner.data(), surroundCorner.size(), surroundSurf.data(), surroundSurf.size(),
             cornerDS.data(), cornerDS.size(), surfDS.data(), surfDS.size(), pose, lastMatch, false);
  twist_to_iso(pose, mappedNew);
  // transformUpdate, LaserMatcher.cpp:342-347
  mappedLast = mappedNew;
  odomLast = odomNew;
  // featureMapUpdate, LaserMappingLocal.cpp:62-76: the frame clouds are transformed IN PLACE by _transformTobeMapped
  Iso tf = iso_identity();
  if (_useMapped) tf = mappedNew;
  else { const float zero[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; twist_to_iso(zero, tf); }   // convertTransform of the untouched Twist
  LocalFrame fr;
  auto xform = [&](const std::vector<PointI>& in, std::vector<PointI>& out) {   // transform_utils.h:601-614
    out = in;
    for (size_t i = 0; i < in.size(); i++) {
      const PointI& p = in[i];
      out[i].x = ((tf.R[0] * p.x + tf.R[1] * p.y) + tf.R[2] * p.z) + tf.t[0];
      out[i].y = ((tf.R[3] * p.x + tf.R[4] * p.y) + tf.R[5] * p.z) + tf.t[1];
      out[i].z = ((tf.R[6] * p.x + tf.R[7] * p.y) + tf.R[8] * p.z) + tf.t[2];
    }
  };
  xform(cornerDS, fr.corner);
  xform(surfDS, fr.surf);
  // LocalFeatureMap::addDataFrame, :62-69 -> FrameUpdater::update, FrameUpdater.hpp:17-42 (Isometry3d of the float pose)
  double R[9], t[3];
  for (int k = 0; k < 9; k++) R[k] = (double)tf.R[k];
  for (int k = 0; k < 3; k++) t[k] = (double)tf.t[k];
  if (_first) {
    _first = false;
  } else {
    // delta = prev_keypose.inverse() * pose: translation = Rp^T * t + (-(Rp^T * tp))
    double a[3], b[3], v[3];
    for (int r = 0; r < 3; r++) {
      a[r] = (_prevR[0 + r] * t[0] + _prevR[3 + r] * t[1]) + _prevR[6 + r] * t[2];
      b[r] = -((_prevR[0 + r] * _prevT[0] + _prevR[3 + r] * _prevT[1]) + _prevR[6 + r] * _prevT[2]);
      v[r] = a[r] + b[r];
    }
    accumDistance += std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  }
  std::memcpy(_prevR, R, sizeof(R)); std::memcpy(_prevT, t, sizeof(t));
  fr.accum = accumDistance;
  window.push_back(std::move(fr));
  // LocalFeatureMap::clean, :70-82 (erases one frame more than it counted)
  int deleteNum = 0;
  for (const LocalFrame& f : window) {
    if (f.accum > (accumDistance - 30.0)) break;   // queue_distance_threshold(30.0), :28
    ++deleteNum;
  }
  if (deleteNum > 0) window.erase(window.begin(), window.begin() + deleteNum + 1);
  return mappedNew;
}


}  // namespace cmo
