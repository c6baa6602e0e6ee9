// oracle_voxel.cpp -- CPU restatement of pcl::VoxelGrid<pcl::PointXYZI>::filter as the reference uses it
// (TEST INFRASTRUCTURE, see cm_oracle.h).  PCL is not under /root/reference; its voxel indexing is stated
// in-tree by L_SLAM/src/util/voxel_grid_partition.hpp:91-272 (a derived copy of VoxelGrid::applyFilter) and
// is followed here line by line: bounding box (:99-107), overflow guard (:108-123), min_b/div_b/divb_mul
// (:124-137), per-point ijk and idx (:212-226), sort by idx (:232-233), one output per occupied cell in idx
// order (:243-258).  The centroid (VoxelGrid.hpp "fourth pass") is the float sum of ALL fields (x, y, z,
// intensity; downsample_all_data default true) in sorted order divided by float(count).
// Call sites: ScanRegistration.cpp:390-399, LaserMatcher.cpp:293-300, FeatureMap.h:289-306, ScanMatch.cpp:381-394.
// Canonical choice: PCL sorts with std::sort (unstable); the oracle uses a STABLE sort, i.e. points of one
// voxel are summed in input order.  Non-finite points are skipped (is_dense == false branch, :204-209).
#include "cm_oracle.h"
#include <algorithm>
#include <cmath>
#include <limits>

namespace cmo {

void voxel_filter(const PointI* in, size_t n, float leaf, std::vector<PointI>& out) {
  out.clear();
  if (n == 0) return;
  const float inv = 1.0f / leaf;   // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
  float minx = std::numeric_limits<float>::max(), miny = minx, minz = minx;
  float maxx = -std::numeric_limits<float>::max(), maxy = maxx, maxz = maxx;
  size_t finite = 0;
  for (size_t i = 0; i < n; i++) {
    const PointI& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    finite++;
    minx = std::min(minx, p.x); miny = std::min(miny, p.y); minz = std::min(minz, p.z);
    maxx = std::max(maxx, p.x); maxy = std::max(maxy, p.y); maxz = std::max(maxz, p.z);
  }
  if (finite == 0) return;
  int64_t dx = static_cast<int64_t>((maxx - minx) * inv) + 1;
  int64_t dy = static_cast<int64_t>((maxy - miny) * inv) + 1;
  int64_t dz = static_cast<int64_t>((maxz - minz) * inv) + 1;
  if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
    // "Leaf size is too small for the input dataset": PCL copies the input through unchanged.
    out.assign(in, in + n);
    return;
  }
  int minb[3] = {static_cast<int>(std::floor(minx * inv)), static_cast<int>(std::floor(miny * inv)),
                 static_cast<int>(std::floor(minz * inv))};
  int maxb[3] = {static_cast<int>(std::floor(maxx * inv)), static_cast<int>(std::floor(maxy * inv)),
                 static_cast<int>(std::floor(maxz * inv))};
  int divb[3] = {maxb[0] - minb[0] + 1, maxb[1] - minb[1] + 1, maxb[2] - minb[2] + 1};
  int mul[3] = {1, divb[0], divb[0] * divb[1]};
  struct Ent { unsigned int idx; unsigned int pt; };
  std::vector<Ent> iv;
  iv.reserve(n);
  for (size_t i = 0; i < n; i++) {
    const PointI& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    int ijk0 = static_cast<int>(std::floor(p.x * inv) - static_cast<float>(minb[0]));
    int ijk1 = static_cast<int>(std::floor(p.y * inv) - static_cast<float>(minb[1]));
    int ijk2 = static_cast<int>(std::floor(p.z * inv) - static_cast<float>(minb[2]));
    int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    iv.push_back(Ent{static_cast<unsigned int>(idx), static_cast<unsigned int>(i)});
  }
  std::stable_sort(iv.begin(), iv.end(), [](const Ent& a, const Ent& b) { return a.idx < b.idx; });
  size_t index = 0;
  while (index < iv.size()) {
    size_t i = index + 1;
    while (i < iv.size() && iv[i].idx == iv[index].idx) ++i;
    float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;
    for (size_t t = index; t < i; t++) {
      const PointI& p = in[iv[t].pt];
      cx += p.x; cy += p.y; cz += p.z; ci += p.intensity;
    }
    float cnt = static_cast<float>(i - index);
    out.push_back(PointI{cx / cnt, cy / cnt, cz / cnt, ci / cnt});
    index = i;
  }
}

}  // namespace cmo
