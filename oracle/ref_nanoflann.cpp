// ref_nanoflann.cpp -- thin C wrapper that compiles the REFERENCE's own vendored KD-tree
// (L_SLAM/src/util/nanoflann.hpp, v1.2.3) from where it lies under /root/reference (include path given by
// oracle/Makefile; the header is never copied into this repo).  TEST INFRASTRUCTURE.  Output: oracle/_ref/.
// The adaptor mirrors nanoflann_pcl.h:92-108,187-210: float xyz accessors, DIM = 3, int index,
// SO3_Adaptor (== L2_Simple, nanoflann.hpp:420-439), default leaf_max_size = 10, SearchParams() defaults.
#include "nanoflann.hpp"
#include <cstddef>
#include <vector>
#include <cfloat>

namespace {
struct CloudAdaptor {
  const float* xyzi; size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, int dim) const {
    if (dim == 0) return xyzi[4 * idx + 0];
    else if (dim == 1) return xyzi[4 * idx + 1];
    else if (dim == 2) return xyzi[4 * idx + 2];
    else return 0.0f;
  }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};
typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::SO3_Adaptor<float, CloudAdaptor>, CloudAdaptor, 3, int> Tree;
struct Handle {
  CloudAdaptor ad; Tree tree;
  Handle(const float* p, size_t n) : ad{p, n}, tree(3, ad) { tree.buildIndex(); }
};
}  // namespace

extern "C" {
// KdTreeFLANN::setInputCloud nanoflann_pcl.h:132-148 (the caller keeps `xyzi` alive)
void* cmref_kd_build(const float* xyzi, size_t n) { return new Handle(xyzi, n); }
void cmref_kd_free(void* h) { delete (Handle*)h; }
// KdTreeFLANN::nearestKSearch nanoflann_pcl.h:150-162.  idx/d2 are zero-filled first like
// ScanMatch.cpp:65-66 (std::vector<int>(5,0)); unfilled slots keep d2[k-1] = FLT_MAX (nanoflann.hpp:91-97).
int cmref_kd_query(void* h, const float* q, int k, int* idx, float* d2) {
  Handle* H = (Handle*)h;
  for (int i = 0; i < k; i++) { idx[i] = 0; d2[i] = 0.f; }
  if (H->ad.n == 0) { d2[k - 1] = FLT_MAX; return 0; }
  nanoflann::KNNResultSet<float, int> rs(k);
  rs.init(idx, d2);
  float qq[4] = {q[0], q[1], q[2], 0.f};
  H->tree.findNeighbors(rs, qq, nanoflann::SearchParams());
  return (int)rs.size();
}
void cmref_kd_query_batch(void* h, const float* q, size_t nq, int k, int* idx, float* d2) {
  for (size_t i = 0; i < nq; i++) cmref_kd_query(h, q + 3 * i, k, idx + (size_t)k * i, d2 + (size_t)k * i);
}
}
