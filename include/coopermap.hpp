// coopermap.hpp -- header-only C++ facade over the C ABI (coopermap.h): the reference's stage classes with the same names and
// the same setup / process split, so that a nodelet written against lidar_slam::{OrganisedScanRegistration,
// MultiScanRegistration, LaserOdometry, LaserMapping, LaserMappingLocal, LaserLocalization, ScanMatch} keeps its shape; LoamPipeline
// is the three stages of the launch file in one object
// (L_SLAM/src/odometry/*.h, scan_to_scan_match/ScanMatch.h, nodelet/*.cpp).  Clouds are std::vector<cm_point> (the payload
// of pcl::PointCloud<pcl::PointXYZI>), poses are cm_iso (Eigen::Isometry3f, row-major rotation + translation) or cm_pose
// (lidar_slam::Twist).  No PCL / Eigen / ROS types: the conversion helpers a nodelet needs are in INTEGRATION.md section 2.
// Every object owns one context = one CUDA stream; like the reference's stages, an object is driven by one thread.
#pragma once
#include "coopermap.h"
#include <stdexcept>
#include <string>
#include <vector>

namespace coopermap {

typedef std::vector<cm_point> PointCloud;

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// RAII owner of a cm_ctx.  There is no CPU fallback: construction throws without a CUDA device.
class Context {
 public:
  explicit Context(const cm_config& cfg) : cfg_(cfg), ctx_(nullptr) {
    const int rc = cm_ctx_create(&cfg_, &ctx_);
    if (rc != CM_OK) throw Error(rc, "cm_ctx_create failed (no CUDA device?)");
  }
  ~Context() { if (ctx_) cm_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  cm_ctx* get() const { return ctx_; }
  const cm_config& config() const { return cfg_; }
  // hard errors (< 0) throw; soft outcomes (>= 0: CM_OK, CM_TOO_FEW_REF, ...) are returned like the reference's bool / warnings
  int check(int rc) const { if (rc < 0) throw Error(rc, cm_last_error(ctx_)); return rc; }
  static cm_config defaults() { cm_config c; cm_config_default(&c); return c; }

 private:
  cm_config cfg_;
  cm_ctx* ctx_;
};

// ---- stage 1 ---------------------------------------------------------------------------------------------------------------
// ScanRegistration (ScanRegistration.h): the four feature clouds of a sweep.
class ScanRegistration {
 public:
  explicit ScanRegistration(const cm_config& cfg = Context::defaults()) : ctx_(cfg) {}
  const PointCloud& cornerPointsSharp() const { return clouds_[0]; }       // /laser_cloud_sharp
  const PointCloud& cornerPointsLessSharp() const { return clouds_[1]; }   // /laser_cloud_less_sharp
  const PointCloud& surfacePointsFlat() const { return clouds_[2]; }       // /laser_cloud_flat
  const PointCloud& surfacePointsLessFlat() const { return clouds_[3]; }   // /laser_cloud_less_flat
  const PointCloud& laserCloud() const { return full_; }                   // /velodyne_cloud_2 (ring-major)
  Context& context() { return ctx_; }

 protected:
  void prepare(size_t n, cm_scanreg_out& out) {
    for (int k = 0; k < 4; k++) { clouds_[k].resize(n ? n : 1); out.pts[k] = clouds_[k].data(); out.cap[k] = (int)clouds_[k].size(); }
    full_.resize(n ? n : 1); curv_.resize(n ? n : 1);
    out.n = n_; out.cloud = full_.data(); out.cloud_curvature = curv_.data();
  }
  void finish(int n_full) {
    for (int k = 0; k < 4; k++) clouds_[k].resize(n_[k]);
    full_.resize(n_full);
  }
  Context ctx_;
  PointCloud clouds_[4], full_;
  std::vector<float> curv_;
  int n_[5] = {0, 0, 0, 0, 0};
};

// OrganisedScanRegistration::process (OrganizedScanRegistration.cpp:82-150): ring = row, missing returns = NaN.
class OrganisedScanRegistration : public ScanRegistration {
 public:
  using ScanRegistration::ScanRegistration;
  void process(const PointCloud& organised, int height, int width) {
    if ((size_t)height * width != organised.size()) throw Error(CM_ERR_ARG, "cloud is not height x width");
    cm_scanreg_out out = cm_scanreg_out();
    prepare(organised.size(), out);
    ctx_.check(cm_scanreg_organised_host(ctx_.get(), organised.data(), 1, height, width, &out));
    // size of the ring-major cloud = returns that are finite and outside the blind radius (OrganizedScanRegistration.cpp:115-123)
    const float b2 = ctx_.config().blind_radius * ctx_.config().blind_radius;
    int kept = 0;
    for (const cm_point& p : organised)
      if (p.x - p.x == 0.f && p.y - p.y == 0.f && p.z - p.z == 0.f && !(p.x * p.x + p.y * p.y + p.z * p.z < b2)) kept++;
    finish(kept);
  }
};

// MultiScanRegistration::process (MultiScanRegistration.cpp:95-200): raw azimuth-major sweep of a spinning multi-beam LiDAR.
class MultiScanRegistration : public ScanRegistration {
 public:
  enum Lidar { VLP16 = 0, HDL32 = 1, HDL64E = 2, Pandar40 = 3 };   // MultiScanRegistration.h:24-42, 90-102
  using ScanRegistration::ScanRegistration;
  void process(const PointCloud& sweep, Lidar lidar) {
    cm_scanreg_out out = cm_scanreg_out();
    prepare(sweep.size(), out);
    int rows = 0, cols = 0;
    ctx_.check(cm_scanreg_sweep_host(ctx_.get(), sweep.data(), sweep.size(), (int)lidar, &out, &rows, &cols));
    finish((int)sweep.size());
  }
  // ScanRegistration::handleIMUMessage (ScanRegistration.cpp:89-121): stamp in seconds, roll / pitch / yaw from tf's getRPY, raw acceleration
  void handleIMUMessage(double stamp, double roll, double pitch, double yaw, double ax, double ay, double az) {
    const cm_imu_sample m = {stamp, roll, pitch, yaw, ax, ay, az};
    ctx_.check(cm_imu_push_host(ctx_.get(), &m));
  }
  // process(in, scanTime) with hasIMUData(): every point projected to the sweep start (ScanRegistration.cpp:123-188)
  void process(const PointCloud& sweep, Lidar lidar, double scanTime) {
    cm_scanreg_out out = cm_scanreg_out();
    prepare(sweep.size(), out);
    int rows = 0, cols = 0;
    ctx_.check(cm_scanreg_sweep_imu_host(ctx_.get(), sweep.data(), sweep.size(), (int)lidar, scanTime, &out, &rows, &cols, imuTrans_));
    finish((int)sweep.size());
  }
  const float* imuTrans() const { return imuTrans_; }   // the four /imu_trans points (x, y, z each), ScanRegistration.cpp:681-708

 private:
  float imuTrans_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

// ---- stage 2 ---------------------------------------------------------------------------------------------------------------
// LaserOdometry::process (LaserOdometry.cpp:288-326)
class LaserOdometry {
 public:
  explicit LaserOdometry(const cm_config& cfg = Context::defaults()) : ctx_(cfg) { ctx_.check(cm_odometry_reset(ctx_.get())); }
  void process(const PointCloud& sharp, const PointCloud& lessSharp, const PointCloud& flat, const PointCloud& lessFlat) {
    corner_.resize(lessSharp.size() ? lessSharp.size() : 1); surf_.resize(lessFlat.size() ? lessFlat.size() : 1);
    ctx_.check(cm_odometry_process_host(ctx_.get(), sharp.data(), (int)sharp.size(), lessSharp.data(), (int)lessSharp.size(), flat.data(),
                                        (int)flat.size(), lessFlat.data(), (int)lessFlat.size(), &sum_, &tf_, corner_.data(), surf_.data(),
                                        &stats_, nullptr));
    corner_.resize(lessSharp.size()); surf_.resize(lessFlat.size());
  }
  const cm_iso& transformSum() const { return sum_; }              // /laser_odom_to_init
  const cm_pose& transform() const { return tf_; }                 // frame-to-frame Twist (_transform)
  const PointCloud& lastCornerCloud() const { return corner_; }    // /laser_cloud_corner_last
  const PointCloud& lastSurfaceCloud() const { return surf_; }     // /laser_cloud_surf_last
  const cm_odom_stats& stats() const { return stats_; }

 private:
  Context ctx_;
  cm_iso sum_ = cm_iso();
  cm_pose tf_ = cm_pose();
  cm_odom_stats stats_ = cm_odom_stats();
  PointCloud corner_, surf_;
};

// ---- stage 3 ---------------------------------------------------------------------------------------------------------------
// LaserMapping::process (LaserMapping.cpp:39-59) with the cube map resident on the GPU.
class LaserMapping {
 public:
  explicit LaserMapping(const cm_config& cfg = Context::defaults(), size_t maxCornerPoints = 400000, size_t maxSurfPoints = 4000000)
      : ctx_(cfg) { ctx_.check(cm_mapping_create(ctx_.get(), 1, maxCornerPoints, maxSurfPoints)); }
  // odom: /laser_odom_to_init; returns /aft_mapped_to_init.  lastStatus(): CM_OK, CM_TOO_FEW_REF, CM_TOO_FEW_MATCHES, CM_NOT_CONVERGED
  cm_iso process(const cm_iso& odom, const PointCloud& cornerLast, const PointCloud& surfLast) {
    cm_iso mapped; int nc = (int)cornerLast.size(), ns = (int)surfLast.size();
    const cm_point dummy = cm_point();
    ctx_.check(cm_mapping_process_host(ctx_.get(), &odom, nc ? cornerLast.data() : &dummy, &nc, nc ? nc : 1, ns ? surfLast.data() : &dummy, &ns,
                                       ns ? ns : 1, &mapped, &stats_));
    return mapped;
  }
  int lastStatus() const { return stats_.status; }
  const cm_match_stats& stats() const { return stats_; }
  // process() returns with the pose; the map insertion of that frame finishes behind it.  sync() waits for it and throws what it hit
  // (map capacity) -- otherwise the next process() reports it
  void sync() { ctx_.check(cm_mapping_sync(ctx_.get())); }
  // saveMap service (LaserMatcher.cpp:357-394 -> FeatureMap::saveCloudToFiles)
  int saveMap(const std::string& dir) { int n = 0; ctx_.check(cm_map_save_host(ctx_.get(), 0, dir.c_str(), &n)); return n; }
  // /laser_cloud_surround_corner, _surf: FeatureMap::getSurroundFeature (FeatureMap.h:256-265), published at LaserMatcher.cpp:357-394
  void surroundClouds(PointCloud& corner, PointCloud& surf) {
    size_t n2[2] = {0, 0};
    ctx_.check(cm_map_surround_host(ctx_.get(), 0, nullptr, 0, nullptr, 0, n2));
    corner.resize(n2[0] ? n2[0] : 1); surf.resize(n2[1] ? n2[1] : 1);
    ctx_.check(cm_map_surround_host(ctx_.get(), 0, corner.data(), corner.size(), surf.data(), surf.size(), n2));
    corner.resize(n2[0]); surf.resize(n2[1]);
  }
  // /FullMap and the ~pubFullMap service: FeatureMap::getFullMap (FeatureMap.h:267-287) at the full-map leaf (map_filter_full)
  PointCloud fullMap(float leaf) {
    size_t n = 0;
    ctx_.check(cm_map_full_host(ctx_.get(), 0, leaf, nullptr, 0, &n));
    PointCloud out(n ? n : 1);
    ctx_.check(cm_map_full_host(ctx_.get(), 0, leaf, out.data(), out.size(), &n));
    out.resize(n);
    return out;
  }
  // every resident point of a class (0 corner, 1 surf), storage order
  PointCloud mapCloud(int cls) {
    size_t n = 0;
    ctx_.check(cm_map_export_host(ctx_.get(), 0, cls, nullptr, nullptr, 0, &n));
    PointCloud out(n ? n : 1);
    ctx_.check(cm_map_export_host(ctx_.get(), 0, cls, out.data(), nullptr, out.size(), &n));
    out.resize(n);
    return out;
  }
  Context& context() { return ctx_; }

 protected:
  Context ctx_;
  cm_match_stats stats_ = cm_match_stats();
};

// The three nodelets of the reference's launch file as ONE object: a sweep in, /laser_odom_to_init and /aft_mapped_to_init out
// (cm_pipeline_chain_step_host).  The feature clouds and the projected clouds the nodelets pass over topics stay in device memory.
// The map insertion of a sweep finishes behind process(); sync() waits for it (the map read-outs below are ordered behind it anyway).
class LoamPipeline : public LaserMapping {
 public:
  LoamPipeline(int rows, int cols, const cm_config& cfg = Context::defaults(), size_t maxCornerPoints = 400000, size_t maxSurfPoints = 4000000)
      : LaserMapping(cfg, maxCornerPoints, maxSurfPoints), rows_(rows), cols_(cols) {
    ctx_.check(cm_pipeline_chain_create(ctx_.get(), rows, cols));
  }
  // organised sweep (rows x cols, ring-major, NaN = no return) -> /aft_mapped_to_init; odometry(): /laser_odom_to_init of the same sweep
  cm_iso process(const PointCloud& organised) {
    if (organised.size() != (size_t)rows_ * cols_) throw Error(CM_ERR_ARG, "sweep size does not match rows x cols");
    cm_iso mapped;
    ctx_.check(cm_pipeline_chain_step_host(ctx_.get(), organised.data(), rows_, cols_, &odom_, &mapped, &ostats_, &stats_));
    return mapped;
  }
  // raw-sweep mode: the unorganised cloud of a spinning LiDAR (MultiScanRegistration in front of the chain); scanTime >= 0 de-skews
  // with the IMU messages given to handleIMUMessage
  LoamPipeline(MultiScanRegistration::Lidar lidar, size_t maxPoints, const cm_config& cfg = Context::defaults(), size_t maxCornerPoints = 400000, size_t maxSurfPoints = 4000000)
      : LaserMapping(cfg, maxCornerPoints, maxSurfPoints), rows_(-1), cols_(0), lidar_((int)lidar) {
    ctx_.check(cm_pipeline_chain_sweep_create(ctx_.get(), maxPoints));
  }
  cm_iso processSweep(const PointCloud& sweep, double scanTime = -1.0) {
    cm_iso mapped;
    ctx_.check(cm_pipeline_chain_step_sweep_host(ctx_.get(), sweep.data(), sweep.size(), lidar_, scanTime, &odom_, &mapped, &ostats_, &stats_));
    return mapped;
  }
  void handleIMUMessage(double stamp, double roll, double pitch, double yaw, double ax, double ay, double az) {
    const cm_imu_sample m = {stamp, roll, pitch, yaw, ax, ay, az};
    ctx_.check(cm_imu_push_host(ctx_.get(), &m));
  }
  const cm_iso& odometry() const { return odom_; }
  const cm_odom_stats& odometryStats() const { return ostats_; }

 private:
  int rows_, cols_, lidar_ = 0;
  cm_iso odom_ = cm_iso();
  cm_odom_stats ostats_ = cm_odom_stats();
};

// LaserLocalization::process (LaserLocalization.cpp:163-188): FeatureMap::scanMatchScan against a prebuilt map, no map update.
class LaserLocalization : public LaserMapping {
 public:
  using LaserMapping::LaserMapping;
  // FeatureMap::loadCloudFromFiles; returns the number of files read
  int loadMap(const std::string& dir) {
    int n = 0; size_t pts = 0, bad = 0;
    ctx_.check(cm_map_load_host(ctx_.get(), 0, dir.c_str(), &n, &pts, &bad));
    return n;
  }
  // dynamic mode (DynamicFeatureMap, DynamicFeatureMap.h:129-161, 504-677): <dir>/index2.txt, a window of cubes around the sensor resident
  int openPagedMap(const std::string& dir, int windowW = 21, int windowH = 11, int windowD = 21) {
    int n = 0;
    ctx_.check(cm_map_page_open_host(ctx_.get(), 0, dir.c_str(), windowW, windowH, windowD, &n));
    return n;
  }
  void updatePagedMap(float x, float y, float z) {
    const float s[3] = {x, y, z};
    ctx_.check(cm_map_page_update_host(ctx_.get(), 0, s, nullptr, nullptr, nullptr));
  }
  cm_iso process(const cm_iso& odom, const PointCloud& cornerLast, const PointCloud& surfLast) {
    cm_iso mapped; int nc = (int)cornerLast.size(), ns = (int)surfLast.size();
    const cm_point dummy = cm_point();
    ctx_.check(cm_localization_process_host(ctx_.get(), &odom, nc ? cornerLast.data() : &dummy, &nc, nc ? nc : 1,
                                            ns ? surfLast.data() : &dummy, &ns, ns ? ns : 1, &mapped, &stats_));
    return mapped;
  }
};

// LaserMappingLocal::process (LaserMappingLocal.cpp:33-78): the mapping stage over LocalFeatureMap, a sliding window of
// voxel-filtered frames.  useMappedPose = false is the reference as written (frames placed with the never-assigned
// _transformTobeMapped = identity), true places them with the mapped pose.
class LaserMappingLocal {
 public:
  explicit LaserMappingLocal(const cm_config& cfg = Context::defaults(), bool useMappedPose = false) : ctx_(cfg) {
    ctx_.check(cm_mapping_local_create(ctx_.get(), useMappedPose ? 1 : 0));
  }
  cm_iso process(const cm_iso& odom, const PointCloud& cornerLast, const PointCloud& surfLast) {
    cm_iso mapped;
    ctx_.check(cm_mapping_local_process_host(ctx_.get(), &odom, cornerLast.data(), cornerLast.size(), surfLast.data(), surfLast.size(),
                                             &mapped, &stats_));
    return mapped;
  }
  // frames in LocalFeatureMap's queue
  int windowFrames() { int n = 0; ctx_.check(cm_mapping_local_window_host(ctx_.get(), &n, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0)); return n; }
  int lastStatus() const { return stats_.status; }
  const cm_match_stats& stats() const { return stats_; }
  Context& context() { return ctx_; }

 private:
  Context ctx_;
  cm_match_stats stats_ = cm_match_stats();
};

// ---- the operator ------------------------------------------------------------------------------------------------------------
// ScanMatch (ScanMatch.h:22-86).  The class defaults are the reference's (ScanMatch.cpp:21-33): 0.05 / 0.05, useScore = true.
class ScanMatch {
 public:
  explicit ScanMatch(int maxIterations = 10) : ctx_(make_cfg(maxIterations)) {}
  explicit ScanMatch(const cm_config& cfg) : ctx_(cfg) {}
  // scanMatchScan(refCorner, refSurf, corner, surf, Twist&): transform is the initial guess and receives the result, also when
  // false is returned (ScanMatch.cpp:342-346)
  bool scanMatchScan(const PointCloud& refCorner, const PointCloud& refSurf, const PointCloud& corner, const PointCloud& surf, cm_pose& transform) {
    ctx_.check(cm_match_stateless_host(ctx_.get(), refCorner.data(), refCorner.size(), refSurf.data(), refSurf.size(), corner.data(),
                                       corner.size(), surf.data(), surf.size(), &transform, &stats_, nullptr, nullptr, nullptr));
    return stats_.ret != 0;
  }
  bool scanMatchScan(const PointCloud& refCorner, const PointCloud& refSurf, const PointCloud& corner, const PointCloud& surf, cm_iso& pose) {
    ctx_.check(cm_match_stateless_iso_host(ctx_.get(), refCorner.data(), refCorner.size(), refSurf.data(), refSurf.size(), corner.data(),
                                           corner.size(), surf.data(), surf.size(), &pose, &stats_));
    return stats_.ret != 0;
  }
  // scanMatchLocal (ScanMatch.cpp:375-398): voxel-filters the four clouds (0.2 / 0.4) first
  bool scanMatchLocal(const PointCloud& refCorner, const PointCloud& refSurf, const PointCloud& corner, const PointCloud& surf, cm_pose& transform) {
    ctx_.check(cm_match_local_host(ctx_.get(), refCorner.data(), refCorner.size(), refSurf.data(), refSurf.size(), corner.data(), corner.size(),
                                   surf.data(), surf.size(), &transform, &stats_));
    return stats_.ret != 0;
  }
  const cm_match_stats& stats() const { return stats_; }

 private:
  static cm_config make_cfg(int maxIterations) {
    cm_config c = Context::defaults();
    c.max_iterations = maxIterations; c.delta_t_abort = 0.05f; c.delta_r_abort = 0.05f; c.use_score = 1;
    return c;
  }
  Context ctx_;
  cm_match_stats stats_ = cm_match_stats();
};

}  // namespace coopermap
