// cm_oracle.h -- CPU oracle for the LOAM hot path of ZhekaiJin/the-Cooper-Mapper (L_SLAM).
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (the-cooper-mapper_b200/) may include, link or call
// this directory; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// PARITY STATUS: "parity unpinned".  The reference ships no tests, golden vectors or fixtures for this path
//  and cannot be built here (needs ROS, PCL, Eigen).  This restatement follows the
// cited reference lines operation by operation in float32; Eigen / PCL arithmetic that lives outside
// /root/reference is restated from the published algorithms (the-cooper-mapper_b200/csrc/cm_math.h for
// Eigen; voxel_grid_partition.hpp:91-272 is the in-tree statement of PCL VoxelGrid indexing).  The one piece
// of the reference that does compile here, the vendored nanoflann KD-tree, is built from where it lies
// (oracle/Makefile -> oracle/_ref/) and is the ground truth for neighbour sets.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <map>
#include <vector>

namespace cmo {

struct PointI { float x, y, z, intensity; };               // pcl::PointXYZI payload
struct PointIN { float x, y, z, intensity, curvature; };   // the fields of pcl::PointXYZINormal the path uses

// ScanRegistration.h:23-40
enum PointLabel {
  SLOP = 8, BLOCKED = 7, UNKNOW = 6, CONER_PICKED_NEAR = 4, SURF_PICKED_NEAR = 3, CORNER_LESS_SHARP = 2,
  CORNER_SHARP = 1, SURFACE_LESS_FLAT = 0, SURFACE_FLAT = -1, ONESIDE_FLAT = 5, EDGE_BROKEN = -2,
  NEAR_BLOCK = -3, BLIND_BLOCK = -4, MESSY = 9
};

// RegistrationParams, ScanRegistration.cpp:32-49 (only the members extractFeatures reads)
struct ScanRegParams {
  float scanPeriod = 0.1f;
  int nFeatureRegions = 6;
  int curvatureRegion = 5;
  int maxCornerSharp = 2;
  int maxSurfaceFlat = 4;
  float lessFlatFilterSize = 0.2f;
  float surfaceCurvatureThreshold = 0.02f;
  float blindDegreeThreshold = 0.5f;
  float blindRadius = 2.5f;  // OrganisedScanRegistration::_blindRaduis, OrganizedScanRegistration.cpp:29
};

struct ScanRegResult {
  std::vector<PointIN> cloud;                  // _laserCloud (ring-major)
  std::vector<int> scanStart, scanEnd;         // _scanIndices (inclusive ranges)
  std::vector<PointI> sharp, lessSharp, flat, lessFlat;     // published feature clouds
  std::vector<PointI> dbgBlind, dbgBlock, dbgSlop, dbgCurv; // debug clouds
  // indices into `cloud`, same order as the clouds above (lessFlatRaw = before the per-ring voxel filter)
  std::vector<int> sharpIdx, lessSharpIdx, flatIdx, lessFlatRawIdx;
  std::vector<int> lessFlatRawRing;            // ring of each lessFlatRaw entry
  std::vector<int> picked;                     // final _scanNeighborPicked per cloud point (0 for skipped rings)
  std::vector<float> curvature;                // region curvature per cloud point (-1 outside any region)
  std::vector<int> classLabel;                 // pointClassify result for points visited in pass 3, else 0x7f
};

void extract_features(const ScanRegParams& prm, ScanRegResult& r);   // ScanRegistration.cpp:190-418
void scanreg_organised(const ScanRegParams& prm, const float* xyzi, int rows, int cols, ScanRegResult& r);
// lidar: 0 VLP-16, 1 HDL-32, 2 HDL-64E  (MultiScanRegistration.h:85-102)
void scanreg_sweep(const ScanRegParams& prm, const float* xyzi, int n, int lidar, ScanRegResult& r);

// IMUState (ScanRegistration.h:122-170) and the IMU history of ScanRegistration (a CircularBuffer of imuHistorySize = 200 states)
struct ImuState {
  double stamp = 0.0;                       // seconds
  float roll = 0.f, pitch = 0.f, yaw = 0.f;   // Angle::rad()
  float pos[3] = {0, 0, 0}, vel[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
};
struct ImuHistory {
  std::vector<ImuState> h;                  // oldest first
  size_t capacity = 200;
  // ScanRegistration::handleIMUMessage (ScanRegistration.cpp:89-121): roll / pitch / yaw as tf's getRPY gives them, raw acceleration
  void push(double stamp, double roll, double pitch, double yaw, double ax, double ay, double az);
};
// MultiScanRegistration::process with hasIMUData() (MultiScanRegistration.cpp:95-200 + ScanRegistration.cpp:123-188): every point is
// projected to the sweep start with the interpolated IMU state.  imuTrans: the four /imu_trans points (ScanRegistration.cpp:681-708)
void scanreg_sweep_imu(const ScanRegParams& prm, const float* xyzi, int n, int lidar, double scanTime, const ImuHistory& imu,
                       ScanRegResult& r, float imuTrans[12]);
int point_classify(const std::vector<PointIN>& cloud, size_t idx, int curvatureRegion);

// pcl::VoxelGrid<PointXYZI>::filter (leaf = cubic)
void voxel_filter(const PointI* in, size_t n, float leaf, std::vector<PointI>& out);

// ---- KNN -------------------------------------------------------------------------------------------------
// k nearest neighbours of q among pts; results sorted by (d2, index); slots beyond the cloud size get
// idx = -1, d2 = FLT_MAX (nanoflann.hpp:91-97 leaves d2[k-1] = max when the tree holds < k points).
typedef void (*KnnBuildFn)(const PointI* pts, size_t n, void** handle);
typedef void (*KnnQueryFn)(void* handle, const float q[3], int k, int* idx, float* d2);
typedef void (*KnnFreeFn)(void* handle);
struct KnnBackend { KnnBuildFn build; KnnQueryFn query; KnnFreeFn free; };
KnnBackend brute_force_backend();
void knn_brute(const PointI* pts, size_t n, const float q[3], int k, int* idx, float* d2);

// ---- scan-to-map solver (ScanMatch.cpp) --------------------------------------------------------------------
struct MatchParams {
  int maxIterations = 10;        // ScanMatch.h:36
  float deltaTAbort = 0.1f;      // LaserMatcher.cpp:94 (class default 0.05, ScanMatch.cpp:22)
  float deltaRAbort = 0.1f;
  bool useScore = false;         // LaserMatcher.cpp:95
  double scoreThreshold = 800;   // ScanMatch.cpp:24
  float matchPercentageThreshold = 0.4f;
  float knnGate = 5.0f;          // ScanMatch.cpp:102,120
  float planeMaxDistance = 0.2f; // ScanMatch.cpp:122
};
struct IterLog {
  float pose_in[6];
  float AtA[36], AtB[6], x[6];
  int rows, lineMatches, planeMatches;
  std::vector<int> nnCorner, nnSurf;   // 5 per query, -1 where the gate rejected
  bool degenerate;
};
struct MatchResult {
  bool ok = false;           // return value of scanMatchScan
  bool converged = false;
  bool tooFewRef = false;
  bool tooFewMatches = false;
  bool degenerate = false;
  int iterations = 0;        // iterations whose update was applied
  int lastRows = 0, lastLine = 0, lastPlane = 0;
  double score = 0;
  std::vector<IterLog> log;  // filled when keepLog
};
void scan_match(const MatchParams& prm, const KnnBackend& knn, const PointI* refCorner, size_t nRefCorner,
                const PointI* refSurf, size_t nRefSurf, const PointI* corner, size_t nCorner,
                const PointI* surf, size_t nSurf, float pose[6], MatchResult& res, bool keepLog);
// neighbour source of one query: fills ind[5] / sq[5] and returns the cloud the indices refer to (NULL: query skipped)
typedef std::function<const PointI*(bool isCorner, const float sel[3], int* ind, float* sq)> NeighbourLookup;
void scan_match_impl(const MatchParams& prm, const NeighbourLookup& lookup, const PointI* corner, size_t nCorner, const PointI* surf,
                     size_t nSurf, float pose[6], MatchResult& res, bool keepLog);
bool find_line(const PointI* cloud, const int* idx, float A[3], float B[3]);                    // feature_utils.h:108-154
bool find_plane(const PointI* cloud, const int* idx, float maxDistance, float plane[4]);        // :157-204
bool corner_coefficients(const float A[3], const float B[3], const float X[3], float coeff[4]); // :63-75
bool surface_coefficients(const float plane[4], const float X[3], float coeff[4]);              // :97-106

// ---- Isometry helpers (host side of the boundary; Eigen::Isometry3f semantics) ---------------------------------
struct Iso { float R[9]; float t[3]; };
Iso iso_identity();
Iso iso_mul(const Iso& a, const Iso& b);
Iso iso_inverse(const Iso& a);
void twist_to_iso(const float pose[6], Iso& it);   // convertTransform(Twist&, Isometry3f&) transform_utils.h:308-311
void iso_to_twist(const Iso& it, float pose[6]);   // convertTransform(Isometry3f&, Twist&) :313-331

// ---- FeatureMap (FeatureMap.h) + LaserMapping::process ---------------------------------------------------------
struct MapParams {
  int cubeW = 121, cubeH = 121, cubeD = 11;   // LaserMatcher.cpp:107-113
  float cubeSize = 50.f;                      // FeatureMap.h:65
  float validDistance = 150.f;                // FeatureMap.h:66
  float mapFilterCorner = 1.0f, mapFilterSurf = 1.0f;   // LaserMatcher.cpp:87-92
  float filterCorner = 1.0f, filterSurf = 1.0f;         // LaserMatcher.cpp:80-85
};
class FeatureMap {
 public:
  explicit FeatureMap(const MapParams& p);
  void update(const float sensor[3]);                                            // FeatureMap.h:232-254
  void getSurroundFeature(std::vector<PointI>& corner, std::vector<PointI>& surf) const;   // :256-265
  void addFeatureCloud(const std::vector<PointI>& corner, const std::vector<PointI>& surf, const Iso& tf);  // :219-230
  // localisation matcher FeatureMap::scanMatchScan(corner, surf, Twist&), FeatureMap.h:490-690: neighbours from the
  // query's own cube (KD-trees per cube, :437,451), fixed 10 iterations and 0.05 / 0.05 thresholds
  void scanMatchScan(const KnnBackend& knn, const std::vector<PointI>& corner, const std::vector<PointI>& surf, float pose[6],
                     MatchResult& res, bool keepLog);
  // saveCloudToFiles / loadCloudFromFiles cube enumeration (:378-462): non-empty (cube, class) pairs in file order
  void fileOrder(std::vector<int>& type, std::vector<int>& ci, std::vector<int>& cj, std::vector<int>& ck) const;
  void loadCube(int type, int i, int j, int k, const std::vector<PointI>& cloud);   // :428-456 (filter, then swap in)
  const std::vector<size_t>& validCubes() const { return _cubeValidInd; }
  int worldToIndex(float x, float y, float z) const;   // :464-473
  size_t totalPoints() const;
  std::vector<std::vector<PointI>> cornerCube, surfCube;
  int originW, originH, originD;
 private:
  bool isIndexValid(int i, int j, int k) const;
  int toIndex(int i, int j, int k) const { return i + j * _p.cubeW + k * _p.cubeW * _p.cubeH; }
  bool worldToCube(float x, float y, float z, int& i, int& j, int& k) const;
  void shift(int di, int dj, int dk);
  void computeActiveArea(const float sensor[3]);
  void downsizeValidCloud();
  MapParams _p;
  int _curW = 0, _curH = 0, _curD = 0;
  std::vector<size_t> _cubeValidInd;
};

// LaserMapping::process (LaserMapping.cpp:39-59) without the ROS plumbing: one frame in, pose out.
class LaserMapping {
 public:
  LaserMapping(const MapParams& mp, const MatchParams& sp, const KnnBackend& knn);
  // LaserLocalization::process (LaserLocalization.cpp:163-188): same frame preparation, FeatureMap::scanMatchScan, no map update
  Iso localize(const Iso& odom, const std::vector<PointI>& corner, const std::vector<PointI>& surf);
  // odom: the odometry pose of this frame (Isometry, /laser_odom_to_init); corner/surf: feature clouds in the
  // sensor frame (/laser_cloud_corner_last, /laser_cloud_surf_last).  Returns the mapped pose.
  Iso process(const Iso& odom, const std::vector<PointI>& corner, const std::vector<PointI>& surf);
  FeatureMap map;
  MatchResult lastMatch;
  std::vector<PointI> cornerDS, surfDS, surroundCorner, surroundSurf;
  Iso mappedLast, mappedNew, odomLast;
  bool keepLog = false;
 private:
  MapParams _mp; MatchParams _sp; KnnBackend _knn;
};

// LaserMappingLocal::process (LaserMappingLocal.cpp:33-78) over LocalFeatureMap (io_module/LocalFeatureMap.h:62-99) and
// FrameUpdater (io_module/FrameUpdater.hpp:17-42): the map is a sliding window of voxel-filtered frames, the surround cloud
// the voxel-filtered (corner 0.2, surf 0.4, LocalFeatureMap.h:29-32) union of the window.
// The reference places a frame in the window with `_transformTobeMapped`, a Twist that is declared (LaserMatcher.h:121) and
// never assigned: frames enter untransformed with the identity as key pose, the travelled distance stays 0 and clean() never
// drops a frame.  useMappedPose = false restates that literally; true uses _lidarMappedNew (the evident intent).
struct LocalFrame { std::vector<PointI> corner, surf; double accum; };
class LaserMappingLocal {
 public:
  LaserMappingLocal(const MapParams& mp, const MatchParams& sp, const KnnBackend& knn, bool useMappedPose);
  Iso process(const Iso& odom, const std::vector<PointI>& corner, const std::vector<PointI>& surf);
  std::vector<LocalFrame> window;     // data_queue
  double accumDistance = 0.0;         // FrameUpdater::accum_distance
  MatchResult lastMatch;
  std::vector<PointI> cornerDS, surfDS, surroundCorner, surroundSurf;
  Iso mappedLast, mappedNew, odomLast;
 private:
  MapParams _mp; MatchParams _sp; KnnBackend _knn; bool _useMapped;
  bool _first = true; double _prevR[9], _prevT[3];   // FrameUpdater::is_first / prev_keypose (Isometry3d)
};

// ---- scan-to-scan odometry (LaserOdometry.cpp) -------------------------------------------------------------------
struct OdomIterLog { float pose_in[6]; float x[6]; int rows; };
class LaserOdometry {
 public:
  LaserOdometry();
  // one frame: the four feature clouds of scan registration (sensor frame, intensity = ring + relTime)
  void process(const std::vector<PointI>& sharp, const std::vector<PointI>& lessSharp, const std::vector<PointI>& flat,
               const std::vector<PointI>& lessFlat, const KnnBackend& knn);
  int maxIterations = 25;          // LaserOdometry.cpp:24
  float deltaTAbort = 0.1f, deltaRAbort = 0.1f;
  bool systemInited = false;
  float transform[6];              // _transform (persists across frames)
  Iso Tsum;                        // /laser_odom_to_init
  std::vector<PointI> lastCorner, lastSurf;   // /laser_cloud_corner_last, /laser_cloud_surf_last
  int iterations = 0, lastRows = 0;
  std::vector<OdomIterLog> log;
  std::vector<int> ind;            // correspondence indices of the last scanMatch: corner {1,2}, surf {1,2,3}
 private:
  void scanMatch(const std::vector<PointI>& sharp, const std::vector<PointI>& flat, const KnnBackend& knn);
};

}  // namespace cmo
