/* coopermap.h -- C ABI of the B200-native LOAM hot path (scan registration + scan-to-map registration).
 *
 * Drop-in boundary for ZhekaiJin/the-Cooper-Mapper (L_SLAM).  Every entry point names the reference interface it
 * replaces (file:line relative to the reference tree).  Plain pointers and sizes only; clouds are arrays of
 * cm_point = the x, y, z, intensity payload of pcl::PointXYZI (16 bytes, not PCL's 32-byte padded struct; see
 * INTEGRATION.md for the one-line repack a nodelet does).  All *_host entry points take HOST buffers and copy
 * inside the call; the *_dev entry points take device pointers (inputs already resident in HBM).
 *
 * Status codes: 0 = ok; > 0 = soft outcome of the algorithm (the reference prints a warning and carries on);
 * < 0 = hard error (cm_last_error() has the text).  No exceptions cross this boundary.  A context is NOT
 * thread-safe: one context <-> one CUDA stream <-> one host thread (the reference drives each stage from one
 * worker thread, LaserMapping.cpp:21,27-37).  There is no CPU fallback: without a CUDA device cm_ctx_create fails.
 */
#ifndef COOPERMAP_H
#define COOPERMAP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM_OK 0
#define CM_TOO_FEW_REF 1        /* "reference cloud points too few."  ScanMatch.cpp:57-61 */
#define CM_TOO_FEW_MATCHES 2    /* "matched cloud points too few."    ScanMatch.cpp:141-145 */
#define CM_NOT_CONVERGED 3      /* iteration cap hit, pose still written back  ScanMatch.cpp:342-346 */
#define CM_LOW_SCORE 4          /* ScanMatch.cpp:323-335 (useScore) */
#define CM_ERR_ARG (-1)
#define CM_ERR_CUDA (-2)
#define CM_ERR_CAPACITY (-3)
#define CM_ERR_UNSUPPORTED (-4)

typedef struct cm_ctx cm_ctx;

typedef struct cm_point { float x, y, z, intensity; } cm_point;      /* pcl::PointXYZI payload */
typedef struct cm_pose { float rx, ry, rz, tx, ty, tz; } cm_pose;    /* lidar_slam::Twist, Twist.h:14-20 (R = Rz Ry Rx) */
typedef struct cm_iso { float R[9]; float t[3]; } cm_iso;            /* Eigen::Isometry3f: row-major rotation + translation */

/* Parameters.  Defaults (cm_config_default) are the reference's: RegistrationParams (ScanRegistration.cpp:32-49),
 * ScanMatch as configured by LaserMatcher (LaserMatcher.cpp:80-118, ScanMatch.cpp:21-33), FeatureMap (FeatureMap.h:59-70). */
typedef struct cm_config {
  int device;                         /* CUDA ordinal */
  /* scan registration */
  float scan_period;                  /* 0.1 */
  int n_feature_regions;              /* 6 */
  int curvature_region;               /* 5 */
  int max_corner_sharp;               /* 2 */
  int max_surface_flat;               /* 4 */
  float less_flat_filter_size;        /* 0.2 */
  float surface_curvature_threshold;  /* 0.02 */
  float blind_degree_threshold;       /* 0.5 */
  float blind_radius;                 /* 2.5  OrganizedScanRegistration.cpp:29 */
  /* scan-to-map solver */
  int max_iterations;                 /* 10   ScanMatch.h:36 */
  float delta_t_abort, delta_r_abort; /* 0.1, 0.1  LaserMatcher.cpp:94 */
  int use_score;                      /* 0    LaserMatcher.cpp:95 */
  double score_threshold;             /* 800  ScanMatch.cpp:24 */
  float match_percentage_threshold;   /* 0.4 */
  /* mapping stage */
  float filter_corner, filter_surf;          /* 1.0, 1.0  LaserMatcher.cpp:80-85 */
  float map_filter_corner, map_filter_surf;  /* 1.0, 1.0  LaserMatcher.cpp:87-92 */
  int cube_w, cube_h, cube_d;                /* 121, 121, 11  LaserMatcher.cpp:107-113 */
  float cube_size, valid_distance;           /* 50, 150  FeatureMap.h:65-66 */
  /* search grid (implementation parameters; results do not depend on them) */
  float cell_corner, cell_surf;       /* edge of the hash cells; <= 0: 8 x / 4 x the matching map leaf */
  /* execution (implementation parameters; results do not depend on them) */
  int gn_groups;                      /* batched mapping: the streams are split into this many groups whose Gauss-Newton
                                         loops run on concurrent CUDA streams (one group's 6x6 solve overlaps another
                                         group's search); <= 0: 1 (measured on B200 with 64 HDL-64 streams: 4 groups = +4 % throughput) */
} cm_config;

typedef struct cm_match_stats {
  int status;        /* CM_OK (converged), CM_TOO_FEW_REF, CM_TOO_FEW_MATCHES, CM_NOT_CONVERGED, CM_LOW_SCORE */
  int ret;           /* the reference's bool return (always 0 when use_score = 0, quirk: ScanMatch.cpp:263,342-346) */
  int converged, degenerate;
  int iterations;    /* pose updates applied */
  int rows, line_matches, plane_matches;   /* counters of the last evaluated iteration */
  double score;
} cm_match_stats;

/* one Gauss-Newton iteration as the oracle logs it (tests / diagnostics) */
typedef struct cm_iter_trace {
  float pose_in[6];
  float AtA[36], AtB[6], x[6];
  int rows, line_matches, plane_matches, degenerate;
} cm_iter_trace;

void cm_config_default(cm_config* cfg);
int cm_ctx_create(const cm_config* cfg, cm_ctx** out);
void cm_ctx_destroy(cm_ctx* ctx);
const char* cm_last_error(const cm_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
unsigned long long cm_launch_count(const cm_ctx* ctx);

/* Exact 5 nearest neighbours of nq map-frame queries (xyz triples) in `map`, squared L2 in float, ordered by
 * (d2, index).  Exact for neighbours within sqrt(gate) of the query (the reference rejects d2[4] >= 5.0 anyway);
 * beyond that idx = -1 / d2 is a lower bound on the true 5th distance that is >= gate.
 * Replaces nanoflann::KdTreeFLANN::setInputCloud + nearestKSearch, nanoflann_pcl.h:132-162 (call sites
 * ScanMatch.cpp:75-76,100-101,119). */
int cm_knn5_host(cm_ctx* ctx, const cm_point* map, size_t n_map, float cell, float gate, const float* queries_xyz,
                 size_t nq, int* idx_out, float* d2_out);

/* ScanMatch::scanMatchScan(refCorner, refSurf, corner, surf, Twist&), ScanMatch.h:51-55 / ScanMatch.cpp:51-347.
 * pose is read as the initial guess and overwritten with the result (also when not converged, like the reference).
 * trace (optional, cfg.max_iterations entries) and nn_corner / nn_surf (optional, max_iterations x n x 5 ints,
 * -1 where the 5.0 gate rejected the query) expose every iteration for parity tests. */
int cm_match_stateless_host(cm_ctx* ctx, const cm_point* ref_corner, size_t n_ref_corner, const cm_point* ref_surf,
                            size_t n_ref_surf, const cm_point* corner, size_t n_corner, const cm_point* surf, size_t n_surf,
                            cm_pose* pose, cm_match_stats* stats, cm_iter_trace* trace, int* nn_corner, int* nn_surf);

/* ScanMatch::scanMatchScan(..., Eigen::Isometry3f&), ScanMatch.h:47-50 / ScanMatch.cpp:349-360: Isometry -> Twist
 * (getEulerAngles, transform_utils.h:54-60) -> solve -> Twist -> Isometry. */
int cm_match_stateless_iso_host(cm_ctx* ctx, const cm_point* ref_corner, size_t n_ref_corner, const cm_point* ref_surf,
                                size_t n_ref_surf, const cm_point* corner, size_t n_corner, const cm_point* surf,
                                size_t n_surf, cm_iso* pose, cm_match_stats* stats);

/* Outputs of scan registration for nstreams independent sweeps.  Clouds carry intensity = ring + relTime, like the
 * reference's published feature clouds (toXYZI, pcl_util.h:30-37).  Optional members may be NULL. */
typedef struct cm_scanreg_out {
  cm_point* pts[4];        /* [nstreams][cap[k]]: 0 /laser_cloud_sharp, 1 /laser_cloud_less_sharp, 2 /laser_cloud_flat,
                              3 /laser_cloud_less_flat (after the per-ring 0.2 m voxel filter) */
  int cap[4];
  int* n;                  /* [nstreams][5]: sizes of the four clouds + number of less-flat points before the filter */
  cm_point* cloud;         /* optional [nstreams][rows*cols]: ring-major full-resolution cloud (/velodyne_cloud_2) */
  float* cloud_curvature;  /* optional, with cloud: its curvature field (= ring + relTime) */
  int* scan_ranges;        /* optional [nstreams][rows][2]: inclusive index range of every ring (_scanIndices) */
  int* idx[4];             /* optional [nstreams][rows*cols]: cloud indices of sharp, less_sharp, flat, UNFILTERED less-flat */
  signed char* picked;     /* optional [nstreams][rows*cols]: final _scanNeighborPicked */
  float* curvature;        /* optional [nstreams][rows*cols]: region curvature, -1 outside the regions */
  signed char* label;      /* optional [nstreams][rows*cols]: pointClassify label where evaluated, 127 elsewhere */
} cm_scanreg_out;

/* OrganisedScanRegistration::process (OrganizedScanRegistration.cpp:82-150) + ScanRegistration::extractFeatures
 * (ScanRegistration.cpp:190-418) for nstreams organised sweeps of rows x cols points (ring = row, missing returns
 * NaN), frames[s][row][col]. */
int cm_scanreg_organised_host(cm_ctx* ctx, const cm_point* frames, int nstreams, int rows, int cols, cm_scanreg_out* out);

/* MultiScanRegistration::process (MultiScanRegistration.cpp:95-200) + extractFeatures for ONE raw azimuth-major sweep of
 * n points of a spinning multi-beam LiDAR (lidar: 0 VLP-16, 1 HDL-32, 2 HDL-64E -- linear ring mappers, MultiScanRegistration.h:90-102;
 * 3 Pandar40 -- MultiScanMapperP, MultiScanRegistration.h:24-42 with lidar_type.h:78-104).  The whole front end (axis swap, ring from
 * the elevation angle, azimuth unwrap with the half-sweep flag, relTime, stable per-ring append) runs on the device; atan / atan2
 * are the library's correctly rounded cm_atanf / cm_atan2f (glibc's float versions differ from them by 1 ulp on ~15 % of the
 * inputs; the oracle uses the same functions).  rows_out / cols_out return the ring-major layout (rings x longest ring); the
 * optional full-resolution outputs of `out` come back in the concatenated _laserCloud order (size them for n entries). */
int cm_scanreg_sweep_host(cm_ctx* ctx, const cm_point* sweep, size_t n, int lidar, cm_scanreg_out* out, int* rows_out, int* cols_out);

/* The IMU branch of scan registration (`hasIMUData()`, ScanRegistration.cpp:89-188).  cm_imu_push_host is handleIMUMessage: one
 * call per sensor_msgs/Imu with the stamp in seconds, roll / pitch / yaw as tf's getRPY gives them and the raw linear acceleration;
 * the context keeps the last 200 states (imuHistorySize) with gravity removed and position / velocity integrated.
 * cm_scanreg_sweep_imu_host is cm_scanreg_sweep_host for a sweep stamped scan_time: every accepted point is projected to the sweep
 * start with the IMU state interpolated at its relTime (setIMUTransformFor + transformToStartIMU, :145-166) on the device -- the
 * forward-only `_imuIdx` of the reference is a prefix maximum over relTime -- and imu_trans (12 floats, optional) returns the four
 * /imu_trans points of publishResult (:681-708).  With an empty history it equals cm_scanreg_sweep_host. */
typedef struct cm_imu_sample { double stamp, roll, pitch, yaw, ax, ay, az; } cm_imu_sample;
int cm_imu_push_host(cm_ctx* ctx, const cm_imu_sample* m);
int cm_imu_clear(cm_ctx* ctx);
int cm_scanreg_sweep_imu_host(cm_ctx* ctx, const cm_point* sweep, size_t n, int lidar, double scan_time, cm_scanreg_out* out, int* rows_out,
                              int* cols_out, float* imu_trans);

/* pcl::VoxelGrid<pcl::PointXYZI>::filter with a cubic leaf, batched over nseg independent clouds: cloud s is
 * in[s*cap_in .. s*cap_in + n_in[s]) and its result out[s*cap_out .. s*cap_out + n_out[s]), ordered by voxel index,
 * every field (x, y, z, intensity) averaged.  Replaces the filter calls at ScanRegistration.cpp:390-399,
 * LaserMatcher.cpp:293-300, ScanMatch.cpp:381-394 (semantics: util/voxel_grid_partition.hpp:91-272). */
int cm_voxel_filter_host(cm_ctx* ctx, const cm_point* in, int nseg, const int* n_in, int cap_in, float leaf, cm_point* out,
                         int* n_out, int cap_out);

/* ---- the scan-to-map STAGE with a device-resident map ------------------------------------------------------------
 * One context drives nstreams independent LiDAR streams (each with its own pose chain and its own map); every call
 * below processes one frame of every stream in one batch of launches.  Arrays are indexed [stream]... */

/* Allocate the per-stream maps (replaces `new FeatureMap<PointI>(map_cube_x, map_cube_y, map_cube_z)` +
 * setupFilterSize, LaserMatcher.cpp:107-116).  Capacities are in points per stream and class. */
int cm_mapping_create(cm_ctx* ctx, int nstreams, size_t max_corner_points, size_t max_surf_points);

/* LaserMapping::process (LaserMapping.cpp:39-59) for one frame per stream.  odom[s]: /laser_odom_to_init pose;
 * corner / surf: /laser_cloud_corner_last and /laser_cloud_surf_last as [nstreams][cap_*] with counts n_*[s];
 * mapped[s]: /aft_mapped_to_init pose.  Steps: transformMerge (LaserMatcher.cpp:333-340), voxel-filter the frame
 * (:288-301), FeatureMap::update + surround selection (:303-325), ScanMatch::scanMatchScan on the map (:327-331),
 * transformUpdate (:342-347), FeatureMap::addFeatureCloud (:349-355).  FeatureMap::shift (sensor within 3 cubes of the grid border,
 * FeatureMap.h:232-245, 354-376) is reproduced literally, including the cubes its in-place pointer swaps move the wrong way. */
int cm_mapping_process_host(cm_ctx* ctx, const cm_iso* odom, const cm_point* corner, const int* n_corner, int cap_corner,
                            const cm_point* surf, const int* n_surf, int cap_surf, cm_iso* mapped, cm_match_stats* stats);
/* cm_mapping_process_host and cm_pipeline_step_* return as soon as the poses are on the host: FeatureMap::addFeatureCloud of that
 * frame (the map insertion) is enqueued behind them on the context's stream and finishes on its own, while the caller publishes
 * the pose and fetches the next sweep.  Everything that reads or changes the map afterwards is ordered behind it.  What the
 * insertion hit (CM_ERR_CAPACITY, voxel range) is reported by the NEXT step -- or by cm_mapping_sync, which waits for it. */
int cm_mapping_sync(cm_ctx* ctx);

/* LaserLocalization::process (LaserLocalization.cpp:163-188) for one frame per stream against the map held by the context
 * (built with cm_map_insert_host / cm_map_load_host): same frame preparation as cm_mapping_process_host, but the pose is
 * refined by the localisation matcher FeatureMap::scanMatchScan (FeatureMap.h:490-690) -- the 5 neighbours of a query
 * come from the query's own 50 m cube only, which must hold >= 5 points (:521-527); no reference-size gate; 10
 * iterations; converged below 0.05 deg and 0.05 cm -- and the map is left untouched (featureMapUpdate is commented out
 * in the reference).  The IMU blending of LaserLocalization::transformUpdate stays with the caller. */
int cm_localization_process_host(cm_ctx* ctx, const cm_iso* odom, const cm_point* corner, const int* n_corner, int cap_corner,
                                 const cm_point* surf, const int* n_surf, int cap_surf, cm_iso* mapped, cm_match_stats* stats);

/* LaserMappingLocal (odometry/LaserMappingLocal.cpp:33-78): the mapping stage over LocalFeatureMap
 * (io_module/LocalFeatureMap.h:28-99) instead of the cube map -- a sliding window of voxel-filtered frames.  One frame:
 * transformMerge, prepareFeatureFrame (cfg.filter_corner / filter_surf), surround = VoxelGrid(0.2 corner / 0.4 surf) of the
 * concatenated window (getSurroundFeature, :84-99), ScanMatch::scanMatchScan, transformUpdate, then the filtered frame is
 * appended to the window (addDataFrame :62-69) with FrameUpdater's travelled distance (io_module/FrameUpdater.hpp:17-42), and
 * clean() (:70-82) drops the frames more than 30 m of travel behind -- one more than it counts, as the reference does.
 * use_mapped_pose = 0 is the reference as written: the frame is placed with `_transformTobeMapped`, a Twist that is declared
 * (LaserMatcher.h:121) and never assigned, i.e. the identity -- frames stay in the sensor frame, the travelled distance stays
 * 0 and nothing is ever dropped.  use_mapped_pose = 1 places the frame with the mapped pose (_lidarMappedNew).
 * The window lives on the device; one window per context. */
int cm_mapping_local_create(cm_ctx* ctx, int use_mapped_pose);
int cm_mapping_local_process_host(cm_ctx* ctx, const cm_iso* odom, const cm_point* corner, size_t n_corner, const cm_point* surf,
                                  size_t n_surf, cm_iso* mapped, cm_match_stats* stats);
/* window state: frames in the queue, their corner / surf point totals, sizes of the last surround clouds (corner, surf) and
 * FrameUpdater's accumulated distance; corner_out / surf_out (may be NULL): the window clouds in queue order */
int cm_mapping_local_window_host(cm_ctx* ctx, int* n_frames, size_t* n_corner, size_t* n_surf, int* n_surround2, double* accum_distance,
                                 cm_point* corner_out, size_t cap_corner, cm_point* surf_out, size_t cap_surf);

/* Scan registration + mapping in one call: frames[s][row][col] organised sweeps -> mapped poses.  The less-sharp and
 * less-flat clouds of cm_scanreg_organised feed cm_mapping_process without leaving the device.  _dev: `frames` is a
 * DEVICE pointer (inputs already resident in HBM); poses and stats stay host arrays. */
/* cm_pipeline_prefetch_host / _dev issue the NEXT step's sweeps ahead of time on a second CUDA stream and return at once:
 * the host variant uploads them (pinned host memory), both run scan registration on them.  The following
 * cm_pipeline_step_host / _dev call with the same `frames` pointer consumes that work instead of redoing it, so the
 * host-to-device transfer and the (issue-bound) feature extraction of step k+1 overlap the (latency-bound) matching and map
 * kernels of step k -- the nodelet receives the next PointCloud2 while the current one is being registered.  At most four
 * sweeps may be in flight (one being consumed, three pending); the buffer must stay untouched until its step has run.  Results are identical either way. */
int cm_pipeline_prefetch_host(cm_ctx* ctx, const cm_point* frames, int rows, int cols);
int cm_pipeline_prefetch_dev(cm_ctx* ctx, const void* d_frames, int rows, int cols);
/* The same, but only REGISTERED now and issued by the next cm_pipeline_step_* call right after it has submitted its
 * Gauss-Newton loop: the host time of the submission (two copies, six launches) hides behind device work instead of delaying
 * the step.  One registration at a time; an error of the deferred prefetch is returned by that step call. */
int cm_pipeline_prefetch_deferred_host(cm_ctx* ctx, const cm_point* frames, int rows, int cols);
int cm_pipeline_prefetch_deferred_dev(cm_ctx* ctx, const void* d_frames, int rows, int cols);
int cm_pipeline_step_host(cm_ctx* ctx, const cm_point* frames, int rows, int cols, const cm_iso* odom, cm_iso* mapped,
                          cm_match_stats* stats);
int cm_pipeline_step_dev(cm_ctx* ctx, const void* d_frames, int rows, int cols, const cm_iso* odom, cm_iso* mapped,
                         cm_match_stats* stats);
/* The same for sweeps the way a nodelet holds them after pcl::fromROSMsg (util/ros_utils.h:27-35): ONE cloud per stream,
 * clouds[s] = &cloud_s->points[0], in ordinary (pageable) host memory, points `stride` bytes apart with x, y, z as three floats at
 * offset 0 -- stride 32 for pcl::PointXYZI, 48 for PointXYZINormal, 16 for cm_point, 12 for packed coordinates, the point_step of a
 * sensor_msgs::PointCloud2 data buffer (any value >= 12, no alignment required).  The library
 * gathers the coordinates into its own pinned staging buffers with worker threads (COOPERMAP_STAGE_THREADS, default: the cores
 * of the calling thread's affinity mask, at most 16), uploads 12 bytes per point chunk by chunk while the next chunk is being
 * packed, and expands them on the device; the intensity is not transferred (scan registration replaces it with ring + relTime,
 * OrganizedScanRegistration.cpp:109-110).  The clouds may be reused as soon as the call returns; the slot is keyed by clouds[0].
 * Special case, the xyz-only entry: stride 12 with the clouds of all streams back to back in ONE buffer needs no packing -- the
 * buffer is uploaded as it is (asynchronously if it is pinned; it must then stay untouched until its step has run, like the
 * packed 16-byte entries) and a quarter fewer bytes cross PCIe.  Results are bit-identical to the packed entries. */
int cm_pipeline_prefetch_strided_host(cm_ctx* ctx, const void* const* clouds, size_t stride, int rows, int cols);
int cm_pipeline_step_strided_host(cm_ctx* ctx, const void* const* clouds, size_t stride, int rows, int cols, const cm_iso* odom,
                                  cm_iso* mapped, cm_match_stats* stats);

/* FeatureMap::update(sensorPose) (FeatureMap.h:232-254) of one stream outside a stage step (the stage entries do it themselves):
 * shift() if the sensor nears the border of the cube grid, then the valid-cube window of computeActiveAera around `sensor_xyz`. */
int cm_map_update_host(cm_ctx* ctx, int stream_index, const float* sensor_xyz);

/* FeatureMap::addFeatureCloud(cornerCloud, surfCloud, tf) (FeatureMap.h:219-230): transform by tf[s], push into the 50 m cubes,
 * downsizeValidCloud (:289-306): the VALID cubes (the window of the last update: cm_map_update_host or a stage step) are voxel-
 * filtered -- incrementally: only the voxels that received points change --, points pushed into other cubes stay unfiltered until
 * their cube is valid during a later insert, exactly like the reference's cube clouds.  Before any update every cube counts as
 * valid (a map that is only being built, like FeatureMap::loadCloudFromFiles filtering every file). */
int cm_map_insert_host(cm_ctx* ctx, const cm_point* corner, const int* n_corner, int cap_corner, const cm_point* surf,
                       const int* n_surf, int cap_surf, const cm_iso* tf);

/* Every resident point of one stream's map (cls 0 corner, 1 surf) with the index of its cube (i + j*W + k*W*H), in
 * storage order; *n_out is the total even when it exceeds cap.  Sorting by (cube, voxel) gives the reference's
 * cube clouds (the content FeatureMap::saveCloudToFiles writes, FeatureMap.h:378-412). */
int cm_map_export_host(cm_ctx* ctx, int stream_index, int cls, cm_point* out, int* cube_index, size_t cap, size_t* n_out);

/* "Map cloud out".  FeatureMap::getSurroundFeature (FeatureMap.h:256-265): the clouds of the valid cubes of the last
 * FeatureMap::update (cm_mapping_process / cm_pipeline_step), concatenated in the order computeActiveAera found them (i, j, k
 * loops, :308-352), every cube in pcl::VoxelGrid order -- what the reference publishes as /laser_cloud_surround_{corner,surf}
 * (LaserMatcher.cpp:357-394).  n_out2[0 / 1] return the full sizes even when they exceed the capacities.
 * cm_map_full_host = FeatureMap::getFullMap (:267-287): every cube in index order, its corner cloud then its surf cloud, each
 * re-filtered with `leaf` (the reference's map_filter_full, LaserMatcher.cpp:116) -- /FullMap and the ~saveMap service (:164-188). */
int cm_map_surround_host(cm_ctx* ctx, int stream_index, cm_point* out_corner, size_t cap_corner, cm_point* out_surf, size_t cap_surf,
                         size_t* n_out2);
int cm_map_full_host(cm_ctx* ctx, int stream_index, float leaf, cm_point* out, size_t cap, size_t* n_out);

/* FeatureMap::saveCloudToFiles / loadCloudFromFiles (FeatureMap.h:378-462): <dir>/index.txt ("count type i j k size" per
 * file, type 0 corner / 1 surf, cubes enumerated i, j, k with corner before surf) + <dir>/<count>.pcd, binary PCD files as
 * pcl::io::savePCDFileBinary writes them for pcl::PointXYZI; a cube's points are stored in VoxelGrid order.  Loading pushes
 * every file through the map voxel filter like the reference; n_misplaced counts points whose coordinates put them into
 * another cube than the index line says (the reference keeps them in the named cube, here they follow their coordinates --
 * zero for files written by either implementation).  ascii and binary PCD are read, binary_compressed is not. */
int cm_map_save_host(cm_ctx* ctx, int stream_index, const char* dir, int* n_files);
int cm_map_load_host(cm_ctx* ctx, int stream_index, const char* dir, int* n_files, size_t* n_points, size_t* n_misplaced);

/* DynamicFeatureMap paging (util/DynamicFeatureMap.h:129-161, 504-677) for a prebuilt map that does not fit (or need not sit) in
 * memory as a whole.  cm_map_page_open_host reads <dir>/index2.txt ("count type i j k size" per <count>.pcd file, type 0 corner /
 * 1 surf, GLOBAL cube indices i = round(x / cube_size) as the reference's indexConvert writes them) and fixes the resident window
 * (the reference's DynamicFeatureMap(21, 11, 21); sizes are odd, the sensor's cube is the centre).  cm_map_page_update_host is
 * DynamicFeatureMap::update: on the first call every catalogued cube of the window around the sensor is read, every file through
 * the map voxel filter; afterwards, whenever the sensor enters another cube, the cubes that entered the window are read and the
 * cubes that left it are dropped.  The localisation / mapping entries then work on the resident cubes.  The device lattice
 * (cm_config cube_w x cube_h x cube_d) is re-centred on the sensor when the window would leave it, so travel is unbounded. */
int cm_map_page_open_host(cm_ctx* ctx, int stream_index, const char* dir, int window_w, int window_h, int window_d, int* n_entries);
int cm_map_page_update_host(cm_ctx* ctx, int stream_index, const float* sensor_xyz, int* n_files_loaded, int* n_cubes_evicted,
                            size_t* n_points_loaded);

/* ---- scan-to-scan odometry ------------------------------------------------------------------------------------------ */
typedef struct cm_odom_stats {
  int initialising;   /* first frame: clouds stored, no motion estimated (LaserOdometry.cpp:295-303) */
  int matched;        /* scanMatch ran (last clouds had > 10 corner and > 100 surf points, :338) */
  int iterations, rows, converged, degenerate;
} cm_odom_stats;

/* LaserOdometry::process (LaserOdometry.cpp:288-326) for one frame of the context's stream: the four feature clouds of scan
 * registration in (intensity = ring + relTime), /laser_odom_to_init (`odom` = _Tsum) and the clouds projected to the sweep
 * end, /laser_cloud_corner_last and /laser_cloud_surf_last (n_less_sharp / n_less_flat points), out.  `transform` returns
 * the frame-to-frame Twist _transform, which persists as the next initial guess; trace (optional, 25 entries) exposes the
 * Gauss-Newton iterations.  cm_odometry_reset forgets the previous frame. */
int cm_odometry_process_host(cm_ctx* ctx, const cm_point* sharp, int n_sharp, const cm_point* less_sharp, int n_less_sharp,
                             const cm_point* flat, int n_flat, const cm_point* less_flat, int n_less_flat, cm_iso* odom,
                             cm_pose* transform, cm_point* corner_last, cm_point* surf_last, cm_odom_stats* stats,
                             cm_iter_trace* trace);
int cm_odometry_reset(cm_ctx* ctx);

/* The odometry stage for EVERY stream of a context in one set of launches (the batch form of cm_odometry_process_host, like
 * cm_mapping_process_host is of the mapping stage): clouds are [nstreams][cap_*] arrays with per-stream counts, each stream keeps
 * its own _transform / _Tsum / last clouds; streams on their first frame only store their clouds, streams whose last clouds are
 * too small skip scanMatch (LaserOdometry.cpp:338), the others iterate together.  corner_last / surf_last (optional) return
 * [nstreams][cap_less_sharp] / [nstreams][cap_less_flat]. */
int cm_odometry_batch_create(cm_ctx* ctx, int nstreams, int cap_sharp, int cap_less_sharp, int cap_flat, int cap_less_flat);
int cm_odometry_batch_process_host(cm_ctx* ctx, const cm_point* sharp, const int* n_sharp, const cm_point* less_sharp, const int* n_less_sharp,
                                   const cm_point* flat, const int* n_flat, const cm_point* less_flat, const int* n_less_flat, cm_iso* odom,
                                   cm_pose* transform, cm_point* corner_last, cm_point* surf_last, cm_odom_stats* stats);

/* The whole LOAM chain for one organised sweep per stream in ONE call: scan registration (MultiScanRegistration /
 * OrganizedScanRegistration) -> LaserOdometry::process (LaserOdometry.cpp:288-326) -> LaserMapping::process (LaserMapping.cpp:39-59).
 * What the three nodelets pass over ROS topics -- /laser_cloud_sharp, /laser_cloud_less_sharp, /laser_cloud_flat,
 * /laser_cloud_less_flat into the odometry; /laser_cloud_corner_last, /laser_cloud_surf_last and /laser_odom_to_init into the
 * mapping -- stays in device memory; the host gets /laser_odom_to_init (odom[s]) and /aft_mapped_to_init (mapped[s]).  Results are
 * those of cm_scanreg_organised_host + cm_odometry_batch_process_host + cm_mapping_process_host called one after the other.
 * cm_pipeline_chain_create (after cm_mapping_create) sizes the odometry stage for rows x cols sweeps; a sweep announced with
 * cm_pipeline_prefetch_host is taken from there. */
int cm_pipeline_chain_create(cm_ctx* ctx, int rows, int cols);
int cm_pipeline_chain_step_host(cm_ctx* ctx, const cm_point* frames, int rows, int cols, cm_iso* odom, cm_iso* mapped, cm_odom_stats* ostats,
                                cm_match_stats* mstats);
int cm_pipeline_chain_step_dev(cm_ctx* ctx, const void* d_frames, int rows, int cols, cm_iso* odom, cm_iso* mapped, cm_odom_stats* ostats,
                               cm_match_stats* mstats);
/* The same chain for ONE stream (cm_mapping_create(ctx, 1, ...)) fed with RAW sweeps, the unorganised azimuth-major cloud a Velodyne /
 * Pandar driver publishes: MultiScanRegistration::process (MultiScanRegistration.cpp:95-200; lidar 0 VLP-16, 1 HDL-32, 2 HDL-64E,
 * 3 Pandar40) runs on the device in front of the three stages.  max_points: the largest sweep.  scan_time >= 0 de-skews with the IMU
 * states given to cm_imu_push_host (ScanRegistration.cpp:89-188), < 0: no IMU.  Results are those of cm_scanreg_sweep_host (or
 * _imu_host) + cm_odometry_batch_process_host + cm_mapping_process_host. */
int cm_pipeline_chain_sweep_create(cm_ctx* ctx, size_t max_points);
int cm_pipeline_chain_step_sweep_host(cm_ctx* ctx, const cm_point* sweep, size_t n, int lidar, double scan_time, cm_iso* odom, cm_iso* mapped,
                                      cm_odom_stats* ostats, cm_match_stats* mstats);


/* ---- sharded-map matching (BASELINE config 4: one map split over ranks) ---------------------------------------------------
 * ScanMatch::scanMatchScan with the reference clouds partitioned in space: rank r holds the map points of its region plus a
 * sqrt(5) m halo (cm_shard_set_map_host), evaluates per iteration only the queries whose map-frame position lies in its
 * own box [own_lo, own_hi) (cm_shard_partial_host -> 32 doubles: 21 A^T A upper triangle, 6 A^T b, rows, line / plane
 * counts, score), the caller all-reduces (sum) those doubles over the ranks (NCCL / gloo) and every rank calls
 * cm_shard_solve_host with the total: identical input, identical pose on every rank, no broadcast.  Because the sums are
 * exact-product double accumulations the result equals cm_match_stateless_host on the unsplit map bit for bit. */
int cm_shard_set_map_host(cm_ctx* ctx, const cm_point* corner, size_t n_corner, const cm_point* surf, size_t n_surf);
int cm_shard_begin_host(cm_ctx* ctx, const cm_point* corner, size_t n_corner, const cm_point* surf, size_t n_surf,
                        const cm_pose* init, size_t total_ref_corner, size_t total_ref_surf);
int cm_shard_partial_host(cm_ctx* ctx, int iter, const float* own_lo, const float* own_hi, double* sums32);
int cm_shard_solve_host(cm_ctx* ctx, int iter, const double* sums32, cm_pose* pose, int* done, cm_match_stats* stats);

/* ---- one map over several GPUs (BASELINE config 4) -------------------------------------------------------------------------------
 * One process and one context per GPU.  Rank 0 calls cm_dist_unique_id (a 128-byte ncclUniqueId), the host program hands it to
 * the other ranks (MPI, torch.distributed, a file), every rank calls cm_dist_init BEFORE cm_mapping_create.  From then on the
 * context's map keeps only the 50 m cubes this rank owns (cube (i, j, k) of the FeatureMap lattice, FeatureMap.h:475-487, belongs
 * to rank (i + 3 j + 5 k) mod nranks, so the cubes around a sensor spread over all ranks) plus the points within sqrt(5) m of them,
 * the gate of ScanMatch.cpp:102,120 -- every accepted 5-NN of a query that falls into an owned cube is then local.  Every rank
 * feeds the SAME sweeps and poses to the ordinary entry points (cm_map_insert_host, cm_mapping_process_host, cm_pipeline_step_*):
 * a rank evaluates the queries whose map-frame position lies in its cubes, the per-rank partial normal equations (32 doubles per
 * stream) are summed over the ranks once per Gauss-Newton iteration by the library's own exchange kernel -- every rank stores its
 * vector into every peer's mailbox over NVLink (CUDA IPC peer mappings) and adds them in rank order; NCCL all-gather where IPC is
 * not available -- and every rank solves the same 6x6 system: identical poses on all ranks, nothing is broadcast.
 * NCCL is loaded with dlopen (libnccl.so.2); without it cm_dist_init returns CM_ERR_UNSUPPORTED. */
int cm_dist_unique_id(void* id128);
int cm_dist_init(cm_ctx* ctx, const void* id128, int rank, int nranks);
int cm_dist_info(cm_ctx* ctx, int* rank, int* nranks, int* p2p);   /* p2p = 1: peer-memory mailboxes, 0: NCCL all-gather transport */
/* the exchange as an operator: vec[0..n) (n <= 8192) summed over the ranks in rank order; repeat > 1 times it (ms per call) */
int cm_dist_allreduce_host(cm_ctx* ctx, double* vec, int n, int repeat, float* ms_per_call);

/* ---- measurement helpers (no reference counterpart; used by bench.py) ------------------------------------------------
 * cm_timer_record(ctx, 0 | 1) records a CUDA event on the context's stream; cm_timer_elapsed_ms returns event 1 - event 0.
 * cm_prof_enable brackets every launch of the dominant kernel (the fused correspondence kernel) with CUDA events on
 * the launching stream; cm_prof_drain returns their summed duration and count since the last drain.
 * cm_last_step_counters: {query-iterations, queries, inserted points, (reserved)} of the last mapping / pipeline step,
 * summed over the streams -- the run-time counts the algorithmic-bytes formula needs. */
int cm_timer_record(cm_ctx* ctx, int which);
/* stage-alone measurements: cm_timer_record_side records on the stream the prefetched scan registration runs on;
 * cm_pipeline_wait blocks until every prefetched sweep is uploaded and registered; cm_pipeline_discard frees a prefetched
 * sweep without running its step */
int cm_timer_record_side(cm_ctx* ctx, int which);
int cm_pipeline_wait(cm_ctx* ctx);
int cm_pipeline_discard(cm_ctx* ctx, const void* frames);
int cm_timer_elapsed_ms(cm_ctx* ctx, float* ms);
int cm_prof_enable(cm_ctx* ctx, int on);
int cm_prof_drain(cm_ctx* ctx, double* kernel_ms, int* launches);
int cm_prof_drain_scanreg(cm_ctx* ctx, double* kernel_ms, int* launches);   /* same for the sr_ring_kernel launches */
int cm_last_step_counters(cm_ctx* ctx, unsigned long long* out4);
/* development aid: bracket EVERY kernel launch with CUDA events and report "name total_us launches" lines (sorted) */
int cm_timeline_enable(cm_ctx* ctx, int on);
int cm_timeline_report(cm_ctx* ctx, char* buf, size_t cap);

/* ScanMatch::scanMatchLocal(refCorner, refSurf, corner, surf, Twist&) (ScanMatch.h:38-46, ScanMatch.cpp:375-398): all four
 * clouds are voxel-filtered first (corner 0.2 m, surf 0.4 m, ScanMatch.cpp:29-30), then scanMatchScan.  This is the call
 * the pose-graph consumers make (pose_graph/loop_detector.hpp:206-208) with the class defaults use_score = 1 and abort
 * thresholds 0.05 / 0.05 -- set them in cm_config; stats->ret is the reference's bool. */
int cm_match_local_host(cm_ctx* ctx, const cm_point* ref_corner, size_t n_ref_corner, const cm_point* ref_surf, size_t n_ref_surf,
                        const cm_point* corner, size_t n_corner, const cm_point* surf, size_t n_surf, cm_pose* pose,
                        cm_match_stats* stats);

/* Development aid (no reference counterpart): per-warp trace of the 5-NN search kernel in Gauss-Newton evaluation `iter` of
 * the following cm_pipeline_step / cm_mapping_process calls (iter < 0: off).  4 words per warp: start ns, end ns,
 * (max << 32 | sum) level-0 candidates over the lanes, (hard queries << 32 | smid << 16 | is_corner). */
int cm_debug_graph_builds(cm_ctx* ctx, unsigned long long* out4);   /* {Gauss-Newton loop graphs built, insert chain captures, insert repeats, 0} */
int cm_debug_read_slots(cm_ctx* ctx, int* out, size_t n_ints);                                      /* neighbour pool slots of the last mapping step */
int cm_debug_read_queries(cm_ctx* ctx, int cls, float* out, size_t n_floats, int* counts);          /* its filtered query clouds */
int cm_debug_read_map_points(cm_ctx* ctx, int stream_index, int cls, const int* slots, int n, float* out4);   /* map points by pool slot */
int cm_debug_graph_info(cm_ctx* ctx, int* n_graphs, int* while_loop);   /* CUDA graphs cached for the Gauss-Newton loop; 1 = WHILE-node graphs */
int cm_debug_search_trace_enable(cm_ctx* ctx, int iter);
int cm_debug_search_trace_read(cm_ctx* ctx, unsigned long long* out, size_t cap_words, size_t* n_words);

/* Self-test hook (no reference counterpart): run one of the shared small-matrix routines (csrc/cm_math.h -- the restated
 * Eigen algorithms) over n packed inputs ON THE DEVICE, so a test can compare with the same header compiled for the host.
 * op: 0 QR-solve 6x6 (42 -> 6 floats), 1 QR-solve 5x3 (20 -> 3), 2 eig 3x3 (6 -> 12), 3 eig 6x6 (36 -> 42),
 * 4 eig 6x6 values (36 -> 6), 5 inverse 6x6 (36 -> 36), 6 pose -> R, sin, cos (6 -> 15). */
int cm_debug_math_host(cm_ctx* ctx, int op, const float* in, size_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif /* COOPERMAP_H */
