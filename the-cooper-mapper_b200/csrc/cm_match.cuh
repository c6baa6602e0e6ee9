// cm_match.cuh -- device-side state of the scan-to-map Gauss-Newton solver and the kernels' launch interface.
// Mirrors the locals of ScanMatch::scanMatchScan (ScanMatch.cpp:51-347).
#pragma once
#include "cm_device.cuh"

namespace cm {

// flags in MatchState::flags
enum { CM_F_CONVERGED = 1, CM_F_TOO_FEW_REF = 2, CM_F_TOO_FEW_MATCHES = 4, CM_F_DEGENERATE = 8 };

struct MatchParamsDev {
  int max_iterations;      // ScanMatch.h:36
  float delta_t_abort;     // ScanMatch.cpp:257 (LaserMatcher.cpp:94 sets 0.1, 0.1)
  float delta_r_abort;
  float knn_gate;          // 5.0, ScanMatch.cpp:102,120
  float plane_max_dist;    // 0.2, ScanMatch.cpp:122
  int min_ref_corner;      // 50, ScanMatch.cpp:57
  int min_ref_surf;        // 100, ScanMatch.cpp:58
  int min_rows;            // 50, ScanMatch.cpp:142
  float eig_threshold;     // 100, ScanMatch.cpp:223 (10 in LaserOdometry.cpp:596)
  int few_rows_continue;   // 0: fewer than min_rows rows ends the loop (ScanMatch.cpp:141-145); 1: the iteration is skipped
                           //    (LaserOdometry.cpp:501-503)
  int own_cube_only;       // 1: a query only sees the map points of its own 50 m cube and needs >= 5 of them there -- the
                           //    localisation matcher FeatureMap::scanMatchScan (FeatureMap.h:490-690); 0: ScanMatch on the surround map
  int nan_guard;           // 1: non-finite pose components are reset to 0 (LaserOdometry.cpp:622-634)
};

// One per stream.  pose = Twist (rot_x, rot_y, rot_z, pos) with the Angle class' cached sin/cos (Angle.h).
struct MatchState {
  float pose[6];
  float sn[3], cs[3];      // sin / cos of rot_x, rot_y, rot_z
  float R[9];              // convertTransform(Twist -> Isometry3f), transform_utils.h:308-311
  float P[36];             // matP, ScanMatch.cpp:234
  int done;                // loop left (converged, too few matches, too few reference points)
  int flags;
  int iterations;          // updates applied
  int rows, line, plane;   // counters of the last evaluated iteration
  double score;            // sum exp(-|w d|), ScanMatch.cpp:42-49 (last evaluated iteration)
};

// Optional per-iteration trace (tests): what the oracle's IterLog holds.
struct IterTrace {
  float pose_in[6];
  float AtA[36], AtB[6], x[6];
  int rows, line, plane, degenerate;
};

// One row of the linearised system per query (ScanMatch.cpp:154-204); flag bit0 = row kept, bit1 = match counted.
struct __align__(16) RowOut {
  float a[6];
  float b;
  int flag;
};

}  // namespace cm
