// cm_ctx.h -- the opaque context behind include/coopermap.h (internal).
#pragma once
#include "../../include/coopermap.h"
#include "cm_host.h"
#include <array>
#include <deque>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace cm {

// Host mirror of the per-stream state LaserMatcher keeps (LaserMatcher.h:100-117): poses + the cube lattice position.
struct HostIso { float R[9]; float t[3]; };
struct MappingStream {
  HostIso mappedLast, mappedNew, odomLast;
  int origin[3];   // _cubeOriginWidth/Height/Depth
  int cur[3];      // _curCubeWidth/Height/Depth
};

// LaserMappingLocal (cm_mapping.cu): one DataFrame of LocalFeatureMap's data_queue, clouds resident on the device
struct LocalFrame { DeviceBuffer corner, surf; int nc = 0, ns = 0; double accum = 0.0; };
struct LocalWindow {
  bool created = false, use_mapped = false;
  std::deque<std::unique_ptr<LocalFrame>> frames;   // LocalFeatureMap::data_queue
  double accum = 0.0; bool first = true; double prevR[9], prevT[3];   // FrameUpdater: accum_distance, is_first, prev_keypose
  HostIso mappedLast, mappedNew, odomLast;
  int n_surround[2] = {0, 0};
};

// batched scan-to-scan odometry (cm_odometry.cu): LaserOdometry's members for every stream of the context
struct OdomBatch {
  int S = 0, cap_sharp = 0, cap_less_sharp = 0, cap_flat = 0, cap_less_flat = 0;
  std::vector<int> inited, n_last_c, n_last_s;
  std::vector<float> tf;            // [S][6] _transform
  std::vector<HostIso> Tsum;        // _Tsum
  DeviceBuffer sharp, flat, last_c, last_s, ints, ind, rows, state, sums, pose, tfinv, box_c, box_s;   // box_*: chunk boxes of the last clouds
  GridBatch grid_c, grid_s;
  OdomGraphCache graphs;
};

// DynamicFeatureMap paging (cm_mapio.cu): the index2.txt catalogue of one stream and the window that is resident
struct PageState {
  bool open = false, first = true;
  std::string dir;
  std::map<std::array<int, 3>, int> files[2];   // global cube index -> file number, [0] corner / [1] surf
  int win[3] = {21, 11, 21};                    // DynamicFeatureMap(cubeWidth_, cubeHeight_, cubeDepth_)
  int sensor[3] = {0, 0, 0};                    // _sensorGloId
};

// ---- multi-GPU state (cm_dist.cu) ------------------------------------------------------------------------------------------
#define CM_DIST_MAX_RANKS 16
#define CM_DIST_NMAX 8192            // doubles per exchange (256 streams x 32)
struct DistDev {                     // what the exchange kernel sees
  int rank, nranks, nmax, p2p;
  double* mbox_peer[CM_DIST_MAX_RANKS];               // every rank's mailbox, mapped into this process (CUDA IPC); [rank] = mine
  unsigned long long* flags_peer[CM_DIST_MAX_RANKS];  // their sequence numbers
  const double* gathered;                             // all-gather transport: [nranks][n] (no IPC)
};
struct StagePool;
struct DistState {
  bool on = false;
  void* comm = nullptr;              // ncclComm_t
  void* mailbox = nullptr;           // this rank's mailbox (cudaMalloc, exported with cudaIpcGetMemHandle)
  void* peer_ptr[CM_DIST_MAX_RANKS] = {};
  DistDev dev;
  unsigned long long seq = 0;        // exchanges so far (all ranks advance together)
  DeviceBuffer err, gathered, scratch;
};

}  // namespace cm

struct cm_ctx {
  cm_config cfg;
  cudaStream_t stream = nullptr;
  int prio_high = 0, prio_low = 0;   // stream priorities: main / aux streams high, prefetch (side) stream low
  std::string err;
  // scratch for the host-buffer entry points
  cm::DeviceBuffer d_ref_corner, d_ref_surf, d_corner, d_surf, d_q, d_idx, d_d2;
  cm::DeviceBuffer d_counts, d_views, d_pose, d_state, d_rows, d_slots, d_sums, d_trace, d_nn;
  cm::GridStorage grid_a, grid_b;
  cm::VoxelFilter voxel;
  // aux_stream: the corner-class half of the mapping stage's voxel filters and map insertion runs here, concurrently with the
  // surf-class half on `stream` (both are chains of small latency-bound kernels; the corner chain hides behind the surf chain)
  cudaStream_t aux_stream = nullptr; cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  cm::ScanRegistrationGpu scanreg;
  cm::DeviceBuffer d_tags, d_l_ds[4], d_sweep;
  cm::SweepFrontEnd frontend;          // raw-sweep front end (cm_scanreg_sweep_host)
  cm::ImuHistoryHost imu;              // cm_imu_push_host
  cm::DeviceBuffer d_frames, d_sr_pts[4], d_sr_idx[4], d_sr_n, d_sr_cloud, d_sr_ccurv, d_sr_picked, d_sr_curv, d_sr_label, d_sr_range;
  cm::DeviceBuffer d_vin, d_vout, d_vn_in, d_vn_out, d_flag;
  // mapping stage (cm_mapping.cu)
  int map_streams = 0;
  cm::DeviceMap map;
  std::vector<cm::MappingStream> mstreams;
  cm::DeviceBuffer m_corner_in, m_surf_in, m_n_in, m_corner_ds, m_surf_ds, m_n_ds, m_pose, m_state, m_rows, m_slots, m_sums, m_tf, m_exp_pts, m_exp_cube, m_exp_n;
  int m_cap_corner = 0, m_cap_surf = 0;
  cm::DeviceBuffer dbg_trace; bool dbg_on = false; int dbg_iter = 0; size_t dbg_words = 0;   // cm_debug_search_trace
  // Gauss-Newton stream groups of the batched mapping stage (cm_mapping.cu)
#define CM_MAX_GN_GROUPS 8
  cudaStream_t gn_stream[CM_MAX_GN_GROUPS] = {}; cudaEvent_t gn_join[CM_MAX_GN_GROUPS] = {}; cudaEvent_t gn_fork = nullptr;
  cm::GraphCache stage_graphs;                      // voxel-filter and map-insert chains of the mapping stage
  cm::MatchGraphCache match_graphs;                 // CUDA graphs of the mapping stage's Gauss-Newton loop
  cm::HardQueue hardq;                              // deferred hard 5-NN queries of the current match (cm_match.cu)
  // scan-to-scan odometry (cm_odometry.cu): LaserOdometry's members
  bool odom_inited = false;
  float odom_tf[6] = {0, 0, 0, 0, 0, 0};           // _transform (persists across frames, LaserOdometry.cpp never resets it)
  cm::HostIso odom_Tsum;                            // _Tsum
  int o_n_last_c = 0, o_n_last_s = 0;
  cm::DeviceBuffer o_last_c, o_last_s, o_sharp, o_flat, o_ind, o_rows, o_state, o_sums, o_pose, o_trace, o_counts, o_views, o_tf, o_inv, o_slots;
  cm::GridStorage o_grid_c, o_grid_s;
  // sharded-map matching (cm_shard_*): persistent grids in grid_a / grid_b
  cm::MatchLaunch shard; size_t shard_nq = 0; bool shard_ready = false;
  cm::DeviceBuffer d_box;
  cm::OdomBatch obatch;                // cm_odometry_batch_*
  int chain_rows = 0, chain_cols = 0;  // cm_pipeline_chain_create: the sweep shape the odometry batch was sized for
  cm::LocalWindow local;               // cm_mapping_local_*
  std::vector<cm::PageState> pages;    // cm_map_page_* (one per stream)
  cm::DeviceBuffer d_drop;
  cm::DistState dist;                  // cm_dist_init: this context is one rank of a sharded map
  // pinned, device-accessible host staging for the per-step parameter uploads of the mapping stage (poses, cube windows): they
  // are copied by a kernel, not by the copy engine that the sweep uploads keep busy
  void* h_stage = nullptr; size_t h_stage_cap = 0;
  // pinned landing area of a mapping step's results (states, filtered counts, flags) + the event behind their copy: the step
  // returns when the poses are in, the map insertion enqueued behind them finishes on its own
  void* h_result = nullptr; size_t h_result_cap = 0; cudaEvent_t result_ready = nullptr;
  std::vector<unsigned char> wins_shadow;   // the cube windows last sent to the device (they rarely change from sweep to sweep)
  cm::KernelProfiler prof, prof_sr;   // search_kernel + search_hard_kernel / sr_ring_kernel launches of the pipeline
  cudaEvent_t timer[2] = {nullptr, nullptr};
  unsigned long long dbg_graph_builds = 0, dbg_stage_captures = 0;
  int est_c = 0, est_s = 0;                 // largest filtered corner / surf cloud of the previous mapping step (sizes the next step's launches)
  unsigned long long n_shifts = 0;          // FeatureMap::shift calls that moved cubes so far
  unsigned long long insert_redos = 0;      // steps whose map insertion had to be repeated with exact sizes
  // per-step counters of the last cm_mapping_process / cm_pipeline_step (for the roofline arithmetic)
  unsigned long long last_query_iters = 0, last_queries = 0, last_inserted = 0, last_features = 0;
  // pipeline (scan registration -> mapping), cm_mapping.cu
  // Slot 2 is the synchronous path; slots 0 / 1 are filled ahead of time by cm_pipeline_prefetch_host / _dev: the NEXT
  // step's sweeps are uploaded (host variant) and run through scan registration on side_stream while the current step's
  // matching and map kernels run on `stream`.  Scan registration is issue-bound, matching is latency-bound: the two overlap.
  struct PipeSlot {
    cm::DeviceBuffer frames, pts[4], n;
    cm::ScanRegistrationGpu scanreg;
    const void* src = nullptr; int rows = 0, cols = 0; bool is_host = false;   // what was prefetched (NULL: free)
    cudaEvent_t done = nullptr, copied = nullptr, copied2 = nullptr, copied_x[2] = {nullptr, nullptr};
    size_t frames_valid = 0;
    // strided / pageable sweeps (cm_stage.cu): packed xyz staging, pinned on the host and its device copy
    cm::DeviceBuffer frames_xyz; void* h_xyz = nullptr; size_t h_xyz_cap = 0;
    // feature counts read back by the prefetch itself (side stream -> pinned host memory): the step that
    // consumes the slot starts without a host round trip
    int* h_n5 = nullptr; int h_streams = 0; bool counts_ready = false;
  };
#define CM_PIPE_SLOTS 4            // prefetch slots (one being consumed + three pending); pipe[CM_PIPE_SLOTS] is the synchronous path
  PipeSlot pipe[CM_PIPE_SLOTS + 1];
  // scan registration ahead of time / sweep upload.  The upload is split over TWO copy streams: one host-to-device stream
  // reaches 36.6 GB/s on the B200 boxes measured, two concurrent ones 53.6 GB/s (tools/h2d_bandwidth.py)
  cudaStream_t side_stream = nullptr, copy_stream = nullptr, copy_stream2 = nullptr;
  cudaStream_t copy_stream_x[2] = {nullptr, nullptr};   // optional third / fourth copy stream (COOPERMAP_COPY_STREAMS)
  int p_cap = 0;
  // a prefetch registered with cm_pipeline_prefetch_deferred_*: issued by the next cm_pipeline_step right after it has submitted
  // its Gauss-Newton loop, so that the host time of the submission hides behind device work
  cm::StagePool* stage_pool = nullptr;   // worker threads that repack strided host sweeps (cm_stage.cu)
  const void* defer_frames = nullptr; int defer_rows = 0, defer_cols = 0; bool defer_is_host = false; int defer_rc = 0;
};

namespace cm {
MatchParamsDev dev_params(const cm_config& c);
int ctx_fail(cm_ctx* ctx, int code, const std::string& msg);
void fill_scanreg_params(const cm_config& c, ScanRegLaunch& L);
void fill_match_stats(const cm_config& cfg, const MatchState& st, size_t nq, cm_match_stats* out);
int dist_allreduce(cm_ctx* ctx, double* d_vec, int n, cudaStream_t stream);   // in-place sum over the ranks (cm_dist.cu)
void dist_destroy(cm_ctx* ctx);
void stage_pool_destroy(cm_ctx* ctx);
int odometry_batch_core(cm_ctx* ctx, const void* sharp, size_t pitch_sharp, const int* n_sharp, const void* less_sharp, size_t pitch_less_sharp,
                        const int* n_less_sharp, const void* flat, size_t pitch_flat, const int* n_flat, const void* less_flat,
                        size_t pitch_less_flat, const int* n_less_flat, cm_iso* odom, cm_pose* transform, cm_point* corner_last,
                        cm_point* surf_last, cm_odom_stats* stats);   // cm_odometry.cu
int stage_upload_strided(cm_ctx* ctx, cm_ctx::PipeSlot& slot, const void* const* clouds, size_t stride, int rows, int cols, cudaStream_t consumer);
}  // namespace cm

extern "C" int cm_match_stateless_dev(cm_ctx* ctx, const float4* d_rc, size_t nrc, const float4* d_rs, size_t nrs, const float4* d_c, size_t nc,
                                      const float4* d_s, size_t ns, cm_pose* pose, cm_match_stats* stats, cm_iter_trace* trace,
                                      int* nn_corner, int* nn_surf);

#define CM_CUDA_CHECK(ctx, expr)                                                                                    \
  do {                                                                                                              \
    cudaError_t e__ = (expr);                                                                                       \
    if (e__ != cudaSuccess) return cm::ctx_fail(ctx, CM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)
