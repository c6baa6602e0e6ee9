// cm_host.h -- host-side plumbing shared by the .cu translation units (device buffers, launch descriptors).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <map>
#include <string>
#include <vector>
#include "cm_match.cuh"

namespace cm {

struct CudaError { cudaError_t code; const char* what; };

// Every kernel launch goes through CM_LAUNCH so that cm_launch_count() reports what actually ran.  With the
// timeline enabled (cm_timeline_enable) every launch is additionally bracketed by CUDA events on its stream and
// cm_timeline_report() aggregates the device time per kernel name (development aid: warm caches, real clocks).
extern unsigned long long g_launch_count;
struct Timeline {
  bool on = false;
  struct Rec { const char* name; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() { if (pool.empty()) { cudaEvent_t e; cudaEventCreate(&e); return e; } cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
  void begin(const char* name, cudaStream_t s) { Rec r; r.name = name; r.a = get(); r.b = get(); cudaEventRecord(r.a, s); recs.push_back(r); }
  void end(cudaStream_t s) { cudaEventRecord(recs.back().b, s); }
};
extern Timeline g_timeline;
#define CM_LAUNCH(kern, grid, block, smem, stream, ...)                 \
  do {                                                                  \
    if (cm::g_timeline.on) cm::g_timeline.begin(#kern, (stream));       \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);           \
    if (cm::g_timeline.on) cm::g_timeline.end((stream));                \
    ++cm::g_launch_count;                                               \
  } while (0)
#define CM_TIMED(name, stream, stmt)                                    \
  do {                                                                  \
    if (cm::g_timeline.on) cm::g_timeline.begin(name, (stream));        \
    stmt;                                                               \
    if (cm::g_timeline.on) cm::g_timeline.end((stream));                \
  } while (0)

// Grow-only device allocation.  g_alloc_generation counts the frees: a CUDA graph that captured a DeviceBuffer pointer is only
// valid while the generation it was captured under is still current (GraphCache / MatchGraphCache check it).
extern unsigned long long g_alloc_generation;
struct DeviceBuffer {
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) { cudaFree(p); ++g_alloc_generation; }
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) throw CudaError{e, "cudaMalloc"};
    cap = want;
  }
  void release() { if (p) { cudaFree(p); ++g_alloc_generation; } p = nullptr; cap = 0; }
  ~DeviceBuffer() { release(); }
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
};

// Capture-once / replay-many helper for fixed launch sequences (a chain of small kernels whose arguments repeat from step to
// step).  The first time a key is seen the sequence runs normally (grow-only buffers reach their size: cudaMalloc cannot be
// captured), the second time it is captured into a CUDA graph, afterwards it is ONE submission.
// The key holds the caller-visible arguments only; the scratch pointers a sequence bakes in (VoxelFilter / DeviceMap::insert
// buffers) are covered by the allocation generation: an entry captured before ANY DeviceBuffer was re-allocated is dropped and
// re-captured, so a graph can never replay a freed pointer.
struct GraphCache {
  struct Entry { std::vector<unsigned long long> key; cudaGraphExec_t exec; unsigned long long launches; unsigned long long gen; };
  std::vector<Entry> entries;
  unsigned long long captures = 0;   // graphs instantiated so far (tests)
  template <class F>
  void run(const std::vector<unsigned long long>& key, cudaStream_t stream, F&& issue) {
    for (Entry& e : entries)
      if (e.key == key) {
        if (e.exec && e.gen != g_alloc_generation) {   // some buffer moved since the capture: run plainly, capture again next time
          cudaGraphExecDestroy(e.exec); e.exec = nullptr;
          issue();
          e.gen = g_alloc_generation;
          return;
        }
        if (!e.exec && e.gen != g_alloc_generation) {  // buffers were still growing when this key was first seen
          issue();
          e.gen = g_alloc_generation;
          return;
        }
        if (e.exec) {
          if (cudaGraphLaunch(e.exec, stream) == cudaSuccess) { g_launch_count += e.launches; return; }
          cudaGetLastError();
          issue();
          return;
        }
        // seen once: capture now
        const unsigned long long before = g_launch_count;
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); issue(); return; }
        issue();
        cudaGraph_t graph = nullptr;
        const bool ok = cudaStreamEndCapture(stream, &graph) == cudaSuccess && graph;
        e.launches = g_launch_count - before;
        g_launch_count = before;
        if (ok && e.gen == g_alloc_generation && cudaGraphInstantiate(&e.exec, graph, 0) == cudaSuccess && cudaGraphLaunch(e.exec, stream) == cudaSuccess) {
          cudaGraphDestroy(graph);
          g_launch_count += e.launches;
          ++captures;
          return;
        }
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        if (e.exec) { cudaGraphExecDestroy(e.exec); e.exec = nullptr; }
        if (e.gen != g_alloc_generation) e.gen = g_alloc_generation;   // a buffer grew during the capture: try again next time
        else e.key.clear();                                           // never try this key again
        issue();
        return;
      }
    if (entries.size() >= 64) clear();
    issue();
    Entry e; e.key = key; e.exec = nullptr; e.launches = 0; e.gen = g_alloc_generation;
    entries.push_back(e);
  }
  void clear() { for (Entry& e : entries) if (e.exec) cudaGraphExecDestroy(e.exec); entries.clear(); }
  ~GraphCache() { clear(); }
};

// Storage behind one GridView built from a caller-supplied cloud (stateless matching / knn).
struct GridStorage {
  DeviceBuffer entries, pts, cell_of, cursor;
  GridView view;
  // keep_w = 0: pts[].w := original index (tie-break + reported neighbour index); 1: keep the input w
  void build(const float4* d_pts, int n, float cell_size, float gate, int keep_w, cudaStream_t stream);
};

int grid_max_level(float cell, float gate);
void launch_knn5(const GridView& g, const float* d_q, int nq, float gate, int* d_idx, float* d_d2, cudaStream_t stream);

#define CM_MAX_EVALS 32          // upper bound of Gauss-Newton evaluations per call (max_iterations <= 32)
#define CM_HARD_ITEM_BYTES 72
// scratch of one MatchLaunch: the deferred-query list and the per-CTA partial sums / tickets of the fused fit + solve
struct HardQueue {
  DeviceBuffer items, count, partials, tickets;
  void attach(struct MatchLaunch& m, size_t capacity);
};
struct MatchLaunch {
  int nstreams;
  const float4* corner; const float4* surf;   // [nstreams][cap]
  const int* n_corner; const int* n_surf;     // device, per stream
  int cap_corner, cap_surf;
  const GridView* grid_corner; const GridView* grid_surf;   // device arrays [nstreams]
  const float* pose_in;                       // device [nstreams][6]
  MatchState* state;                          // device [nstreams]
  RowOut* rows;                               // device [nstreams][cap_corner + cap_surf]
  int* nn_slot;                               // device [nstreams][cap_corner + cap_surf][5]
  double* sums;                               // device [nstreams][32] (A^T A / A^T b partial sums)
  IterTrace* trace;                           // optional device [nstreams][max_iterations]
  int* nn;                                    // optional device [max_iterations][nstreams][cap][5]
  int orig_idx;                               // grids carry original indices in pts[].w
  int max_queries = 0;                        // what the grids are sized for: n_corner[s] + n_surf[s] of the largest stream, exact or an
                                              // ESTIMATE (the kernels loop when a stream has more); 0: use the capacities
  int dist_rank = 0, dist_nranks = 1;         // sharded map: evaluate only the queries whose map-frame cube this rank owns
  const int* skip = nullptr;                  // optional device flag: non-zero = the grids above were too small for some stream (estimate
                                              // missed): every kernel of the match returns at once and the caller repeats it with exact sizes
  int bound_queries = 0;                      // host-known UPPER BOUND of n_corner[s] + n_surf[s]: sizes the scratch (0: max_queries)
  const float* own_box = nullptr;             // device {lo[3], hi[3]}: evaluate only queries inside (sharded map), else all
  void* hard = nullptr;                       // optional device list of deferred "hard" queries (hard_cap * CM_HARD_ITEM_BYTES)
  int* hard_count = nullptr;                  // device [CM_MAX_EVALS] counters, one per Gauss-Newton evaluation
  int hard_cap = 0, hard_blocks = 0;
  double* partials = nullptr; int* tickets = nullptr; int partial_blocks = 0;   // fused fit + reduce + solve (launch_match only)
  unsigned long long* dbg = nullptr; int dbg_iter = 0;   // per-warp trace of search_kernel in evaluation dbg_iter (development aid)
  MatchParamsDev prm;
};
// optional per-kernel timing of the dominant kernel (corr_kernel): event pairs recorded on the launching stream
struct KernelProfiler {
  bool enabled = false;
  std::vector<cudaEvent_t> ev;   // pairs
  size_t used = 0;
  void begin(cudaStream_t s) { if (!enabled) return; grow(); cudaEventRecord(ev[used], s); }
  void end(cudaStream_t s) { if (!enabled) return; cudaEventRecord(ev[used + 1], s); used += 2; }
  void grow() { while (ev.size() < used + 2) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); } }
  // sums the recorded intervals (call after the stream is synchronized) and resets
  double drain_ms(int* launches) {
    double tot = 0; int n = 0;
    for (size_t i = 0; i + 1 < used; i += 2) { float ms = 0; if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) { tot += ms; n++; } }
    used = 0; if (launches) *launches = n; return tot;
  }
};
inline void HardQueue::attach(MatchLaunch& m, size_t capacity) {
  if (capacity < 1) capacity = 1;
  items.reserve(capacity * CM_HARD_ITEM_BYTES); count.reserve(sizeof(int) * CM_MAX_EVALS * 8);   // x8: one counter row per stream group
  m.hard = items.p; m.hard_count = (int*)count.p; m.hard_cap = (int)capacity;
  const int maxq = m.bound_queries > 0 ? m.bound_queries : (m.max_queries > 0 ? m.max_queries : m.cap_corner + m.cap_surf);
  m.partial_blocks = (maxq + 32 + 255) / 256;
  partials.reserve((size_t)m.nstreams * m.partial_blocks * 32 * sizeof(double)); tickets.reserve(sizeof(int) * m.nstreams);
  m.partials = (double*)partials.p; m.tickets = (int*)tickets.p;
}
void launch_match(const MatchLaunch& m, cudaStream_t stream, KernelProfiler* prof = nullptr);
// The same launch sequence replayed from a CUDA graph: the <= 10 x (search, hard-search, fit+solve) kernels of one match are
// captured once per distinct argument set (buffers are grow-only, so the set is stable after the first steps) and then
// submitted with ONE call -- the Gauss-Newton loop no longer pays a host launch per kernel, which matters most while the
// PCIe link is busy with the next sweeps (tools/pcie_interference.py: +25 % step time with per-kernel launches).
struct MatchGraphCache {
  struct Entry { std::vector<unsigned long long> key; cudaGraphExec_t exec; unsigned long long launches; unsigned long long gen; };
  unsigned long long builds = 0;   // graphs built so far (tests: stays small over many frames)
  std::vector<Entry> entries;
  // use_while: build the loop as a conditional WHILE node (CUDA 12.4+) whose body reads the evaluation index from device memory
  // and whose last node (gn_advance_kernel) calls cudaGraphSetConditional -- the graph then runs exactly as many evaluations as
  // the slowest stream needs instead of max_iterations mostly empty ones; falls back to the unrolled graph when unavailable
  bool use_while = true;
  int* d_iter = nullptr;
  bool launch(const MatchLaunch& m, cudaStream_t stream);   // false: capture failed, nothing was launched
  void clear();
  ~MatchGraphCache();
};
// the same with the streams split into `ngroups` groups whose iteration loops run concurrently on gs[0..ngroups)
void launch_match_groups(const MatchLaunch& m, cudaStream_t stream, int ngroups, cudaStream_t* gs, cudaEvent_t fork, cudaEvent_t* join,
                         KernelProfiler* prof = nullptr);
void launch_match_init(const MatchLaunch& m, cudaStream_t stream);
// defer_solve (fused only): stop after the per-stream sums -- the caller exchanges them between ranks and calls launch_match_solve_warp
void launch_match_partial(const MatchLaunch& m, int it, cudaStream_t stream, KernelProfiler* prof = nullptr, bool fused = false, bool defer_solve = false);
void launch_match_solve_warp(const MatchLaunch& m, int it, cudaStream_t stream);
void launch_match_solve(const MatchLaunch& m, int it, const double* sums, cudaStream_t stream);
void launch_match_reduce(const MatchLaunch& m, int it, cudaStream_t stream);
void launch_match_reduce_solve(const MatchLaunch& m, int it, cudaStream_t stream, const int* d_iter = nullptr);

// K8: scan-to-scan odometry (cm_odom.inl)
struct OdomLaunch {
  const float4* sharp; const float4* flat; int n_sharp, n_flat;
  const float4* last_corner; const float4* last_surf; int bound_corner, bound_surf;
  GridView grid_corner, grid_surf;
  const MatchState* state; int* ind; RowOut* rows;
};
void launch_odom_corr(const OdomLaunch& o, int iter, cudaStream_t stream);
// batched odometry (one launch set for all streams of a context)
struct GridBatch {            // nstreams voxel-cell grids of equal capacity, rebuilt together (streams with on[s] == 0 keep theirs)
  DeviceBuffer entries, pts, cell_of, cursor, views;
  int nstreams = 0, cap = 0; unsigned int tcap = 0;
  void create(int nstreams, int cap, cudaStream_t stream);
  void build(const float4* d_pts, const int* d_n, int max_n, const int* d_on, float cell_size, float gate, cudaStream_t stream);
};
struct OdomBatchLaunch {
  int nstreams;
  const float4* sharp; const float4* flat; int cap_sharp, cap_flat; const int* n_sharp; const int* n_flat; int max_sharp, max_flat;
  const float4* last_corner; const float4* last_surf; int cap_last_corner, cap_last_surf; const int* bound_corner; const int* bound_surf;
  const GridView* grid_corner; const GridView* grid_surf;
  const MatchState* state; int* ind; RowOut* rows;
  const void* box_corner = nullptr; const void* box_surf = nullptr; int box_cap_corner = 0, box_cap_surf = 0;   // chunk boxes (cm_odom.inl), optional
};
void launch_odom_boxes_batch(const float4* d_cloud, int cap, const int* d_n, int max_n, int nstreams, void* d_boxes, int box_cap, cudaStream_t stream);
void launch_odom_corr_batch(const OdomBatchLaunch& o, int iter, cudaStream_t stream, const int* d_iter = nullptr);
// The batch odometry's Gauss-Newton loop as ONE submission: init -> WHILE { correspondences + rows, reduction, 6x6 step, advance },
// as many evaluations as the slowest stream needs (25 x 3 launches otherwise, most of them empty once the streams have converged).
struct OdomGraphCache {
  struct Entry { std::vector<unsigned long long> key; cudaGraphExec_t exec; unsigned long long launches; unsigned long long gen; };
  std::vector<Entry> entries;
  bool usable = true;
  int* d_iter = nullptr;
  bool launch(const MatchLaunch& m, const OdomBatchLaunch& o, const int* d_active, cudaStream_t stream);   // false: nothing was launched
  void clear();
  ~OdomGraphCache();
};
void launch_odom_gate(MatchState* d_state, const int* d_active, int nstreams, cudaStream_t stream);
void launch_odom_to_end_batch(float4* d_cloud, int cap, const int* d_n, int max_n, int nstreams, const float* d_tf6, const float* d_inv12,
                              const int* d_on, cudaStream_t stream);
void launch_odom_to_end(float4* d_cloud, int n, const float* d_tf6, const float* d_inv12, cudaStream_t stream);

// K3: batched pcl::VoxelGrid-equivalent filter (cm_voxel.cu).  Segment s reads in[s*cap_in .. +n_in[s]) and writes
// out[s*cap_out .. +n_out[s]).
struct VoxBox {          // per segment: PCL's min_b_ / divb_mul_ of the cloud's bounding box
  int minb[3];
  int mul1, mul2;        // divb_mul_[1], divb_mul_[2]
  int passthrough;       // index overflow: copy input to output
  int nfinite;
  long long cells;       // dx * dy * dz: the voxel index of the segment is < cells
};
int vox_index_bits(long long cells);
struct VoxelFilter {
  DeviceBuffer scratch;
  // max_n: host-known upper bound of n_in[s] (<= 0: cap_in); it only sizes the scratch, the counts themselves stay on the device
  void run(int nseg, const float4* d_in, const int* d_n_in, int cap_in, int max_n, float leaf, float4* d_out, int* d_n_out, int cap_out,
           int* d_overflow, cudaStream_t stream);
  // two batches with different leaves (the corner and the surf clouds of every stream) in ONE launch
  void run2(int nseg, const float4* d_in0, const int* d_n_in0, int cap_in0, float leaf0, float4* d_out0, int* d_n_out0, int cap_out0,
            const float4* d_in1, const int* d_n_in1, int cap_in1, float leaf1, float4* d_out1, int* d_n_out1, int cap_out1,
            int max_n, int* d_overflow, cudaStream_t stream);
};

void launch_step_guard(const int* d_n2, int nstreams, int max_c, int max_s, int* d_flag, cudaStream_t stream);

// Small parameter uploads without the copy engine: `pinned` is device-accessible pinned host memory, a kernel reads it over
// PCIe and writes `d_dst`.  (cm_map.cu)
void staged_upload(void* d_dst, const void* pinned, size_t bytes, cudaStream_t stream);

// K7: device-resident local map (cm_map.cu)
struct MapClassDev {            // one class (corner / surf) of one stream, lives in device memory
  CellEntry* entries; unsigned int* cellcap; unsigned int* pending; unsigned int mask;
  float4* pts; unsigned int pool_cap; unsigned int* cursor; int* total;
  int* cube_count;              // points per 50 m cube [W*H*D]
  float leaf, inv_leaf; int kdiv;
  float cube_size; int dims[3]; int origin[3];
  // FeatureMap::shift bookkeeping (cm_map.cu): epoch per pool slot, displacement of every epoch [256][3], the epoch new points get
  unsigned char* epoch; int* eoff; int cur_epoch;
  // cubes that received points while they were outside the valid window: their clouds are unfiltered until they become valid again
  // (FeatureMap::downsizeValidCloud only visits valid cubes, FeatureMap.h:289-306)
  unsigned char* dirty;         // [W*H*D]
};
struct MapConfig {
  size_t max_corner, max_surf;  // capacity in points per stream
  float leaf_corner, leaf_surf; // map_filter_corner / map_filter_surf
  int kdiv_corner, kdiv_surf;   // search cell = kdiv voxels
  float cube_size; int dims[3]; int origin[3];
};
struct DeviceMap {
  int nstreams = 0;
  MapConfig cfg;
  DeviceBuffer entries[2], cellcap[2], pending_cnt[2], pts[2], cube_count[2], cursor[2], dev[2], views[2];
  DeviceBuffer windows, flags;
  DeviceBuffer epoch[2], eoff;                    // see MapClassDev
  DeviceBuffer cube_dirty[2], need[2];            // unfiltered cubes per stream; per-stream "a dirty cube is valid now" flags of an insert
  bool windows_valid = false;                     // set_windows has run at least once (before that every cube counts as valid)
  std::vector<MapClassDev> hdev[2];               // host mirror of dev[] (pointers and constants; origin / cur_epoch as of the last shift)
  std::vector<int> h_eoff;                        // [nstreams][256][3]
  std::vector<int> cur_epoch;                     // per stream
  DeviceBuffer n_pending[2], world[2], keys_a[2], keys_b[2], vals_a[2], vals_b[2], pending[2];   // insert scratch per class: the two classes may run on different streams
  unsigned int table_cap[2] = {0, 0}, pool_cap[2] = {0, 0};
  int shard_rank = 0, shard_nranks = 1;   // > 1 ranks: keep only the cubes cube_owner() gives this rank, plus a sqrt(5) m halo (cm_dist.cu)
  void create(int nstreams, const MapConfig& c, cudaStream_t stream);
  // also refreshes the GridViews.  staged: h_windows is pinned, device-accessible host memory -> copied by a kernel instead of
  // a host-to-device memcpy (a small memcpy queues behind the sweep uploads that keep the copy engine busy)
  // h_windows == NULL: the windows on the device are still current (only the views are refreshed)
  void set_windows(const CubeWindow* h_windows, float gate, cudaStream_t stream, bool staged = false);
  // sharded map: views[].npts (points of the active cubes THIS rank owns) <-> a vector of 2 * nstreams doubles for the exchange
  void pack_npts(double* d_vec, cudaStream_t stream);
  void unpack_npts(const double* d_vec, cudaStream_t stream);
  // transform by the per-stream pose in d_state (or by d_tf: [S][12] = R row-major + t) and merge into the map
  // max_n: what the launches are sized for (an estimate is fine: if a stream has more, the insert does nothing and sets flags[4 + cls]);
  // step_skip: optional device flag, non-zero = insert nothing (the caller is about to repeat the step)
  // filter_all: every cube is filtered (FeatureMap::loadCloudFromFiles filters each file, :428-456); otherwise the reference's
  // addFeatureCloud: points pushed into a cube outside the valid window stay unfiltered until the cube is valid during an insert
  void insert(int cls, const float4* d_pts, const int* d_n, int cap, int max_n, const MatchState* d_state, const float* d_tf, cudaStream_t stream,
              const int* step_skip = nullptr, bool filter_all = false);
  size_t export_points(int cls, int s, float4* d_out, int* d_cube, unsigned int* d_n, unsigned int cap, cudaStream_t stream);
  // FeatureMap::shift(d) of stream s followed by `origin += d` (FeatureMap.h:232-245, 354-376): relabels the stored points, drops
  // the cubes that leave the grid, recounts cube_count.  new_origin = the stream's origin after the shift.  False: more than 255
  // wrong-way shifts (the epoch counter is a byte)
  bool shift(int s, const int d[3], const int new_origin[3], cudaStream_t stream, bool literal = true, const unsigned char* d_drop = nullptr);
};

// K1/K2: scan registration for organised sweeps (cm_scanreg.cu)
struct ScanRegLaunch {
  KernelProfiler* prof;                       // optional: event pair around sr_ring_kernel (bench.py's per-kernel timing)
  int nstreams, rows, cols;
  const float4* frames;                       // device [S][rows][cols]
  const float* tags;                          // optional device [S][rows][cols] (raw-sweep front end), else NULL
  float blind_sq_override;                    // >= 0: use instead of blind_radius^2 (the sweep front end filtered already)
  float scan_period, blind_radius, blind_thr, curv_thr, less_flat_leaf;
  int R, nregions, max_sharp, max_flat;
  double cos175, cos5, cos135, cos45;
  float4* out_pts[4]; int cap[4];             // device [S][cap[k]]: sharp, lessSharp, flat, lessFlat
  int* out_n;                                 // device [S][5]
  int* overflow;                              // device flag (optional)
  int want_idx; int* out_idx[4];              // device [S][rows*cols] (parity / diagnostics)
  float4* cloud; float* cloud_curv;           // optional device [S][rows*cols]
  signed char* picked; float* curvature; signed char* label;   // optional device [S][rows*cols]
  int* scan_range;                            // optional device [S][rows][2]
};
struct ScanRegistrationGpu {
  DeviceBuffer ring_count, ring_n, ring_pts[4], ring_idx[4];
  void run(const ScanRegLaunch& L, cudaStream_t stream);
};
size_t scanreg_smem_bytes(int cols);

int debug_math_dims(int op, int* nin, int* nout);
void launch_debug_math(int op, const float* d_in, int nin, float* d_out, int nout, int n, cudaStream_t stream);

// Raw-sweep front end on the device (cm_frontend.cu): MultiScanRegistration::process up to the per-ring clouds.
// ScanRegistration's IMU history (a CircularBuffer of imuHistorySize states, ScanRegistration.cpp:53,89-121), oldest first
struct ImuHistoryHost {
  std::vector<double> stamp;      // seconds
  std::vector<float> state;       // [n][9]: roll, pitch, yaw, position, velocity
  size_t capacity = 200;
  void push(double stamp, double roll, double pitch, double yaw, double ax, double ay, double az);   // handleIMUMessage
  void clear() { stamp.clear(); state.clear(); }
};
struct SweepFrontEnd {
  DeviceBuffer ring_of, hist, frame, tags, rel, imu_buf;
  // imu (optional, non-empty): de-skew every point to the sweep start (hasIMUData() branch); imu_trans12: the /imu_trans points
  void run(const float4* d_sweep, int n, const float4& first, const float4& last, int lidar, float scan_period, cudaStream_t st,
           int* rows_out, int* cols_out, const ImuHistoryHost* imu = nullptr, double scan_time = 0.0, float* imu_trans12 = nullptr);
};
bool frontend_mapper(int lidar, float* lower, float* upper, int* nrings);

}  // namespace cm
