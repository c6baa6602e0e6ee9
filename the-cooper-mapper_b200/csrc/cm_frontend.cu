// cm_frontend.cu -- the raw-sweep front end of scan registration ON THE DEVICE.
//
// Replaces the per-point loop of MultiScanRegistration::process (MultiScanRegistration.cpp:95-190): axis swap, validity, ring
// from the elevation angle (linear mappers MultiScanRegistration.h:57-102, Pandar40 lidar_type.h:78-104), azimuth with the
// half-sweep unwrap, relTime, and the stable per-ring append (`laserCloudScans[scanID].push_back(point)`), for ONE unordered
// azimuth-major sweep.  The reference's loop is sequential in two places; both have a parallel form:
//  * `halfPassed` flips once, at the first ACCEPTED point whose first-half orientation exceeds startOri + pi; every point up to
//    and including that one uses the first-half rule, every later point the second-half rule -> an atomicMin over the indices;
//  * a point's position inside its ring is the number of earlier accepted points of the same ring -> per-CTA ring histograms,
//    a scan over the CTAs, `match_any` ranks inside a warp.
// atan / atan2 are cm_atanf / cm_atan2f (cm_math.h), the same functions the oracle calls, so the device result is bit-identical
// to the host restatement.  Output: ring-major rows [nRings][cols] (NaN padded) + the curvature field ring + relTime per slot,
// the layout sr_ring_kernel consumes with ScanRegArgs::tags.
#include "cm_host.h"
#include "cm_math.h"
#include <float.h>
#include <vector>

namespace cm {

#define FE_THREADS 256
#define FE_CHUNKS 4                       // a CTA owns FE_THREADS * FE_CHUNKS consecutive points of the sweep
#define FE_MAXRINGS 64

struct FrontEndArgs {
  const float4* sweep; int n;
  int lidar, nrings; float lower, factor, scan_period, start_ori, end_ori;
  int* ring_of;          // [n] ring id, -1: dropped
  int* first_half;       // [1] index of the accepted point that sets halfPassed (n: never)
  int* block_hist;       // [nblocks][nrings]; after fe_scan_kernel: exclusive prefix over the CTAs
  int* ring_total;       // [nrings]
  float4* frame; float* tags; int cols;   // outputs of fe_place_kernel
  // IMU de-skew (ScanRegistration.cpp:123-188), nimu == 0: none
  int nimu;
  const double* imu_dt;      // [nimu] scanTime - stamp[j]
  const double* imu_dstamp;  // [nimu] stamp[j] - stamp[j - 1] (j >= 1)
  const float* imu_state;    // [nimu][9] roll, pitch, yaw, position, velocity
  float imu_start[9];        // _imuStart
  float* rel; float* bmax;   // [n] relTime of the accepted points (-FLT_MAX: dropped), [nblocks] its maximum per CTA
  int* last_idx;             // [1] index of the last accepted point
  float* last_out;           // [9] _imuCur (roll, pitch, yaw), _imuPositionShift, _imuCur.velocity of that point
};

// Angle.h: radian + buffered cos / sin; rotations of math_utils.h:115-236 on a float[3]
struct FeAng { float c, s; };
__host__ __device__ __forceinline__ FeAng fe_ang(float r) { FeAng a; cm_sincosf(r, &a.s, &a.c); return a; }
__host__ __device__ __forceinline__ FeAng fe_neg(FeAng a) { a.s = -a.s; return a; }
__host__ __device__ __forceinline__ void fe_rotX(float v[3], FeAng a) { const float y = v[1]; v[1] = a.c * y - a.s * v[2]; v[2] = a.s * y + a.c * v[2]; }
__host__ __device__ __forceinline__ void fe_rotY(float v[3], FeAng a) { const float x = v[0]; v[0] = a.c * x + a.s * v[2]; v[2] = a.c * v[2] - a.s * x; }
__host__ __device__ __forceinline__ void fe_rotZ(float v[3], FeAng a) { const float x = v[0]; v[0] = a.c * x - a.s * v[1]; v[1] = a.s * x + a.c * v[1]; }
// IMUState::interpolate (ScanRegistration.h:151-169) on {roll, pitch, yaw, pos[3], vel[3]}
__host__ __device__ __forceinline__ void fe_imu_interpolate(const float* start, const float* end, float ratio, float* r) {
  const float invRatio = 1 - ratio;
  r[0] = start[0] * invRatio + end[0] * ratio;
  r[1] = start[1] * invRatio + end[1] * ratio;
  if ((double)(start[2] - end[2]) > M_PI) r[2] = (float)((double)(start[2] * invRatio) + ((double)end[2] + 2 * M_PI) * (double)ratio);
  else if ((double)(start[2] - end[2]) < -M_PI) r[2] = (float)((double)(start[2] * invRatio) + ((double)end[2] - 2 * M_PI) * (double)ratio);
  else r[2] = start[2] * invRatio + end[2] * ratio;
  for (int k = 3; k < 9; k++) r[k] = start[k] * invRatio + end[k] * ratio;
}
// interpolateIMUStateFor (ScanRegistration.cpp:168-186) given where the forward-only _imuIdx stands: idx = first state at or after
// scanTime + rmax (rmax = the largest relTime seen so far, 0 included), the point's own relTime decides the interpolation
__host__ __device__ __forceinline__ void fe_imu_state_for(int nimu, const double* dt, const double* dstamp, const float* states, float rmax,
                                                          float rel, float* out) {
  int idx = 0;
  while (idx < nimu - 1 && dt[idx] + (double)rmax > 0) idx++;
  const double timeDiff = dt[idx] + (double)rel;
  if (idx == 0 || timeDiff > 0) { for (int k = 0; k < 9; k++) out[k] = states[9 * idx + k]; return; }
  const float ratio = (float)(-timeDiff / dstamp[idx]);
  fe_imu_interpolate(states + 9 * idx, states + 9 * (idx - 1), ratio, out);
}

__device__ __forceinline__ int fe_scan_id_pandar(float angle) {   // lidar_type.h:78-104, double comparisons and arithmetic
  int scanID = 0;
  if ((double)angle < -15.0) scanID = 0;
  else if ((double)angle > -15.0 && (double)angle < -5.8) scanID = (int)((double)angle + 16.0 + 0.5);
  else if ((double)angle > -5.8 && (double)angle < 2.8) scanID = (int)(((double)angle + 5.667) / 0.33 + 0.5) + 10;
  else if ((double)angle > 1.8 && (double)angle < 7.5) scanID = (int)((double)angle - 2.0 + 0.5) + 34;
  return scanID;
}

// point i of the sweep -> (swapped point, ring or -1, first-half orientation)
__device__ __forceinline__ int fe_classify(const FrontEndArgs& a, int i, float4* pt, float* ori0) {
  const float4 in = a.sweep[i];
  float4 p = make_float4(in.y, in.z, in.x, in.w);                                        // :120-123
  *pt = p; *ori0 = 0.f;
  if (!isfinite(p.x) || !isfinite(p.y) || !isfinite(p.z)) return -1;
  if (p.x * p.x + p.y * p.y + p.z * p.z < 0.0001) return -1;                             // double compare, like the reference
  const float angle = cm_atanf(p.y / sqrtf(p.x * p.x + p.z * p.z));
  int ring;
  if (a.lidar == 3) ring = fe_scan_id_pandar((float)((double)angle * 180.0 / M_PI));     // rad2deg(float), math_utils.h:23
  else ring = (int)((((double)(angle * 180.f) / M_PI) - (double)a.lower) * (double)a.factor + 0.5);   // MultiScanRegistration.h:85-87 (angle * 180 is a float product)
  if (ring >= a.nrings || ring < 0) return -1;
  *ori0 = -cm_atan2f(p.x, p.z);
  return ring;
}
// the two unwrap rules (:139-156), double arithmetic where the reference mixes float with M_PI
__device__ __forceinline__ float fe_first_half(float ori, float start_ori) {
  if ((double)ori < (double)start_ori - M_PI / 2) ori = (float)((double)ori + 2 * M_PI);
  else if ((double)ori > (double)start_ori + M_PI * 3 / 2) ori = (float)((double)ori - 2 * M_PI);
  return ori;
}
__device__ __forceinline__ float fe_second_half(float ori, float end_ori) {
  ori = (float)((double)ori + 2 * M_PI);
  if ((double)ori < (double)end_ori - M_PI * 3 / 2) ori = (float)((double)ori + 2 * M_PI);
  else if ((double)ori > (double)end_ori + M_PI / 2) ori = (float)((double)ori - 2 * M_PI);
  return ori;
}

__global__ void __launch_bounds__(FE_THREADS) fe_classify_kernel(FrontEndArgs a) {
  __shared__ int hist[FE_MAXRINGS];
  if (threadIdx.x < FE_MAXRINGS) hist[threadIdx.x] = 0;
  __syncthreads();
  int first = a.n;
  for (int c = 0; c < FE_CHUNKS; c++) {
    const int i = (blockIdx.x * FE_CHUNKS + c) * FE_THREADS + threadIdx.x;
    if (i >= a.n) break;
    float4 p; float ori0;
    const int ring = fe_classify(a, i, &p, &ori0);
    a.ring_of[i] = ring;
    if (ring >= 0) {
      atomicAdd(&hist[ring], 1);
      const float o1 = fe_first_half(ori0, a.start_ori);
      if ((double)(o1 - a.start_ori) > M_PI && i < first) first = i;                    // `ori - startOri > M_PI`: float difference
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  if ((threadIdx.x & 31) == 0 && first < a.n) atomicMin(a.first_half, first);
  __syncthreads();
  if (threadIdx.x < a.nrings) a.block_hist[blockIdx.x * a.nrings + threadIdx.x] = hist[threadIdx.x];
}

// IMU de-skew, pass 1: relTime of every accepted point (needs the half-sweep index, i.e. a finished fe_classify_kernel), its
// maximum per CTA and the index of the last accepted point
__global__ void __launch_bounds__(FE_THREADS) fe_reltime_kernel(FrontEndArgs a) {
  __shared__ float wmax[FE_THREADS / 32];
  const int half_at = *a.first_half;
  float m = -FLT_MAX; int last = -1;
  for (int c = 0; c < FE_CHUNKS; c++) {
    const int i = (blockIdx.x * FE_CHUNKS + c) * FE_THREADS + threadIdx.x;
    if (i >= a.n) break;
    float r = -FLT_MAX;
    if (a.ring_of[i] >= 0) {
      float4 p; float ori0;
      fe_classify(a, i, &p, &ori0);
      const float ori = (i <= half_at) ? fe_first_half(ori0, a.start_ori) : fe_second_half(ori0, a.end_ori);
      r = a.scan_period * (ori - a.start_ori) / (a.end_ori - a.start_ori);
      last = i;
    }
    a.rel[i] = r;
    m = fmaxf(m, r);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); last = max(last, __shfl_xor_sync(0xffffffffu, last, o)); }
  if ((threadIdx.x & 31) == 0) { wmax[threadIdx.x >> 5] = m; if (last >= 0) atomicMax(a.last_idx, last); }
  __syncthreads();
  if (threadIdx.x == 0) { for (int w = 1; w < FE_THREADS / 32; w++) m = fmaxf(m, wmax[w]); a.bmax[blockIdx.x] = m; }
}

__global__ void __launch_bounds__(FE_MAXRINGS) fe_scan_kernel(FrontEndArgs a, int nblocks) {
  const int r = threadIdx.x;
  if (r >= a.nrings) return;
  int acc = 0;
  for (int b = 0; b < nblocks; b++) { const int c = a.block_hist[b * a.nrings + r]; a.block_hist[b * a.nrings + r] = acc; acc += c; }
  a.ring_total[r] = acc;
}

__global__ void __launch_bounds__(256) fe_fill_kernel(float4* frame, float* tags, size_t n) {
  const float qnan = __int_as_float(0x7fc00000);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    frame[i] = make_float4(qnan, qnan, qnan, 0.f);
    tags[i] = 0.f;
  }
}

__global__ void __launch_bounds__(FE_THREADS) fe_place_kernel(FrontEndArgs a) {
  __shared__ int run[FE_MAXRINGS];                       // points of this CTA already placed, per ring
  __shared__ int wcnt[FE_THREADS / 32][FE_MAXRINGS];     // this chunk's points per warp and ring
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < FE_MAXRINGS) run[threadIdx.x] = (threadIdx.x < a.nrings) ? a.block_hist[blockIdx.x * a.nrings + threadIdx.x] : 0;
  for (int k = threadIdx.x; k < (FE_THREADS / 32) * FE_MAXRINGS; k += FE_THREADS) (&wcnt[0][0])[k] = 0;
  __syncthreads();
  const int half_at = *a.first_half;
  // IMU de-skew: the largest relTime of all EARLIER accepted points (0 included: reset() interpolates the start state at relTime 0)
  __shared__ float s_wmax[FE_THREADS / 32];
  float carry = 0.f;
  if (a.nimu > 0) for (int b = 0; b < (int)blockIdx.x; b++) carry = fmaxf(carry, a.bmax[b]);
  const int last_idx = a.nimu > 0 ? *a.last_idx : -1;
  for (int c = 0; c < FE_CHUNKS; c++) {
    const int i = (blockIdx.x * FE_CHUNKS + c) * FE_THREADS + threadIdx.x;
    const int ring = i < a.n ? a.ring_of[i] : -1;
    float rmax = 0.f, rown = 0.f;
    if (a.nimu > 0) {   // inclusive prefix maximum of relTime over the points of this chunk, in index order
      rown = i < a.n ? a.rel[i] : -FLT_MAX;
      float x = rown;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x = fmaxf(x, y); }
      if (lane == 31) s_wmax[warp] = x;
      __syncthreads();
      float before = carry;
      for (int w = 0; w < warp; w++) before = fmaxf(before, s_wmax[w]);
      rmax = fmaxf(before, x);
      float chunk = carry;
      for (int w = 0; w < FE_THREADS / 32; w++) chunk = fmaxf(chunk, s_wmax[w]);
      __syncthreads();
      carry = chunk;
    }
    const unsigned int same = __match_any_sync(0xffffffffu, ring);
    const int below = __popc(same & ((1u << lane) - 1u));
    if (ring >= 0 && below == 0) wcnt[warp][ring] = __popc(same);
    __syncthreads();
    if (ring >= 0) {
      int pos = run[ring] + below;
      for (int w = 0; w < warp; w++) pos += wcnt[w][ring];
      float4 p; float ori0;
      fe_classify(a, i, &p, &ori0);                      // recomputed: cheaper than 20 bytes per point through global memory
      const float ori = (i <= half_at) ? fe_first_half(ori0, a.start_ori) : fe_second_half(ori0, a.end_ori);
      const float relTime = a.scan_period * (ori - a.start_ori) / (a.end_ori - a.start_ori);   // :159-160
      const size_t slot = (size_t)ring * a.cols + pos;
      if (a.nimu > 0) {   // setIMUTransformFor + transformToStartIMU (ScanRegistration.cpp:145-166)
        float cur[9];
        fe_imu_state_for(a.nimu, a.imu_dt, a.imu_dstamp, a.imu_state, rmax, relTime, cur);
        const float relSweepTime = relTime;             // (_scanTime - _sweepStart).toSec() + relTime at the start of a sweep
        float shift[3];
        for (int k = 0; k < 3; k++) shift[k] = (cur[3 + k] - a.imu_start[3 + k]) - a.imu_start[6 + k] * relSweepTime;
        float v[3] = {p.x, p.y, p.z};
        fe_rotZ(v, fe_ang(cur[0])); fe_rotX(v, fe_ang(cur[1])); fe_rotY(v, fe_ang(cur[2]));              // rotateZXY(point, roll, pitch, yaw)
        v[0] += shift[0]; v[1] += shift[1]; v[2] += shift[2];
        fe_rotY(v, fe_neg(fe_ang(a.imu_start[2]))); fe_rotX(v, fe_neg(fe_ang(a.imu_start[1]))); fe_rotZ(v, fe_neg(fe_ang(a.imu_start[0])));
        p.x = v[0]; p.y = v[1]; p.z = v[2];
        if (i == last_idx) {
          for (int k = 0; k < 3; k++) { a.last_out[k] = cur[k]; a.last_out[3 + k] = shift[k]; a.last_out[6 + k] = cur[6 + k]; }
        }
      }
      a.frame[slot] = p;
      a.tags[slot] = (float)ring + relTime;              // point.curvature = scanID + relTime
    }
    __syncthreads();
    if (threadIdx.x < a.nrings) {
      int add = 0;
      for (int w = 0; w < FE_THREADS / 32; w++) { add += wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = 0; }
      run[threadIdx.x] += add;
    }
    __syncthreads();
  }
}

// host side ------------------------------------------------------------------------------------------------------------------
bool frontend_mapper(int lidar, float* lower, float* upper, int* nrings) {
  if (lidar == 0) { *lower = -15; *upper = 15; *nrings = 16; }
  else if (lidar == 1) { *lower = -30.67f; *upper = 10.67f; *nrings = 32; }
  else if (lidar == 2) { *lower = -24.9f; *upper = 2; *nrings = 64; }
  else if (lidar == 3) { *lower = -15.444f; *upper = 6.96f; *nrings = 40; }   // MultiScanMapperP::Pandar40
  else return false;
  return true;
}

// d_sweep: n raw points on the device; first / last: the same sweep's first and last point (host copies, they fix startOri / endOri).
// Returns the ring-major frame and tags in fe.frame / fe.tags (device), rows = nrings, cols = longest ring.
// ScanRegistration::handleIMUMessage (ScanRegistration.cpp:89-121): gravity removed in the IMU frame, position / velocity integrated
// in the world frame.  Host code (one call per IMU message).
void ImuHistoryHost::push(double t, double roll, double pitch, double yaw, double ax, double ay, double az) {
  float acc[3] = {float(ay - sin(roll) * cos(pitch) * 9.81), float(az - cos(roll) * cos(pitch) * 9.81), float(ax + sin(pitch) * 9.81)};
  float nw[9] = {(float)roll, (float)pitch, (float)yaw, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (!stamp.empty()) {
    fe_rotZ(acc, fe_ang(nw[0])); fe_rotX(acc, fe_ang(nw[1])); fe_rotY(acc, fe_ang(nw[2]));   // rotateZXY(acc, roll, pitch, yaw)
    const float* prev = &state[9 * (stamp.size() - 1)];
    const float timeDiff = float(t - stamp.back());
    for (int k = 0; k < 3; k++) {
      nw[3 + k] = (prev[3 + k] + (prev[6 + k] * timeDiff)) + (((0.5f * acc[k]) * timeDiff) * timeDiff);
      nw[6 + k] = prev[6 + k] + acc[k] * timeDiff;
    }
  }
  if (stamp.size() >= capacity) { stamp.erase(stamp.begin()); state.erase(state.begin(), state.begin() + 9); }
  stamp.push_back(t);
  state.insert(state.end(), nw, nw + 9);
}

void SweepFrontEnd::run(const float4* d_sweep, int n, const float4& first, const float4& last, int lidar, float scan_period, cudaStream_t st,
                        int* rows_out, int* cols_out, const ImuHistoryHost* imu, double scan_time, float* imu_trans12) {
  float lower, upper; int nrings;
  frontend_mapper(lidar, &lower, &upper, &nrings);
  FrontEndArgs a;
  a.sweep = d_sweep; a.n = n; a.lidar = lidar; a.nrings = nrings; a.lower = lower;
  a.factor = (nrings - 1) / (upper - lower);                                              // MultiScanRegistration.h:63
  a.scan_period = scan_period;
  float startOri = -cm_atan2f(first.y, first.x);                                          // :103-110
  float endOri = -cm_atan2f(last.y, last.x) + 2 * float(M_PI);
  if (endOri - startOri > 3 * M_PI) endOri -= 2 * M_PI;
  else if (endOri - startOri < M_PI) endOri += 2 * M_PI;
  a.start_ori = startOri; a.end_ori = endOri;
  const int per_block = FE_THREADS * FE_CHUNKS;
  const int nblocks = (n + per_block - 1) / per_block;
  ring_of.reserve(sizeof(int) * (size_t)(n > 0 ? n : 1));
  hist.reserve(sizeof(int) * ((size_t)(nblocks > 0 ? nblocks : 1) * nrings + nrings + 1));
  a.ring_of = (int*)ring_of.p;
  a.block_hist = (int*)hist.p; a.ring_total = a.block_hist + (size_t)nblocks * nrings; a.first_half = a.ring_total + nrings;
  a.frame = nullptr; a.tags = nullptr; a.cols = 0;
  // ---- IMU history -> device: scanTime - stamp, stamp differences, states; _imuStart = the state at relTime 0 (reset()) ----
  const int nimu = (imu && n > 0) ? (int)imu->stamp.size() : 0;
  a.nimu = nimu; a.imu_dt = nullptr; a.imu_dstamp = nullptr; a.imu_state = nullptr; a.rel = nullptr; a.bmax = nullptr; a.last_idx = nullptr; a.last_out = nullptr;
  for (int k = 0; k < 9; k++) a.imu_start[k] = 0.f;
  if (imu_trans12) for (int k = 0; k < 12; k++) imu_trans12[k] = 0.f;
  if (nimu > 0) {
    std::vector<double> hd(2 * (size_t)nimu);
    for (int j = 0; j < nimu; j++) { hd[j] = scan_time - imu->stamp[j]; hd[nimu + j] = j ? imu->stamp[j] - imu->stamp[j - 1] : 0.0; }
    fe_imu_state_for(nimu, hd.data(), hd.data() + nimu, imu->state.data(), 0.f, 0.f, a.imu_start);
    const size_t bytes_d = 2 * (size_t)nimu * sizeof(double), bytes_s = 9 * (size_t)nimu * sizeof(float);
    imu_buf.reserve(bytes_d + bytes_s + 64);
    cudaMemcpyAsync(imu_buf.p, hd.data(), bytes_d, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync((char*)imu_buf.p + bytes_d, imu->state.data(), bytes_s, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);                        // hd goes out of scope
    a.imu_dt = (const double*)imu_buf.p; a.imu_dstamp = a.imu_dt + nimu; a.imu_state = (const float*)((const char*)imu_buf.p + bytes_d);
    rel.reserve(((size_t)n + (size_t)nblocks + 16) * sizeof(float));
    a.rel = (float*)rel.p; a.bmax = a.rel + n; a.last_out = a.bmax + nblocks; a.last_idx = (int*)(a.last_out + 9);
    const int minus1 = -1;
    cudaMemcpyAsync(a.last_idx, &minus1, sizeof(int), cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(a.last_out, 0, 9 * sizeof(float), st);
  }
  int totals[FE_MAXRINGS];
  for (int r = 0; r < nrings; r++) totals[r] = 0;
  if (n > 0) {
    cudaMemcpyAsync(a.first_half, &n, sizeof(int), cudaMemcpyHostToDevice, st);
    CM_LAUNCH(fe_classify_kernel, nblocks, FE_THREADS, 0, st, a);
    CM_LAUNCH(fe_scan_kernel, 1, FE_MAXRINGS, 0, st, a, nblocks);
    if (nimu > 0) CM_LAUNCH(fe_reltime_kernel, nblocks, FE_THREADS, 0, st, a);
    // the longest ring fixes the row pitch (and the shared-memory size of sr_ring_kernel): one small read-back
    cudaError_t e = cudaMemcpyAsync(totals, a.ring_total, sizeof(int) * nrings, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) throw CudaError{e, "front end: ring totals"};
  }
  int cols = 1;
  for (int r = 0; r < nrings; r++) cols = totals[r] > cols ? totals[r] : cols;
  const size_t slots = (size_t)nrings * cols;
  frame.reserve(slots * sizeof(float4)); tags.reserve(slots * sizeof(float));
  a.frame = (float4*)frame.p; a.tags = (float*)tags.p; a.cols = cols;
  CM_LAUNCH(fe_fill_kernel, (int)((slots + 255) / 256 < 1184 ? (slots + 255) / 256 : 1184), 256, 0, st, a.frame, a.tags, slots);
  if (n > 0) CM_LAUNCH(fe_place_kernel, nblocks, FE_THREADS, 0, st, a);
  if (nimu > 0 && imu_trans12) {   // publishResult (ScanRegistration.cpp:681-708): _imuCur / _imuPositionShift are those of the last point
    float lo[9];
    cudaError_t e = cudaMemcpyAsync(lo, a.last_out, sizeof(lo), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) throw CudaError{e, "front end: imu state of the last point"};
    const FeAng ny = fe_neg(fe_ang(a.imu_start[2])), np_ = fe_neg(fe_ang(a.imu_start[1])), nr = fe_neg(fe_ang(a.imu_start[0]));
    imu_trans12[0] = a.imu_start[1]; imu_trans12[1] = a.imu_start[2]; imu_trans12[2] = a.imu_start[0];   // pitch, yaw, roll
    imu_trans12[3] = lo[1]; imu_trans12[4] = lo[2]; imu_trans12[5] = lo[0];
    float s3[3] = {lo[3], lo[4], lo[5]};
    fe_rotY(s3, ny); fe_rotX(s3, np_); fe_rotZ(s3, nr);                                                      // rotateYXZ(-yaw, -pitch, -roll)
    float v3[3] = {lo[6] - a.imu_start[6], lo[7] - a.imu_start[7], lo[8] - a.imu_start[8]};
    fe_rotY(v3, ny); fe_rotX(v3, np_); fe_rotZ(v3, nr);
    for (int k = 0; k < 3; k++) { imu_trans12[6 + k] = s3[k]; imu_trans12[9 + k] = v3[k]; }
  }
  *rows_out = nrings; *cols_out = cols;
}

}  // namespace cm
