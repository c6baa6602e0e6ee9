// cm_stage.cu -- sweeps that arrive the way a nodelet holds them: one pcl::PointCloud<pcl::PointXYZI> per LiDAR stream, i.e.
// PAGEABLE host memory with a 32-byte point stride (x, y, z, pad, intensity, pad[3]; the result of pcl::fromROSMsg,
// util/ros_utils.h:27-35).  The copy engine cannot read pageable memory at full speed and the pipeline never looks at the padding
// or the intensity (scan registration overwrites it with ring + relTime, OrganizedScanRegistration.cpp:109-110), so the library
// repacks the sweeps itself: worker threads gather x, y, z (12 bytes per point) into library-owned pinned staging buffers, chunk by
// chunk, each chunk's upload starts as soon as it is packed (the copy of chunk c overlaps the packing of chunk c + 1), and a small
// kernel expands the packed coordinates to the float4 layout scan registration reads.  12 instead of 16 bytes per point cross PCIe.
#include "cm_ctx.h"
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <sched.h>
#include <thread>

namespace cm {

// ---- a small persistent worker pool (parallel_for over task indices) ---------------------------------------------------------
struct StagePool {
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  const std::function<void(int)>* fn = nullptr;
  int ntasks = 0;
  std::atomic<int> next{0};
  int running = 0;
  unsigned long long epoch = 0;
  bool stop = false;

  explicit StagePool(int nthreads) {
    for (int i = 0; i < nthreads; i++) threads.emplace_back([this]() { loop(); });
  }
  ~StagePool() {
    { std::lock_guard<std::mutex> l(mu); stop = true; }
    cv_work.notify_all();
    for (auto& t : threads) t.join();
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void(int)>* f;
      int n;
      {
        std::unique_lock<std::mutex> l(mu);
        cv_work.wait(l, [&]() { return stop || epoch != seen; });
        if (stop) return;
        seen = epoch; f = fn; n = ntasks;
        if (!f) continue;          // woke up after that parallel_for had already finished: must not touch `next`
        running++;
      }
      for (int t; (t = next.fetch_add(1)) < n;) (*f)(t);
      {
        std::lock_guard<std::mutex> l(mu);
        if (--running == 0) cv_done.notify_all();
      }
    }
  }
  // runs f(0 .. n-1) on the workers and the calling thread; returns when all are done
  void parallel_for(int n, const std::function<void(int)>& f) {
    if (threads.empty() || n <= 1) { for (int t = 0; t < n; t++) f(t); return; }
    {
      std::lock_guard<std::mutex> l(mu);
      fn = &f; ntasks = n; next.store(0); epoch++;
    }
    cv_work.notify_all();
    for (int t; (t = next.fetch_add(1)) < n;) f(t);
    std::unique_lock<std::mutex> l(mu);
    // every worker that picked this epoch up has left its task loop; workers that never woke find no task left
    cv_done.wait(l, [&]() { return running == 0; });
    fn = nullptr; ntasks = 0;
  }
};

static int stage_threads() {
  static const int n = []() {
    if (const char* e = getenv("COOPERMAP_STAGE_THREADS")) return std::max(1, atoi(e));
    cpu_set_t set; CPU_ZERO(&set);
    int avail = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
    if (avail <= 0) avail = 4;
    return std::max(1, std::min(avail, 16));
  }();
  return n;
}

StagePool* stage_pool(cm_ctx* ctx) {
  if (!ctx->stage_pool) ctx->stage_pool = new StagePool(stage_threads() - 1);   // the calling thread works too
  return ctx->stage_pool;
}
void stage_pool_destroy(cm_ctx* ctx) { delete ctx->stage_pool; ctx->stage_pool = nullptr; }

// x, y, z of `count` points, `stride` bytes apart, packed to 12 bytes each
static void pack_xyz(const unsigned char* src, size_t stride, size_t count, float* dst) {
#if defined(__SSE2__)
  // four points per round: four 16-byte loads (x, y, z and the float behind them -- inside the point for every stride >= 16), three
  // shuffled 16-byte stores; non-temporal when the destination is aligned (the staging buffer is read next by the copy engine, not
  // by this core: no read-for-ownership, no cache pollution)
  static const bool simd = getenv("COOPERMAP_STAGE_SCALAR") == nullptr;
  if (simd && stride >= 16 && count >= 4) {
    const bool nt = ((uintptr_t)dst & 15) == 0;
    size_t i = 0;
    for (; i + 4 <= count; i += 4) {
      const unsigned char* q = src + i * stride;
      const __m128 p0 = _mm_loadu_ps(reinterpret_cast<const float*>(q)), p1 = _mm_loadu_ps(reinterpret_cast<const float*>(q + stride));
      const __m128 p2 = _mm_loadu_ps(reinterpret_cast<const float*>(q + 2 * stride)), p3 = _mm_loadu_ps(reinterpret_cast<const float*>(q + 3 * stride));
      const __m128 t01 = _mm_shuffle_ps(p0, p1, _MM_SHUFFLE(0, 0, 2, 2));   // z0 z0 x1 x1
      const __m128 o0 = _mm_shuffle_ps(p0, t01, _MM_SHUFFLE(2, 0, 1, 0));   // x0 y0 z0 x1
      const __m128 o1 = _mm_shuffle_ps(p1, p2, _MM_SHUFFLE(1, 0, 2, 1));    // y1 z1 x2 y2
      const __m128 t23 = _mm_shuffle_ps(p2, p3, _MM_SHUFFLE(0, 0, 2, 2));   // z2 z2 x3 x3
      const __m128 o2 = _mm_shuffle_ps(t23, p3, _MM_SHUFFLE(2, 1, 2, 0));   // z2 x3 y3 z3
      float* d = dst + 3 * i;
      if (nt) { _mm_stream_ps(d, o0); _mm_stream_ps(d + 4, o1); _mm_stream_ps(d + 8, o2); }
      else { _mm_storeu_ps(d, o0); _mm_storeu_ps(d + 4, o1); _mm_storeu_ps(d + 8, o2); }
    }
    if (nt) _mm_sfence();
    for (; i < count; i++) memcpy(dst + 3 * i, src + i * stride, 12);
    return;
  }
#endif
  if (stride == 32 && ((uintptr_t)src & 3) == 0) {   // pcl::PointXYZI: two points per 64-byte line, read 12 of every 32 bytes
    for (size_t i = 0; i < count; i++) {
      const float* p = reinterpret_cast<const float*>(src + i * 32);
      dst[3 * i] = p[0]; dst[3 * i + 1] = p[1]; dst[3 * i + 2] = p[2];
    }
    return;
  }
  for (size_t i = 0; i < count; i++) memcpy(dst + 3 * i, src + i * stride, 12);   // any stride (a PointCloud2 point_step of 22: unaligned floats)
}

// packed xyz (12 B) -> float4 (x, y, z, 0): four points per thread, three 16-byte loads and four 16-byte stores
__global__ void __launch_bounds__(256) unpack_xyz_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t nquads) {
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (size_t)gridDim.x * blockDim.x) {
    const float4 a = in[3 * q], b = in[3 * q + 1], c = in[3 * q + 2];
    out[4 * q] = make_float4(a.x, a.y, a.z, 0.f);
    out[4 * q + 1] = make_float4(a.w, b.x, b.y, 0.f);
    out[4 * q + 2] = make_float4(b.z, b.w, c.x, 0.f);
    out[4 * q + 3] = make_float4(c.y, c.z, c.w, 0.f);
  }
}

// Packs the S sweeps into the slot's pinned staging buffer and uploads them; on return every copy has been ISSUED (the caller's
// clouds are no longer needed) and `copied` is recorded behind the last one on copy stream 0.  d_out: [S][rows*cols] float4.
int stage_upload_strided(cm_ctx* ctx, cm_ctx::PipeSlot& slot, const void* const* clouds, size_t stride, int rows, int cols,
                         cudaStream_t consumer) {
  const int S = ctx->map_streams;
  const size_t npts = (size_t)rows * cols;                 // per stream
  const size_t total = npts * S;
  if (total % 4) return ctx_fail(ctx, CM_ERR_ARG, "rows * cols * streams must be a multiple of 4");
  const size_t bytes = total * 12;
  bool packed_in_place = stride == 12;
  for (int s = 1; s < S && packed_in_place; s++) packed_in_place = (const char*)clouds[s] == (const char*)clouds[0] + (size_t)s * npts * 12;
  if (!packed_in_place && slot.h_xyz_cap < bytes) {
    if (slot.h_xyz) cudaFreeHost(slot.h_xyz);
    slot.h_xyz = nullptr; slot.h_xyz_cap = 0;
    static const bool wc = getenv("COOPERMAP_STAGE_WC") != nullptr;   // write-combined staging: faster for the copy engine, slower to fill
    CM_CUDA_CHECK(ctx, cudaHostAlloc(&slot.h_xyz, bytes, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    slot.h_xyz_cap = bytes;
  }
  slot.frames_xyz.reserve(bytes);
  slot.frames.reserve(total * sizeof(float4));
  cudaStream_t cs2[2] = {ctx->copy_stream, ctx->copy_stream2};
  // packed coordinates of all streams in ONE buffer (stride 12, cloud s right behind cloud s - 1): nothing to pack -- the caller's
  // buffer goes to the device as it is (asynchronously when it is pinned), 12 bytes per point, in two halves on the two copy streams
  bool contiguous = stride == 12;
  for (int s = 1; s < S && contiguous; s++) contiguous = (const char*)clouds[s] == (const char*)clouds[0] + (size_t)s * npts * 12;
  if (contiguous) {
    const size_t half = (bytes / 2) & ~(size_t)255;
    CM_TIMED("h2d_upload(xyz half)", cs2[0], CM_CUDA_CHECK(ctx, cudaMemcpyAsync(slot.frames_xyz.p, clouds[0], half, cudaMemcpyHostToDevice, cs2[0])));
    CM_TIMED("h2d_upload(xyz half)", cs2[1], CM_CUDA_CHECK(ctx, cudaMemcpyAsync((char*)slot.frames_xyz.p + half, (const char*)clouds[0] + half, bytes - half,
                                                                                  cudaMemcpyHostToDevice, cs2[1])));
    CM_CUDA_CHECK(ctx, cudaEventRecord(slot.copied, cs2[0]));
    CM_CUDA_CHECK(ctx, cudaEventRecord(slot.copied2, cs2[1]));
    CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(consumer, slot.copied, 0));
    CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(consumer, slot.copied2, 0));
    const size_t nq = total / 4;
    CM_LAUNCH(unpack_xyz_kernel, (int)std::min<size_t>((nq + 255) / 256, 148 * 8), 256, 0, consumer, (const float4*)slot.frames_xyz.p, (float4*)slot.frames.p, nq);
    return CM_OK;
  }
  StagePool* pool = stage_pool(ctx);
  // chunks of whole tasks; a task = `tpts` consecutive points of one stream
  const int NCH = 4;
  const size_t tpts = 16384;
  const int tasks_per_stream = (int)((npts + tpts - 1) / tpts);
  const int ntasks = S * tasks_per_stream;
  float* h = reinterpret_cast<float*>(slot.h_xyz);
  cudaStream_t cs[2] = {ctx->copy_stream, ctx->copy_stream2};
  size_t done_bytes = 0;
  for (int ch = 0; ch < NCH; ch++) {
    const int t0 = (int)((long long)ntasks * ch / NCH), t1 = (int)((long long)ntasks * (ch + 1) / NCH);
    if (t1 <= t0) continue;
    std::function<void(int)> f = [&](int k) {
      const int t = t0 + k, s = t / tasks_per_stream, part = t % tasks_per_stream;
      const size_t p0 = (size_t)part * tpts, cnt = std::min(tpts, npts - p0);
      pack_xyz(reinterpret_cast<const unsigned char*>(clouds[s]) + p0 * stride, stride, cnt, h + 3 * ((size_t)s * npts + p0));
    };
    pool->parallel_for(t1 - t0, f);
    // tasks t0 .. t1-1 cover a contiguous byte range of the staging buffer (streams and parts are laid out in task order)
    const size_t end_pts = (t1 == ntasks) ? total : ((size_t)(t1 / tasks_per_stream) * npts + std::min(npts, (size_t)(t1 % tasks_per_stream) * tpts));
    const size_t end_bytes = end_pts * 12;
    CM_TIMED("h2d_upload(xyz chunk)", cs[ch & 1],
             CM_CUDA_CHECK(ctx, cudaMemcpyAsync((char*)slot.frames_xyz.p + done_bytes, (const char*)slot.h_xyz + done_bytes, end_bytes - done_bytes,
                                                cudaMemcpyHostToDevice, cs[ch & 1])));
    done_bytes = end_bytes;
  }
  CM_CUDA_CHECK(ctx, cudaEventRecord(slot.copied, cs[0]));
  CM_CUDA_CHECK(ctx, cudaEventRecord(slot.copied2, cs[1]));
  CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(consumer, slot.copied, 0));
  CM_CUDA_CHECK(ctx, cudaStreamWaitEvent(consumer, slot.copied2, 0));
  const size_t nquads = total / 4;
  const int grid = (int)std::min<size_t>((nquads + 255) / 256, 148 * 8);
  CM_LAUNCH(unpack_xyz_kernel, grid, 256, 0, consumer, (const float4*)slot.frames_xyz.p, (float4*)slot.frames.p, nquads);
  return CM_OK;
}

}  // namespace cm
