// cm_dist.cu -- multi-GPU plumbing inside the library (BASELINE config 4: one map split over the ranks).
//
// One process per GPU.  cm_dist_init joins the ranks (NCCL, loaded with dlopen so that libcoopermap.so itself has no NCCL
// dependency) and, where CUDA IPC works between the processes (same box, NVLink / NVSwitch), maps a small mailbox of every
// peer into this process.  The only data-path exchange of the sharded scan-to-map loop is the sum of the per-rank partial
// normal equations (ScanMatch.cpp:134-208 distributed: 32 doubles per stream and Gauss-Newton iteration).  It is done by ONE
// kernel, dist_exchange_kernel: every rank stores its vector into every peer's mailbox over NVLink (plain st.global on the
// mapped peer pointers), publishes a sequence number, waits for the sequence numbers of all peers and adds the nranks vectors
// in RANK ORDER -- the same total, bit for bit, on every rank, so the 6x6 solve that follows is redundant and no pose is ever
// broadcast.  256 bytes per rank: latency is everything, a 2 us one-shot exchange replaces a ~15 us library all-reduce.
// Without IPC (or across boxes) the same kernel sums what ncclAllGather delivered.
#include "cm_ctx.h"
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace cm {

// ---- the few NCCL entry points used, bound at run time -------------------------------------------------------------------
typedef struct { char internal[128]; } NcclUniqueId;                     // nccl.h: ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
enum { kNcclChar = 0, kNcclFloat64 = 8 };                                // nccl.h: ncclDataType_t
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string* why) {
    if (lib) return true;
    const char* names[] = {getenv("COOPERMAP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { if (why) *why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return false; }
    GetUniqueId = (int (*)(NcclUniqueId*))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
    AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllGather");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather) { if (why) *why = "libnccl lacks an entry point"; return false; }
    return true;
  }
};
static NcclApi g_nccl;

// ---- the exchange kernel ---------------------------------------------------------------------------------------------------
// mailbox of one rank: [2 slots][nranks] sequence numbers (u64, 256-byte aligned block) + [2 slots][nranks][nmax] doubles.
// Exchange number `seq` uses slot seq & 1; a rank can only start exchange seq + 2 after every peer has finished reading seq
// (it has seen their seq + 1), so two slots never collide.
__global__ void __launch_bounds__(1024) dist_exchange_kernel(DistDev d, double* vec, int n, unsigned long long seq, int* err) {
  const int slot = (int)(seq & 1ull);
  const int tid = threadIdx.x;
  if (d.p2p) {
    // 1. my vector into every rank's mailbox (mine included), then the sequence number behind a system-wide fence
    for (int i = tid; i < d.nranks * n; i += blockDim.x) {
      const int r = i / n, k = i - r * n;
      d.mbox_peer[r][((size_t)slot * d.nranks + d.rank) * d.nmax + k] = vec[k];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < d.nranks) {
      volatile unsigned long long* f = d.flags_peer[tid] + (size_t)slot * d.nranks + d.rank;
      *f = seq;
    }
    // 2. wait for everybody's vector (bounded: a dead peer must not hang the GPU)
    if (tid < d.nranks) {
      volatile unsigned long long* f = d.flags_peer[d.rank] + (size_t)slot * d.nranks + tid;
      unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      while (*f != seq) {
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) { atomicExch(err, 1); break; }   // 4 s
      }
    }
    __threadfence_system();
    __syncthreads();
  }
  // 3. the total, ranks added in rank order (p2p: from my mailbox; else from the all-gathered copy the host passed in d.gathered)
  const double* src = d.p2p ? d.mbox_peer[d.rank] + (size_t)slot * d.nranks * d.nmax : d.gathered;
  const size_t stride = d.p2p ? (size_t)d.nmax : (size_t)n;
  for (int k = tid; k < n; k += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < d.nranks; r++) t += ((const volatile double*)src)[(size_t)r * stride + k];
    vec[k] = t;
  }
}

// sum d_vec[0..n) over the ranks, in place, on `stream` (every rank must call with the same n, in the same order)
int dist_allreduce(cm_ctx* ctx, double* d_vec, int n, cudaStream_t stream) {
  DistState& ds = ctx->dist;
  if (!ds.on || ds.dev.nranks <= 1) return CM_OK;
  if (n > ds.dev.nmax) return ctx_fail(ctx, CM_ERR_ARG, "dist_allreduce: vector longer than the mailbox");
  ++ds.seq;
  if (!ds.dev.p2p) {
    const int rc = g_nccl.AllGather(d_vec, ds.gathered.p, (size_t)n, kNcclFloat64, (NcclComm)ds.comm, stream);
    if (rc != 0) return ctx_fail(ctx, CM_ERR_CUDA, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
  }
  CM_LAUNCH(dist_exchange_kernel, 1, 1024, 0, stream, ds.dev, d_vec, n, ds.seq, (int*)ds.err.p);
  return CM_OK;
}

}  // namespace cm

using namespace cm;

extern "C" {

int cm_dist_unique_id(void* id128) {
  if (!id128) return CM_ERR_ARG;
  std::string why;
  if (!g_nccl.load(&why)) { fprintf(stderr, "coopermap: %s\n", why.c_str()); return CM_ERR_UNSUPPORTED; }
  NcclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return CM_ERR_CUDA;
  memcpy(id128, &id, sizeof(id));
  return CM_OK;
}

int cm_dist_init(cm_ctx* ctx, const void* id128, int rank, int nranks) {
  if (!ctx || !id128 || nranks < 1 || nranks > CM_DIST_MAX_RANKS || rank < 0 || rank >= nranks) return ctx_fail(ctx, CM_ERR_ARG, "bad argument");
  DistState& ds = ctx->dist;
  if (ds.on) return ctx_fail(ctx, CM_ERR_ARG, "cm_dist_init was already called on this context");
  std::string why;
  if (!g_nccl.load(&why)) return ctx_fail(ctx, CM_ERR_UNSUPPORTED, why);
  try {
    cudaSetDevice(ctx->cfg.device);
    NcclUniqueId id; memcpy(&id, id128, sizeof(id));
    NcclComm comm = nullptr;
    int rc = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (rc != 0) return ctx_fail(ctx, CM_ERR_CUDA, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    ds.comm = comm;
    DistDev& d = ds.dev;
    memset(&d, 0, sizeof(d));
    d.rank = rank; d.nranks = nranks; d.nmax = CM_DIST_NMAX;
    const size_t flag_bytes = 256 * ((2 * (size_t)nranks * sizeof(unsigned long long) + 255) / 256);
    const size_t box_bytes = flag_bytes + 2 * (size_t)nranks * d.nmax * sizeof(double);
    CM_CUDA_CHECK(ctx, cudaMalloc(&ds.mailbox, box_bytes));
    CM_CUDA_CHECK(ctx, cudaMemset(ds.mailbox, 0, box_bytes));
    ds.err.reserve(sizeof(int)); CM_CUDA_CHECK(ctx, cudaMemset(ds.err.p, 0, sizeof(int)));
    ds.gathered.reserve((size_t)nranks * d.nmax * sizeof(double));
    // exchange the IPC handles of the mailboxes (and whether creating one worked at all) through the communicator
    struct Hello { cudaIpcMemHandle_t h; int ok; int pad[15]; };
    static_assert(sizeof(Hello) == 128, "Hello is 128 bytes");
    Hello mine; memset(&mine, 0, sizeof(mine));
    mine.ok = (!getenv("COOPERMAP_DIST_NO_P2P") && cudaIpcGetMemHandle(&mine.h, ds.mailbox) == cudaSuccess) ? 1 : 0;
    cudaGetLastError();
    DeviceBuffer d_hello;
    d_hello.reserve(sizeof(Hello) * (size_t)(nranks + 1));
    CM_CUDA_CHECK(ctx, cudaMemcpy(d_hello.p, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    rc = g_nccl.AllGather(d_hello.p, (char*)d_hello.p + sizeof(Hello), sizeof(Hello), kNcclChar, comm, ctx->stream);
    if (rc != 0) return ctx_fail(ctx, CM_ERR_CUDA, "ncclAllGather (handles) failed");
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<Hello> all(nranks);
    CM_CUDA_CHECK(ctx, cudaMemcpy(all.data(), (char*)d_hello.p + sizeof(Hello), sizeof(Hello) * nranks, cudaMemcpyDeviceToHost));
    bool p2p = true;
    for (int r = 0; r < nranks; r++) p2p = p2p && all[r].ok;
    int opened = 1;
    if (p2p) {
      for (int r = 0; r < nranks && opened; r++) {
        void* p = ds.mailbox;
        if (r != rank) {
          if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; break; }
          ds.peer_ptr[r] = p;
        }
        d.flags_peer[r] = (unsigned long long*)p;
        d.mbox_peer[r] = (double*)((char*)p + flag_bytes);
      }
    }
    // everybody must agree: one failed open anywhere -> the all-gather transport everywhere
    Hello second; memset(&second, 0, sizeof(second)); second.ok = (p2p && opened) ? 1 : 0;
    CM_CUDA_CHECK(ctx, cudaMemcpy(d_hello.p, &second, sizeof(second), cudaMemcpyHostToDevice));
    rc = g_nccl.AllGather(d_hello.p, (char*)d_hello.p + sizeof(Hello), sizeof(Hello), kNcclChar, comm, ctx->stream);
    if (rc != 0) return ctx_fail(ctx, CM_ERR_CUDA, "ncclAllGather (agreement) failed");
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    CM_CUDA_CHECK(ctx, cudaMemcpy(all.data(), (char*)d_hello.p + sizeof(Hello), sizeof(Hello) * nranks, cudaMemcpyDeviceToHost));
    bool agreed = true;
    for (int r = 0; r < nranks; r++) agreed = agreed && all[r].ok;
    d.p2p = agreed ? 1 : 0;
    d.gathered = (const double*)ds.gathered.p;
    if (!d.p2p) { d.flags_peer[rank] = (unsigned long long*)ds.mailbox; d.mbox_peer[rank] = (double*)((char*)ds.mailbox + flag_bytes); }
    ds.seq = 0;
    ds.on = true;
    // the map of this context now keeps only this rank's cubes (+ halo)
    ctx->map.shard_rank = rank; ctx->map.shard_nranks = nranks;
  } catch (const CudaError& e) {
    return ctx_fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_dist_info(cm_ctx* ctx, int* rank, int* nranks, int* p2p) {
  if (!ctx) return CM_ERR_ARG;
  if (rank) *rank = ctx->dist.on ? ctx->dist.dev.rank : 0;
  if (nranks) *nranks = ctx->dist.on ? ctx->dist.dev.nranks : 1;
  if (p2p) *p2p = ctx->dist.on ? ctx->dist.dev.p2p : 0;
  return CM_OK;
}

/* sum n doubles over the ranks (n <= 8192), host buffer in and out: the library's exchange kernel as an operator (tests, timing) */
int cm_dist_allreduce_host(cm_ctx* ctx, double* vec, int n, int repeat, float* ms_per_call) {
  if (!ctx || !vec || n <= 0 || !ctx->dist.on) return ctx_fail(ctx, CM_ERR_ARG, "bad argument / cm_dist_init not called");
  try {
    cudaSetDevice(ctx->cfg.device);
    DeviceBuffer& b = ctx->dist.scratch;
    b.reserve(sizeof(double) * (size_t)n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (repeat < 1) repeat = 1;
    float ms = 0.f;
    if (repeat > 1) {   // timing: the exchange kernel alone, back to back, on a vector of zeros (sums stay zero)
      CM_CUDA_CHECK(ctx, cudaMemsetAsync(b.p, 0, sizeof(double) * n, ctx->stream));
      for (int k = 0; k < 3; k++) { const int rc = dist_allreduce(ctx, (double*)b.p, n, ctx->stream); if (rc < 0) return rc; }
      cudaEventRecord(e0, ctx->stream);
      for (int k = 0; k < repeat; k++) { const int rc = dist_allreduce(ctx, (double*)b.p, n, ctx->stream); if (rc < 0) return rc; }
      cudaEventRecord(e1, ctx->stream);
      CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
      cudaEventElapsedTime(&ms, e0, e1);
    }
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(b.p, vec, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    { const int rc = dist_allreduce(ctx, (double*)b.p, n, ctx->stream); if (rc < 0) return rc; }
    std::vector<double> out(n);
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(out.data(), b.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    memcpy(vec, out.data(), sizeof(double) * n);
    if (ms_per_call) *ms_per_call = ms / (float)repeat;
    int err = 0;
    CM_CUDA_CHECK(ctx, cudaMemcpy(&err, ctx->dist.err.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return ctx_fail(ctx, CM_ERR_CUDA, "dist exchange timed out waiting for a peer");
  } catch (const CudaError& e) {
    return ctx_fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

}  // extern "C"

namespace cm {
void dist_destroy(cm_ctx* ctx) {
  DistState& ds = ctx->dist;
  if (!ds.on && !ds.mailbox) return;
  for (int r = 0; r < CM_DIST_MAX_RANKS; r++) if (ds.peer_ptr[r]) { cudaIpcCloseMemHandle(ds.peer_ptr[r]); ds.peer_ptr[r] = nullptr; }
  if (ds.mailbox) { cudaFree(ds.mailbox); ds.mailbox = nullptr; }
  if (ds.comm && g_nccl.CommDestroy) { g_nccl.CommDestroy((NcclComm)ds.comm); ds.comm = nullptr; }
  ds.on = false;
}
}  // namespace cm
