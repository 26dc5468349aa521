// multi_gpu.cu -- C-ABI of the multi-GPU path: independent reference keyframes, one per
// GPU, no data-path collective; the only exchange is the final gather of the depth (f32) and convergence
// (i32) maps to one root GPU, as grouped ncclSend / ncclRecv over NVLink (NCCL has no gather primitive).
//
// The reference has no multi-GPU path at all (one SeedMatrix on the current device, legacy default stream:
// src/check_cuda_device.cu:109, src/seed_matrix.cu).  Two ways to drive it from C / C++:
//   * one process, one host thread, n GPUs (what a ROS node that owns several rmd::Depthmap objects does):
//     rmd_multi_create -> ncclCommInitAll;
//   * one process per GPU (torchrun / mpirun style): rmd_multi_unique_id on one rank, ship the 128 bytes by
//     any means, rmd_multi_create_rank on every rank -> ncclCommInitRank.
// NCCL is loaded with dlopen ("libnccl.so.2": the system's 2.27 or the one a host framework already
// loaded), so the library itself has no link-time dependency on it; without NCCL these entry points return
// RMD_ERR_UNSUPPORTED and everything else works.
#include <dlfcn.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "rmd_common.cuh"

namespace
{

using namespace rmdb;

// ---- the slice of nccl.h this file needs (ABI-stable since NCCL 2.7: ncclSend / ncclRecv)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // ncclSuccess = 0
enum { kNcclInt32 = 2 };    // ncclDataType_t: ncclInt8 0, ncclUint8 1, ncclInt32 2, ...

struct NcclApi
{
  void *handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
};

const NcclApi *nccl()
{
  static NcclApi api;
  static bool tried = false, ok = false;
  static std::mutex mutex;
  std::lock_guard<std::mutex> lock(mutex);
  if(!tried)
  {
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for(int i = 0; i < 2 && !api.handle; ++i)
      api.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if(api.handle)
    {
      bool all = true;
#define RMD_NCCL_SYM(field, sym) \
      do { *(void**)(&api.field) = dlsym(api.handle, sym); all = all && (api.field != NULL); } while(0)
      RMD_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
      RMD_NCCL_SYM(CommInitAll, "ncclCommInitAll");
      RMD_NCCL_SYM(CommInitRank, "ncclCommInitRank");
      RMD_NCCL_SYM(CommDestroy, "ncclCommDestroy");
      RMD_NCCL_SYM(GroupStart, "ncclGroupStart");
      RMD_NCCL_SYM(GroupEnd, "ncclGroupEnd");
      RMD_NCCL_SYM(Send, "ncclSend");
      RMD_NCCL_SYM(Recv, "ncclRecv");
      RMD_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef RMD_NCCL_SYM
      ok = all;
    }
  }
  return ok ? &api : NULL;
}

int fail_nccl(const NcclApi *api, ncclResult_t res, const char *what)
{
  set_last_error(std::string(what) + ": NCCL error " + std::to_string(res) + " (" +
                 (api && api->GetErrorString ? api->GetErrorString(res) : "?") + ")");
  return RMD_ERR_UNSUPPORTED;
}

#define RMD_NCCL_TRY(expr)                                                 \
  do {                                                                     \
    const ncclResult_t rmd_res_ = (expr);                                  \
    if(rmd_res_ != 0) return fail_nccl(api, rmd_res_, #expr);              \
  } while(0)

} // namespace

struct rmd_multi
{
  int n_ranks;                  // keyframes = GPUs of the job
  int n_local;                  // of which this process drives n_local (all of them, or one)
  int first_rank;               // rank of local member 0 (members are consecutive ranks)
  int width, height;
  std::vector<int> device;      // per local member
  std::vector<ncclComm_t> comm;
  std::vector<cudaStream_t> stream;
  std::vector<cudaEvent_t> ready;
  std::vector<int32_t*> send;   // packed [depth bits | convergence], 2 * w * h int32 per member
  std::vector<int32_t*> recv;   // per local member: n_ranks * 2 * w * h, allocated when it first is the root
};

namespace
{

int multi_alloc(rmd_multi *m)
{
  const size_t count = 2 * (size_t)m->width * m->height;
  for(int i = 0; i < m->n_local; ++i)
  {
    DeviceGuard guard(m->device[i]);
    RMD_CUDA_TRY(cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking));
    RMD_CUDA_TRY(cudaEventCreateWithFlags(&m->ready[i], cudaEventDisableTiming));
    RMD_CUDA_TRY(cudaMalloc(&m->send[i], count * sizeof(int32_t)));
  }
  return 0;
}

rmd_multi *multi_new(int n_ranks, int n_local, int first_rank, int width, int height)
{
  rmd_multi *m = new(std::nothrow) rmd_multi();
  if(!m) return NULL;
  m->n_ranks = n_ranks; m->n_local = n_local; m->first_rank = first_rank;
  m->width = width; m->height = height;
  m->device.assign(n_local, 0);
  m->comm.assign(n_local, (ncclComm_t)NULL);
  m->stream.assign(n_local, (cudaStream_t)NULL);
  m->ready.assign(n_local, (cudaEvent_t)NULL);
  m->send.assign(n_local, (int32_t*)NULL);
  m->recv.assign(n_local, (int32_t*)NULL);
  return m;
}

} // namespace

extern "C"
{

int rmd_multi_create(const int *devices, int n, int width, int height, rmd_multi_t **out)
{
  RMD_REQUIRE(out, "rmd_multi_create: out is null");
  *out = NULL;
  RMD_REQUIRE(devices && n >= 1 && n <= 64 && width > 0 && height > 0, "rmd_multi_create: bad argument");
  const NcclApi *api = nccl();
  if(!api) return fail(RMD_ERR_UNSUPPORTED, "rmd_multi_create: libnccl.so.2 could not be loaded");
  rmd_multi *m = multi_new(n, n, 0, width, height);
  if(!m) return fail((int)cudaErrorMemoryAllocation, "rmd_multi_create: host allocation failed");
  for(int i = 0; i < n; ++i) m->device[i] = devices[i];
  int rc = multi_alloc(m);
  if(!rc)
  {
    const ncclResult_t res = api->CommInitAll(m->comm.data(), n, devices);
    if(res != 0) rc = fail_nccl(api, res, "ncclCommInitAll");
  }
  if(rc)
  {
    rmd_multi_destroy(m);
    return rc;
  }
  *out = m;
  return 0;
}

int rmd_multi_unique_id(char id[128])
{
  RMD_REQUIRE(id, "rmd_multi_unique_id: null argument");
  const NcclApi *api = nccl();
  if(!api) return fail(RMD_ERR_UNSUPPORTED, "rmd_multi_unique_id: libnccl.so.2 could not be loaded");
  ncclUniqueId uid;
  RMD_NCCL_TRY(api->GetUniqueId(&uid));
  memcpy(id, uid.internal, 128);
  return 0;
}

int rmd_multi_create_rank(const char id[128], int n_ranks, int rank, int device, int width, int height,
                          rmd_multi_t **out)
{
  RMD_REQUIRE(out, "rmd_multi_create_rank: out is null");
  *out = NULL;
  RMD_REQUIRE(id && n_ranks >= 1 && rank >= 0 && rank < n_ranks && width > 0 && height > 0,
              "rmd_multi_create_rank: bad argument");
  const NcclApi *api = nccl();
  if(!api) return fail(RMD_ERR_UNSUPPORTED, "rmd_multi_create_rank: libnccl.so.2 could not be loaded");
  if(device < 0) RMD_CUDA_TRY(cudaGetDevice(&device));
  rmd_multi *m = multi_new(n_ranks, 1, rank, width, height);
  if(!m) return fail((int)cudaErrorMemoryAllocation, "rmd_multi_create_rank: host allocation failed");
  m->device[0] = device;
  int rc = multi_alloc(m);
  if(!rc)
  {
    DeviceGuard guard(device);
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    const ncclResult_t res = api->CommInitRank(&m->comm[0], n_ranks, uid, rank);
    if(res != 0) rc = fail_nccl(api, res, "ncclCommInitRank");
  }
  if(rc)
  {
    rmd_multi_destroy(m);
    return rc;
  }
  *out = m;
  return 0;
}

int rmd_multi_destroy(rmd_multi_t *m)
{
  if(!m) return 0;
  const NcclApi *api = nccl();
  for(int i = 0; i < m->n_local; ++i)
  {
    DeviceGuard guard(m->device[i]);
    if(m->stream[i]) cudaStreamSynchronize(m->stream[i]);
    if(m->comm[i] && api) api->CommDestroy(m->comm[i]);
    if(m->stream[i]) cudaStreamDestroy(m->stream[i]);
    if(m->ready[i]) cudaEventDestroy(m->ready[i]);
    cudaFree(m->send[i]);
    cudaFree(m->recv[i]);
    cudaGetLastError();
  }
  delete m;
  return 0;
}

int rmd_multi_size(rmd_multi_t *m, int *n_ranks, int *n_local, int *first_rank)
{
  RMD_REQUIRE(m, "rmd_multi_size: null handle");
  if(n_ranks) *n_ranks = m->n_ranks;
  if(n_local) *n_local = m->n_local;
  if(first_rank) *first_rank = m->first_rank;
  return 0;
}

int rmd_multi_gather_maps(rmd_multi_t *m, rmd_seeds_t *const *seeds, const float *const *dev_depth,
                          const size_t *dev_depth_pitch, int root, float *host_depth, int32_t *host_conv)
{
  RMD_REQUIRE(m && seeds, "rmd_multi_gather_maps: null argument");
  RMD_REQUIRE(root >= 0 && root < m->n_ranks, "rmd_multi_gather_maps: root out of range");
  const NcclApi *api = nccl();
  if(!api) return fail(RMD_ERR_UNSUPPORTED, "rmd_multi_gather_maps: libnccl.so.2 could not be loaded");
  const size_t px = (size_t)m->width * m->height, count = 2 * px;
  const int root_local = root - m->first_rank;               // index of the root among the local members, if local
  const bool have_root = root_local >= 0 && root_local < m->n_local;
  RMD_REQUIRE(!have_root || (host_depth && host_conv), "rmd_multi_gather_maps: the root needs output buffers");
  // 1. every local keyframe packs its final maps on its own stream; the communication stream waits for it
  for(int i = 0; i < m->n_local; ++i)
  {
    rmd_seeds_t *s = seeds[i];
    RMD_REQUIRE(s, "rmd_multi_gather_maps: null seeds handle");
    int w = 0, h = 0;
    rmd_seeds_size(s, &w, &h, NULL);
    RMD_REQUIRE(w == m->width && h == m->height, "rmd_multi_gather_maps: image size differs from the communicator's");
    DeviceGuard guard(m->device[i]);
    void *seeds_stream = NULL;
    int rc = rmd_seeds_get_stream(s, &seeds_stream);
    if(rc) return rc;
    if(dev_depth && dev_depth[i])
    {
      // e.g. the denoised map (rmd_denoiser_run_seeds_to_device); the caller has synchronised its producer
      RMD_CUDA_TRY(cudaMemcpy2DAsync(m->send[i], sizeof(float) * (size_t)w, dev_depth[i], dev_depth_pitch[i],
                                     sizeof(float) * (size_t)w, h, cudaMemcpyDeviceToDevice,
                                     (cudaStream_t)seeds_stream));
    }
    else
    {
      rc = rmd_seeds_copy_field_to_device(s, RMD_FIELD_MU, m->send[i], sizeof(float) * (size_t)w);
      if(rc) return rc;
    }
    rc = rmd_seeds_copy_field_to_device(s, RMD_FIELD_CONVERGENCE, m->send[i] + px, sizeof(int32_t) * (size_t)w);
    if(rc) return rc;
    RMD_CUDA_TRY(cudaEventRecord(m->ready[i], (cudaStream_t)seeds_stream));
    RMD_CUDA_TRY(cudaStreamWaitEvent(m->stream[i], m->ready[i], 0));
    if(i == root_local && have_root && !m->recv[i])
      RMD_CUDA_TRY(cudaMalloc(&m->recv[i], sizeof(int32_t) * count * (size_t)m->n_ranks));
  }
  // 2. the gather: one grouped exchange, every rank sends, the root receives n_ranks blocks
  RMD_NCCL_TRY(api->GroupStart());
  for(int i = 0; i < m->n_local; ++i)
  {
    const ncclResult_t res = api->Send(m->send[i], count, kNcclInt32, root, m->comm[i], m->stream[i]);
    if(res != 0) { api->GroupEnd(); return fail_nccl(api, res, "ncclSend"); }
  }
  if(have_root)
  {
    for(int q = 0; q < m->n_ranks; ++q)
    {
      const ncclResult_t res = api->Recv(m->recv[root_local] + (size_t)q * count, count, kNcclInt32, q,
                                         m->comm[root_local], m->stream[root_local]);
      if(res != 0) { api->GroupEnd(); return fail_nccl(api, res, "ncclRecv"); }
    }
  }
  RMD_NCCL_TRY(api->GroupEnd());
  // 3. root: [depth | convergence] blocks -> the caller's two arrays (rank-major)
  if(have_root)
  {
    DeviceGuard guard(m->device[root_local]);
    for(int q = 0; q < m->n_ranks; ++q)
    {
      const int32_t *blk = m->recv[root_local] + (size_t)q * count;
      RMD_CUDA_TRY(cudaMemcpyAsync(host_depth + (size_t)q * px, blk, sizeof(float) * px, cudaMemcpyDeviceToHost,
                                   m->stream[root_local]));
      RMD_CUDA_TRY(cudaMemcpyAsync(host_conv + (size_t)q * px, blk + px, sizeof(int32_t) * px, cudaMemcpyDeviceToHost,
                                   m->stream[root_local]));
    }
  }
  for(int i = 0; i < m->n_local; ++i)
  {
    DeviceGuard guard(m->device[i]);
    RMD_CUDA_TRY(cudaStreamSynchronize(m->stream[i]));   // the send buffers are free again; the root's output is complete
  }
  return 0;
}

} // extern "C"
