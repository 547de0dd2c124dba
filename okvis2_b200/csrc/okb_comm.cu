// okb_comm.cu -- the one collective of the path behind the C ABI: NCCL all-gather of the fixed-capacity feature blocks of
// the cameras that live on different GPUs (BASELINE config 4: cameras of an NCameraSystem sharded one per GPU,
// descriptors exchanged only where Frontend::matchStereo needs the peer camera, reference Frontend.cpp:1990-2074).
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": inside a PyTorch process that is torch's own copy, already mapped; in a
// plain C++ host the system library), so libokvis_b200.so has no link-time dependency on it and single-GPU hosts never touch it.
// Two ways to form the communicator:
//   okb_comm_init_all   one process drives all GPUs (the reference is one process; ncclCommInitAll)
//   okb_comm_init_rank  one process per GPU (torchrun): rank 0 makes the id with okb_comm_unique_id and shares it out of band
// The all-gather is enqueued on the camera's own stream right behind okb_export_features, so it overlaps the map / motion
// matchers of the other cameras; the stereo matcher's stream waits on the event okb_comm_wait places.
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; every entry point is resolved with dlsym
#include <string.h>

#include <vector>

#include "okb_internal.h"

namespace okb {
namespace {
struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
Nccl g_nccl;
std::mutex g_nccl_mutex;

bool load_nccl()
{
  std::lock_guard<std::mutex> lk(g_nccl_mutex);
  if (g_nccl.ok) return true;
  if (!g_nccl.handle) {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      g_nccl.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
    }
  }
  if (!g_nccl.handle) { set_error("NCCL: libnccl.so.2 cannot be loaded (%s)", dlerror()); return false; }
  auto sym = [&](const char* n) { return dlsym(g_nccl.handle, n); };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.AllGather && g_nccl.GroupStart &&
              g_nccl.GroupEnd && g_nccl.GetErrorString;
  if (!g_nccl.ok) set_error("NCCL: libnccl.so.2 lacks an expected entry point");
  return g_nccl.ok;
}
}  // namespace
}  // namespace okb

// one communicator handle: n local ranks (n = 1 in the one-process-per-GPU mode)
struct okb_comm {
  int world = 0;
  std::vector<int> device, rank;
  std::vector<ncclComm_t> comm;
  std::vector<cudaEvent_t> done;   // recorded behind the last all-gather of each local rank
};

using namespace okb;

#define OKB_NCCL(call)                                                                              \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess) { set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); return OKB_ERR_NCCL; } \
  } while (0)

extern "C" {

int okb_comm_unique_id(void* id128)
{
  if (!id128) { set_error("okb_comm_unique_id: null"); return OKB_ERR_ARGUMENT; }
  if (!load_nccl()) return OKB_ERR_NCCL;
  ncclUniqueId id;
  OKB_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return OKB_OK;
}

int okb_comm_init_all(int n_devices, const int* devices, okb_comm_t** out)
{
  if (!out || n_devices < 1 || !devices) { set_error("okb_comm_init_all: bad arguments"); return OKB_ERR_ARGUMENT; }
  *out = nullptr;
  if (!load_nccl()) return OKB_ERR_NCCL;
  okb_comm* c = new okb_comm();
  c->world = n_devices; c->device.assign(devices, devices + n_devices); c->comm.resize(n_devices); c->done.resize(n_devices);
  for (int i = 0; i < n_devices; i++) c->rank.push_back(i);
  ncclResult_t r = g_nccl.CommInitAll(c->comm.data(), n_devices, devices);
  if (r != ncclSuccess) { set_error("ncclCommInitAll: %s", g_nccl.GetErrorString(r)); delete c; return OKB_ERR_NCCL; }
  for (int i = 0; i < n_devices; i++) {
    cudaSetDevice(devices[i]);
    if (cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming) != cudaSuccess) { set_error("okb_comm_init_all: event"); delete c; return OKB_ERR_CUDA; }
  }
  *out = c;
  return OKB_OK;
}

int okb_comm_init_rank(int world, int rank, const void* id128, int device, okb_comm_t** out)
{
  if (!out || world < 1 || rank < 0 || rank >= world || !id128) { set_error("okb_comm_init_rank: bad arguments"); return OKB_ERR_ARGUMENT; }
  *out = nullptr;
  if (!load_nccl()) return OKB_ERR_NCCL;
  OKB_CUDA(cudaSetDevice(device));
  okb_comm* c = new okb_comm();
  c->world = world; c->device.push_back(device); c->rank.push_back(rank); c->comm.resize(1); c->done.resize(1);
  ncclUniqueId id; memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&c->comm[0], world, id, rank);
  if (r != ncclSuccess) { set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r)); delete c; return OKB_ERR_NCCL; }
  if (cudaEventCreateWithFlags(&c->done[0], cudaEventDisableTiming) != cudaSuccess) { set_error("okb_comm_init_rank: event"); delete c; return OKB_ERR_CUDA; }
  *out = c;
  return OKB_OK;
}

void okb_comm_destroy(okb_comm_t* c)
{
  if (!c) return;
  for (size_t i = 0; i < c->comm.size(); i++) {
    cudaSetDevice(c->device[i]);
    if (c->done[i]) cudaEventDestroy(c->done[i]);
    if (c->comm[i] && g_nccl.ok) g_nccl.CommDestroy(c->comm[i]);
  }
  delete c;
}

int okb_comm_world(const okb_comm_t* c) { return c ? c->world : 0; }
int okb_comm_local_ranks(const okb_comm_t* c) { return c ? (int)c->comm.size() : 0; }

int okb_allgather_features(okb_comm_t* c, int n_local, okb_context_t* const* ctxs, const int* cams, const void* const* d_send, void* const* d_recv,
                           size_t bytes_per_rank)
{
  if (!c || n_local != (int)c->comm.size() || !ctxs || !cams || !d_send || !d_recv || bytes_per_rank == 0) {
    set_error("okb_allgather_features: bad arguments (%d local ranks expected)", c ? (int)c->comm.size() : -1); return OKB_ERR_ARGUMENT;
  }
  for (int i = 0; i < n_local; i++)
    if (!ctxs[i] || ctxs[i]->device != c->device[i] || cams[i] < 0 || cams[i] >= ctxs[i]->n_cams || !d_send[i] || !d_recv[i]) {
      set_error("okb_allgather_features: local rank %d: context / camera / buffers do not belong to device %d", i, c->device[i]); return OKB_ERR_ARGUMENT;
    }
  if (n_local > 1) OKB_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n_local; i++) {
    OKB_CUDA(cudaSetDevice(c->device[i]));
    cudaStream_t st = ctxs[i]->cams[cams[i]].stream;   // behind okb_export_features of that camera
    OKB_NCCL(g_nccl.AllGather(d_send[i], d_recv[i], bytes_per_rank, ncclUint8, c->comm[i], st));
  }
  if (n_local > 1) OKB_NCCL(g_nccl.GroupEnd());
  for (int i = 0; i < n_local; i++) {
    OKB_CUDA(cudaSetDevice(c->device[i]));
    OKB_CUDA(cudaEventRecord(c->done[i], ctxs[i]->cams[cams[i]].stream));
  }
  return OKB_OK;
}

int okb_comm_wait(okb_comm_t* c, int local_rank, void* stream)
{
  if (!c || local_rank < 0 || local_rank >= (int)c->comm.size()) { set_error("okb_comm_wait: bad arguments"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(c->device[local_rank]));
  OKB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, c->done[local_rank], 0));
  return OKB_OK;
}

}  // extern "C"
