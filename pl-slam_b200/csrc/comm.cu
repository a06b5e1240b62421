// The one exchange step of the path: an all-gather of fixed-size per-frame pose records over NCCL
// (NVLink 5 / NVSwitch inside one box), so that the rank that runs the sequential Tracking logic sees every frame's pose
// in frame order.  Everything else is embarrassingly parallel over frames and never leaves its GPU.
//
// NCCL is resolved at run time from the libnccl.so.2 that is already in the process (torch's bundled copy when the host
// is a torchrun worker, the system library for a C++ host): the library links no NCCL, so a communicator created here and
// the host's own NCCL are always the same implementation.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>

namespace {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclSuccess = 0, ncclFloat32 = 7 };
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
Nccl g_nccl;
std::mutex g_mu;

int load_nccl() {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_nccl.AllGather) return PL_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);   // an already loaded libnccl.so.2 (torch's) is returned as is
  if (!h) { pl::set_error("NCCL: cannot load libnccl.so.2 (%s)", dlerror()); return PL_ERR_CUDA; }
  g_nccl.lib = h;
  *(void**)&g_nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
  *(void**)&g_nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
  *(void**)&g_nccl.AllGather = dlsym(h, "ncclAllGather");
  *(void**)&g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
  *(void**)&g_nccl.GetErrorString = dlsym(h, "ncclGetErrorString");
  *(void**)&g_nccl.GetVersion = dlsym(h, "ncclGetVersion");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
    g_nccl.AllGather = nullptr;
    pl::set_error("NCCL: libnccl.so.2 lacks the expected entry points");
    return PL_ERR_CUDA;
  }
  return PL_OK;
}
#define PL_NCCL(expr)                                                                                        \
  do {                                                                                                       \
    int _r = (expr);                                                                                         \
    if (_r != ncclSuccess) {                                                                                 \
      pl::set_error("%s:%d %s -> NCCL error %d (%s)", __FILE__, __LINE__, #expr, _r,                         \
                    g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");                                \
      return PL_ERR_CUDA;                                                                                    \
    }                                                                                                        \
  } while (0)
}  // namespace

struct PLComm { ncclComm_t comm = nullptr; int nranks = 0, rank = 0; };

extern "C" int pl_comm_unique_id(void* id128) {
  PL_ARG(id128);
  int rc = load_nccl(); if (rc) return rc;
  ncclUniqueId id;
  PL_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return PL_OK;
}
extern "C" int pl_comm_create(const void* id128, int nranks, int rank, PLComm** out) {
  PL_ARG(id128 && out && nranks >= 1 && rank >= 0 && rank < nranks);
  int rc = pl::require_device(); if (rc) return rc;
  if ((rc = load_nccl())) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  PLComm* c = new PLComm;
  c->nranks = nranks; c->rank = rank;
  int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { pl::set_error("ncclCommInitRank -> %d (%s)", r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); delete c; return PL_ERR_CUDA; }
  *out = c;
  return PL_OK;
}
extern "C" void pl_comm_destroy(PLComm* c) {
  if (!c) return;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
}
extern "C" int pl_nccl_version(void) {
  if (load_nccl()) return -1;
  int v = 0;
  if (g_nccl.GetVersion) g_nccl.GetVersion(&v);
  return v;
}
// every rank contributes floats_per_rank floats (its block of [frames][16] row-major Tcw, padded to the common block size);
// recv = [nranks][floats_per_rank] on every rank.  Asynchronous on `stream`.
extern "C" int pl_allgather_poses_nccl(void* nccl_comm, const float* send_dev, float* recv_dev, size_t floats_per_rank, void* stream) {
  PL_ARG(nccl_comm && send_dev && recv_dev && floats_per_rank > 0);
  int rc = load_nccl(); if (rc) return rc;
  PL_NCCL(g_nccl.AllGather(send_dev, recv_dev, floats_per_rank, ncclFloat32, (ncclComm_t)nccl_comm, (cudaStream_t)stream));
  pl::count_launch();
  return PL_OK;
}
extern "C" int pl_allgather_poses(PLComm* c, const float* send_dev, float* recv_dev, size_t floats_per_rank, void* stream) {
  PL_ARG(c && c->comm);
  return pl_allgather_poses_nccl(c->comm, send_dev, recv_dev, floats_per_rank, stream);
}
