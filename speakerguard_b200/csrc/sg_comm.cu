// The one collective of the path: an all-reduce (sum) over NVLink / NVSwitch of a handful of metric scalars
// at the end of an attack - success count, sum of SNR / L2 / Linf, utterance count (reference metric/metric.py:10-42 semantics).
// Utterances shard independently, so there is no data-path collective to fuse with a kernel.
//
// NCCL is bound at run time with dlopen("libnccl.so.2"): inside a PyTorch process that resolves to the copy torch already
// loaded (one NCCL per process, no link-time dependency and no second copy), in a plain C host to the system library.
#include <dlfcn.h>
#include <string.h>

#include "sg_handle.cuh"

typedef struct { char internal[128]; } SgNcclUniqueId;          // ncclUniqueId: 128 opaque bytes (nccl.h NCCL_UNIQUE_ID_BYTES)
typedef void* SgNcclComm;
enum { SG_NCCL_SUM = 0, SG_NCCL_FLOAT64 = 8 };                    // ncclSum, ncclFloat64 (nccl.h; stable since NCCL 2.0)

struct SgNccl {
  void* lib = nullptr;
  int (*GetUniqueId)(SgNcclUniqueId*) = nullptr;
  int (*CommInitRank)(SgNcclComm*, int, SgNcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, SgNcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(SgNcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
static SgNccl g_nccl;

static int nccl_bind() {
  if (g_nccl.lib) return SG_OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { sg_set_error("sg_comm: libnccl.so.2 not found (%s)", dlerror()); return SG_EUNSUPPORTED; }
  SgNccl n;
  n.lib = lib;
  n.GetUniqueId = (int (*)(SgNcclUniqueId*))dlsym(lib, "ncclGetUniqueId");
  n.CommInitRank = (int (*)(SgNcclComm*, int, SgNcclUniqueId, int))dlsym(lib, "ncclCommInitRank");
  n.AllReduce = (int (*)(const void*, void*, size_t, int, int, SgNcclComm, cudaStream_t))dlsym(lib, "ncclAllReduce");
  n.CommDestroy = (int (*)(SgNcclComm))dlsym(lib, "ncclCommDestroy");
  n.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
  n.GetVersion = (int (*)(int*))dlsym(lib, "ncclGetVersion");
  if (!n.GetUniqueId || !n.CommInitRank || !n.AllReduce || !n.CommDestroy) { sg_set_error("sg_comm: NCCL symbols missing"); return SG_EUNSUPPORTED; }
  g_nccl = n;
  return SG_OK;
}
#define SG_NCCL_CHECK(expr)                                                                                                   \
  do {                                                                                                                        \
    int _r = (expr);                                                                                                          \
    if (_r != 0) { sg_set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error"); return SG_ECUDA; } \
  } while (0)

extern "C" int sg_comm_unique_id(void* id128) {
  if (!id128) { sg_set_error("sg_comm_unique_id: null output"); return SG_EINVAL; }
  SG_TRY(nccl_bind());
  SgNcclUniqueId id;
  SG_NCCL_CHECK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return SG_OK;
}

extern "C" int sg_comm_init(sg_handle* h, const void* id128, int rank, int world) {
  SG_TRY(sg_check_handle(h, false));
  if (!id128 || world < 1 || rank < 0 || rank >= world) { sg_set_error("sg_comm_init: bad argument (rank %d of %d)", rank, world); return SG_EINVAL; }
  if (h->comm) { sg_set_error("sg_comm_init: communicator already initialised on this handle"); return SG_ESTATE; }
  SG_TRY(nccl_bind());
  SgNcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  SgNcclComm c = nullptr;
  SG_NCCL_CHECK(g_nccl.CommInitRank(&c, world, id, rank));
  h->comm = c; h->comm_rank = rank; h->comm_world = world;
  return SG_OK;
}

extern "C" int sg_allreduce_metrics(sg_handle* h, double* v, int n, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!v || n < 1) { sg_set_error("sg_allreduce_metrics: bad argument"); return SG_EINVAL; }
  if (!h->comm) { sg_set_error("sg_allreduce_metrics: call sg_comm_init first"); return SG_ESTATE; }
  SG_NCCL_CHECK(g_nccl.AllReduce(v, v, (size_t)n, SG_NCCL_FLOAT64, SG_NCCL_SUM, (SgNcclComm)h->comm, (cudaStream_t)stream));
  return SG_OK;
}

extern "C" int sg_comm_destroy(sg_handle* h) {
  if (!h || !h->comm) return SG_OK;
  if (g_nccl.CommDestroy) g_nccl.CommDestroy((SgNcclComm)h->comm);
  h->comm = nullptr; h->comm_world = 0;
  return SG_OK;
}

extern "C" int sg_comm_nccl_version(void) {
  if (nccl_bind() != SG_OK || !g_nccl.GetVersion) return 0;
  int v = 0;
  return g_nccl.GetVersion(&v) == 0 ? v : 0;
}
