// nccl_dyn.cpp — lazy binding of the handful of NCCL entry points the library needs.
// The only collective on the hot path is one 4-float ncclMin all-reduce ({min, -max} per metric
// component) for the global-normalisation extension ("multi-GPU").
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "pxz_host.h"

namespace pxz {

// minimal declarations (ABI-stable across NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };
enum { ncclMinOp = 3 };

struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*);
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(int);
};

static NcclApi g_api;
static bool g_ok = false;
static std::string g_err;
static std::once_flag g_once;

static void load_once() {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    g_err = std::string("cannot dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
    return;
  }
  g_api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
  g_api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
  g_api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_api.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
  g_api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.AllReduce || !g_api.CommDestroy) {
    g_err = "libnccl is missing a required symbol";
    return;
  }
  g_ok = true;
}

const NcclApi* nccl_api(std::string* err) {
  std::call_once(g_once, load_once);
  if (!g_ok) {
    if (err) *err = g_err;
    return nullptr;
  }
  return &g_api;
}

static std::string nccl_msg(const NcclApi* a, const char* what, int rc) {
  return std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(rc) : "nccl error");
}

int nccl_get_unique_id(uint8_t id[PXZ_COMM_ID_BYTES], std::string* err) {
  const NcclApi* a = nccl_api(err);
  if (!a) return -1;
  ncclUniqueId u;
  int rc = a->GetUniqueId(&u);
  if (rc != ncclSuccess) { if (err) *err = nccl_msg(a, "ncclGetUniqueId", rc); return -1; }
  static_assert(sizeof(u) == PXZ_COMM_ID_BYTES, "id size");
  memcpy(id, &u, sizeof(u));
  return 0;
}

int nccl_comm_init(void** comm, int nranks, int rank, const uint8_t id[PXZ_COMM_ID_BYTES], std::string* err) {
  const NcclApi* a = nccl_api(err);
  if (!a) return -1;
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t c = nullptr;
  int rc = a->CommInitRank(&c, nranks, u, rank);
  if (rc != ncclSuccess) { if (err) *err = nccl_msg(a, "ncclCommInitRank", rc); return -1; }
  *comm = c;
  return 0;
}

int nccl_allreduce_min_f32(void* comm, float* dev_buf, size_t count, cudaStream_t stream, std::string* err) {
  const NcclApi* a = nccl_api(err);
  if (!a) return -1;
  int rc = a->AllReduce(dev_buf, dev_buf, count, ncclFloat32, ncclMinOp, (ncclComm_t)comm, stream);
  if (rc != ncclSuccess) { if (err) *err = nccl_msg(a, "ncclAllReduce", rc); return -1; }
  return 0;
}

void nccl_comm_destroy(void* comm) {
  const NcclApi* a = nccl_api(nullptr);
  if (a && comm) a->CommDestroy((ncclComm_t)comm);
}

}  // namespace pxz
