// ds_comm_* / ds_allgather: the single collective of the sharded sampling job (one all-gather of the rank-local
// waveforms, text2sound's batch split over GPUs) on NCCL.  NCCL is bound at run time with dlopen so that the library carries no
// link-time dependency and shares the copy a host process has already loaded (PyTorch ships its own libnccl.so.2).
#include "common.cuh"
#include "../../include/diffusynth_b200.h"
#include <dlfcn.h>

namespace ds {
namespace {
struct NcclId { char internal[128]; };
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(NcclId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static Nccl* nccl() {
  static Nccl n;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy already in the process, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      n.lib = h;
      n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
      n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
      n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(h, "ncclAllGather"));
      n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
      n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    }
  }
  return (n.lib && n.GetUniqueId && n.CommInitRank && n.AllGather && n.CommDestroy) ? &n : nullptr;
}
}  // namespace
}  // namespace ds

struct ds_comm { ds::ncclComm_t comm; int rank, world; };

#define DS_CHECK_NCCL(N, expr)                                                                                   \
  do {                                                                                                           \
    int r_ = (expr);                                                                                             \
    if (r_ != 0) {                                                                                               \
      ds::set_error("%s -> NCCL error %d (%s)", #expr, r_, (N)->GetErrorString ? (N)->GetErrorString(r_) : "?"); \
      return ds::DS_ERR_NCCL;                                                                                    \
    }                                                                                                            \
  } while (0)

extern "C" {

int ds_comm_unique_id(void* id128) {
  DS_REQUIRE(id128 != nullptr, "ds_comm_unique_id: null pointer");
  ds::Nccl* n = ds::nccl();
  if (!n) { ds::set_error("ds_comm_unique_id: libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing"); return ds::DS_ERR_NCCL; }
  DS_CHECK_NCCL(n, n->GetUniqueId(reinterpret_cast<ds::NcclId*>(id128)));
  return ds::DS_OK;
}

int ds_comm_init(int rank, int world, const void* id128, ds_comm** out) {
  DS_REQUIRE(out && id128 && world >= 1 && rank >= 0 && rank < world, "ds_comm_init: bad arguments (rank %d of %d)", rank, world);
  ds::Nccl* n = ds::nccl();
  if (!n) { ds::set_error("ds_comm_init: libnccl.so.2 could not be loaded"); return ds::DS_ERR_NCCL; }
  ds::NcclId id;
  memcpy(&id, id128, sizeof(id));
  ds::ncclComm_t c = nullptr;
  DS_CHECK_NCCL(n, n->CommInitRank(&c, world, id, rank));
  *out = new ds_comm{c, rank, world};
  return ds::DS_OK;
}

int ds_allgather(ds_comm* c, const void* d_send, void* d_recv, long long count, int dtype, void* stream) {
  DS_REQUIRE(c && d_send && d_recv && count > 0, "ds_allgather: bad arguments");
  ds::Nccl* n = ds::nccl();
  if (!n) { ds::set_error("ds_allgather: NCCL not loaded"); return ds::DS_ERR_NCCL; }
  // ncclDataType_t: ncclInt64 = 4, ncclFloat16 = 6, ncclFloat32 = 7, ncclBfloat16 = 9
  int nt;
  if (dtype == 0) nt = 7;
  else if (dtype == 1) nt = ds::kOperandIsFp16 ? 6 : 9;
  else if (dtype == 2) nt = 4;
  else { ds::set_error("ds_allgather: dtype %d (0 = fp32, 1 = act16, 2 = int64)", dtype); return ds::DS_ERR_INVALID; }
  DS_CHECK_NCCL(n, n->AllGather(d_send, d_recv, (size_t)count, nt, c->comm, (cudaStream_t)stream));
  return ds::DS_OK;
}

void ds_comm_destroy(ds_comm* c) {
  if (!c) return;
  ds::Nccl* n = ds::nccl();
  if (n && c->comm) n->CommDestroy(c->comm);
  delete c;
}

}  // extern "C"
