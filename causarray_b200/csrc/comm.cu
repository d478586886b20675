#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "../../include/causarray_b200.h"
#include "comm.h"
#include "common.cuh"

namespace ca {

namespace {
struct Api {
  void* lib = nullptr;
  ncclResult_t (*get_unique_id)(ncclUniqueId*) = nullptr;
  ncclResult_t (*comm_init_rank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*all_reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*comm_destroy)(ncclComm_t) = nullptr;
  const char* (*error_string)(ncclResult_t) = nullptr;
  bool tried = false;
};
Api g_api;

int load_api() {
  if (g_api.lib) return 0;
  if (g_api.tried) {
    set_error("NCCL is not available (libnccl.so.2 could not be loaded)");
    return -1;
  }
  g_api.tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_api.lib) break;
  }
  if (!g_api.lib) {
    set_error("NCCL is not available: %s", dlerror());
    return -1;
  }
  g_api.get_unique_id = (decltype(g_api.get_unique_id))dlsym(g_api.lib, "ncclGetUniqueId");
  g_api.comm_init_rank = (decltype(g_api.comm_init_rank))dlsym(g_api.lib, "ncclCommInitRank");
  g_api.all_reduce = (decltype(g_api.all_reduce))dlsym(g_api.lib, "ncclAllReduce");
  g_api.comm_destroy = (decltype(g_api.comm_destroy))dlsym(g_api.lib, "ncclCommDestroy");
  g_api.error_string = (decltype(g_api.error_string))dlsym(g_api.lib, "ncclGetErrorString");
  if (!g_api.get_unique_id || !g_api.comm_init_rank || !g_api.all_reduce || !g_api.comm_destroy) {
    set_error("libnccl lacks a required symbol");
    g_api.lib = nullptr;
    return -1;
  }
  return 0;
}

int nccl_fail(ncclResult_t r, const char* what) {
  set_error("NCCL error in %s: %s", what, g_api.error_string ? g_api.error_string(r) : "?");
  return -1;
}
}  // namespace

int comm_unique_id(unsigned char out[128]) {
  if (load_api()) return -1;
  ncclUniqueId id;
  ncclResult_t r = g_api.get_unique_id(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  static_assert(sizeof(id) == 128, "ncclUniqueId size");
  memcpy(out, &id, 128);
  return 0;
}

int comm_init(void** comm, int rank, int world, const unsigned char idb[128]) {
  if (load_api()) return -1;
  ncclUniqueId id;
  memcpy(&id, idb, 128);
  ncclComm_t c = nullptr;
  ncclResult_t r = g_api.comm_init_rank(&c, world, id, rank);
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
  *comm = c;
  return 0;
}

void comm_destroy(void* comm) {
  if (comm && g_api.comm_destroy) g_api.comm_destroy((ncclComm_t)comm);
}

int comm_allreduce_sum(void* comm, double* buf, size_t count, cudaStream_t st) {
  if (!comm || count == 0) return 0;
  ncclResult_t r = g_api.all_reduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)comm, st);
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
  return 0;
}

}  // namespace ca
