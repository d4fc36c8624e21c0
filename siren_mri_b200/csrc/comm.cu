// Gradient all-reduce of the C ABI: one ncclAllReduce(sum, fp32) over the flat gradient buffer.
// Replaces the bucketed DDP-Reducer path of train_mri_neural_process_ddp.py:238 /
// training_ddp.py:155-164 for the single-scene configurations.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2"): inside a PyTorch process that is the
// copy torch already loaded (same soname), so there is exactly one NCCL in the process and no
// link-time dependency.  The unique id is created on rank 0 and handed to the other ranks by the
// host (the Python side broadcasts it with torch.distributed).
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include <cuda_runtime.h>

#include "../../include/siren_b200.h"

namespace {

struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int /*dtype*/, int /*op*/, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char* (*ErrStrFn)(int);

struct NcclApi {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  ErrStrFn err = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) return;
    a.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(a.handle, "ncclGetUniqueId"));
    a.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(a.handle, "ncclCommInitRank"));
    a.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(a.handle, "ncclAllReduce"));
    a.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(a.handle, "ncclCommDestroy"));
    a.err = reinterpret_cast<ErrStrFn>(dlsym(a.handle, "ncclGetErrorString"));
    a.ok = a.get_unique_id && a.comm_init_rank && a.all_reduce && a.comm_destroy;
  });
  return a;
}

thread_local char g_comm_err[256] = "";

int comm_fail(const char* what, int rc) {
  NcclApi& a = api();
  snprintf(g_comm_err, sizeof(g_comm_err), "%s: %s", what, (a.err && rc > 0) ? a.err(rc) : "NCCL not available");
  return SIREN_ERR_CUDA;
}

}  // namespace

extern "C" {

const char* siren_b200_comm_last_error(void) { return g_comm_err; }

int siren_b200_comm_unique_id(void* id_out) {
  NcclApi& a = api();
  if (!a.ok || !id_out) return comm_fail("siren_b200_comm_unique_id", -1);
  NcclUniqueId id;
  const int rc = a.get_unique_id(&id);
  if (rc != 0) return comm_fail("ncclGetUniqueId", rc);
  memcpy(id_out, &id, sizeof(id));
  return SIREN_OK;
}

int siren_b200_comm_init(int rank, int world, const void* id_in, void** comm_out) {
  NcclApi& a = api();
  if (!a.ok || !id_in || !comm_out || world < 1 || rank < 0 || rank >= world)
    return comm_fail("siren_b200_comm_init", -1);
  NcclUniqueId id;
  memcpy(&id, id_in, sizeof(id));
  NcclComm comm = nullptr;
  const int rc = a.comm_init_rank(&comm, world, id, rank);
  if (rc != 0) return comm_fail("ncclCommInitRank", rc);
  *comm_out = comm;
  return SIREN_OK;
}

int siren_b200_allreduce(void* comm, float* buf, long n, void* stream) {
  NcclApi& a = api();
  if (!a.ok || !comm || !buf || n <= 0) return comm_fail("siren_b200_allreduce", -1);
  const int rc = a.all_reduce(buf, buf, size_t(n), 7 /*ncclFloat32*/, 0 /*ncclSum*/, comm,
                              reinterpret_cast<cudaStream_t>(stream));
  if (rc != 0) return comm_fail("ncclAllReduce", rc);
  return SIREN_OK;
}

int siren_b200_comm_destroy(void* comm) {
  NcclApi& a = api();
  if (!a.ok || !comm) return SIREN_OK;
  a.comm_destroy(comm);
  return SIREN_OK;
}

}  // extern "C"
