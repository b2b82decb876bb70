// nccl_dyn.h -- NCCL bound at run time (dlopen), so that librscape_b200.so has no link-time dependency on it: a single-GPU host
// never needs it, a torch process reuses the libnccl.so.2 torch has already loaded, a plain C host gets the system's.
// Only the handful of entry points the covariation path uses; their ABI has been stable since NCCL 2.0 (nccl.h).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace rsb_nccl {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;             // NCCL_UNIQUE_ID_BYTES = 128
enum { ncclSuccess = 0 };
enum { ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };         // ncclDataType_t
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };    // ncclRedOp_t

struct Api {
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;     // optional (peer-memory setup)
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;  // optional (rsb_pool_broadcast)
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
  char why[256] = "";
};

inline Api &api()
{
  static Api a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { snprintf(a.why, sizeof(a.why), "libnccl.so.2 not found (%s)", dlerror()); return a; }
  a.GetUniqueId    = (int (*)(ncclUniqueId *)) dlsym(h, "ncclGetUniqueId");
  a.CommInitRank   = (int (*)(ncclComm_t *, int, ncclUniqueId, int)) dlsym(h, "ncclCommInitRank");
  a.CommInitAll    = (int (*)(ncclComm_t *, int, const int *)) dlsym(h, "ncclCommInitAll");
  a.CommDestroy    = (int (*)(ncclComm_t)) dlsym(h, "ncclCommDestroy");
  a.AllReduce      = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t)) dlsym(h, "ncclAllReduce");
  a.AllGather      = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t)) dlsym(h, "ncclAllGather");
  a.Broadcast      = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t)) dlsym(h, "ncclBroadcast");
  a.GroupStart     = (int (*)()) dlsym(h, "ncclGroupStart");
  a.GroupEnd       = (int (*)()) dlsym(h, "ncclGroupEnd");
  a.GetErrorString = (const char *(*)(int)) dlsym(h, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString;
  if (!a.ok) snprintf(a.why, sizeof(a.why), "libnccl.so.2 lacks an expected entry point");
  return a;
}

} // namespace rsb_nccl
