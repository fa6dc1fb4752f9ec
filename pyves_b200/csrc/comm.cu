// The one collective of the path behind the C ABI: the sum of the moment vector [sum_w, sum_w f (K), sum_w f f (K*K)?]
// over the ranks that share one ensemble (the reference has no multi-walker code at all).  NCCL is bound at
// run time (dlopen of libnccl.so.2: inside a torch process that is the library torch itself loaded, elsewhere the
// system one), so the shared library has no link-time dependency and single-GPU users never touch it.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

#include <cuda_runtime.h>

#include "../../include/pyves_b200.h"

namespace {

// the few NCCL declarations used here (nccl.h, stable since 2.0)
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId_ {
  char internal[128];
};
constexpr int kNcclSuccess = 0, kNcclFloat64 = 8, kNcclSum = 0;
typedef int (*fn_get_unique_id)(ncclUniqueId_*);
typedef int (*fn_comm_init_rank)(ncclComm_t*, int, ncclUniqueId_, int);
typedef int (*fn_comm_destroy)(ncclComm_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char* (*fn_error_string)(int);
typedef int (*fn_get_version)(int*);

struct Nccl {
  void* lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_error_string error_string = nullptr;
  fn_get_version get_version = nullptr;
  std::string why;
};

Nccl& nccl() {
  static Nccl N;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      N.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (N.lib) break;
    }
    if (!N.lib) {
      N.why = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    N.get_unique_id = (fn_get_unique_id)dlsym(N.lib, "ncclGetUniqueId");
    N.comm_init_rank = (fn_comm_init_rank)dlsym(N.lib, "ncclCommInitRank");
    N.comm_destroy = (fn_comm_destroy)dlsym(N.lib, "ncclCommDestroy");
    N.all_reduce = (fn_all_reduce)dlsym(N.lib, "ncclAllReduce");
    N.error_string = (fn_error_string)dlsym(N.lib, "ncclGetErrorString");
    N.get_version = (fn_get_version)dlsym(N.lib, "ncclGetVersion");
    if (!N.get_unique_id || !N.comm_init_rank || !N.comm_destroy || !N.all_reduce || !N.error_string) {
      N.why = "libnccl.so.2 lacks a required symbol";
      N.lib = nullptr;
    }
  });
  return N;
}

thread_local std::string g_comm_err;

int comm_fail(int code, const std::string& msg) {
  g_comm_err = msg;
  return code;
}

}  // namespace

struct pyves_comm {
  ncclComm_t comm = nullptr;
  int device = 0, nranks = 1, rank = 0;
};

extern "C" {

const char* pyves_comm_last_error(void) { return g_comm_err.c_str(); }

int pyves_comm_unique_id(void* id128) {
  if (!id128) return comm_fail(PYVES_EINVAL, "id128 is NULL");
  Nccl& N = nccl();
  if (!N.lib) return comm_fail(PYVES_EUNSUPPORTED, N.why);
  ncclUniqueId_ id;
  const int rc = N.get_unique_id(&id);
  if (rc != kNcclSuccess) return comm_fail(PYVES_ECUDA, std::string("ncclGetUniqueId: ") + N.error_string(rc));
  std::memcpy(id128, id.internal, sizeof(id.internal));
  return PYVES_OK;
}

int pyves_comm_create(pyves_comm_t** out, int device, int nranks, int rank, const void* id128) {
  if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return comm_fail(PYVES_EINVAL, "comm_create: bad arguments");
  *out = nullptr;
  Nccl& N = nccl();
  if (!N.lib) return comm_fail(PYVES_EUNSUPPORTED, N.why);
  int prev = -1;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) {
    cudaGetLastError();
    return comm_fail(PYVES_EINVAL, "invalid device ordinal");
  }
  ncclUniqueId_ id;
  std::memcpy(id.internal, id128, sizeof(id.internal));
  pyves_comm* c = new pyves_comm();
  c->device = device;
  c->nranks = nranks;
  c->rank = rank;
  const int rc = N.comm_init_rank(&c->comm, nranks, id, rank);
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  if (rc != kNcclSuccess) {
    delete c;
    return comm_fail(PYVES_ECUDA, std::string("ncclCommInitRank: ") + N.error_string(rc));
  }
  *out = c;
  return PYVES_OK;
}

int pyves_comm_destroy(pyves_comm_t* c) {
  if (!c) return PYVES_OK;
  Nccl& N = nccl();
  if (N.lib && c->comm) N.comm_destroy(c->comm);
  delete c;
  return PYVES_OK;
}

int pyves_allreduce_moments(pyves_comm_t* c, double* moments_dev, int64_t count, void* stream) {
  if (!c || !moments_dev || count < 1) return comm_fail(PYVES_EINVAL, "allreduce_moments: bad arguments");
  Nccl& N = nccl();
  if (!N.lib) return comm_fail(PYVES_EUNSUPPORTED, N.why);
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != c->device) cudaSetDevice(c->device);
  const int rc = N.all_reduce(moments_dev, moments_dev, (size_t)count, kNcclFloat64, kNcclSum, c->comm, static_cast<cudaStream_t>(stream));
  if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  if (rc != kNcclSuccess) return comm_fail(PYVES_ECUDA, std::string("ncclAllReduce: ") + N.error_string(rc));
  return PYVES_OK;
}

int pyves_comm_nccl_version(int* version) {
  if (!version) return PYVES_EINVAL;
  Nccl& N = nccl();
  if (!N.lib || !N.get_version) return comm_fail(PYVES_EUNSUPPORTED, N.lib ? "ncclGetVersion missing" : N.why);
  return N.get_version(version) == kNcclSuccess ? PYVES_OK : PYVES_ECUDA;
}

}  // extern "C"
