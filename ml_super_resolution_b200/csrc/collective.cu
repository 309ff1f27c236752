// The data-parallel exchange step behind the C ABI: one NCCL sum all-reduce of the flat fp32 gradient
// arena per training step over NVLink / NVSwitch.  NCCL is bound at run time (dlopen of the library the host process already
// uses -- torch bundles one), so libsrk has no link-time dependency on it and single-GPU users never load it.
// Absent in the reference (one P100, vdsr/README.md:14); the call site it extends is the `minimize` of
// vdsr/vdsr/model_vdsr.py:146-148, whose gradients are summed across ranks before the optimiser applies them.
#include <dlfcn.h>

#include "srk_common.cuh"

namespace srk {

// the handful of NCCL entry points used, with the library's own C types spelled out (no nccl.h at build time)
struct NcclApi {
  int (*GetUniqueId)(void* id128);
  int (*CommInitRank)(void** comm, int nranks, const void* id128_by_value_hack, int rank);
  int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t stream);
  int (*CommDestroy)(void* comm);
  const char* (*GetErrorString)(int);
};

struct NcclId {
  char bytes[128];
};
typedef int (*CommInitRankFn)(void** comm, int nranks, NcclId id, int rank);  // ncclUniqueId is passed BY VALUE

static void* g_lib = nullptr;
static NcclApi g_api{};
static CommInitRankFn g_init = nullptr;

static int load_nccl() {
  if (g_lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_lib) break;
  }
  SRK_REQUIRE(g_lib != nullptr, "srk_comm_init: cannot load libnccl.so.2 (%s); import torch first or put NCCL on the loader path", dlerror());
  g_api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(g_lib, "ncclGetUniqueId"));
  g_init = reinterpret_cast<CommInitRankFn>(dlsym(g_lib, "ncclCommInitRank"));
  g_api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(g_lib, "ncclAllReduce"));
  g_api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(g_lib, "ncclCommDestroy"));
  g_api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(g_lib, "ncclGetErrorString"));
  SRK_REQUIRE(g_api.GetUniqueId && g_init && g_api.AllReduce && g_api.CommDestroy && g_api.GetErrorString, "srk_comm_init: NCCL symbols missing");
  return 0;
}

#define SRK_CHECK_NCCL(expr)                                                                              \
  do {                                                                                                    \
    int r_ = (expr);                                                                                      \
    if (r_ != 0) {                                                                                        \
      ::srk::set_error("%s failed: %s (%s:%d)", #expr, ::srk::g_api.GetErrorString(r_), __FILE__, __LINE__); \
      return -4;                                                                                          \
    }                                                                                                     \
  } while (0)

}  // namespace srk

using namespace srk;

extern "C" int srk_comm_unique_id(void* id128) {
  SRK_REQUIRE(id128 != nullptr, "srk_comm_unique_id: null argument");
  if (int rc = load_nccl()) return rc;
  SRK_CHECK_NCCL(g_api.GetUniqueId(id128));
  return 0;
}

extern "C" int srk_comm_init(srk_handle_t h, const void* id128, int rank, int world) {
  SRK_REQUIRE(h && id128 && world >= 1 && rank >= 0 && rank < world, "srk_comm_init: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (int rc = load_nccl()) return rc;
  if (h->comm) {
    SRK_CHECK_NCCL(g_api.CommDestroy(h->comm));
    h->comm = nullptr;
  }
  NcclId id;
  memcpy(id.bytes, id128, sizeof id.bytes);
  void* comm = nullptr;
  SRK_CHECK_NCCL(g_init(&comm, world, id, rank));
  h->comm = comm;
  h->comm_world = world;
  return 0;
}

extern "C" int srk_allreduce_grads(srk_handle_t h, float* flat, size_t n, srk_stream_t stream) {
  SRK_REQUIRE(h && flat, "srk_allreduce_grads: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE(h->comm != nullptr, "srk_allreduce_grads: srk_comm_init has not been called on this handle");
  if (n == 0 || h->comm_world == 1) return 0;
  SRK_CHECK_NCCL(g_api.AllReduce(flat, flat, n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, h->comm, as_stream(stream)));
  return 0;
}

extern "C" int srk_comm_destroy(srk_handle_t h) {
  SRK_REQUIRE(h != nullptr, "srk_comm_destroy: null handle");
  if (h->comm) {
    SRK_CHECK_NCCL(g_api.CommDestroy(h->comm));
    h->comm = nullptr;
  }
  return 0;
}
