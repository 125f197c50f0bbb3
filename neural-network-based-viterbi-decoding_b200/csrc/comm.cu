// comm.cu — the one exchange step of the path: independent frames shard over GPUs by global frame
// range, and only the [n_snr][4] uint64 error counters are combined, with ncclAllReduce over NVLink, queued on the
// context's stream directly behind the kernels that produce them (no host bounce, no extra synchronisation).
//
// NCCL is bound at RUN time (dlopen), not linked: libnvd.so must load on a box without NCCL (single-GPU use), and a
// Python host that also imports torch already has torch's bundled libnccl.so.2 mapped — a second, different NCCL with
// the same soname must not be pulled in beside it.  Order: NVD_NCCL_LIB if set, the copy already in the process
// (RTLD_NOLOAD), the system library.  Only five entry points that have been stable since NCCL 2.0 are used.
#include <dlfcn.h>

#include <cstdlib>

#include "common.cuh"

namespace nvd {

namespace {

// the parts of nccl.h this file needs (ABI-stable since 2.0; checked against /usr/include/nccl.h 2.27.3)
struct NcclUniqueId { char internal[128]; };
static_assert(sizeof(NcclUniqueId) == NVD_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
typedef struct ncclComm* NcclComm;
constexpr int kNcclSuccess = 0;
constexpr int kNcclUint64 = 5;   // ncclDataType_t: int8 0, uint8 1, int32 2, uint32 3, int64 4, uint64 5
constexpr int kNcclSum = 0;      // ncclRedOp_t

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return NVD_OK;
  void* h = nullptr;
  if (const char* path = getenv("NVD_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy already mapped (e.g. torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("NCCL not found (%s); set NVD_NCCL_LIB to libnccl.so.2", dlerror());
    return NVD_ERR_UNSUPPORTED;
  }
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString) {
    set_error("libnccl.so.2 lacks a required entry point");
    return NVD_ERR_UNSUPPORTED;
  }
  g_nccl = a;
  return NVD_OK;
}

#define NVD_NCCL(expr)                                                                                  \
  do {                                                                                                  \
    const int _r = (expr);                                                                              \
    if (_r != kNcclSuccess) {                                                                           \
      ::nvd::set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), __FILE__, __LINE__);  \
      return NVD_ERR_CUDA;                                                                              \
    }                                                                                                   \
  } while (0)

}  // namespace

int comm_allreduce_u64(nvd_ctx* ctx, uint64_t* values_dev, size_t n) {
  if (!ctx->comm || ctx->comm_world <= 1 || n == 0) return NVD_OK;
  NVD_NCCL(g_nccl.AllReduce(values_dev, values_dev, n, kNcclUint64, kNcclSum, static_cast<NcclComm>(ctx->comm), ctx->stream));
  return NVD_OK;
}

void comm_release(nvd_ctx* ctx) {
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<NcclComm>(ctx->comm));
  ctx->comm = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
}

}  // namespace nvd

using namespace nvd;

extern "C" {

int nvd_comm_unique_id(void* id_out) {
  NVD_REQUIRE(id_out != nullptr, "null id_out");
  NVD_TRY(load_nccl());
  NcclUniqueId id;
  NVD_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return NVD_OK;
}

int nvd_comm_init(nvd_ctx* ctx, int rank, int world, const void* id) {
  NVD_REQUIRE(ctx != nullptr && id != nullptr, "null argument");
  NVD_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank %d / world %d out of range", rank, world);
  NVD_REQUIRE(ctx->comm == nullptr, "the context already has a communicator");
  NVD_CUDA(cudaSetDevice(ctx->device));
  NVD_TRY(load_nccl());
  NcclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  NcclComm comm = nullptr;
  NVD_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return NVD_OK;
}

int nvd_comm_destroy(nvd_ctx* ctx) {
  NVD_REQUIRE(ctx != nullptr, "null context");
  if (ctx->comm) {
    NVD_CUDA(cudaSetDevice(ctx->device));
    NVD_CUDA(cudaStreamSynchronize(ctx->stream));
    comm_release(ctx);
  }
  return NVD_OK;
}

int nvd_comm_info(nvd_ctx* ctx, int* rank, int* world, int* nccl_version) {
  NVD_REQUIRE(ctx != nullptr, "null context");
  if (rank) *rank = ctx->comm_rank;
  if (world) *world = ctx->comm_world;
  if (nccl_version) {
    *nccl_version = 0;
    if (g_nccl.GetVersion) g_nccl.GetVersion(nccl_version);
  }
  return NVD_OK;
}

int nvd_comm_allreduce_u64(nvd_ctx* ctx, uint64_t* values_dev, size_t n) {
  NVD_REQUIRE(ctx != nullptr && (n == 0 || values_dev != nullptr), "null argument");
  NVD_CUDA(cudaSetDevice(ctx->device));
  return comm_allreduce_u64(ctx, values_dev, n);
}

}  // extern "C"
