// comm.cu — the one exchange step of the path: independent frames shard over GPUs by global frame
// range, and only the [n_snr][4] uint64 error counters are combined over NVLink, queued on the context's stream
// directly behind the kernels that produce them (no host bounce, no extra synchronisation).  Two transports:
//   * peer memory (nvd_comm_p2p_*): every rank owns an 8 KB exchange block, exported with CUDA IPC and mapped by all
//     the others.  One small kernel adds the rank's values into EVERY rank's block with system-scope atomics over
//     NVLink / NVSwitch and then raises every rank's arrival counter; a second one waits until this rank's counter
//     shows that all ranks have arrived and hands the sums back.  ~10 us for 8 ranks, a fifth of an ncclAllReduce of
//     32 bytes; no communicator, no proxy thread.  Two alternating data areas make the protocol re-usable without a
//     barrier: the push of exchange e clears this rank's area of exchange e + 1 BEFORE it signals, and nobody can
//     start exchange e + 1 before it has seen that signal.
//   * ncclAllReduce (nvd_comm_init) where peer mapping is not available (GPUs hidden from each other, no P2P).
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

// ---- peer-memory exchange ---------------------------------------------------------------------------------------------
struct XchgPeers {
  uint64_t* p[kMaxPeers];
};
constexpr int kXchgFlag = 7168 / 8;          // uint64 index of the arrival counter inside a block
constexpr uint64_t kXchgWaitNs = 5000000000ull;

__global__ void __launch_bounds__(512) xchg_push_kernel(XchgPeers peers, int world, int rank, const uint64_t* __restrict__ values,
                                                        int n, int area) {
  const int tid = threadIdx.x;
  if (tid < kXchgValues) peers.p[rank][(area ^ 1) * kXchgValues + tid] = 0ull;     // the area of the NEXT exchange
  if (tid < n) {
    const uint64_t v = values[tid];
    for (int r = 0; r < world; ++r) atomicAdd_system(reinterpret_cast<unsigned long long*>(peers.p[r] + area * kXchgValues + tid), static_cast<unsigned long long>(v));
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) atomicAdd_system(reinterpret_cast<unsigned long long*>(peers.p[tid] + kXchgFlag), 1ull);
}

__global__ void __launch_bounds__(512) xchg_wait_kernel(uint64_t* local, unsigned long long target, uint64_t* __restrict__ values,
                                                        int n, int area, int* abort_flag) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    ok = 1;
    for (;;) {
      unsigned long long seen;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(local + kXchgFlag) : "memory");
      if (seen >= target) break;
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > kXchgWaitNs) { ok = 0; atomicExch(abort_flag, 11); break; }
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (ok && static_cast<int>(threadIdx.x) < n) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(local + area * kXchgValues + threadIdx.x) : "memory");
    values[threadIdx.x] = v;
  }
}

int xchg_allreduce_u64(nvd_ctx* ctx, uint64_t* values_dev, size_t n) {
  XchgPeers peers;
  for (int r = 0; r < kMaxPeers; ++r) peers.p[r] = ctx->xchg_peer[r];
  const int area = static_cast<int>(ctx->xchg_epoch & 1u);
  ctx->xchg_epoch += 1;
  xchg_push_kernel<<<1, 512, 0, ctx->stream>>>(peers, ctx->xchg_world, ctx->comm_rank, values_dev, static_cast<int>(n), area);
  NVD_CHECK_LAUNCH();
  xchg_wait_kernel<<<1, 512, 0, ctx->stream>>>(ctx->xchg_local, static_cast<unsigned long long>(ctx->xchg_epoch) * ctx->xchg_world,
                                               values_dev, static_cast<int>(n), area, ctx->abort_flag);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

void xchg_release(nvd_ctx* ctx) {
  for (int r = 0; r < ctx->xchg_world; ++r)
    if (ctx->xchg_peer[r] && ctx->xchg_peer[r] != ctx->xchg_local) cudaIpcCloseMemHandle(ctx->xchg_peer[r]);
  for (int r = 0; r < kMaxPeers; ++r) ctx->xchg_peer[r] = nullptr;
  ctx->xchg_world = 0;
  ctx->xchg_epoch = 0;
  cudaFree(ctx->xchg_local);
  ctx->xchg_local = nullptr;
}

int comm_allreduce_u64(nvd_ctx* ctx, uint64_t* values_dev, size_t n) {
  if (n == 0) return NVD_OK;
  if (ctx->xchg_world >= 1 && n <= static_cast<size_t>(kXchgValues)) return xchg_allreduce_u64(ctx, values_dev, n);
  if (ctx->xchg_world > 1 && !ctx->comm) {       // more values than an exchange block holds, and no second transport
    set_error("nvd_comm_allreduce_u64: %zu values exceed the peer exchange block (%d) and there is no NCCL communicator", n, kXchgValues);
    return NVD_ERR_UNSUPPORTED;
  }
  if (!ctx->comm || ctx->comm_world <= 1) return NVD_OK;
  NVD_NCCL(g_nccl.AllReduce(values_dev, values_dev, n, kNcclUint64, kNcclSum, static_cast<NcclComm>(ctx->comm), ctx->stream));
  return NVD_OK;
}

void comm_release(nvd_ctx* ctx) {
  xchg_release(ctx);
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

int nvd_comm_p2p_handle(nvd_ctx* ctx, void* handle_out) {
  NVD_REQUIRE(ctx != nullptr && handle_out != nullptr, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == NVD_COMM_IPC_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  NVD_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->xchg_local) {
    NVD_CUDA(cudaMalloc(&ctx->xchg_local, kXchgBytes));
    NVD_CUDA(cudaMemset(ctx->xchg_local, 0, kXchgBytes));
  }
  cudaIpcMemHandle_t h;
  NVD_CUDA(cudaIpcGetMemHandle(&h, ctx->xchg_local));
  memcpy(handle_out, &h, sizeof(h));
  return NVD_OK;
}

int nvd_comm_p2p_open(nvd_ctx* ctx, int rank, int world, const void* handles) {
  NVD_REQUIRE(ctx != nullptr && handles != nullptr, "null argument");
  NVD_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "rank %d / world %d out of range (max %d ranks)", rank, world, kMaxPeers);
  NVD_REQUIRE(ctx->xchg_local != nullptr, "call nvd_comm_p2p_handle first");
  NVD_REQUIRE(ctx->xchg_world == 0, "the peer exchange is already open");
  NVD_REQUIRE(ctx->comm == nullptr || (ctx->comm_rank == rank && ctx->comm_world == world), "rank / world differ from the communicator's");
  NVD_CUDA(cudaSetDevice(ctx->device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) { ctx->xchg_peer[r] = ctx->xchg_local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + static_cast<size_t>(r) * sizeof(h), sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int q = 0; q < r; ++q)
        if (q != rank && ctx->xchg_peer[q]) { cudaIpcCloseMemHandle(ctx->xchg_peer[q]); ctx->xchg_peer[q] = nullptr; }
      ctx->xchg_peer[rank] = nullptr;
      set_error("cannot map the exchange block of rank %d: %s", r, cudaGetErrorString(e));
      return NVD_ERR_UNSUPPORTED;
    }
    ctx->xchg_peer[r] = static_cast<uint64_t*>(p);
  }
  ctx->xchg_world = world;
  ctx->xchg_epoch = 0;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return NVD_OK;
}

int nvd_comm_p2p_close(nvd_ctx* ctx) {
  NVD_REQUIRE(ctx != nullptr, "null context");
  NVD_CUDA(cudaSetDevice(ctx->device));
  NVD_CUDA(cudaStreamSynchronize(ctx->stream));
  xchg_release(ctx);
  return NVD_OK;
}

int nvd_comm_allreduce_u64(nvd_ctx* ctx, uint64_t* values_dev, size_t n) {
  NVD_REQUIRE(ctx != nullptr && (n == 0 || values_dev != nullptr), "null argument");
  NVD_CUDA(cudaSetDevice(ctx->device));
  return comm_allreduce_u64(ctx, values_dev, n);
}

}  // extern "C"
