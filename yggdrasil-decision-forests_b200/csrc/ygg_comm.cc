// ygg_comm.cc — NCCL communicator behind include/ygg_b200_comm.h.
//
// NCCL is bound at run time (dlopen) through the handful of entry points below, declared here with
// their stable 2.x ABI so that neither nccl.h nor libnccl is needed to build the library.  The two
// collectives are enqueued on the caller's CUDA stream; nothing here synchronises with the host, so
// the level loop of ygg_engine.cu stays free of host round trips on every rank.
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ygg_b200_comm.h"
#include "ygg_internal.h"

namespace {

struct NcclUniqueId { char internal[YGG_COMM_UNIQUE_ID_BYTES]; };
using NcclComm = void*;
enum : int { kNcclUint8 = 1, kNcclUint32 = 3, kNcclUint64 = 5, kNcclFloat64 = 8 };  // ncclDataType_t
enum : int { kNcclSum = 0, kNcclMax = 2 };                                             // ncclRedOp_t

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  char error[256] = {0};
};

NcclApi* api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {std::getenv("YGG_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (n == nullptr || *n == 0) continue;
      a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib != nullptr) break;
      std::snprintf(a.error, sizeof(a.error), "%s", dlerror());
    }
    if (a.lib == nullptr) return;
    auto sym = [&](const char* s) {
      void* p = dlsym(a.lib, s);
      if (p == nullptr) std::snprintf(a.error, sizeof(a.error), "symbol %s not found in the NCCL library", s);
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.ReduceScatter = reinterpret_cast<decltype(a.ReduceScatter)>(sym("ncclReduceScatter"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &a;
}

int require_api(NcclApi** out) {
  NcclApi* a = api();
  if (a->lib == nullptr || !a->GetUniqueId || !a->CommInitRank || !a->CommDestroy || !a->AllReduce || !a->AllGather || !a->ReduceScatter) {
    char msg[384];
    std::snprintf(msg, sizeof(msg), "NCCL is not available: %s (set YGG_B200_NCCL_LIB to libnccl.so.2)",
                  a->error[0] ? a->error : "library not found");
    return ygg_set_error_msg(YGG_ERR_UNIMPLEMENTED, msg);
  }
  *out = a;
  return YGG_OK;
}

int nccl_error(NcclApi* a, const char* what, int rc) {
  char msg[256];
  std::snprintf(msg, sizeof(msg), "%s failed: %s (ncclResult %d)", what,
                a->GetErrorString ? a->GetErrorString(rc) : "?", rc);
  return ygg_set_error_msg(YGG_ERR_CUDA, msg);
}

}  // namespace

struct ygg_comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1, device = 0;
  std::vector<void*> local_windows;   // cudaMalloc'ed by this rank
  std::vector<void*> peer_mappings;   // cudaIpcOpenMemHandle'd
};

extern "C" {

int ygg_comm_unique_id(uint8_t out[YGG_COMM_UNIQUE_ID_BYTES]) {
  if (out == nullptr) return ygg_set_error_msg(YGG_ERR_INVALID_ARGUMENT, "null argument");
  NcclApi* a = nullptr;
  if (int st = require_api(&a)) return st;
  NcclUniqueId id;
  if (int rc = a->GetUniqueId(&id)) return nccl_error(a, "ncclGetUniqueId", rc);
  std::memcpy(out, id.internal, YGG_COMM_UNIQUE_ID_BYTES);
  return YGG_OK;
}

int ygg_comm_create(ygg_comm** out, const uint8_t unique_id[YGG_COMM_UNIQUE_ID_BYTES], int32_t rank, int32_t world,
                    int32_t device) {
  if (out == nullptr || unique_id == nullptr || world < 1 || rank < 0 || rank >= world)
    return ygg_set_error_msg(YGG_ERR_INVALID_ARGUMENT, "bad communicator arguments");
  NcclApi* a = nullptr;
  if (int st = require_api(&a)) return st;
  if (cudaSetDevice(device) != cudaSuccess) return ygg_set_error_msg(YGG_ERR_CUDA, "cudaSetDevice failed");
  NcclUniqueId id;
  std::memcpy(id.internal, unique_id, YGG_COMM_UNIQUE_ID_BYTES);
  ygg_comm* c = new ygg_comm;
  c->rank = rank; c->world = world; c->device = device;
  if (int rc = a->CommInitRank(&c->comm, world, id, rank)) {
    delete c;
    return nccl_error(a, "ncclCommInitRank", rc);
  }
  *out = c;
  return YGG_OK;
}

int ygg_comm_destroy(ygg_comm* c) {
  if (c == nullptr) return YGG_OK;
  NcclApi* a = api();
  cudaSetDevice(c->device);
  for (void* m : c->peer_mappings) cudaIpcCloseMemHandle(m);
  for (void* w : c->local_windows) cudaFree(w);
  if (c->comm != nullptr && a->CommDestroy != nullptr) a->CommDestroy(c->comm);
  delete c;
  return YGG_OK;
}

int ygg_comm_window_create(ygg_comm* c, int64_t bytes, void** peers) {
  if (c == nullptr || c->comm == nullptr || peers == nullptr || bytes <= 0) return ygg_set_error_msg(YGG_ERR_INVALID_ARGUMENT, "bad window arguments");
  NcclApi* a = nullptr;
  if (int st = require_api(&a)) return st;
  if (cudaSetDevice(c->device) != cudaSuccess) return ygg_set_error_msg(YGG_ERR_CUDA, "cudaSetDevice failed");
  void* local = nullptr;
  if (cudaMalloc(&local, static_cast<size_t>(bytes)) != cudaSuccess || cudaMemset(local, 0, static_cast<size_t>(bytes)) != cudaSuccess)
    return ygg_set_error_msg(YGG_ERR_CUDA, "window allocation failed");
  c->local_windows.push_back(local);
  // exchange the IPC handles with the communicator itself (one all-gather of 64 bytes per rank)
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, local) != cudaSuccess) return ygg_set_error_msg(YGG_ERR_CUDA, "cudaIpcGetMemHandle failed");
  cudaIpcMemHandle_t* d_handles = nullptr;
  std::vector<cudaIpcMemHandle_t> handles(c->world);
  cudaStream_t stream = nullptr;
  bool ok = cudaMalloc(&d_handles, sizeof(mine) * c->world) == cudaSuccess && cudaStreamCreate(&stream) == cudaSuccess &&
            cudaMemcpyAsync(d_handles + c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream) == cudaSuccess;
  if (ok) {
    const int rc = a->AllGather(d_handles + c->rank, d_handles, sizeof(mine), kNcclUint8, c->comm, stream);
    ok = rc == 0 && cudaMemcpyAsync(handles.data(), d_handles, sizeof(mine) * c->world, cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
         cudaStreamSynchronize(stream) == cudaSuccess;
  }
  if (stream) cudaStreamDestroy(stream);
  cudaFree(d_handles);
  if (!ok) return ygg_set_error_msg(YGG_ERR_CUDA, "exchange of the window handles failed");
  for (int r = 0; r < c->world; r++) {
    if (r == c->rank) { peers[r] = local; continue; }
    void* m = nullptr;
    if (cudaIpcOpenMemHandle(&m, handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      char msg[160];
      std::snprintf(msg, sizeof(msg), "cudaIpcOpenMemHandle of rank %d's window failed: %s", r, cudaGetErrorString(cudaGetLastError()));
      return ygg_set_error_msg(YGG_ERR_CUDA, msg);
    }
    c->peer_mappings.push_back(m);
    peers[r] = m;
  }
  return YGG_OK;
}

int ygg_comm_allreduce(void* ctx, void* buf, int64_t count, int32_t dtype, int32_t op, void* stream) {
  ygg_comm* c = static_cast<ygg_comm*>(ctx);
  if (c == nullptr || c->comm == nullptr || buf == nullptr || count < 0) return 1;
  static const int kTypes[3] = {kNcclUint32, kNcclUint64, kNcclFloat64};
  if (dtype < 0 || dtype > 2 || op < 0 || op > 1) return 1;
  return api()->AllReduce(buf, buf, static_cast<size_t>(count), kTypes[dtype], op == 0 ? kNcclSum : kNcclMax, c->comm,
                          static_cast<cudaStream_t>(stream));
}

int ygg_comm_reducescatter(void* ctx, void* buf, int64_t count_per_rank, int32_t dtype, int32_t op, void* stream) {
  ygg_comm* c = static_cast<ygg_comm*>(ctx);
  if (c == nullptr || c->comm == nullptr || buf == nullptr || count_per_rank < 0) return 1;
  static const int kTypes[3] = {kNcclUint32, kNcclUint64, kNcclFloat64};
  static const size_t kSize[3] = {4, 8, 8};
  if (dtype < 0 || dtype > 2 || op < 0 || op > 1) return 1;
  // in place: the receive buffer is this rank's chunk of the send buffer
  char* recv = static_cast<char*>(buf) + static_cast<size_t>(c->rank) * static_cast<size_t>(count_per_rank) * kSize[dtype];
  return api()->ReduceScatter(buf, recv, static_cast<size_t>(count_per_rank), kTypes[dtype], op == 0 ? kNcclSum : kNcclMax,
                              c->comm, static_cast<cudaStream_t>(stream));
}

int ygg_comm_allgather(void* ctx, const void* send, void* recv, int64_t bytes, void* stream) {
  ygg_comm* c = static_cast<ygg_comm*>(ctx);
  if (c == nullptr || c->comm == nullptr || send == nullptr || recv == nullptr || bytes < 0) return 1;
  // in place when send == recv + rank * bytes (how the engine lays out its shard-best table)
  return api()->AllGather(send, recv, static_cast<size_t>(bytes), kNcclUint8, c->comm, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
