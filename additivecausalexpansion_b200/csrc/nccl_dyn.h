// nccl_dyn.h -- NCCL entry points resolved at run time (dlopen), so that libace_b200.so has no link-time
// dependency on NCCL: it loads on a CPU-only box, and inside a process that already loaded PyTorch's bundled
// libnccl.so.2 the very same library instance is reused (same SONAME).  Only the multi-GPU sharded mode of a
// fit (ace_fit_shard) needs it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace ace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string why;
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    api.why = std::string("cannot load libnccl.so.2: ") + dlerror();
    return api;
  }
  auto sym = [&](const char* n) { return dlsym(h, n); };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Broadcast && api.AllGather &&
           api.GroupStart && api.GroupEnd && api.GetErrorString;
  if (!api.ok) api.why = "libnccl.so.2 lacks a required symbol";
  return api;
}

#define ACE_NCCL(expr)                                                                              \
  do {                                                                                              \
    ncclResult_t _r = (expr);                                                                       \
    if (_r != ncclSuccess) {                                                                        \
      ::ace::set_error(std::string(#expr) + ": " + ::ace::nccl_api().GetErrorString(_r));           \
      return -2000 - (int)_r;                                                                       \
    }                                                                                               \
  } while (0)

// Column-block plan of the sharded kernel build: the padded matrix is cut into 2*world column blocks of
// equal width; rank r owns blocks r and 2*world-1-r.  Only rows >= the block's first column are computed
// (potrf reads the lower triangle), so the pairing balances the trapezoids: every rank builds
// ~ n^2 (1 + 1/(2 world)) / (2 world) pairs.
inline int shard_block_width(int n_pad, int world) {
  if (world < 1 || n_pad % (2 * world * 64) != 0) return 0;
  return n_pad / (2 * world);
}
inline int shard_block_owner(int block, int world) { return block < world ? block : 2 * world - 1 - block; }

}  // namespace ace
