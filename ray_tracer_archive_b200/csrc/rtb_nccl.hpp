// rtb_nccl.hpp — the handful of NCCL entry points librtb200 uses for the framebuffer reduce (one
// ncclReduce(float32, 4 W H, sum, root 0) per frame over NVLink / NVSwitch), bound at run time with dlopen so that a
// single-GPU caller needs no NCCL at all and a process that already carries a libnccl.so.2 (PyTorch bundles one) shares
// it instead of loading a second copy.  Declarations restate the public nccl.h ABI (NCCL 2.x).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

#include <mutex>
#include <string>

namespace rtb {

struct NcclUniqueId { char internal[128]; };  // NCCL_UNIQUE_ID_BYTES
typedef void* NcclComm;
enum { kNcclSuccess = 0, kNcclFloat32 = 7, kNcclSum = 0 };

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string error;
  bool ok = false;
};

inline const NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
      return;
    }
    auto sym = [&](const char* n) -> void* {
      void* p = dlsym(h, n);
      if (!p && api.error.empty()) api.error = std::string("libnccl is missing ") + n;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.ok = api.error.empty();
  });
  return api;
}

}  // namespace rtb
