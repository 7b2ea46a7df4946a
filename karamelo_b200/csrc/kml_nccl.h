// kml_nccl.h - NCCL entry points resolved at run time.
//
// libkml.so does not link libnccl: a single-GPU run never needs it, and inside a Python process torch has
// usually loaded its own bundled libnccl.so.2 already - a second, link-time copy of a different version
// under the same soname breaks whichever library is loaded second.  kml_comm_init() therefore binds the
// few calls the slab exchange uses with dlopen/dlsym: first the copy already in the process
// (RTLD_NOLOAD), then $KML_NCCL_LIB, then the system search path.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <string>

namespace kml {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;

  // returns an empty string on success, else the reason
  std::string load() {
    if (handle) return "";
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) { const char *p = getenv("KML_NCCL_LIB"); if (p && *p) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL); }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return std::string("cannot load libnccl.so.2: ") + dlerror();
    bool ok = true;
    auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) ok = false; return p; };
    GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    Send = (decltype(Send))sym("ncclSend");
    Recv = (decltype(Recv))sym("ncclRecv");
    AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!ok) return "libnccl.so.2 lacks a required entry point";
    handle = h; return "";
  }
};

inline NcclApi &nccl() { static NcclApi api; return api; }

} // namespace kml
