// NCCL is bound lazily with dlopen so that (a) the library loads on a machine without a GPU and
// (b) it shares the NCCL build already loaded by the embedding process (PyTorch bundles its own
// libnccl.so.2; linking the system copy at load time would clash with it).
// Search order: $SPIRK_NCCL_LIB, an already-loaded libnccl.so.2, then the system libnccl.so.2.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>
#include <string>

namespace spirk
{
  struct NcclApi
  {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *)                                                           = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int)                                    = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                                               = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)    = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)            = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                  = nullptr;
    ncclResult_t (*GroupStart)()                                                                          = nullptr;
    ncclResult_t (*GroupEnd)()                                                                            = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *)                         = nullptr;
    const char *(*GetErrorString)(ncclResult_t)                                                           = nullptr;
    bool        ok = false;
    std::string error;
  };

  inline const NcclApi &nccl_api()
  {
    static NcclApi        api;
    static std::once_flag once;
    std::call_once(once, [] {
      void       *h   = nullptr;
      const char *env = std::getenv("SPIRK_NCCL_LIB");
      if (env && *env)
        h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
      if (!h)
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (!h)
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h)
        {
          api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
          return;
        }
#define SPIRK_NCCL_SYM(field, name)                          \
  api.field = (decltype(api.field))dlsym(h, name);           \
  if (!api.field)                                            \
    {                                                        \
      api.error = std::string("missing NCCL symbol ") + name; \
      return;                                                \
    }
      SPIRK_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
      SPIRK_NCCL_SYM(CommInitRank, "ncclCommInitRank")
      SPIRK_NCCL_SYM(CommDestroy, "ncclCommDestroy")
      SPIRK_NCCL_SYM(AllReduce, "ncclAllReduce")
      SPIRK_NCCL_SYM(AllGather, "ncclAllGather")
      SPIRK_NCCL_SYM(Send, "ncclSend")
      SPIRK_NCCL_SYM(Recv, "ncclRecv")
      SPIRK_NCCL_SYM(GroupStart, "ncclGroupStart")
      SPIRK_NCCL_SYM(GroupEnd, "ncclGroupEnd")
      SPIRK_NCCL_SYM(CommSplit, "ncclCommSplit")
      SPIRK_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SPIRK_NCCL_SYM
      api.ok = true;
    });
    return api;
  }
} // namespace spirk
