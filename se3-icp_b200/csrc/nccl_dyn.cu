#include <dlfcn.h>

#include <mutex>

#include "common.cuh"
#include "nccl_dyn.h"
#include "../../include/se3icp.h"

namespace se3 {

const NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // RTLD_LOCAL: a later `import torch` in the same process must stay free to load its own bundled NCCL
        void* h = dlopen("libnccl.so.2", RTLD_NOLOAD | RTLD_NOW | RTLD_LOCAL);  // already in the process?
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
        api.Broadcast = (decltype(api.Broadcast))dlsym(h, "ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.Broadcast &&
                 api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    if (!api.ok) {
        set_last_error("libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
        return nullptr;
    }
    return &api;
}

}  // namespace se3
