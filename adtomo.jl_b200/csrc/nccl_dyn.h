// nccl_dyn.h -- NCCL reached through dlopen, so that libadtomo_b200.so loads on machines without
// NCCL and shares the copy a host framework (e.g. PyTorch) has already loaded.  Only the handful
// of entry points the multi-GPU path needs; ABI constants copied from nccl.h 2.x (stable).
#pragma once
#include <dlfcn.h>
#include <cstddef>

namespace adtomo {

struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;   // ncclUniqueId, NCCL_UNIQUE_ID_BYTES = 128
    typedef void *Comm;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void *handle = nullptr;
    static constexpr int kFloat64 = 8;   // ncclFloat64 / ncclDouble
    static constexpr int kSum = 0;       // ncclSum

    bool load(const char **why) {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { *why = dlerror(); return false; }
        GetUniqueId = (int (*)(UniqueId *))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (int (*)(Comm *, int, UniqueId, int))dlsym(handle, "ncclCommInitRank");
        CommDestroy = (int (*)(Comm))dlsym(handle, "ncclCommDestroy");
        AllReduce = (int (*)(const void *, void *, size_t, int, int, Comm, void *))dlsym(handle, "ncclAllReduce");
        GetErrorString = (const char *(*)(int))dlsym(handle, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) {
            *why = "libnccl lacks a required symbol";
            dlclose(handle);       // a later call must not take the half-resolved table for a loaded library
            handle = nullptr;
            GetUniqueId = nullptr; CommInitRank = nullptr; CommDestroy = nullptr; AllReduce = nullptr; GetErrorString = nullptr;
            return false;
        }
        return true;
    }
};

}  // namespace adtomo
