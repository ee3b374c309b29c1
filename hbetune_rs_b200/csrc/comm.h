// NCCL, loaded at run time.  libhbegp.so has no link-time dependency on libnccl: a single-GPU user never needs it, and
// inside a Python process the library must share the NCCL that PyTorch already loaded instead of pulling a second
// copy (the image has 2.27.3 in /usr/lib and 2.28.9 bundled with torch; two libnccl.so.2 in one process do not mix).
// Resolution order: a libnccl.so.2 that is already loaded, $HBEGP_NCCL_LIB, then the system library.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>
#include <string>

namespace hbegp {

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    static Nccl& get() {
        static Nccl n;
        static std::once_flag once;
        std::call_once(once, [] { n.load(); });
        return n;
    }
    bool ok() const { return lib != nullptr; }

private:
    template <typename F>
    bool sym(F& f, const char* name) {
        f = reinterpret_cast<F>(dlsym(lib, name));
        if (!f) error = std::string("libnccl: missing symbol ") + name;
        return f != nullptr;
    }
    void load() {
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!lib)
            if (const char* path = getenv("HBEGP_NCCL_LIB")) lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        const bool all = sym(GetVersion, "ncclGetVersion") && sym(GetUniqueId, "ncclGetUniqueId") && sym(CommInitRank, "ncclCommInitRank") &&
                         sym(CommInitAll, "ncclCommInitAll") && sym(CommDestroy, "ncclCommDestroy") && sym(CommAbort, "ncclCommAbort") &&
                         sym(AllReduce, "ncclAllReduce") && sym(AllGather, "ncclAllGather") && sym(Broadcast, "ncclBroadcast") &&
                         sym(GroupStart, "ncclGroupStart") && sym(GroupEnd, "ncclGroupEnd") && sym(GetErrorString, "ncclGetErrorString");
        if (!all) lib = nullptr;
    }
};

}  // namespace hbegp
