// colq_nccl.h -- NCCL bound at run time with dlopen so that
//   (a) libcolq.so loads (and single-GPU contexts work) on hosts without NCCL, and
//   (b) inside a process that already loaded an NCCL (torch's bundled libnccl.so.2) we bind to that very
//       copy instead of a second one: dlopen by SONAME returns the already-mapped library.
// Only the handful of entry points the query path needs are bound (one mask exchange per
// cross-shard hop, one final gather of matched indices).
#pragma once

#include <dlfcn.h>
#include <stddef.h>
#include <string>

namespace colq {

// Minimal mirror of the public NCCL C API (stable since NCCL 2.0); we avoid a build-time dependency on nccl.h.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;       // 0 == ncclSuccess
typedef int ncclDataType_t;     // ncclUint8 = 1, ncclUint32 = 3, ncclInt32 = 2, ncclUint64 = 5
constexpr ncclDataType_t kNcclUint8 = 1, kNcclInt32 = 2, kNcclUint32 = 3, kNcclUint64 = 5;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, void* /*cudaStream_t*/) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, void*) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, void*) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    bool load(std::string& err) {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?");
            return false;
        }
#define COLQ_BIND(field, sym)                                            \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, sym));       \
    if (!field) { err = std::string("NCCL symbol missing: ") + sym; handle = nullptr; return false; }
        COLQ_BIND(GetUniqueId, "ncclGetUniqueId")
        COLQ_BIND(CommInitRank, "ncclCommInitRank")
        COLQ_BIND(CommDestroy, "ncclCommDestroy")
        COLQ_BIND(AllGather, "ncclAllGather")
        COLQ_BIND(Send, "ncclSend")
        COLQ_BIND(Recv, "ncclRecv")
        COLQ_BIND(GroupStart, "ncclGroupStart")
        COLQ_BIND(GroupEnd, "ncclGroupEnd")
        COLQ_BIND(GetErrorString, "ncclGetErrorString")
#undef COLQ_BIND
        return true;
    }
};

}  // namespace colq
