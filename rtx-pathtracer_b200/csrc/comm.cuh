// Multi-GPU plumbing shared by api.cu and guiding_fit.cu: NCCL entry points resolved with dlopen (the library loads on
// machines without NCCL) and one context's view of its communicator.  The reference is single-GPU; the exchange steps
// are SURVEY.md §8(e): image all-reduce, and the sample exchange of a guiding refit.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <dlfcn.h>

namespace b200pt {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) return false;
#define B200PT_NCCL_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(handle, name)); if (!field) { dlclose(handle); handle = nullptr; return false; }
        B200PT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") B200PT_NCCL_SYM(CommInitRank, "ncclCommInitRank") B200PT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        B200PT_NCCL_SYM(AllReduce, "ncclAllReduce") B200PT_NCCL_SYM(AllGather, "ncclAllGather") B200PT_NCCL_SYM(GroupStart, "ncclGroupStart")
        B200PT_NCCL_SYM(GroupEnd, "ncclGroupEnd") B200PT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
        B200PT_NCCL_SYM(Send, "ncclSend") B200PT_NCCL_SYM(Recv, "ncclRecv")
#undef B200PT_NCCL_SYM
        return true;
    }
};
extern NcclApi g_nccl;

#define B200PT_MAX_RANKS 16

// one context's communicator.  `peerMode`: the sorted sample buffers of all ranks are mapped into this process with
// CUDA IPC, so the exchange kernel of the region-sharded refit reads its records straight from the peers' HBM over
// NVLink; otherwise the records travel by grouped ncclSend / ncclRecv into a staging buffer.
struct RankComm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    bool peerMode = false;
};

}  // namespace b200pt
