// NCCL inside the product (replaces the reference's only cross-worker exchange, the list splice under a mutex,
// /root/reference/lib/malbac/Malbac.cpp:105-141, and its shared counters Malbac.h:29-41): one communicator per context,
// ncclAllReduce for the small per-pass vectors and the cell-wide weight vector, ncclAllGather for the replication of the packed
// genome and the amplicon table over NVLink.
//
// libnccl.so.2 is resolved at run time (dlopen): a process that already carries an NCCL (torch's bundled copy in bench.py) is
// joined to that copy, the stand-alone CLI loads the system library. Nothing here is a fallback: without NCCL a multi-rank
// context can only use caller-supplied hooks (scs_set_collectives).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <string>

namespace scs {

struct NcclApi {
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false; std::string why;
};
// loads libnccl.so.2 on first use; ok == false with `why` when it cannot be found
NcclApi& nccl_api();

}  // namespace scs
