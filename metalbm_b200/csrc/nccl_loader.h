// nccl_loader.h -- NCCL entry points resolved at run time (dlopen), so that the library loads on
// machines without NCCL and shares the libnccl already mapped by the host process when there is one.
// Used for the x-slab halo exchange (replaces MPI_Isend/Irecv, Communication.h:134-180) and the
// scalar reductions (replaces MPI_Reduce, Communication.h:76-89).
#pragma once

#include <nccl.h>

namespace mlbm {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  // optional (NCCL >= 2.18; nullptr otherwise): a second communicator over the same ranks for the analysis stream
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*);
};

// nullptr (and a message in *error) when libnccl cannot be loaded
const NcclApi* loadNccl(const char** error);

}  // namespace mlbm
