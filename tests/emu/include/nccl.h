// tests/emu/include/nccl.h -- TEST INFRASTRUCTURE ONLY: the NCCL types csrc/nccl_loader.h names, for the emulated
// (single-rank) build of the library; no NCCL function is ever called there.
#pragma once
#include <cstddef>
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclFloat = 7, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2 } ncclRedOp_t;
