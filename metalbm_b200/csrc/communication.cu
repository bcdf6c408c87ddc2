// communication.cu -- multi-GPU wiring of the C-ABI: NCCL (resolved at run time), the x-slab halo exchange as a message list
// and as grouped send / recv (Communication.h:134-180, 494-500), and the CUDA IPC mappings of the direct peer halos.
#include <dlfcn.h>
#include <unistd.h>

#include <cstring>

#include "context.h"

namespace mlbm {

const NcclApi* loadNccl(const char** error) {
  static NcclApi api;
  static bool tried = false, ok = false;
  static std::string message;
  if (!tried) {
    tried = true;
    void* handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) {
      message = std::string("cannot load libnccl: ") + dlerror();
    } else {
      ok = true;
      auto resolve = [&](const char* name) -> void* {
        void* symbol = dlsym(handle, name);
        if (!symbol) { ok = false; message = std::string("libnccl lacks ") + name; }
        return symbol;
      };
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(resolve("ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(resolve("ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(resolve("ncclCommDestroy"));
      api.Send = reinterpret_cast<decltype(api.Send)>(resolve("ncclSend"));
      api.Recv = reinterpret_cast<decltype(api.Recv)>(resolve("ncclRecv"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(resolve("ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(resolve("ncclGroupEnd"));
      api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(resolve("ncclAllReduce"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(resolve("ncclGetErrorString"));
      api.CommSplit = reinterpret_cast<decltype(api.CommSplit)>(dlsym(handle, "ncclCommSplit"));  // optional
    }
  }
  if (!ok) { if (error) *error = message.c_str(); return nullptr; }
  return &api;
}

// Communication::communicateHalos (Communication.h:494-500) as a list of messages: the last interior plane of
// the c_x > 0 populations goes to the right neighbour's plane 0, the first interior plane of the c_x < 0
// populations to the left neighbour's plane LX+1 (Communication.h:134-180).
int haloPlan(const mlbm_config* config, std::vector<mlbm_halo_message>* plan) {
  SlabGeometry g;
  if (!slabGeometry(config, &g)) return MLBM_ERR_INVALID;
  plan->clear();
  if (config->nranks == 1) return MLBM_OK;
  const int left = (config->rank + config->nranks - 1) % config->nranks;  // MPIInitializer.h:56
  const int right = (config->rank + 1) % config->nranks;                  // MPIInitializer.h:57
  auto add = [&](int q, int peer, int isSend, long long xPlane) {
    mlbm_halo_message message;
    message.population = q;
    message.peer = peer;
    message.is_send = isSend;
    message.reserved = 0;
    message.offset = (uint64_t)(q * g.stride + xPlane * g.plane);
    message.count = (uint64_t)(g.H * g.plane);   // dimH adjacent planes travel together (Communication.h:145-150: sizeStripeX)
    plan->push_back(message);
  };
  for (int q = g.faceQ + 1; q < 2 * g.faceQ + 1; ++q) {
    add(q, right, 1, g.LX);  // last H interior planes (interior planes are H .. LX + H - 1)
    add(q, left, 0, 0);      // left halo planes
  }
  for (int q = 1; q < g.faceQ + 1; ++q) {
    add(q, left, 1, g.H);          // first H interior planes
    add(q, right, 0, g.LX + g.H);  // right halo planes
  }
  return MLBM_OK;
}

int exchangeHalos(mlbm_ctx* ctx, int which, cudaStream_t stream) {
  if (ctx->config.nranks == 1) return MLBM_OK;
  if (!ctx->comm) return fail(MLBM_ERR_STATE, "nranks > 1 but mlbm_comm_init was not called");
  const ncclDataType_t type = ctx->config.dtype == MLBM_F64 ? ncclDouble : ncclFloat;
  void* base = ctx->populations[which];
  MLBM_NCCL(ctx, ctx->nccl->GroupStart());
  for (const mlbm_halo_message& message : ctx->haloMessages) {
    void* pointer = offsetElements(base, (long long)message.offset, ctx->elementSize);
    if (message.is_send) MLBM_NCCL(ctx, ctx->nccl->Send(pointer, message.count, type, message.peer, ctx->comm, stream));
    else MLBM_NCCL(ctx, ctx->nccl->Recv(pointer, message.count, type, message.peer, ctx->comm, stream));
  }
  MLBM_NCCL(ctx, ctx->nccl->GroupEnd());
  ctx->launches += 1;
  return MLBM_OK;
}

}  // namespace mlbm

using namespace mlbm;

extern "C" {

int mlbm_halo_plan(const mlbm_config* config, mlbm_halo_message* out, int capacity, int* count) {
  if (!config || !count) return fail(MLBM_ERR_INVALID, "null argument");
  std::vector<mlbm_halo_message> plan;
  if (haloPlan(config, &plan)) return fail(MLBM_ERR_INVALID, "bad lattice or nranks does not divide globalLengthX");
  *count = (int)plan.size();
  if (out) {
    if (capacity < (int)plan.size()) return fail(MLBM_ERR_INVALID, "capacity %d < %d messages", capacity, (int)plan.size());
    memcpy(out, plan.data(), plan.size() * sizeof(mlbm_halo_message));
  }
  return MLBM_OK;
}

int mlbm_comm_unique_id(void* id128) {
  if (!id128) return fail(MLBM_ERR_INVALID, "null argument");
  const char* error = nullptr;
  const NcclApi* api = loadNccl(&error);
  if (!api) return fail(MLBM_ERR_COMM, "%s", error);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ncclResult_t result = api->GetUniqueId(&id);
  if (result != ncclSuccess) return fail(MLBM_ERR_COMM, "ncclGetUniqueId: %s", api->GetErrorString(result));
  memcpy(id128, &id, sizeof(id));
  return MLBM_OK;
}

int mlbm_comm_init(mlbm_ctx* ctx, const void* id128) {
  if (!ctx || !id128) return fail(MLBM_ERR_INVALID, "null argument");
  if (ctx->comm) return fail(MLBM_ERR_STATE, "communicator already initialised");
  const char* error = nullptr;
  ctx->nccl = loadNccl(&error);
  if (!ctx->nccl) return fail(MLBM_ERR_COMM, "%s", error);
  MLBM_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  MLBM_NCCL(ctx, ctx->nccl->CommInitRank(&ctx->comm, ctx->config.nranks, id, ctx->config.rank));
  // The spectral analyses of a stored step run on their own stream next to the following steps (context.cu: enqueueStep):
  // their all-to-all gets its own communicator, so that it never queues in front of a step's halo exchange or of the
  // all-reduce of the observables on ctx->comm.  Without ncclCommSplit the one communicator serves both (NCCL orders them).
  ctx->analysisComm = ctx->comm;
  static const bool shareComm = getenv("MLBM_ANALYSIS_COMM") && atoi(getenv("MLBM_ANALYSIS_COMM")) == 0;
  if (ctx->nccl->CommSplit && !shareComm) {
    ncclComm_t second = nullptr;
    if (ctx->nccl->CommSplit(ctx->comm, 0, ctx->config.rank, &second, nullptr) == ncclSuccess && second) ctx->analysisComm = second;
  }
  return MLBM_OK;
}

// ---- direct peer halos ---------------------------------------------------------------------------
struct PeerBlob {  // MLBM_PEER_HANDLE_BYTES = 256
  cudaIpcMemHandle_t populations[2];
  cudaIpcMemHandle_t flags;
  uint64_t bufferBytes;
  int32_t rank, device;
  int64_t process;
  char padding[256 - 3 * sizeof(cudaIpcMemHandle_t) - 8 - 8 - 8];
};
static_assert(sizeof(PeerBlob) == MLBM_PEER_HANDLE_BYTES, "peer handle blob size");

int mlbm_comm_peer_export(mlbm_ctx* ctx, void* handle) {
  if (!ctx || !handle) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->peerFlags) {
    MLBM_CUDA(cudaMalloc(&ctx->peerFlags, 256));
    MLBM_CUDA(cudaMemset(ctx->peerFlags, 0, 256));
    MLBM_CUDA(cudaHostAlloc(&ctx->peerTimedOut, sizeof(int), cudaHostAllocMapped));
    *ctx->peerTimedOut = 0;
  }
  PeerBlob blob;
  memset(&blob, 0, sizeof(blob));
  for (int i = 0; i < 2; ++i) {
    cudaError_t error = cudaIpcGetMemHandle(&blob.populations[i], ctx->populations[i]);
    if (error != cudaSuccess) return fail(MLBM_ERR_COMM, "cudaIpcGetMemHandle: %s", cudaGetErrorString(error));
  }
  cudaError_t error = cudaIpcGetMemHandle(&blob.flags, ctx->peerFlags);
  if (error != cudaSuccess) return fail(MLBM_ERR_COMM, "cudaIpcGetMemHandle: %s", cudaGetErrorString(error));
  blob.bufferBytes = (uint64_t)ctx->stride * ctx->Q * ctx->elementSize;
  blob.rank = ctx->config.rank;
  blob.device = ctx->device;
  blob.process = (int64_t)getpid();
  memcpy(handle, &blob, sizeof(blob));
  return MLBM_OK;
}

int mlbm_comm_peer_attach(mlbm_ctx* ctx, const void* leftHandle, const void* rightHandle) {
  if (!ctx || !leftHandle || !rightHandle) return fail(MLBM_ERR_INVALID, "null argument");
  if (ctx->config.nranks < 2) return fail(MLBM_ERR_STATE, "a single rank has no neighbours");
  if (ctx->peerAttached) return fail(MLBM_ERR_STATE, "peer halos already attached");
  if (!ctx->peerFlags) return fail(MLBM_ERR_STATE, "mlbm_comm_peer_export has to be called first");
  if (!ctx->comm) return fail(MLBM_ERR_STATE, "mlbm_comm_init has to be called first (initial halo exchange and shutdown barrier)");
  if (ctx->H > 1) return fail(MLBM_ERR_INVALID, "direct peer halos are built for the single-speed lattices (one halo plane); the multi-speed ones exchange over NCCL");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  PeerBlob blobs[2];
  memcpy(&blobs[0], leftHandle, sizeof(PeerBlob));
  memcpy(&blobs[1], rightHandle, sizeof(PeerBlob));
  const int expected[2] = {(ctx->config.rank + ctx->config.nranks - 1) % ctx->config.nranks, (ctx->config.rank + 1) % ctx->config.nranks};
  const uint64_t bufferBytes = (uint64_t)ctx->stride * ctx->Q * ctx->elementSize;
  for (int side = 0; side < 2; ++side) {
    if (blobs[side].rank != expected[side]) return fail(MLBM_ERR_INVALID, "handle of rank %d where the %s neighbour %d was expected", blobs[side].rank, side ? "right" : "left", expected[side]);
    if (blobs[side].bufferBytes != bufferBytes) return fail(MLBM_ERR_INVALID, "neighbour %d has a different slab geometry", blobs[side].rank);
    if (blobs[side].process == (int64_t)getpid()) return fail(MLBM_ERR_INVALID, "peer halos need one process per rank (CUDA IPC)");
  }
  auto closeAll = [&]() {
    for (int side = 0; side < 2; ++side) {
      if (ctx->mappedOwned[side]) for (void*& pointer : ctx->mapped[side]) if (pointer) cudaIpcCloseMemHandle(pointer);
      for (void*& pointer : ctx->mapped[side]) pointer = nullptr;
      ctx->mappedOwned[side] = false;
    }
  };
  for (int side = 0; side < 2; ++side) {
    if (side == 1 && expected[1] == expected[0]) {  // two ranks: both neighbours are the same process, map it once
      for (int i = 0; i < 3; ++i) ctx->mapped[1][i] = ctx->mapped[0][i];
      break;
    }
    ctx->mappedOwned[side] = true;
    const cudaIpcMemHandle_t* handles[3] = {&blobs[side].populations[0], &blobs[side].populations[1], &blobs[side].flags};
    for (int i = 0; i < 3; ++i) {
      cudaError_t error = cudaIpcOpenMemHandle(&ctx->mapped[side][i], *handles[i], cudaIpcMemLazyEnablePeerAccess);
      if (error != cudaSuccess) {
        cudaGetLastError();
        closeAll();
        return fail(MLBM_ERR_COMM, "cudaIpcOpenMemHandle (rank %d, device %d): %s", blobs[side].rank, blobs[side].device, cudaGetErrorString(error));
      }
    }
  }
  ctx->peerAttached = true;
  ctx->peerEpoch = 0;
  ctx->halosValid = false;
  return MLBM_OK;
}

}  // extern "C"
