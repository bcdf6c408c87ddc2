// context.h -- internal to the library: the per-rank context behind the opaque `mlbm_ctx` of include/metalbm_b200.h, the
// error plumbing and the few helpers shared by context.cu (life cycle, step orchestration, copies), communication.cu (NCCL,
// peer mappings, halo exchange) and analysis.cu (observables, spectra, diagnostics).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/metalbm_b200.h"
#include "nccl_loader.h"
#include "shell_force.h"
#include "spectral.h"
#include "step_kernel.cuh"

namespace mlbm {

// stores the message behind mlbm_last_error() for the calling thread and returns `status`
int fail(int status, const char* format, ...);
const char* lastError();

// geometry shared by the context and the (device-free) halo plan
struct SlabGeometry {
  int D, Q, faceQ, LX, NM, NR;
  int H;                 // halo planes per side in x (Lattice dimH: 1 but for the multi-speed lattices)
  long long plane, stride;
};

bool slabGeometry(const mlbm_config* config, SlabGeometry* g);

// Communication::communicateHalos (Communication.h:494-500) as a list of messages / as grouped NCCL send + recv (communication.cu)
int haloPlan(const mlbm_config* config, std::vector<mlbm_halo_message>* plan);

constexpr int kReduceBlocks = 296;  // two per SM (reduceObservablesKernel)

}  // namespace mlbm

#define MLBM_CUDA(call)                                                                                  \
  do {                                                                                                   \
    cudaError_t error_ = (call);                                                                         \
    if (error_ != cudaSuccess)                                                                           \
      return mlbm::fail(error_ == cudaErrorMemoryAllocation ? MLBM_ERR_NOMEM : MLBM_ERR_CUDA, "[%s:%d] CUDA failed with %s", \
                  __FILE__, __LINE__, cudaGetErrorString(error_));                                       \
  } while (0)

#define MLBM_NCCL(ctx, call)                                                                             \
  do {                                                                                                   \
    ncclResult_t result_ = (call);                                                                       \
    if (result_ != ncclSuccess)                                                                          \
      return mlbm::fail(MLBM_ERR_COMM, "[%s:%d] NCCL failed with %s", __FILE__, __LINE__,                      \
                  (ctx)->nccl->GetErrorString(result_));                                                 \
  } while (0)


// ------------------------------------------------------------------------------------------------
// the context
// ------------------------------------------------------------------------------------------------
struct mlbm_ctx {
  mlbm_config config;
  int device = 0;
  int D = 0, Q = 0, faceQ = 0;
  int LX = 0, NM = 0, NR = 0;           // local extents on the kernel axes (x, m, r)
  size_t elementSize = 8;
  long long plane = 0, stride = 0, fieldStride = 0, nodes = 0;
  int H = 1;                            // halo planes per side in x
  long long interior = 0;               // elements from the start of a population to its first interior plane (H * plane)
  long long partialBlocks = 0;
  int gridR = 0;

  void* populations[2] = {nullptr, nullptr};  // ping-pong SoA pair (Distribution.h:19-20)
  int current = 0;                             // buffer the next step reads ("previous")
  void* alpha = nullptr;
  void* density = nullptr;
  void* velocity = nullptr;
  void* force = nullptr;
  bool fieldsStored = false;
  double* partials = nullptr;
  unsigned long long* newtonCounters = nullptr;  // mlbm_newton_statistics: {nodes solved, evaluations}, counted while non-null
  void* staging = nullptr;               // staged pack / unpack: one padded population block
  size_t stagingBytes = 0;
  double* reduceStage = nullptr;         // [kReduceBlocks][kObservableSlots] second-stage partials
  unsigned* reduceTicket = nullptr;
  double* deviceObservables = nullptr;   // [energy sum, mass, max speed^2, enstrophy sum]
  mlbm::SpectralEnstrophy* spectral = nullptr; // created on the first step that stores the fields
  mlbm::ShellForce* shell = nullptr;           // ConstantShell / EnergyRemoval / Turbulent2D (2-D): maker of the force field
  bool forceStale = false;               // the fields changed since the force field was made (Force::update is due)
  bool enstrophyValid = false;           // the last stored step stored the velocity field (bit 0 of isStored)
  double* forceTables[3] = {nullptr, nullptr, nullptr};
  int forceAxis[3] = {-1, -1, -1};
  bool observablesValid = false;

  mlbm::StepKernel kernel = nullptr;
  int sharedBytes = 0;  // dynamic shared memory of the fused kernel (entropic kernels stage f / fNeq there)
  int hydroShift = 0;
  cudaStream_t computeStream = nullptr;
  cudaStream_t commStream = nullptr;
  cudaStream_t analysisStream = nullptr;   // spectral enstrophy of a stored step, next to the steps that follow it
  cudaEvent_t fieldsReady = nullptr, analysisDone = nullptr;
  bool poisoned = false;                   // a synchronisation failed (sticky CUDA error): no collectives at shutdown
  bool analysisPending = false;            // analysisDone was recorded and nobody has waited for it on the compute stream yet
  ncclComm_t analysisComm = nullptr;       // == comm where ncclCommSplit is unavailable
  cudaEvent_t boundaryDone = nullptr, exchangeDone = nullptr, bulkDone = nullptr;
  cudaEvent_t timeStart = nullptr, timeMid = nullptr, timeStop = nullptr;
  cudaEvent_t marks[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double lastCommunication = 0.0, lastComputation = 0.0;
  unsigned long long launches = 0;

  // per-launch timing of the fused kernel (mlbm_kernel_time)
  bool profiling = false;
  std::vector<cudaEvent_t> profileEvents;
  size_t profileUsed = 0;
  double profileMs = 0.0;
  unsigned long long profileLaunches = 0;

  const mlbm::NcclApi* nccl = nullptr;
  ncclComm_t comm = nullptr;
  // direct peer halos (mlbm_comm_peer_export / _attach)
  unsigned long long* peerFlags = nullptr;       // this rank's two handshake words (+ padding), written by the neighbours
  int* peerTimedOut = nullptr;
  void* mapped[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // [left, right][populations 0, 1, flags]
  bool mappedOwned[2] = {false, false};          // right shares left's mapping when both are the same rank
  bool peerAttached = false;
  unsigned long long peerEpoch = 0;
  cudaEvent_t stepStart = nullptr;
  bool halosValid = false;  // halo planes of populations[current] hold the neighbours' data
  std::vector<mlbm_halo_message> haloMessages;
};

namespace mlbm {
int exchangeHalos(mlbm_ctx* ctx, int which, cudaStream_t stream);
int joinAnalysis(mlbm_ctx* ctx);  // the compute stream waits for the spectral analysis of the last stored step
inline void* offsetElements(void* base, long long elements, size_t elementSize) {
  return static_cast<char*>(base) + elements * (long long)elementSize;
}
}  // namespace mlbm
