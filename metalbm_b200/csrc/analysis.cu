// analysis.cu -- what the C-ABI reports about a state: the scalar observables of the last stored step (AnalysisList.h:55-73),
// the energy / forcing spectra (AnalysisList.h:99-202), the alpha statistics, and the two host-side helpers
// (Communication::reduce, the logarithm self-test).
#include <cmath>
#include <cstring>

#include "context.h"

namespace mlbm {

// one block: count of alpha != 2, min and max of the alpha field (a diagnostic, not on the step path)
template <typename StoreT>
__global__ void __launch_bounds__(256) alphaStatisticsKernel(const StoreT* __restrict__ alpha, long long nodes, double* __restrict__ out) {
  __shared__ double scratch[3][256];
  double count = 0.0, low = 1e300, high = -1e300;
  for (long long i = threadIdx.x; i < nodes; i += blockDim.x) {
    const double a = (double)alpha[i];
    count += a != 2.0 ? 1.0 : 0.0;
    low = fmin(low, a);
    high = fmax(high, a);
  }
  scratch[0][threadIdx.x] = count; scratch[1][threadIdx.x] = low; scratch[2][threadIdx.x] = high;
  __syncthreads();
  for (int width = 128; width > 0; width >>= 1) {
    if ((int)threadIdx.x < width) {
      scratch[0][threadIdx.x] += scratch[0][threadIdx.x + width];
      scratch[1][threadIdx.x] = fmin(scratch[1][threadIdx.x], scratch[1][threadIdx.x + width]);
      scratch[2][threadIdx.x] = fmax(scratch[2][threadIdx.x], scratch[2][threadIdx.x + width]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = scratch[0][0]; out[1] = -scratch[1][0]; out[2] = scratch[2][0]; }   // -min: reduced with max over ranks
}

// both table formats of fastLogCore (step_kernel.cuh): even blocks the split one, odd blocks the {invc, logc} pairs
__global__ void fastLogKernel(const double* __restrict__ in, double* __restrict__ out, long long count) {
  __shared__ double2 pairs[kLogTableEntries];   // large enough for either format
  unsigned char* storage = reinterpret_cast<unsigned char*>(pairs);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x & 1) {
    LogTable<false> table;
    table.stage(storage, threadIdx.x, blockDim.x);
    __syncthreads();
    if (i < count) out[i] = fastLog(in[i], table);
  } else {
    LogTable<true> table;
    table.stage(storage, threadIdx.x, blockDim.x);
    __syncthreads();
    if (i < count) out[i] = fastLog(in[i], table);
  }
}

}  // namespace mlbm

using namespace mlbm;

extern "C" {

int mlbm_observables(mlbm_ctx* ctx, double out[4]) {
  if (!ctx || !out) return fail(MLBM_ERR_INVALID, "null argument");
  if (!ctx->observablesValid) return fail(MLBM_ERR_STATE, "no stored step yet (Algorithm::isStored was never set)");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (int status = joinAnalysis(ctx)) return status;  // the enstrophy of the last stored step comes from the analysis stream
  double local[4];
  if (ctx->config.nranks > 1) {
    if (!ctx->comm) return fail(MLBM_ERR_STATE, "nranks > 1 but mlbm_comm_init was not called");
    // Communication::reduce (Communication.h:76-89): sums over ranks (max for the Mach number)
    double* values = ctx->deviceObservables;
    const ncclDataType_t type = ncclDouble;
    MLBM_NCCL(ctx, ctx->nccl->GroupStart());
    MLBM_NCCL(ctx, ctx->nccl->AllReduce(values, values, 2, type, ncclSum, ctx->comm, ctx->computeStream));
    MLBM_NCCL(ctx, ctx->nccl->AllReduce(values + 2, values + 2, 1, type, ncclMax, ctx->comm, ctx->computeStream));
    MLBM_NCCL(ctx, ctx->nccl->AllReduce(values + 3, values + 3, 1, type, ncclSum, ctx->comm, ctx->computeStream));
    MLBM_NCCL(ctx, ctx->nccl->GroupEnd());
    ctx->observablesValid = false;  // reduced in place: valid again after the next stored step
  }
  MLBM_CUDA(cudaMemcpyAsync(local, ctx->deviceObservables, sizeof(local), cudaMemcpyDeviceToHost, ctx->computeStream));
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  double globalVolume = 1.0;
  for (int d = 0; d < ctx->D; ++d) globalVolume *= ctx->config.global_length[d];
  out[0] = local[0] / globalVolume;          // AnalysisScalar::normalize (Analysis.h:30)
  out[1] = ctx->enstrophyValid ? local[3] / globalVolume : NAN;  // Analysis.h:85-93; needs the stored velocity field
  out[2] = sqrt(local[2] * latticeInvCs2(ctx->config.lattice));  // |u| / c_s, c_s^2 = 1 / inv_cs2 (1/3 but for the multi-speed lattices)
  out[3] = local[1];
  return MLBM_OK;
}

int mlbm_power_spectra(mlbm_ctx* ctx, double* energySpectrum, double* forcingSpectrum, int capacity, int* count) {
  if (!ctx || !count) return fail(MLBM_ERR_INVALID, "null argument");
  const int* L = ctx->config.global_length;
  const int rest = L[1] < L[2] ? L[1] : L[2];                     // unused dimensions are 1
  const int bins = (L[0] > rest ? L[0] : rest) / 2;               // gFD::maxWaveNumber(): arrayMax is max(first, MIN of the rest)
  *count = bins;
  if (!energySpectrum && !forcingSpectrum) return MLBM_OK;
  if (capacity < bins) return fail(MLBM_ERR_INVALID, "capacity %d < %d wave numbers", capacity, bins);
  if (!ctx->fieldsStored || !ctx->velocity) return fail(MLBM_ERR_STATE, "no stored step yet (Algorithm::isStored was never set)");
  if (bins == 0) return MLBM_OK;
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (ctx->config.nranks > 1 && !ctx->comm) return fail(MLBM_ERR_STATE, "nranks > 1 but mlbm_comm_init was not called");
  if (int status = joinAnalysis(ctx)) return status;  // the transforms below reuse the buffers of the enstrophy analysis
  std::string error;
  if (!ctx->spectral) {
    SpectralGeometry geometry = {ctx->D, ctx->LX, ctx->NM, ctx->NR, ctx->config.rank, ctx->config.nranks, (int)ctx->elementSize};
    ctx->spectral = spectralCreate(geometry, ctx->nccl, ctx->analysisComm, &error);
    if (!ctx->spectral) return fail(MLBM_ERR_CUDA, "power spectra: %s", error.c_str());
  }
  double* device = nullptr;
  MLBM_CUDA(cudaMalloc(&device, sizeof(double) * 2 * (size_t)bins));
  int status = MLBM_OK;
  cudaStream_t stream = ctx->computeStream;
  if (spectralPowerSpectrum(ctx->spectral, ctx->velocity, ctx->fieldStride, bins, device, stream, &ctx->launches, &error) ||
      spectralPowerSpectrum(ctx->spectral, ctx->force, ctx->fieldStride, bins, device + bins, stream, &ctx->launches, &error))
    status = fail(MLBM_ERR_CUDA, "power spectra: %s", error.c_str());
  if (status == MLBM_OK && ctx->config.nranks > 1) {
    const ncclResult_t result = ctx->nccl->AllReduce(device, device, 2 * (size_t)bins, ncclDouble, ncclSum, ctx->comm, stream);
    if (result != ncclSuccess) status = fail(MLBM_ERR_COMM, "ncclAllReduce: %s", ctx->nccl->GetErrorString(result));
  }
  std::vector<double> host(2 * (size_t)bins);
  cudaError_t copyError = cudaSuccess;
  if (status == MLBM_OK) copyError = cudaMemcpyAsync(host.data(), device, sizeof(double) * host.size(), cudaMemcpyDeviceToHost, stream);
  const cudaError_t syncError = cudaStreamSynchronize(stream);
  cudaFree(device);
  if (status != MLBM_OK) return status;
  if (copyError != cudaSuccess || syncError != cudaSuccess)
    return fail(MLBM_ERR_CUDA, "power spectra: %s", cudaGetErrorString(copyError != cudaSuccess ? copyError : syncError));
  double volume = 1.0;
  for (int d = 0; d < ctx->D; ++d) volume *= L[d];
  for (int k = 0; k < bins; ++k) {
    if (energySpectrum) energySpectrum[k] = host[k] / volume;      // normalizeAnalyses: the energy spectrum only (AnalysisList.h:189)
    if (forcingSpectrum) forcingSpectrum[k] = host[bins + k];
  }
  return MLBM_OK;
}

int mlbm_alpha_statistics(mlbm_ctx* ctx, double out[3]) {
  if (!ctx || !out) return fail(MLBM_ERR_INVALID, "null argument");
  if (!ctx->alpha) { out[0] = 0.0; out[1] = 2.0; out[2] = 2.0; return MLBM_OK; }   // BGK: alpha == 2 (Collision.h:121)
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (ctx->config.nranks > 1 && !ctx->comm) return fail(MLBM_ERR_STATE, "nranks > 1 but mlbm_comm_init was not called");
  double* device = nullptr;
  MLBM_CUDA(cudaMalloc(&device, 3 * sizeof(double)));
  cudaStream_t stream = ctx->computeStream;
  if (ctx->config.dtype == MLBM_F64) alphaStatisticsKernel<double><<<1, 256, 0, stream>>>(static_cast<const double*>(ctx->alpha), ctx->nodes, device);
  else alphaStatisticsKernel<float><<<1, 256, 0, stream>>>(static_cast<const float*>(ctx->alpha), ctx->nodes, device);
  ctx->launches += 1;
  int status = MLBM_OK;
  if (ctx->config.nranks > 1) {
    ncclResult_t result = ctx->nccl->GroupStart();
    if (result == ncclSuccess) result = ctx->nccl->AllReduce(device, device, 1, ncclDouble, ncclSum, ctx->comm, stream);
    if (result == ncclSuccess) result = ctx->nccl->AllReduce(device + 1, device + 1, 2, ncclDouble, ncclMax, ctx->comm, stream);
    const ncclResult_t end = ctx->nccl->GroupEnd();
    if (result == ncclSuccess) result = end;
    if (result != ncclSuccess) status = fail(MLBM_ERR_COMM, "ncclAllReduce: %s", ctx->nccl->GetErrorString(result));
  }
  double host[3] = {0.0, 0.0, 0.0};
  cudaError_t error = cudaGetLastError();
  if (error == cudaSuccess && status == MLBM_OK) error = cudaMemcpyAsync(host, device, sizeof(host), cudaMemcpyDeviceToHost, stream);
  const cudaError_t syncError = cudaStreamSynchronize(stream);
  cudaFree(device);
  if (status != MLBM_OK) return status;
  if (error != cudaSuccess || syncError != cudaSuccess) return fail(MLBM_ERR_CUDA, "alpha statistics: %s", cudaGetErrorString(error != cudaSuccess ? error : syncError));
  double globalNodes = 1.0;
  for (int d = 0; d < ctx->D; ++d) globalNodes *= ctx->config.global_length[d];
  out[0] = host[0] / globalNodes;
  out[1] = -host[1];
  out[2] = host[2];
  return MLBM_OK;
}

int mlbm_newton_statistics(mlbm_ctx* ctx, int mode, unsigned long long out[2]) {
  if (!ctx) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (mode == 1) {
    if (!ctx->alpha) return MLBM_OK;  // BGK: nothing is solved
    if (!ctx->newtonCounters) MLBM_CUDA(cudaMalloc(&ctx->newtonCounters, 2 * sizeof(unsigned long long)));
    MLBM_CUDA(cudaMemsetAsync(ctx->newtonCounters, 0, 2 * sizeof(unsigned long long), ctx->computeStream));
    return MLBM_OK;
  }
  if (!out) return fail(MLBM_ERR_INVALID, "null argument");
  out[0] = out[1] = 0ull;
  if (!ctx->newtonCounters) return MLBM_OK;
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  MLBM_CUDA(cudaStreamSynchronize(ctx->commStream));
  MLBM_CUDA(cudaMemcpy(out, ctx->newtonCounters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  MLBM_CUDA(cudaFree(ctx->newtonCounters));
  ctx->newtonCounters = nullptr;
  return MLBM_OK;
}

int mlbm_reduce_sum(mlbm_ctx* ctx, double* values, int count) {
  if (!ctx || !values || count < 0) return fail(MLBM_ERR_INVALID, "null argument");
  if (ctx->config.nranks == 1 || count == 0) return MLBM_OK;
  if (!ctx->comm) return fail(MLBM_ERR_STATE, "nranks > 1 but mlbm_comm_init was not called");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  double* staging = nullptr;
  MLBM_CUDA(cudaMalloc(&staging, sizeof(double) * (size_t)count));
  int status = MLBM_OK;
  cudaError_t error = cudaMemcpyAsync(staging, values, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, ctx->computeStream);
  if (error == cudaSuccess) {
    ncclResult_t result = ctx->nccl->AllReduce(staging, staging, (size_t)count, ncclDouble, ncclSum, ctx->comm, ctx->computeStream);
    if (result != ncclSuccess) status = fail(MLBM_ERR_COMM, "ncclAllReduce: %s", ctx->nccl->GetErrorString(result));
  }
  if (error == cudaSuccess && status == MLBM_OK)
    error = cudaMemcpyAsync(values, staging, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, ctx->computeStream);
  if (error == cudaSuccess) error = cudaStreamSynchronize(ctx->computeStream);
  cudaFree(staging);
  if (error != cudaSuccess) return fail(MLBM_ERR_CUDA, "mlbm_reduce_sum: %s", cudaGetErrorString(error));
  ctx->launches += 1;
  return status;
}

int mlbm_selftest_log(const double* in, double* out, size_t count) {
  if (!in || !out) return fail(MLBM_ERR_INVALID, "null argument");
  if (!count) return MLBM_OK;
  double *deviceIn = nullptr, *deviceOut = nullptr;
  MLBM_CUDA(cudaMalloc(&deviceIn, count * sizeof(double)));
  cudaError_t error = cudaMalloc(&deviceOut, count * sizeof(double));
  if (error == cudaSuccess) error = cudaMemcpy(deviceIn, in, count * sizeof(double), cudaMemcpyHostToDevice);
  if (error == cudaSuccess) {
    fastLogKernel<<<(unsigned)((count + 127) / 128), 128>>>(deviceIn, deviceOut, (long long)count);
    error = cudaGetLastError();
  }
  if (error == cudaSuccess) error = cudaMemcpy(out, deviceOut, count * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(deviceIn);
  cudaFree(deviceOut);
  if (error != cudaSuccess) return fail(MLBM_ERR_CUDA, "mlbm_selftest_log: %s", cudaGetErrorString(error));
  return MLBM_OK;
}

}  // extern "C"
