// shell_force.cu -- the reference's spectral body forces on 2-D lattices without a distributed transform.
//
// What the reference does (CPU, FFTW-MPI):
//   ConstantShell (Force.h:296-420): stream function psi^(k) = A on the shell kMin^2 <= |k|^2 <= kMax^2 of the r2c half
//     spectrum (integer wave numbers k = i <= N/2 ? i : i - N), MakeIncompressible (Transformer.h:300-384):
//     F^ = (i k_y psi^, -i k_x psi^), c2r, divided by the volume V (Transformer.h:101-108).  Made once (Collision.h:51-54).
//   EnergyRemoval (Force.h:423-561): momentum rho u_d of fieldList (the fields of the last STORED step), r2c,
//     F^_d = -A_d (rho u_d)^ on the shell and 0 elsewhere, c2r, divided by V.  Remade in every iterate (Force.h:552-558).
//   Turbulent2D (Force.h:564-616): ConstantShell(forceAmplitude, forcekMin/Max) + EnergyRemoval(removalForce*).
//
// Both transforms only ever touch the few modes of a shell, so they are evaluated as what they are -- sums over those modes:
//   synthesis   F(x, y) = sum_{k in stored shell} w_k Re[ F^(k) e^{ i theta_k(x, y)} ],  theta_k = 2 pi (k_x x / N_x + k_y y / N_y),
//               w_k = 2 for 0 < k_y < N_y / 2 (the mode and its conjugate), 1 on the two self-conjugate columns of the half
//               spectrum (whose non-Hermitian parts a c2r transform drops: they cancel in the sum over +-k_x);
//   projection  (rho u_d)^(k) = sum_{x, y} rho u_d e^{-i theta_k}: block partial sums over a fixed grid, a second stage in
//               index order, and ONE all-reduce of 4 doubles per mode over the ranks (x-slabs) -- no all-to-all.
// The angle is reduced in integer arithmetic (k x mod N), so the fields do not depend on the decomposition beyond the
// summation order of the projection.  Pinned against arrays and populations the reference itself produced
// (tests/golden/*constantshell*, *energyremoval*, *turbulent2d*) through the literal FFT restatement in oracle/oracle.py.
#include "shell_force.h"

#include <algorithm>
#include <vector>

namespace mlbm {

namespace {

constexpr int kBlock = 256;
constexpr int kProjectBlocks = 592;  // four per SM, fixed: the summation order does not depend on timing

struct Mode {
  int kx, ky;
  double weight;
};

__device__ __forceinline__ double angleOverPi(const Mode& mode, long long x, long long y, int globalX, int globalY) {
  const long long px = ((mode.kx * x) % globalX + globalX) % globalX, py = (mode.ky * y) % globalY;
  return 2.0 * (double)px / globalX + 2.0 * (double)py / globalY;
}

// ConstantShell: modes[i].weight already carries w_k A / V.  out = [2][nodes] doubles.
__global__ void injectionKernel(double* __restrict__ out, long long nodes, int NR, int xOffset, int globalX, int globalY,
                                const Mode* __restrict__ modes, int modeCount) {
  const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= nodes) return;
  const long long x = node / NR + xOffset, y = node % NR;
  double fx = 0.0, fy = 0.0;
  for (int i = 0; i < modeCount; ++i) {
    const Mode mode = modes[i];
    const double s = sinpi(angleOverPi(mode, x, y, globalX, globalY)) * mode.weight;
    fx -= (double)mode.ky * s;   // Re[ i k_y A e^{i theta}] = -k_y A sin theta
    fy += (double)mode.kx * s;   // Re[-i k_x A e^{i theta}] =  k_x A sin theta
  }
  out[node] = fx;
  out[nodes + node] = fy;
}

// partials[(mode * gridDim.x + block) * 4 + c]: this block's share of sum rho u_d e^{-i theta}, c = (Re x, Im x, Re y, Im y)
template <typename StoreT>
__global__ void __launch_bounds__(kBlock)
projectKernel(const StoreT* __restrict__ density, const StoreT* __restrict__ velocity, long long fieldStride, long long nodes, int NR,
              int xOffset, int globalX, int globalY, const Mode* __restrict__ modes, double* __restrict__ partials) {
  const Mode mode = modes[blockIdx.y];
  double sum[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x; node < nodes; node += (long long)gridDim.x * blockDim.x) {
    const long long x = node / NR + xOffset, y = node % NR;
    double sine, cosine;
    sincospi(angleOverPi(mode, x, y, globalX, globalY), &sine, &cosine);
    const double rho = (double)density[node];
    const double mx = rho * (double)velocity[node], my = rho * (double)velocity[fieldStride + node];  // Force.h:466-474
    sum[0] += mx * cosine; sum[1] -= mx * sine;
    sum[2] += my * cosine; sum[3] -= my * sine;
  }
  __shared__ double scratch[4][kBlock / 32];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double value = sum[c];
    for (int offset = 16; offset > 0; offset >>= 1) value += __shfl_xor_sync(0xffffffffu, value, offset);
    if ((threadIdx.x & 31) == 0) scratch[c][threadIdx.x >> 5] = value;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double total = 0.0;
    for (int w = 0; w < kBlock / 32; ++w) total += scratch[threadIdx.x][w];
    partials[((long long)blockIdx.y * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = total;
  }
}

// projections[mode * 4 + c] = sum over the blocks, in index order (one thread per (mode, c))
__global__ void reduceKernel(const double* __restrict__ partials, int blocks, int modeCount, double* __restrict__ projections) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= modeCount * 4) return;
  const int mode = item / 4, c = item % 4;
  double total = 0.0;
  for (int b = 0; b < blocks; ++b) total += partials[((long long)mode * blocks + b) * 4 + c];
  projections[item] = total;
}

// force_d = constant_d + scale_d sum_k w_k (Re P_d cos theta - Im P_d sin theta), scale_d = -A_d / V
template <typename StoreT>
__global__ void synthesisKernel(StoreT* __restrict__ force, long long fieldStride, long long nodes, int NR, int xOffset, int globalX,
                                int globalY, const Mode* __restrict__ modes, int modeCount, const double* __restrict__ projections,
                                double scaleX, double scaleY, const double* __restrict__ constant) {
  const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= nodes) return;
  const long long x = node / NR + xOffset, y = node % NR;
  double fx = 0.0, fy = 0.0;
  for (int i = 0; i < modeCount; ++i) {
    const Mode mode = modes[i];
    double sine, cosine;
    sincospi(angleOverPi(mode, x, y, globalX, globalY), &sine, &cosine);
    fx += mode.weight * (projections[4 * i + 0] * cosine - projections[4 * i + 1] * sine);
    fy += mode.weight * (projections[4 * i + 2] * cosine - projections[4 * i + 3] * sine);
  }
  fx *= scaleX;
  fy *= scaleY;
  if (constant) {  // Turbulent2D: forcePtr[index] += removalForcePtr[index] (Force.h:594-600)
    fx = constant[node] + fx;
    fy = constant[nodes + node] + fy;
  }
  force[node] = (StoreT)fx;
  force[fieldStride + node] = (StoreT)fy;
}

template <typename StoreT>
__global__ void copyConstantKernel(StoreT* __restrict__ force, long long fieldStride, long long nodes, const double* __restrict__ constant) {
  const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= nodes) return;
  force[node] = (StoreT)constant[node];
  force[fieldStride + node] = (StoreT)constant[nodes + node];
}

// the stored half spectrum's modes inside a shell, in (i_x, i_y) index order (Force.h:349-354, 385-386)
std::vector<Mode> shellModes(int globalX, int globalY, int kMin, int kMax, double scale) {
  std::vector<Mode> modes;
  for (int ix = 0; ix < globalX; ++ix) {
    const int kx = ix <= globalX / 2 ? ix : ix - globalX;
    for (int iy = 0; iy <= globalY / 2; ++iy) {
      const long long k2 = (long long)kx * kx + (long long)iy * iy;
      if (k2 < (long long)kMin * kMin || k2 > (long long)kMax * kMax) continue;
      const bool selfConjugate = iy == 0 || (globalY % 2 == 0 && iy == globalY / 2);
      modes.push_back({kx, iy, (selfConjugate ? 1.0 : 2.0) * scale});
    }
  }
  return modes;
}

bool check(cudaError_t status, const char* what, std::string* error) {
  if (status == cudaSuccess) return true;
  if (error) *error = std::string(what) + ": " + cudaGetErrorString(status);
  return false;
}

}  // namespace

class ShellForce {
 public:
  ShellForceGeometry g;
  ShellForceSpec spec;
  long long nodes = 0;
  double* constant = nullptr;      // [2][nodes] injection part (ConstantShell), made once
  Mode* removalModes = nullptr;
  int removalModeCount = 0;
  double* partials = nullptr;      // [removalModeCount][kProjectBlocks][4]
  double* projections = nullptr;   // [removalModeCount][4]
  int projectBlocks = 0;

  ~ShellForce() {
    for (void* pointer : {(void*)constant, (void*)removalModes, (void*)partials, (void*)projections})
      if (pointer) cudaFree(pointer);
  }
};

ShellForce* shellForceCreate(const ShellForceGeometry& geometry, const ShellForceSpec& spec, std::string* error) {
  ShellForce* s = new ShellForce();
  s->g = geometry;
  s->spec = spec;
  s->nodes = (long long)geometry.LX * geometry.NR;
  auto failed = [&]() { delete s; return (ShellForce*)nullptr; };
  const double volume = (double)geometry.globalX * geometry.globalY;
  const unsigned grid = (unsigned)((s->nodes + kBlock - 1) / kBlock);
  if (spec.injection) {
    const std::vector<Mode> modes = shellModes(geometry.globalX, geometry.globalY, spec.injectionKMin, spec.injectionKMax,
                                               spec.injectionAmplitude / volume);
    Mode* deviceModes = nullptr;
    if (!check(cudaMalloc(&s->constant, sizeof(double) * 2 * s->nodes), "cudaMalloc", error)) return failed();
    if (!check(cudaMalloc(&deviceModes, sizeof(Mode) * (modes.size() + 1)), "cudaMalloc", error)) return failed();
    bool ok = check(cudaMemcpy(deviceModes, modes.data(), sizeof(Mode) * modes.size(), cudaMemcpyHostToDevice), "cudaMemcpy", error);
    if (ok) {
      injectionKernel<<<grid, kBlock>>>(s->constant, s->nodes, geometry.NR, geometry.rank * geometry.LX, geometry.globalX, geometry.globalY,
                                        deviceModes, (int)modes.size());
      ok = check(cudaGetLastError(), "injectionKernel", error) && check(cudaDeviceSynchronize(), "injectionKernel", error);
    }
    cudaFree(deviceModes);
    if (!ok) return failed();
  }
  if (spec.removal) {
    const std::vector<Mode> modes = shellModes(geometry.globalX, geometry.globalY, spec.removalKMin, spec.removalKMax, 1.0);
    s->removalModeCount = (int)modes.size();
    s->projectBlocks = (int)std::min<long long>(kProjectBlocks, (s->nodes + kBlock - 1) / kBlock);
    if (!check(cudaMalloc(&s->removalModes, sizeof(Mode) * (modes.size() + 1)), "cudaMalloc", error)) return failed();
    if (!check(cudaMemcpy(s->removalModes, modes.data(), sizeof(Mode) * modes.size(), cudaMemcpyHostToDevice), "cudaMemcpy", error)) return failed();
    if (!check(cudaMalloc(&s->partials, sizeof(double) * 4 * ((size_t)s->removalModeCount * s->projectBlocks + 1)), "cudaMalloc", error)) return failed();
    if (!check(cudaMalloc(&s->projections, sizeof(double) * 4 * ((size_t)s->removalModeCount + 1)), "cudaMalloc", error)) return failed();
  }
  return s;
}

void shellForceDestroy(ShellForce* plan) { delete plan; }

bool shellForceIsTimeDependent(const ShellForce* plan) { return plan->spec.removal; }

int shellForceInitial(ShellForce* s, void* force, long long fieldStride, cudaStream_t stream, unsigned long long* launches, std::string* error) {
  const ShellForceGeometry& g = s->g;
  const unsigned grid = (unsigned)((s->nodes + kBlock - 1) / kBlock);
  if (s->constant) {
    if (g.elementSize == 8) copyConstantKernel<double><<<grid, kBlock, 0, stream>>>(static_cast<double*>(force), fieldStride, s->nodes, s->constant);
    else copyConstantKernel<float><<<grid, kBlock, 0, stream>>>(static_cast<float*>(force), fieldStride, s->nodes, s->constant);
    if (launches) *launches += 1;
  } else if (!check(cudaMemsetAsync(force, 0, (size_t)(fieldStride + s->nodes) * g.elementSize, stream), "cudaMemsetAsync", error)) {
    return -1;
  }
  return check(cudaGetLastError(), "spectral force kernels", error) ? 0 : -1;
}

int shellForceUpdate(ShellForce* s, const void* density, const void* velocity, void* force, long long fieldStride, const NcclApi* nccl,
                     ncclComm_t comm, cudaStream_t stream, unsigned long long* launches, std::string* error) {
  const ShellForceGeometry& g = s->g;
  const unsigned grid = (unsigned)((s->nodes + kBlock - 1) / kBlock);
  const int xOffset = g.rank * g.LX;
  unsigned long long count = 0;
  if (!s->spec.removal || s->removalModeCount == 0) {
    return shellForceInitial(s, force, fieldStride, stream, launches, error);  // nothing depends on the fields
  } else {
    const dim3 projectGrid((unsigned)s->projectBlocks, (unsigned)s->removalModeCount);
    if (g.elementSize == 8)
      projectKernel<double><<<projectGrid, kBlock, 0, stream>>>(static_cast<const double*>(density), static_cast<const double*>(velocity), fieldStride,
                                                              s->nodes, g.NR, xOffset, g.globalX, g.globalY, s->removalModes, s->partials);
    else
      projectKernel<float><<<projectGrid, kBlock, 0, stream>>>(static_cast<const float*>(density), static_cast<const float*>(velocity), fieldStride,
                                                             s->nodes, g.NR, xOffset, g.globalX, g.globalY, s->removalModes, s->partials);
    reduceKernel<<<(unsigned)((s->removalModeCount * 4 + 127) / 128), 128, 0, stream>>>(s->partials, s->projectBlocks, s->removalModeCount, s->projections);
    count += 2;
    if (g.nranks > 1) {
      if (!nccl || !comm) {
        if (error) *error = "the spectral force on several ranks needs the NCCL communicator (mlbm_comm_init)";
        return -1;
      }
      const ncclResult_t result = nccl->AllReduce(s->projections, s->projections, (size_t)s->removalModeCount * 4, ncclDouble, ncclSum, comm, stream);
      if (result != ncclSuccess) {
        if (error) *error = std::string("spectral force all-reduce: ") + nccl->GetErrorString(result);
        return -1;
      }
      ++count;
    }
    const double volume = (double)g.globalX * g.globalY;
    const double scaleX = -s->spec.removalAmplitude[0] / volume, scaleY = -s->spec.removalAmplitude[1] / volume;
    if (g.elementSize == 8)
      synthesisKernel<double><<<grid, kBlock, 0, stream>>>(static_cast<double*>(force), fieldStride, s->nodes, g.NR, xOffset, g.globalX, g.globalY,
                                                          s->removalModes, s->removalModeCount, s->projections, scaleX, scaleY, s->constant);
    else
      synthesisKernel<float><<<grid, kBlock, 0, stream>>>(static_cast<float*>(force), fieldStride, s->nodes, g.NR, xOffset, g.globalX, g.globalY,
                                                         s->removalModes, s->removalModeCount, s->projections, scaleX, scaleY, s->constant);
    ++count;
  }
  if (!check(cudaGetLastError(), "spectral force kernels", error)) return -1;
  if (launches) *launches += count;
  return 0;
}

}  // namespace mlbm
