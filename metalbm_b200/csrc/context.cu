// context.cu -- the C-ABI of include/metalbm_b200.h: device memory, the per-step orchestration
// (Algorithm::iterate, Algorithm.h:326-358 / 392-447), the x-slab halo exchange (Communication.h:134-180)
// and the scalar observables (AnalysisList.h:55-73) for one rank == one GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <string>
#include <vector>

#include "context.h"

namespace mlbm {

StepKernel lookupStepKernel_d2q5_f64(int, int, int);
StepKernel lookupStepKernel_d2q5_f32(int, int, int);
StepKernel lookupStepKernel_d2q9_f64(int, int, int);
StepKernel lookupStepKernel_d2q9_f32(int, int, int);
StepKernel lookupStepKernel_d3q15_f64(int, int, int);
StepKernel lookupStepKernel_d3q15_f32(int, int, int);
StepKernel lookupStepKernel_d3q19_f64(int, int, int);
StepKernel lookupStepKernel_d3q19_f32(int, int, int);
StepKernel lookupStepKernel_d3q27_f64(int, int, int);
StepKernel lookupStepKernel_d3q27_f32(int, int, int);
StepKernel lookupStepKernel_d2q13_f64(int, int, int);
StepKernel lookupStepKernel_d2q13_f32(int, int, int);
StepKernel lookupStepKernel_d2q17_f64(int, int, int);
StepKernel lookupStepKernel_d2q17_f32(int, int, int);
StepKernel lookupStepKernel_d2q21_f64(int, int, int);
StepKernel lookupStepKernel_d2q21_f32(int, int, int);
StepKernel lookupStepKernel_d3q33_f64(int, int, int);
StepKernel lookupStepKernel_d3q33_f32(int, int, int);

StepKernel lookupStepKernel(int lattice, int collision, int equilibrium, int scheme, int dtype) {
  const bool f64 = dtype == MLBM_F64;
  switch (lattice) {
    case kD2Q5: return f64 ? lookupStepKernel_d2q5_f64(collision, equilibrium, scheme) : lookupStepKernel_d2q5_f32(collision, equilibrium, scheme);
    case kD2Q9: return f64 ? lookupStepKernel_d2q9_f64(collision, equilibrium, scheme) : lookupStepKernel_d2q9_f32(collision, equilibrium, scheme);
    case kD3Q15: return f64 ? lookupStepKernel_d3q15_f64(collision, equilibrium, scheme) : lookupStepKernel_d3q15_f32(collision, equilibrium, scheme);
    case kD3Q19: return f64 ? lookupStepKernel_d3q19_f64(collision, equilibrium, scheme) : lookupStepKernel_d3q19_f32(collision, equilibrium, scheme);
    case kD3Q27: return f64 ? lookupStepKernel_d3q27_f64(collision, equilibrium, scheme) : lookupStepKernel_d3q27_f32(collision, equilibrium, scheme);
    case kD2Q13: return f64 ? lookupStepKernel_d2q13_f64(collision, equilibrium, scheme) : lookupStepKernel_d2q13_f32(collision, equilibrium, scheme);
    case kD2Q17: return f64 ? lookupStepKernel_d2q17_f64(collision, equilibrium, scheme) : lookupStepKernel_d2q17_f32(collision, equilibrium, scheme);
    case kD2Q21: return f64 ? lookupStepKernel_d2q21_f64(collision, equilibrium, scheme) : lookupStepKernel_d2q21_f32(collision, equilibrium, scheme);
    case kD3Q33: return f64 ? lookupStepKernel_d3q33_f64(collision, equilibrium, scheme) : lookupStepKernel_d3q33_f32(collision, equilibrium, scheme);
    default: return nullptr;
  }
}


// ------------------------------------------------------------------------------------------------
// small device kernels around the fused step
// ------------------------------------------------------------------------------------------------

// Deterministic reduction of the per-block partials written by the fused kernel on stored steps: every block sums a
// fixed contiguous share into `stage`, the block that finishes last (atomic ticket) adds the shares in index order.
constexpr int kReduceThreads = 256;

__global__ void __launch_bounds__(kReduceThreads)
reduceObservablesKernel(const double* __restrict__ partials, long long blocks, double* __restrict__ stage,
                        unsigned* __restrict__ ticket, double* __restrict__ out) {
  __shared__ double scratch[kObservableSlots][kReduceThreads];
  __shared__ bool last;
  const long long share = (blocks + gridDim.x - 1) / gridDim.x;
  const long long begin = share * blockIdx.x;
  const long long end = begin + share < blocks ? begin + share : blocks;
  double e = 0.0, ms = 0.0, s2 = 0.0;
  for (long long i = begin + threadIdx.x; i < end; i += blockDim.x) {
    e += partials[i * kObservableSlots + 0];
    ms += partials[i * kObservableSlots + 1];
    s2 = fmax(s2, partials[i * kObservableSlots + 2]);
  }
  auto blockReduce = [&](double& a, double& b, double& c) {
    scratch[0][threadIdx.x] = a;
    scratch[1][threadIdx.x] = b;
    scratch[2][threadIdx.x] = c;
    __syncthreads();
    for (int width = kReduceThreads / 2; width > 0; width >>= 1) {
      if ((int)threadIdx.x < width) {
        scratch[0][threadIdx.x] += scratch[0][threadIdx.x + width];
        scratch[1][threadIdx.x] += scratch[1][threadIdx.x + width];
        scratch[2][threadIdx.x] = fmax(scratch[2][threadIdx.x], scratch[2][threadIdx.x + width]);
      }
      __syncthreads();
    }
    a = scratch[0][0]; b = scratch[1][0]; c = scratch[2][0];
    __syncthreads();
  };
  blockReduce(e, ms, s2);
  if (threadIdx.x == 0) {
    stage[blockIdx.x * kObservableSlots + 0] = e;
    stage[blockIdx.x * kObservableSlots + 1] = ms;
    stage[blockIdx.x * kObservableSlots + 2] = s2;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  e = 0.0; ms = 0.0; s2 = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    e += __ldcg(stage + i * kObservableSlots + 0);
    ms += __ldcg(stage + i * kObservableSlots + 1);
    s2 = fmax(s2, __ldcg(stage + i * kObservableSlots + 2));
  }
  blockReduce(e, ms, s2);
  if (threadIdx.x == 0) {
    out[0] = e;
    out[1] = ms;
    out[2] = s2;
    *ticket = 0;
  }
}

// Direct peer halos: per-step handshake through two 64-bit words in every rank's memory.
//   flags[0] = number of steps whose halo plane 0 the LEFT neighbour has delivered, flags[1] likewise for plane LX+1 / RIGHT.
// A neighbour raises the word only after its boundary kernel has finished, so `flags >= step - 1` also says that the
// neighbour is done reading the halo planes this rank is about to overwrite.
__global__ void waitPeerFlagsKernel(const unsigned long long* flags, unsigned long long target, int* timedOut) {
  const volatile unsigned long long* word = flags + threadIdx.x;
  const long long begin = clock64();
  while (*word < target) {
    if (clock64() - begin > 40000000000ll) {  // ~20 s at 2 GHz: a lost neighbour must not hang the box
      *timedOut = 1;
      __threadfence_system();
      asm volatile("trap;");
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

__global__ void signalPeersKernel(unsigned long long* leftNeighbourFlags, unsigned long long* rightNeighbourFlags, unsigned long long step) {
  __threadfence_system();  // the boundary kernel before this one in the stream has completed: order its peer stores first
  if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(leftNeighbourFlags + 1) = step;   // I am its RIGHT neighbour
  if (threadIdx.x == 1) *reinterpret_cast<volatile unsigned long long*>(rightNeighbourFlags + 0) = step;  // I am its LEFT neighbour
  __threadfence_system();
}

// Staged pack / unpack (large blocks, or MLBM_STAGED_COPY=1): the host keeps every population as rows of NR values with a pitch of
// `pitch` >= NR values (the FFTW padding of lSD, Domain.h:53-57).  Instead of one pitched DMA per population (2 KB rows),
// the padded block crosses PCIe as ONE contiguous copy and the padding is stripped / added on the device at HBM speed.
template <typename StoreT>
__global__ void stripPaddingKernel(const StoreT* __restrict__ padded, StoreT* __restrict__ dense, long long rows, int NR, long long pitch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * NR) return;
  const long long row = i / NR;
  dense[i] = padded[row * pitch + (i - row * NR)];
}
template <typename StoreT>
__global__ void addPaddingKernel(const StoreT* __restrict__ dense, StoreT* __restrict__ padded, long long rows, int NR, long long pitch) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= rows * pitch) return;
  const long long row = j / pitch, r = j - row * pitch;
  padded[j] = r < NR ? dense[row * NR + r] : (StoreT)0;   // the padding values arrive as zeros (they are scratch: Domain.h:53-57)
}

template <typename StoreT> __global__ void fillKernel(StoreT* data, long long count, StoreT value) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) data[i] = value;
}

// f = feq(rho, u) (initDistribution, Initialize.h:106-117) into the interior of an SoA buffer
template <class L, int EQ, typename StoreT>
__global__ void initEquilibriumKernel(StoreT* __restrict__ populations, const StoreT* __restrict__ density,
                                      const StoreT* __restrict__ velocity, long long stride, long long plane,
                                      long long fieldStride, long long nodes) {
  const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= nodes) return;
  double u[3] = {0.0, 0.0, 0.0};
  double u2 = 0.0;
#pragma unroll
  for (int d = 0; d < L::D; ++d) {
    u[d] = (double)velocity[d * fieldStride + node];
    u2 += u[d] * u[d];
  }
  const double rho = (double)density[node];
  EquilibriumCoefficients<L, EQ> eq;
  eq.set(u, u2);
  staticFor<0, L::Q>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    populations[q * stride + plane + node] = (StoreT)(rho * L::w(q) * eq.template shape<q>());
  });
}

// The synthetic initial field of the benchmark (SURVEY.md 8d "Init B": density ripple + Taylor-Green-like velocity)
// evaluated on the device at GLOBAL coordinates and written as f = feq(rho, u): no host field arrays, which is what
// the slabs that fill a GPU (1024^3 on two GPUs) need.
template <class L, int EQ, typename StoreT>
__global__ void initSyntheticKernel(StoreT* __restrict__ populations, long long stride, long long plane, long long interior, long long nodes,
                                    int NR, int xOffset, int globalX, int globalY, int globalZ, double densityAmplitude,
                                    double velocityAmplitude) {
  const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= nodes) return;
  const int x = (int)(node / plane);
  const long long inPlane = node - (long long)x * plane;
  const int m = (int)(inPlane / NR), r = (int)(inPlane - (long long)m * NR);
  double sx, cx, sy, cy, sz = 0.0, cz = 1.0;
  sincospi(2.0 * (x + xOffset) / globalX, &sx, &cx);
  sincospi(2.0 * (L::D == 3 ? m : r) / globalY, &sy, &cy);
  if (L::D == 3) sincospi(2.0 * r / globalZ, &sz, &cz);
  const double rho = 1.0 + densityAmplitude * sx * cy * cz;
  double u[3] = {0.0, 0.0, 0.0};
  if (L::D == 3) {
    u[0] = velocityAmplitude * sx * cy * cz;
    u[1] = -velocityAmplitude * cx * sy * cz;
    u[2] = 0.5 * velocityAmplitude * cx * cy * sz;
  } else {
    u[0] = velocityAmplitude * sy;
    u[1] = velocityAmplitude * cx;
  }
  double u2 = 0.0;
#pragma unroll
  for (int d = 0; d < L::D; ++d) u2 += u[d] * u[d];
  EquilibriumCoefficients<L, EQ> eq;
  eq.set(u, u2);
  staticFor<0, L::Q>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    populations[q * stride + interior + node] = (StoreT)(rho * L::w(q) * eq.template shape<q>());
  });
}

// f *= 1 + eps * n, n uniform with unit variance from a counter-based hash of (seed, population, GLOBAL node):
// the synthetic non-equilibrium initial fields of the benchmark, independent of the decomposition.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <typename StoreT>
__global__ void perturbKernel(StoreT* __restrict__ populations, long long stride, long long plane, long long nodes,
                              long long globalNodeOffset, int Q, double eps, unsigned long long seed) {
  const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= nodes) return;
  for (int q = 0; q < Q; ++q) {
    const unsigned long long bits = splitmix64(splitmix64(seed + (unsigned long long)q) ^ (unsigned long long)(globalNodeOffset + node));
    const double uniform = (double)(bits >> 11) * (1.0 / 9007199254740992.0);  // [0, 1)
    const double n = (2.0 * uniform - 1.0) * 1.7320508075688772;
    StoreT* p = populations + q * stride + plane + node;
    *p = (StoreT)((double)*p * (1.0 + eps * n));
  }
}


template <int EQ, typename StoreT>
static void launchInitEquilibrium(int lattice, cudaStream_t stream, StoreT* populations, const StoreT* density,
                                  const StoreT* velocity, long long stride, long long plane, long long fieldStride,
                                  long long nodes) {
  const int block = 128;
  const unsigned grid = (unsigned)((nodes + block - 1) / block);
  switch (lattice) {
    case kD2Q5:
      if (EQ == kTruncationMa3) initEquilibriumKernel<Lattice<kD2Q5>, kTruncationMa3, StoreT><<<grid, block, 0, stream>>>(populations, density, velocity, stride, plane, fieldStride, nodes);
      break;
    case kD2Q9: initEquilibriumKernel<Lattice<kD2Q9>, EQ, StoreT><<<grid, block, 0, stream>>>(populations, density, velocity, stride, plane, fieldStride, nodes); break;
    case kD3Q15:
      if (EQ == kTruncationMa3) initEquilibriumKernel<Lattice<kD3Q15>, kTruncationMa3, StoreT><<<grid, block, 0, stream>>>(populations, density, velocity, stride, plane, fieldStride, nodes);
      break;
    case kD3Q19:
      if (EQ == kTruncationMa3) initEquilibriumKernel<Lattice<kD3Q19>, kTruncationMa3, StoreT><<<grid, block, 0, stream>>>(populations, density, velocity, stride, plane, fieldStride, nodes);
      break;
    case kD3Q27: initEquilibriumKernel<Lattice<kD3Q27>, EQ, StoreT><<<grid, block, 0, stream>>>(populations, density, velocity, stride, plane, fieldStride, nodes); break;
#define MLBM_INIT_WIDE(LATTICE)                                                                                        \
    case LATTICE:                                                                                                        \
      if (EQ == kTruncationMa3) initEquilibriumKernel<Lattice<LATTICE>, kTruncationMa3, StoreT><<<grid, block, 0, stream>>>(populations, density, velocity, stride, plane, fieldStride, nodes); \
      break;
    MLBM_INIT_WIDE(kD2Q13) MLBM_INIT_WIDE(kD2Q17) MLBM_INIT_WIDE(kD2Q21) MLBM_INIT_WIDE(kD3Q33)
#undef MLBM_INIT_WIDE
  }
}

template <int EQ, typename StoreT>
static void launchInitSynthetic(int lattice, cudaStream_t stream, StoreT* populations, long long stride, long long plane,
                                long long interior, long long nodes, int NR, int xOffset, const int* global, double densityAmplitude,
                                double velocityAmplitude) {
  const int block = 128;
  const unsigned grid = (unsigned)((nodes + block - 1) / block);
#define MLBM_INIT_SYNTHETIC(LATTICE, EQUILIBRIUM)                                                                      \
  initSyntheticKernel<Lattice<LATTICE>, EQUILIBRIUM, StoreT><<<grid, block, 0, stream>>>(                                \
      populations, stride, plane, interior, nodes, NR, xOffset, global[0], global[1], global[2], densityAmplitude, velocityAmplitude)
  switch (lattice) {
    case kD2Q5: if (EQ == kTruncationMa3) MLBM_INIT_SYNTHETIC(kD2Q5, kTruncationMa3); break;
    case kD2Q9: MLBM_INIT_SYNTHETIC(kD2Q9, EQ); break;
    case kD3Q15: if (EQ == kTruncationMa3) MLBM_INIT_SYNTHETIC(kD3Q15, kTruncationMa3); break;
    case kD3Q19: if (EQ == kTruncationMa3) MLBM_INIT_SYNTHETIC(kD3Q19, kTruncationMa3); break;
    case kD3Q27: MLBM_INIT_SYNTHETIC(kD3Q27, EQ); break;
    case kD2Q13: if (EQ == kTruncationMa3) MLBM_INIT_SYNTHETIC(kD2Q13, kTruncationMa3); break;
    case kD2Q17: if (EQ == kTruncationMa3) MLBM_INIT_SYNTHETIC(kD2Q17, kTruncationMa3); break;
    case kD2Q21: if (EQ == kTruncationMa3) MLBM_INIT_SYNTHETIC(kD2Q21, kTruncationMa3); break;
    case kD3Q33: if (EQ == kTruncationMa3) MLBM_INIT_SYNTHETIC(kD3Q33, kTruncationMa3); break;
  }
#undef MLBM_INIT_SYNTHETIC
}

}  // namespace mlbm

using namespace mlbm;

// ------------------------------------------------------------------------------------------------
// error plumbing and slab geometry (declared in context.h)
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_lastError;

namespace mlbm {

int fail(int status, const char* format, ...) {
  char buffer[1024];
  va_list arguments;
  va_start(arguments, format);
  vsnprintf(buffer, sizeof(buffer), format, arguments);
  va_end(arguments);
  g_lastError = buffer;
  return status;
}

const char* lastError() { return g_lastError.c_str(); }

bool slabGeometry(const mlbm_config* config, SlabGeometry* g) {
  g->Q = latticeQ(config->lattice);
  if (!g->Q || config->nranks < 1 || config->global_length[0] % config->nranks) return false;
  g->H = latticeHalo(config->lattice);
  if (config->nranks > 1 && config->global_length[0] / config->nranks < g->H) return false;  // a slab thinner than the halo
  g->D = latticeDim(config->lattice);
  g->faceQ = latticeFaceQ(config->lattice);
  g->LX = config->global_length[0] / config->nranks;
  g->NM = g->D == 3 ? config->global_length[1] : 1;
  g->NR = g->D == 3 ? config->global_length[2] : config->global_length[1];
  g->plane = (long long)g->NM * g->NR;
  const long long perPopulation = g->plane * (g->LX + 2 * g->H);
  g->stride = (perPopulation + 31) / 32 * 32;  // keep every population 128-byte aligned
  // the kernels address a node INSIDE one population with 32-bit element offsets (NodeIndex, step_kernel.cuh); the
  // offsets BETWEEN populations are 64-bit.  1024^3 on 2 GPUs has 5.4e8 elements per population.
  if (perPopulation >= (1LL << 32)) return false;
  return true;
}

}  // namespace mlbm

// ------------------------------------------------------------------------------------------------
// launches of the fused kernel and the per-step orchestration
// ------------------------------------------------------------------------------------------------
static int collectProfile(mlbm_ctx* ctx) {
  for (size_t i = 0; i + 1 < ctx->profileUsed; i += 2) {
    float ms = 0.f;
    MLBM_CUDA(cudaEventElapsedTime(&ms, ctx->profileEvents[i], ctx->profileEvents[i + 1]));
    ctx->profileMs += ms;
    ctx->profileLaunches += 1;
  }
  ctx->profileUsed = 0;
  return MLBM_OK;
}

// Everything of a launch that does not depend on device memory: the scalar kernel parameters and the grid.  Shared by
// launchStep and by mlbm_launch_plan_for, the device-free mirror the CPU test-suite checks (a launch whose scalars are
// silently wrong -- beta = 0, no periodic wrap -- still runs at full speed and only a GPU parity test would notice).
static void fillLaunchScalars(const mlbm_config& config, const SlabGeometry& g, int gridR, bool entropic, int hydroShift, int x0,
                              int x1, int isStored, int planeStep, StepParams* p, dim3* grid) {
  p->stride = g.stride;
  p->plane = g.plane;
  p->LX = g.LX; p->NM = g.NM; p->NR = g.NR;
  p->x0 = x0;
  p->planeStep = planeStep;
  p->planeCount = x1 - x0;
  p->wrapX = config.nranks == 1 ? 1 : 0;
  p->isStored = isStored;
  p->hydroShift = hydroShift;
  p->hasForce = config.force == MLBM_FORCE_NONE ? 0 : (config.force >= MLBM_FORCE_FIELD ? 2 : 1);  // array-type forces read the field
  p->beta = 1.0 / (2.0 * config.tau);
  p->guoFactor = (1.0 - 1.0 / (2.0 * config.tau)) * latticeInvCs2(config.lattice);
  // entropic kernels stage their logarithm table and constants once per block: let a block walk up to 16 planes
  // (measured: +30 % on D2Q9 8192^2, +11 % on D3Q27 512^3 against one plane per block) while the grid keeps >= ~20 waves
  static const int planesOverride = getenv("MLBM_PLANES_PER_BLOCK") ? atoi(getenv("MLBM_PLANES_PER_BLOCK")) : 0;  // experiments
  p->planesPerBlock = 1;
  if (entropic && planeStep == 1) {
    const long long blocks = (long long)gridR * g.NM * p->planeCount;
    const long long wanted = 148LL * 4 * 20;
    long long planes = blocks / wanted;
    planes = planes < 1 ? 1 : (planes > 16 ? 16 : planes);
    p->planesPerBlock = planesOverride > 0 ? planesOverride : (int)planes;
  }
  *grid = dim3((unsigned)gridR, (unsigned)g.NM, (unsigned)((p->planeCount + p->planesPerBlock - 1) / p->planesPerBlock));
}

// forcing scheme -> (kernel scheme, hydrodynamic velocity shift); ShanChen shares the kernel of None (ForcingScheme.h:141-151)
static bool schemeOf(int forcingScheme, int* scheme, int* hydroShift) {
  switch (forcingScheme) {
    case MLBM_SCHEME_NONE: *scheme = kSchemeNone; *hydroShift = 0; return true;
    case MLBM_SHAN_CHEN: *scheme = kSchemeNone; *hydroShift = 1; return true;
    case MLBM_GUO: *scheme = kSchemeGuo; *hydroShift = 1; return true;
    case MLBM_EXACT_DIFFERENCE: *scheme = kSchemeEDM; *hydroShift = 1; return true;
    default: return false;
  }
}

// one launch of the fused kernel over local planes [x0, x1)
static int launchStep(mlbm_ctx* ctx, cudaStream_t stream, int x0, int x1, int isStored, bool profile, int planeStep = 1,
                      void* peerLow = nullptr, void* peerHigh = nullptr) {
  if (x1 <= x0) return MLBM_OK;
  StepParams p;
  memset(&p, 0, sizeof(p));
  p.prev = ctx->populations[ctx->current];
  p.next = ctx->populations[ctx->current ^ 1];
  p.alpha = ctx->alpha;
  p.density = ctx->density;
  p.velocity = ctx->velocity;
  p.force = ctx->force;
  p.partials = ctx->partials;
  p.newtonCounters = ctx->newtonCounters;
  for (int d = 0; d < 3; ++d) { p.forceTable[d] = ctx->forceTables[d]; p.forceAxis[d] = ctx->forceAxis[d]; }
  p.fieldStride = ctx->fieldStride;
  p.peerLow = peerLow;
  p.peerHigh = peerHigh;
  SlabGeometry geometry;
  slabGeometry(&ctx->config, &geometry);
  dim3 grid;
  fillLaunchScalars(ctx->config, geometry, ctx->gridR, ctx->alpha != nullptr, ctx->hydroShift, x0, x1, isStored, planeStep, &p, &grid);
  cudaEvent_t start = nullptr, stop = nullptr;
  if (profile) {
    if (ctx->profileUsed + 2 > ctx->profileEvents.size()) {
      if (ctx->profileEvents.size() >= 8192) {
        MLBM_CUDA(cudaStreamSynchronize(stream));
        if (int status = collectProfile(ctx)) return status;
      } else {
        for (int i = 0; i < 2; ++i) {
          cudaEvent_t event;
          MLBM_CUDA(cudaEventCreate(&event));
          ctx->profileEvents.push_back(event);
        }
      }
    }
    start = ctx->profileEvents[ctx->profileUsed];
    stop = ctx->profileEvents[ctx->profileUsed + 1];
    ctx->profileUsed += 2;
    MLBM_CUDA(cudaEventRecord(start, stream));
  }
  ctx->kernel<<<grid, kStepBlock, ctx->sharedBytes, stream>>>(p);
  if (profile) MLBM_CUDA(cudaEventRecord(stop, stream));
  MLBM_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return MLBM_OK;
}

static int ensureFields(mlbm_ctx* ctx) {
  if (ctx->density) return MLBM_OK;
  const size_t bytes = (size_t)ctx->fieldStride * ctx->elementSize;
  MLBM_CUDA(cudaMalloc(&ctx->density, bytes));
  MLBM_CUDA(cudaMalloc(&ctx->velocity, bytes * ctx->D));
  MLBM_CUDA(cudaMalloc(&ctx->force, bytes * ctx->D));
  MLBM_CUDA(cudaMemsetAsync(ctx->density, 0, bytes, ctx->computeStream));
  MLBM_CUDA(cudaMemsetAsync(ctx->velocity, 0, bytes * ctx->D, ctx->computeStream));
  MLBM_CUDA(cudaMemsetAsync(ctx->force, 0, bytes * ctx->D, ctx->computeStream));
  return MLBM_OK;
}

static int ensurePartials(mlbm_ctx* ctx) {
  if (ctx->partials) return MLBM_OK;
  MLBM_CUDA(cudaMalloc(&ctx->partials, sizeof(double) * kObservableSlots * (size_t)ctx->partialBlocks));
  MLBM_CUDA(cudaMalloc(&ctx->reduceStage, sizeof(double) * kObservableSlots * kReduceBlocks));
  MLBM_CUDA(cudaMalloc(&ctx->reduceTicket, sizeof(unsigned)));
  MLBM_CUDA(cudaMemsetAsync(ctx->reduceTicket, 0, sizeof(unsigned), ctx->computeStream));
  return MLBM_OK;
}

// the compute stream waits for the analysis of the last stored step (a no-op when none is in flight)
int mlbm::joinAnalysis(mlbm_ctx* ctx) {
  if (!ctx->analysisPending) return MLBM_OK;
  MLBM_CUDA(cudaStreamWaitEvent(ctx->computeStream, ctx->analysisDone, 0));
  ctx->analysisPending = false;
  return MLBM_OK;
}

// enqueue one Algorithm::iterate; `timed` records the events behind mlbm_timers
static int enqueueStep(mlbm_ctx* ctx, int isStored, bool timed, bool profile) {
  const bool multi = ctx->config.nranks > 1;
  if (isStored & 1) { if (int status = ensureFields(ctx)) return status; }
  if (isStored) { if (int status = ensurePartials(ctx)) return status; }
  cudaStream_t compute = ctx->computeStream;
  // the analysis of the previous stored step still reads the velocity field this step is about to overwrite
  if ((isStored & 1) && ctx->analysisPending) { if (int status = joinAnalysis(ctx)) return status; }
  if (ctx->shell && ctx->forceStale) {
    // Collision::update -> Force::update at the top of iterate (Algorithm.h:338, Collision.h:97-100, Force.h:552-558)
    std::string error;
    if (shellForceUpdate(ctx->shell, ctx->density, ctx->velocity, ctx->force, ctx->fieldStride, ctx->nccl, ctx->comm, compute, &ctx->launches, &error))
      return fail(MLBM_ERR_CUDA, "spectral force: %s", error.c_str());
    ctx->forceStale = false;
  }
  if (timed) MLBM_CUDA(cudaEventRecord(ctx->timeStart, compute));

  if (!multi) {
    if (timed) MLBM_CUDA(cudaEventRecord(ctx->timeMid, compute));
    if (int status = launchStep(ctx, compute, 0, ctx->LX, isStored, profile)) return status;
  } else if (ctx->peerAttached && ctx->config.overlap == MLBM_OVERLAP_ON) {
    // Direct peer halos: ONE kernel computes the two boundary planes and stores their outgoing populations straight
    // into the neighbours' halo planes over NVLink (no pack, no send/recv, no staging); it runs on the high-priority
    // stream next to the bulk kernel.  Replaces Communication::communicateHalos (Communication.h:134-180, 494-500).
    const unsigned long long step = ++ctx->peerEpoch;
    if (!ctx->halosValid) {
      // first step after an upload: the halo planes of the buffer about to be read were never delivered
      if (int status = exchangeHalos(ctx, ctx->current, compute)) return status;
    }
    MLBM_CUDA(cudaEventRecord(ctx->stepStart, compute));
    MLBM_CUDA(cudaStreamWaitEvent(ctx->commStream, ctx->stepStart, 0));
    if (ctx->halosValid) {
      waitPeerFlagsKernel<<<1, 2, 0, ctx->commStream>>>(ctx->peerFlags, step - 1, ctx->peerTimedOut);
      ctx->launches += 1;
    }
    if (timed) MLBM_CUDA(cudaEventRecord(ctx->timeMid, compute));
    const int next = ctx->current ^ 1;
    const bool twoPlanes = ctx->LX >= 2;
    {
      // launchStep reads ctx->current; the boundary launch goes to the communication stream
      if (int status = launchStep(ctx, ctx->commStream, 0, twoPlanes ? 2 : 1, isStored, false, twoPlanes ? ctx->LX - 1 : 1,
                                  ctx->mapped[0][next], ctx->mapped[1][next])) return status;
    }
    signalPeersKernel<<<1, 2, 0, ctx->commStream>>>(static_cast<unsigned long long*>(ctx->mapped[0][2]),
                                                    static_cast<unsigned long long*>(ctx->mapped[1][2]), step);
    ctx->launches += 1;
    MLBM_CUDA(cudaEventRecord(ctx->boundaryDone, ctx->commStream));
    if (int status = launchStep(ctx, compute, 1, ctx->LX - 1, isStored, profile)) return status;
    MLBM_CUDA(cudaStreamWaitEvent(compute, ctx->boundaryDone, 0));
    ctx->halosValid = true;  // the neighbours deliver the halo planes of the buffer that becomes current; waited for next step
  } else if (ctx->config.overlap == MLBM_OVERLAP_OFF || ctx->LX < 2 * ctx->H + 1) {
    // the reference's order (Algorithm.h:336-355): exchange the halos of the buffer about to be read, then compute
    if (!ctx->halosValid) { if (int status = exchangeHalos(ctx, ctx->current, compute)) return status; }
    if (timed) MLBM_CUDA(cudaEventRecord(ctx->timeMid, compute));
    if (int status = launchStep(ctx, compute, 0, ctx->LX, isStored, profile)) return status;
    ctx->halosValid = false;
  } else {
    // overlap (the intent of Algorithm.h:392-447): the two boundary planes first, their exchange on the
    // communication stream while the bulk planes are computed.
    if (!ctx->halosValid) {
      if (int status = exchangeHalos(ctx, ctx->current, compute)) return status;
    }
    if (timed) MLBM_CUDA(cudaEventRecord(ctx->timeMid, compute));
    if (int status = launchStep(ctx, compute, 0, ctx->H, isStored, false)) return status;
    if (int status = launchStep(ctx, compute, ctx->LX - ctx->H, ctx->LX, isStored, false)) return status;
    MLBM_CUDA(cudaEventRecord(ctx->boundaryDone, compute));
    MLBM_CUDA(cudaStreamWaitEvent(ctx->commStream, ctx->boundaryDone, 0));
    if (int status = exchangeHalos(ctx, ctx->current ^ 1, ctx->commStream)) return status;
    MLBM_CUDA(cudaEventRecord(ctx->exchangeDone, ctx->commStream));
    if (int status = launchStep(ctx, compute, ctx->H, ctx->LX - ctx->H, isStored, profile)) return status;
    MLBM_CUDA(cudaStreamWaitEvent(compute, ctx->exchangeDone, 0));
    ctx->halosValid = true;  // of the buffer that becomes current below
  }
  ctx->current ^= 1;  // std::swap(previous, next) (Algorithm.h:336), done after the step instead of before

  if (isStored) {
    reduceObservablesKernel<<<kReduceBlocks, kReduceThreads, 0, compute>>>(ctx->partials, ctx->partialBlocks, ctx->reduceStage,
                                                                          ctx->reduceTicket, ctx->deviceObservables);
    ctx->launches += 1;
    ctx->enstrophyValid = false;
    if (isStored & 1) {
      // Routine.h:129-132 + TotalEnstrophy (Analysis.h:68-98): spectral vorticity of the stored velocity
      std::string error;
      if (!ctx->spectral) {
        SpectralGeometry geometry = {ctx->D, ctx->LX, ctx->NM, ctx->NR, ctx->config.rank, ctx->config.nranks, (int)ctx->elementSize};
        ctx->spectral = spectralCreate(geometry, ctx->nccl, ctx->analysisComm, &error);
        if (!ctx->spectral) return fail(MLBM_ERR_CUDA, "spectral enstrophy: %s", error.c_str());
      }
      // The transforms, their all-to-all and the vorticity norm run on the analysis stream, next to the steps that follow
      // (the fields are only written on stored steps); whoever needs the result or the fields' buffers joins it first
      // (joinAnalysis: mlbm_observables, mlbm_power_spectra, the next stored step, mlbm_sync).  MLBM_ASYNC_ANALYSIS=0 keeps
      // everything on the compute stream.
      static const bool asyncAnalysis = !(getenv("MLBM_ASYNC_ANALYSIS") && atoi(getenv("MLBM_ASYNC_ANALYSIS")) == 0);
      cudaStream_t analysis = asyncAnalysis ? ctx->analysisStream : compute;
      if (asyncAnalysis) {
        MLBM_CUDA(cudaEventRecord(ctx->fieldsReady, compute));
        MLBM_CUDA(cudaStreamWaitEvent(analysis, ctx->fieldsReady, 0));
      }
      if (spectralEnqueue(ctx->spectral, ctx->velocity, ctx->fieldStride, ctx->deviceObservables + 3, analysis, &ctx->launches, &error))
        return fail(MLBM_ERR_CUDA, "spectral enstrophy: %s", error.c_str());
      if (asyncAnalysis) {
        MLBM_CUDA(cudaEventRecord(ctx->analysisDone, analysis));
        ctx->analysisPending = true;
      }
      ctx->fieldsStored = true;
      ctx->enstrophyValid = true;
      if (ctx->shell && shellForceIsTimeDependent(ctx->shell)) ctx->forceStale = true;  // fieldList changed
    }
    MLBM_CUDA(cudaGetLastError());
    ctx->observablesValid = true;
  }
  if (timed) MLBM_CUDA(cudaEventRecord(ctx->timeStop, compute));
  return MLBM_OK;
}

// A synchronisation that failed: say so, and say WHY when the direct peer halos are in use and the handshake kernel gave up
// (waitPeerFlagsKernel traps after ~20 s without the neighbour's flag; it raises *peerTimedOut first).  The context is
// unusable afterwards (sticky CUDA error); mlbm_destroy then skips its shutdown barrier, which the lost neighbour would hang.
static int synchronizeOrDiagnose(mlbm_ctx* ctx, cudaStream_t stream) {
  const cudaError_t error = cudaStreamSynchronize(stream);
  if (error == cudaSuccess) return MLBM_OK;
  ctx->poisoned = true;
  if (ctx->peerTimedOut && *ctx->peerTimedOut)
    return fail(MLBM_ERR_COMM, "peer halo handshake timed out on rank %d: a neighbour (rank %d or %d) never delivered its halo planes (%s)",
                ctx->config.rank, (ctx->config.rank + ctx->config.nranks - 1) % ctx->config.nranks, (ctx->config.rank + 1) % ctx->config.nranks,
                cudaGetErrorString(error));
  return fail(MLBM_ERR_CUDA, "CUDA failed with %s", cudaGetErrorString(error));
}

// ------------------------------------------------------------------------------------------------
// C entry points
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* mlbm_last_error(void) { return g_lastError.c_str(); }
int mlbm_abi_version(void) { return MLBM_ABI_VERSION; }

int mlbm_destroy(mlbm_ctx* ctx) {
  if (!ctx) return MLBM_OK;
  cudaSetDevice(ctx->device);
  if (ctx->computeStream) cudaStreamSynchronize(ctx->computeStream);
  if (ctx->commStream) cudaStreamSynchronize(ctx->commStream);
  if (ctx->analysisStream) cudaStreamSynchronize(ctx->analysisStream);
  if (ctx->peerAttached && ctx->comm && ctx->nccl && ctx->deviceObservables && !ctx->poisoned) {
    // the neighbours store into this rank's halo planes and flags: nobody frees before everybody has drained its streams
    if (ctx->nccl->AllReduce(ctx->deviceObservables, ctx->deviceObservables, 1, ncclDouble, ncclSum, ctx->comm, ctx->computeStream) == ncclSuccess)
      cudaStreamSynchronize(ctx->computeStream);
  }
  for (int side = 0; side < 2; ++side)
    if (ctx->mappedOwned[side])
      for (void* pointer : ctx->mapped[side]) if (pointer) cudaIpcCloseMemHandle(pointer);
  if (ctx->peerFlags) cudaFree(ctx->peerFlags);
  if (ctx->peerTimedOut) cudaFreeHost(ctx->peerTimedOut);
  if (ctx->spectral) spectralDestroy(ctx->spectral);
  if (ctx->shell) shellForceDestroy(ctx->shell);
  if (ctx->analysisComm && ctx->analysisComm != ctx->comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->analysisComm);
  if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->comm);
  for (void* pointer : {ctx->populations[0], ctx->populations[1], ctx->alpha, ctx->density, ctx->velocity, ctx->force,
                        (void*)ctx->partials, (void*)ctx->newtonCounters, ctx->staging, (void*)ctx->reduceStage, (void*)ctx->reduceTicket, (void*)ctx->deviceObservables,
                        (void*)ctx->forceTables[0], (void*)ctx->forceTables[1], (void*)ctx->forceTables[2]})
    if (pointer) cudaFree(pointer);
  for (cudaEvent_t event : {ctx->boundaryDone, ctx->exchangeDone, ctx->bulkDone, ctx->stepStart, ctx->timeStart, ctx->timeMid, ctx->timeStop, ctx->fieldsReady, ctx->analysisDone})
    if (event) cudaEventDestroy(event);
  for (cudaEvent_t event : ctx->profileEvents) cudaEventDestroy(event);
  for (cudaEvent_t event : ctx->marks) if (event) cudaEventDestroy(event);
  if (ctx->computeStream) cudaStreamDestroy(ctx->computeStream);
  if (ctx->commStream) cudaStreamDestroy(ctx->commStream);
  if (ctx->analysisStream) cudaStreamDestroy(ctx->analysisStream);
  delete ctx;
  return MLBM_OK;
}

int mlbm_create(const mlbm_config* config, mlbm_ctx** out) {
  if (!config || !out) return fail(MLBM_ERR_INVALID, "null argument");
  *out = nullptr;
  if (config->abi_version != MLBM_ABI_VERSION) return fail(MLBM_ERR_INVALID, "mlbm_config.abi_version %d != %d", config->abi_version, MLBM_ABI_VERSION);
  const int Q = latticeQ(config->lattice);
  if (!Q) return fail(MLBM_ERR_INVALID, "unknown lattice %d", config->lattice);
  const int D = latticeDim(config->lattice);
  if (config->dtype != MLBM_F64 && config->dtype != MLBM_F32) return fail(MLBM_ERR_INVALID, "unknown dtype %d", config->dtype);
  if (!(config->tau > 0.5)) return fail(MLBM_ERR_INVALID, "relaxation time must exceed 0.5 (got %g)", config->tau);
  if (config->nranks < 1 || config->rank < 0 || config->rank >= config->nranks) return fail(MLBM_ERR_INVALID, "bad rank %d of %d", config->rank, config->nranks);
  for (int d = 0; d < D; ++d)
    if (config->global_length[d] < 1) return fail(MLBM_ERR_INVALID, "global_length[%d] = %d", d, config->global_length[d]);
  if (config->global_length[0] % config->nranks) return fail(MLBM_ERR_INVALID, "nranks %d does not divide globalLengthX %d (Domain.h:22-24)", config->nranks, config->global_length[0]);
  if (config->nranks > 1 && config->global_length[0] / config->nranks < latticeHalo(config->lattice))
    return fail(MLBM_ERR_INVALID, "slabs of %d planes are thinner than the lattice's halo of %d", config->global_length[0] / config->nranks, latticeHalo(config->lattice));

  int collision, scheme, hydroShift;
  switch (config->collision) {
    case MLBM_BGK: collision = kBGK; break;
    // identical to ELBM in the reference snapshot: their calculateAlpha overrides are dead code (Collision.h:239, 705-723)
    case MLBM_ELBM: case MLBM_FORCED_NR_ELBM: case MLBM_APPROACHED_ELBM: case MLBM_MALASPINAS_ELBM:
    case MLBM_ESSENTIALLY1_ELBM: case MLBM_ESSENTIALLY2_ELBM: case MLBM_FORCED_BNR_ELBM: collision = kELBM; break;
    case MLBM_FORCED_NR_ELBM_FORCING: collision = kELBMForcing; break;  // alpha solved on the forced populations (Collision.h:727-857)
    default: return fail(MLBM_ERR_INVALID, "unknown collision %d", config->collision);
  }
  if (!schemeOf(config->forcing_scheme, &scheme, &hydroShift)) return fail(MLBM_ERR_INVALID, "unknown forcing scheme %d", config->forcing_scheme);
  if (config->force < MLBM_FORCE_NONE || config->force > MLBM_FORCE_TURBULENT_2D) return fail(MLBM_ERR_INVALID, "unknown force %d", config->force);
  if (config->force >= MLBM_FORCE_CONSTANT_SHELL) {
    if (D != 2) return fail(MLBM_ERR_INVALID, "the spectral forces are rebuilt for 2-D lattices only (the reference's 3-D variants corrupt its heap, Force.h:341-355); use MLBM_FORCE_FIELD");
    if (config->force_k_min < 0 || config->force_k_max < config->force_k_min) return fail(MLBM_ERR_INVALID, "bad shell [%d, %d]", config->force_k_min, config->force_k_max);
    if (config->force == MLBM_FORCE_TURBULENT_2D && (config->removal_k_min < 0 || config->removal_k_max < config->removal_k_min))
      return fail(MLBM_ERR_INVALID, "bad removal shell [%d, %d]", config->removal_k_min, config->removal_k_max);
  }
  if (config->equilibrium != MLBM_TRUNCATION_MA3 && config->equilibrium != MLBM_EXACT) return fail(MLBM_ERR_INVALID, "unknown equilibrium %d", config->equilibrium);
  StepKernel kernel = lookupStepKernel(config->lattice, collision, config->equilibrium, scheme, config->dtype);
  if (!kernel) return fail(MLBM_ERR_INVALID, "no kernel for this lattice/equilibrium combination (the exact equilibrium is built for D2Q9 and D3Q27, Equilibrium.h:60-81, 106-126; the reference also defines it for D1Q3 and D2Q13, :36-58, 83-103, which this library does not cover)");
  // the launch grid is (ceil(NR / 128), NM, x planes [/ planes per block]): CUDA caps grid.y and grid.z at 65535
  {
    SlabGeometry check;
    if (slabGeometry(config, &check) && (check.NM > 65535 || (collision == kBGK && check.LX > 65535) || check.LX > 65535 * 16))
      return fail(MLBM_ERR_INVALID, "local extents %d x %d planes exceed the launch grid (65535 rows; 65535 x planes per rank for BGK, 16 times that for the entropic collisions): use more ranks or the transposed orientation", check.LX, check.NM);
  }

  int deviceCount = 0;
  cudaError_t error = cudaGetDeviceCount(&deviceCount);
  if (error != cudaSuccess || deviceCount == 0)
    return fail(MLBM_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                error == cudaSuccess ? "device count is 0" : cudaGetErrorString(error));
  const int device = config->device >= 0 ? config->device : config->rank % deviceCount;  // CUDAInitializer.h:23-26
  if (device >= deviceCount) return fail(MLBM_ERR_INVALID, "device %d of %d", device, deviceCount);
  MLBM_CUDA(cudaSetDevice(device));

  mlbm_ctx* ctx = new mlbm_ctx();
  ctx->config = *config;
  ctx->device = device;
  ctx->D = D; ctx->Q = Q; ctx->faceQ = latticeFaceQ(config->lattice);
  SlabGeometry geometry;
  if (!slabGeometry(config, &geometry)) {
    delete ctx;
    return fail(MLBM_ERR_INVALID, "a population of this slab has 2^32 elements or more (32-bit in-population offsets); use more ranks");
  }
  ctx->LX = geometry.LX;
  ctx->NM = geometry.NM;
  ctx->NR = geometry.NR;
  ctx->elementSize = config->dtype == MLBM_F64 ? 8 : 4;
  ctx->plane = geometry.plane;
  ctx->H = geometry.H;
  ctx->interior = geometry.H * geometry.plane;
  ctx->nodes = ctx->plane * ctx->LX;
  ctx->fieldStride = (ctx->nodes + 31) / 32 * 32;  // every field component 128-byte aligned (cuFFT reads them as double2)
  ctx->stride = geometry.stride;
  haloPlan(config, &ctx->haloMessages);
  ctx->gridR = (ctx->NR + kStepBlock - 1) / kStepBlock;
  ctx->partialBlocks = (long long)ctx->gridR * ctx->NM * ctx->LX;
  ctx->kernel = kernel;
  ctx->hydroShift = hydroShift;
  if (collision != kBGK) {
    ctx->sharedBytes = entropicSharedBytes(Q, logTableInShared(Q));
    cudaError_t attributeError = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->sharedBytes);
    if (attributeError != cudaSuccess) {
      const int bytes = ctx->sharedBytes;
      delete ctx;
      return fail(MLBM_ERR_CUDA, "cudaFuncSetAttribute(%d bytes of shared memory): %s", bytes, cudaGetErrorString(attributeError));
    }
  }

  auto cleanup = [&](int status) { mlbm_destroy(ctx); return status; };
#define MLBM_CREATE_CUDA(call)                                                                          \
  do {                                                                                                  \
    cudaError_t error_ = (call);                                                                        \
    if (error_ != cudaSuccess)                                                                          \
      return cleanup(fail(error_ == cudaErrorMemoryAllocation ? MLBM_ERR_NOMEM : MLBM_ERR_CUDA,         \
                          "[%s:%d] CUDA failed with %s", __FILE__, __LINE__, cudaGetErrorString(error_))); \
  } while (0)

  int leastPriority = 0, greatestPriority = 0;
  MLBM_CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&leastPriority, &greatestPriority));
  MLBM_CREATE_CUDA(cudaStreamCreateWithPriority(&ctx->computeStream, cudaStreamNonBlocking, leastPriority));
  MLBM_CREATE_CUDA(cudaStreamCreateWithPriority(&ctx->commStream, cudaStreamNonBlocking, greatestPriority));
  MLBM_CREATE_CUDA(cudaStreamCreateWithPriority(&ctx->analysisStream, cudaStreamNonBlocking, leastPriority));
  for (cudaEvent_t* event : {&ctx->fieldsReady, &ctx->analysisDone}) MLBM_CREATE_CUDA(cudaEventCreateWithFlags(event, cudaEventDisableTiming));
  for (cudaEvent_t* event : {&ctx->boundaryDone, &ctx->exchangeDone, &ctx->bulkDone, &ctx->stepStart})
    MLBM_CREATE_CUDA(cudaEventCreateWithFlags(event, cudaEventDisableTiming));
  for (cudaEvent_t* event : {&ctx->timeStart, &ctx->timeMid, &ctx->timeStop}) MLBM_CREATE_CUDA(cudaEventCreate(event));

  const size_t bufferBytes = (size_t)ctx->stride * Q * ctx->elementSize;
  for (int i = 0; i < 2; ++i) {
    MLBM_CREATE_CUDA(cudaMalloc(&ctx->populations[i], bufferBytes));
    MLBM_CREATE_CUDA(cudaMemsetAsync(ctx->populations[i], 0, bufferBytes, ctx->computeStream));
  }
  if (collision != kBGK) {
    // initAlpha: the alpha field starts at 2 (Initialize.h:82-88)
    MLBM_CREATE_CUDA(cudaMalloc(&ctx->alpha, (size_t)ctx->nodes * ctx->elementSize));
    const unsigned grid = (unsigned)((ctx->nodes + 255) / 256);
    if (config->dtype == MLBM_F64) fillKernel<double><<<grid, 256, 0, ctx->computeStream>>>(static_cast<double*>(ctx->alpha), ctx->nodes, 2.0);
    else fillKernel<float><<<grid, 256, 0, ctx->computeStream>>>(static_cast<float*>(ctx->alpha), ctx->nodes, 2.0f);
    MLBM_CREATE_CUDA(cudaGetLastError());
  }
  MLBM_CREATE_CUDA(cudaMalloc(&ctx->deviceObservables, 4 * sizeof(double)));
  MLBM_CREATE_CUDA(cudaMemsetAsync(ctx->deviceObservables, 0, 4 * sizeof(double), ctx->computeStream));

  // Force profiles, evaluated on the host with libm exactly like Force::setForce does per node
  // (Force.h:154-159 Constant, :208-215 Sinusoidal, :262-267 Kolmogorov) at LOCAL coordinates (Collision.h:86).
  if (config->force == MLBM_FORCE_FIELD) {
    // the kernel reads the force field from the first step on: it exists (zero) before mlbm_set_force_field fills it
    if (int status = ensureFields(ctx)) return cleanup(status);
  } else if (config->force >= MLBM_FORCE_CONSTANT_SHELL) {
    // setForceArray in the Collision ctor (Collision.h:51-54): the array exists from the start; its time-dependent part is
    // remade from the stored fields at the top of the step that follows a change of them (Force::update, Algorithm.h:338)
    if (int status = ensureFields(ctx)) return cleanup(status);
    ShellForceGeometry shellGeometry = {ctx->LX, ctx->NR, config->rank, config->nranks, config->global_length[0], config->global_length[1], (int)ctx->elementSize};
    ShellForceSpec spec = {};
    spec.injection = config->force == MLBM_FORCE_CONSTANT_SHELL || config->force == MLBM_FORCE_TURBULENT_2D;
    spec.injectionAmplitude = config->force_amplitude[0];
    spec.injectionKMin = config->force_k_min;
    spec.injectionKMax = config->force_k_max;
    spec.removal = config->force != MLBM_FORCE_CONSTANT_SHELL;
    const bool turbulent = config->force == MLBM_FORCE_TURBULENT_2D;
    for (int d = 0; d < 2; ++d) spec.removalAmplitude[d] = turbulent ? config->removal_amplitude[d] : config->force_amplitude[d];
    spec.removalKMin = turbulent ? config->removal_k_min : config->force_k_min;
    spec.removalKMax = turbulent ? config->removal_k_max : config->force_k_max;
    std::string shellError;
    ctx->shell = shellForceCreate(shellGeometry, spec, &shellError);
    if (!ctx->shell) return cleanup(fail(MLBM_ERR_CUDA, "spectral force: %s", shellError.c_str()));
    if (shellForceInitial(ctx->shell, ctx->force, ctx->fieldStride, ctx->computeStream, &ctx->launches, &shellError))
      return cleanup(fail(MLBM_ERR_CUDA, "spectral force: %s", shellError.c_str()));
    ctx->fieldsStored = true;  // the force array is meaningful from the start (writeForce, Collision.h:51-54)
  } else if (config->force != MLBM_FORCE_NONE) {
    const int extent[3] = {ctx->LX, D == 3 ? config->global_length[1] : config->global_length[1], D == 3 ? config->global_length[2] : 1};
    const int kernelAxisOf[3] = {0, D == 3 ? 1 : 2, 2};  // physical axis -> kernel axis
    for (int d = 0; d < D; ++d) {
      int physicalAxis = -1;
      if (config->force == MLBM_FORCE_CONSTANT) physicalAxis = 0;
      if (config->force == MLBM_FORCE_SINUSOIDAL) physicalAxis = d;
      if (config->force == MLBM_FORCE_KOLMOGOROV && d == 0) physicalAxis = 1;
      if (physicalAxis < 0) continue;
      std::vector<double> table((size_t)extent[physicalAxis]);
      for (unsigned i = 0; i < (unsigned)table.size(); ++i) {
        if (config->force == MLBM_FORCE_CONSTANT) table[i] = config->force_amplitude[d];
        else if (config->force == MLBM_FORCE_SINUSOIDAL) table[i] = config->force_amplitude[d] * sin(i * 2 * M_PI / config->force_wavelength[d]);
        else table[i] = config->force_amplitude[0] * sin(i * 2 * M_PI / config->force_wavelength[0]);
      }
      MLBM_CREATE_CUDA(cudaMalloc(&ctx->forceTables[d], table.size() * sizeof(double)));
      MLBM_CREATE_CUDA(cudaMemcpy(ctx->forceTables[d], table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice));
      ctx->forceAxis[d] = kernelAxisOf[physicalAxis];
    }
  }
  MLBM_CREATE_CUDA(cudaStreamSynchronize(ctx->computeStream));
#undef MLBM_CREATE_CUDA
  *out = ctx;
  return MLBM_OK;
}


int mlbm_launch_plan_for(const mlbm_config* config, int x0, int x1, int isStored, int planeStep, mlbm_launch_plan* out) {
  if (!config || !out) return fail(MLBM_ERR_INVALID, "null argument");
  SlabGeometry g;
  if (!slabGeometry(config, &g)) return fail(MLBM_ERR_INVALID, "bad lattice or nranks does not divide globalLengthX");
  int scheme, hydroShift;
  if (!schemeOf(config->forcing_scheme, &scheme, &hydroShift)) return fail(MLBM_ERR_INVALID, "unknown forcing scheme %d", config->forcing_scheme);
  if (!(config->tau > 0.5)) return fail(MLBM_ERR_INVALID, "relaxation time must exceed 0.5 (got %g)", config->tau);
  if (x0 < 0 || x1 > g.LX || x1 <= x0 || planeStep < 1 || x0 + (x1 - x0 - 1) * planeStep >= g.LX) return fail(MLBM_ERR_INVALID, "bad plane range");
  const bool entropic = config->collision != MLBM_BGK;
  const int gridR = (g.NR + kStepBlock - 1) / kStepBlock;
  StepParams p;
  memset(&p, 0, sizeof(p));
  dim3 grid;
  fillLaunchScalars(*config, g, gridR, entropic, hydroShift, x0, x1, isStored, planeStep, &p, &grid);
  memset(out, 0, sizeof(*out));
  out->grid[0] = (int32_t)grid.x; out->grid[1] = (int32_t)grid.y; out->grid[2] = (int32_t)grid.z;
  out->block = kStepBlock;
  out->shared_bytes = entropic ? entropicSharedBytes(g.Q, logTableInShared(g.Q)) : 0;
  out->x0 = p.x0; out->plane_step = p.planeStep; out->plane_count = p.planeCount; out->planes_per_block = p.planesPerBlock;
  out->local_length[0] = p.LX; out->local_length[1] = p.NM; out->local_length[2] = p.NR;
  out->wrap_x = p.wrapX; out->is_stored = p.isStored; out->hydro_shift = p.hydroShift; out->has_force = p.hasForce;
  out->stride = (uint64_t)p.stride; out->plane = (uint64_t)p.plane;
  out->beta = p.beta; out->guo_factor = p.guoFactor;
  return MLBM_OK;
}

// host (reference local-padded SoA) <-> device interior planes, one strided copy per population
static int copyDistribution(mlbm_ctx* ctx, void* host, size_t componentStride, size_t paddedY, size_t paddedZ, bool upload) {
  if (!ctx || !host) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  const size_t es = ctx->elementSize;
  const bool threeD = ctx->D == 3;
  const size_t hostPitch = threeD ? paddedZ : paddedY;  // elements between consecutive host rows
  if (hostPitch < (size_t)ctx->NR || (threeD && paddedY < (size_t)ctx->NM)) return fail(MLBM_ERR_INVALID, "padded lengths smaller than the local lengths");
  const bool uniform = !threeD || paddedY == (size_t)ctx->NM;
  // MLBM_STAGED_COPY=1 / 0 forces the staged route (see stripPaddingKernel) on / off; unset: staged for blocks of 32 MB and
  // more, where one contiguous DMA per population beats hundreds of thousands of 2 KB rows, pitched copies below
  static const int stagedCopy = getenv("MLBM_STAGED_COPY") ? (atoi(getenv("MLBM_STAGED_COPY")) != 0 ? 1 : 0) : -1;
  const long long stagedRows = (long long)ctx->LX * ctx->NM;
  const bool staged = uniform && (stagedCopy == 1 || (stagedCopy == -1 && (size_t)stagedRows * hostPitch * es >= ((size_t)32 << 20)));
  if (staged) {
    const long long rows = stagedRows;
    const size_t blockBytes = (size_t)rows * hostPitch * es;
    if (ctx->stagingBytes < blockBytes) {
      if (ctx->staging) MLBM_CUDA(cudaFree(ctx->staging));
      ctx->staging = nullptr;
      ctx->stagingBytes = 0;
      MLBM_CUDA(cudaMalloc(&ctx->staging, blockBytes));
      ctx->stagingBytes = blockBytes;
    }
    const unsigned denseGrid = (unsigned)((rows * ctx->NR + 255) / 256), paddedGrid = (unsigned)((rows * (long long)hostPitch + 255) / 256);
    for (int q = 0; q < ctx->Q; ++q) {
      void* device = static_cast<char*>(ctx->populations[ctx->current]) + ((size_t)q * ctx->stride + ctx->interior) * es;
      char* hostQ = static_cast<char*>(host) + (size_t)q * componentStride * es;
      if (upload) {
        MLBM_CUDA(cudaMemcpyAsync(ctx->staging, hostQ, blockBytes, cudaMemcpyHostToDevice, ctx->computeStream));
        if (es == 8) stripPaddingKernel<double><<<denseGrid, 256, 0, ctx->computeStream>>>(static_cast<const double*>(ctx->staging), static_cast<double*>(device), rows, ctx->NR, (long long)hostPitch);
        else stripPaddingKernel<float><<<denseGrid, 256, 0, ctx->computeStream>>>(static_cast<const float*>(ctx->staging), static_cast<float*>(device), rows, ctx->NR, (long long)hostPitch);
      } else {
        if (es == 8) addPaddingKernel<double><<<paddedGrid, 256, 0, ctx->computeStream>>>(static_cast<const double*>(device), static_cast<double*>(ctx->staging), rows, ctx->NR, (long long)hostPitch);
        else addPaddingKernel<float><<<paddedGrid, 256, 0, ctx->computeStream>>>(static_cast<const float*>(device), static_cast<float*>(ctx->staging), rows, ctx->NR, (long long)hostPitch);
        MLBM_CUDA(cudaMemcpyAsync(hostQ, ctx->staging, blockBytes, cudaMemcpyDeviceToHost, ctx->computeStream));
      }
      ctx->launches += 1;
    }
    MLBM_CUDA(cudaGetLastError());
    MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
    if (upload) ctx->halosValid = false;
    return MLBM_OK;
  }
  for (int q = 0; q < ctx->Q; ++q) {
    char* device = static_cast<char*>(ctx->populations[ctx->current]) + ((size_t)q * ctx->stride + ctx->interior) * es;
    char* hostQ = static_cast<char*>(host) + (size_t)q * componentStride * es;
    const int chunks = uniform ? 1 : ctx->LX;
    const size_t rows = uniform ? (size_t)ctx->LX * ctx->NM : (size_t)ctx->NM;
    for (int chunk = 0; chunk < chunks; ++chunk) {
      char* d = device + (size_t)chunk * ctx->plane * es;
      char* h = hostQ + (size_t)chunk * paddedY * paddedZ * es;
      if (upload) MLBM_CUDA(cudaMemcpy2DAsync(d, ctx->NR * es, h, hostPitch * es, ctx->NR * es, rows, cudaMemcpyHostToDevice, ctx->computeStream));
      else MLBM_CUDA(cudaMemcpy2DAsync(h, hostPitch * es, d, ctx->NR * es, ctx->NR * es, rows, cudaMemcpyDeviceToHost, ctx->computeStream));
    }
  }
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  if (upload) ctx->halosValid = false;
  return MLBM_OK;
}

int mlbm_upload_distribution(mlbm_ctx* ctx, const void* host, size_t componentStride, size_t paddedY, size_t paddedZ) {
  return copyDistribution(ctx, const_cast<void*>(host), componentStride, paddedY, paddedZ, true);
}

int mlbm_download_distribution(mlbm_ctx* ctx, void* host, size_t componentStride, size_t paddedY, size_t paddedZ) {
  return copyDistribution(ctx, host, componentStride, paddedY, paddedZ, false);
}

int mlbm_download_halo_distribution(mlbm_ctx* ctx, void* host, size_t capacity) {
  if (!ctx || !host) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  const int H = ctx->H, D = ctx->D;
  const long long NY = D == 3 ? ctx->NM : (D == 2 ? ctx->NR : 1), NZ = D == 3 ? ctx->NR : 1;
  const long long HX = ctx->LX + 2 * H, HY = D >= 2 ? NY + 2 * H : 1, HZ = D == 3 ? NZ + 2 * H : 1;
  const long long volume = HX * HY * HZ;
  if (capacity < (size_t)(volume * ctx->Q)) return fail(MLBM_ERR_INVALID, "halo array of %zu elements, needs dimQ * hSD::volume() = %lld", capacity, volume * ctx->Q);
  const bool multi = ctx->config.nranks > 1;
  if (multi && !ctx->halosValid) {
    // the x halo planes of the buffer about to be read are delivered like at the top of iterate (Algorithm.h:339-341)
    if (int status = exchangeHalos(ctx, ctx->current, ctx->computeStream)) return status;
    ctx->halosValid = true;
  }
  const size_t es = ctx->elementSize;
  std::vector<char> device((size_t)ctx->stride * ctx->Q * es);
  MLBM_CUDA(cudaStreamSynchronize(ctx->commStream));
  MLBM_CUDA(cudaMemcpyAsync(device.data(), ctx->populations[ctx->current], device.size(), cudaMemcpyDeviceToHost, ctx->computeStream));
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  auto wrap = [](long long v, long long n) { v %= n; return v < 0 ? v + n : v; };
  for (int q = 0; q < ctx->Q; ++q)
    for (long long hx = 0; hx < HX; ++hx) {
      // one rank: x is periodic inside the slab; several: the device halo planes hold the neighbours' planes
      const long long xs = multi ? hx : wrap(hx - H, ctx->LX) + H;
      for (long long hy = 0; hy < HY; ++hy) {
        const long long ys = D >= 2 ? wrap(hy - H, NY) : 0;
        for (long long hz = 0; hz < HZ; ++hz) {
          const long long zs = D == 3 ? wrap(hz - H, NZ) : 0;
          const long long m = D == 3 ? ys : 0, r = D == 3 ? zs : ys;
          const size_t from = ((size_t)q * ctx->stride + (size_t)xs * ctx->plane + (size_t)m * ctx->NR + (size_t)r) * es;
          const size_t to = ((size_t)q * volume + (size_t)(HZ * (HY * hx + hy) + hz)) * es;
          memcpy(static_cast<char*>(host) + to, device.data() + from, es);
        }
      }
    }
  return MLBM_OK;
}

// ---- checkpoint container ------------------------------------------------------------------------------
// DistributionWriter::writeDistribution (Writer.h:400-445) / DistributionReader (Reader.h:119-157) move dimQ data sets
// "distribution<iQ>": each the PADDED GLOBAL box gSD::pLength() of doubles (H5T_NATIVE_DOUBLE), rank r owning the hyperslab at
// gSD::pOffset(r) of extent lSD::pLength(), filled from / into the local padded array.  HDF5 is not in this image, so the
// container is a flat file: a 4096-byte text header, then the dimQ data sets back to back in exactly that order and layout
// (tools/checkpoint_to_hdf5.py wraps them into the reference's .h5 where h5py exists).  x is the slowest index, so a rank's
// hyperslab of a data set is ONE contiguous range: every rank writes / reads its part with pwrite / pread, any number of
// ranks at once, and a file written by n ranks restarts on m.
namespace {
constexpr size_t kCheckpointHeaderBytes = 4096;

struct CheckpointLayout {
  long long paddedY, paddedZ, localElements, globalElements;  // per data set
};

CheckpointLayout checkpointLayout(const mlbm_ctx* ctx) {
  CheckpointLayout c;
  const int* L = ctx->config.global_length;
  const long long ny = ctx->D >= 2 ? L[1] : 1, nz = ctx->D == 3 ? L[2] : 1;
  c.paddedY = ctx->D == 2 ? 2 * (ny / 2 + 1) : ny;   // lSD::pLength: the last used dimension is padded (Domain.h:53-57)
  c.paddedZ = ctx->D == 3 ? 2 * (nz / 2 + 1) : nz;
  c.localElements = (long long)ctx->LX * c.paddedY * c.paddedZ;
  c.globalElements = c.localElements * ctx->config.nranks;
  return c;
}
}  // namespace

int mlbm_checkpoint_write(mlbm_ctx* ctx, const char* path, unsigned iteration) {
  if (!ctx || !path) return fail(MLBM_ERR_INVALID, "null argument");
  const CheckpointLayout c = checkpointLayout(ctx);
  const size_t es = ctx->elementSize;
  std::vector<char> local((size_t)c.localElements * ctx->Q * es);
  if (int status = mlbm_download_distribution(ctx, local.data(), (size_t)c.localElements, (size_t)c.paddedY, (size_t)c.paddedZ)) return status;
  const int fd = open(path, O_CREAT | O_WRONLY, 0644);
  if (fd < 0) return fail(MLBM_ERR_INVALID, "cannot open %s for writing: %s", path, strerror(errno));
  int status = MLBM_OK;
  if (ctx->config.rank == 0) {
    char header[kCheckpointHeaderBytes];
    memset(header, ' ', sizeof(header));
    const int* L = ctx->config.global_length;
    const int written = snprintf(header, sizeof(header),
                                 "{\"format\": \"metalbm_b200 checkpoint 1\", \"datasets\": \"distribution<iQ>, iQ = 0 .. dimQ - 1 (Writer.h:400-445)\", "
                                 "\"dtype\": \"float64\", \"dimD\": %d, \"dimQ\": %d, \"global_length\": [%d, %d, %d], "
                                 "\"padded_global_length\": [%d, %lld, %lld], \"header_bytes\": %zu, \"iteration\": %u, \"written_by_ranks\": %d}",
                                 ctx->D, ctx->Q, L[0], ctx->D >= 2 ? L[1] : 1, ctx->D == 3 ? L[2] : 1, L[0], c.paddedY, c.paddedZ,
                                 kCheckpointHeaderBytes, iteration, ctx->config.nranks);
    header[written] = ' ';
    header[sizeof(header) - 1] = '\n';
    if (pwrite(fd, header, sizeof(header), 0) != (ssize_t)sizeof(header)) status = fail(MLBM_ERR_INVALID, "short write to %s", path);
  }
  std::vector<double> widened;
  for (int q = 0; q < ctx->Q && status == MLBM_OK; ++q) {
    const char* source = local.data() + (size_t)q * c.localElements * es;
    if (es == 4) {   // the data sets are doubles whatever the context stores
      widened.resize((size_t)c.localElements);
      for (long long i = 0; i < c.localElements; ++i) widened[(size_t)i] = (double)reinterpret_cast<const float*>(source)[i];
      source = reinterpret_cast<const char*>(widened.data());
    }
    const off_t offset = (off_t)kCheckpointHeaderBytes + (off_t)sizeof(double) * ((off_t)q * c.globalElements + (off_t)ctx->config.rank * c.localElements);
    const size_t bytes = sizeof(double) * (size_t)c.localElements;
    size_t done = 0;
    while (done < bytes) {
      const ssize_t step = pwrite(fd, source + done, bytes - done, offset + (off_t)done);
      if (step <= 0) { status = fail(MLBM_ERR_INVALID, "write to %s failed: %s", path, strerror(errno)); break; }
      done += (size_t)step;
    }
  }
  if (close(fd) != 0 && status == MLBM_OK) status = fail(MLBM_ERR_INVALID, "closing %s: %s", path, strerror(errno));
  return status;
}

int mlbm_checkpoint_read(mlbm_ctx* ctx, const char* path, unsigned* iteration) {
  if (!ctx || !path) return fail(MLBM_ERR_INVALID, "null argument");
  const CheckpointLayout c = checkpointLayout(ctx);
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(MLBM_ERR_INVALID, "cannot open %s: %s", path, strerror(errno));
  char header[kCheckpointHeaderBytes];
  int status = MLBM_OK;
  if (pread(fd, header, sizeof(header), 0) != (ssize_t)sizeof(header)) status = fail(MLBM_ERR_INVALID, "%s is not a checkpoint (short header)", path);
  header[sizeof(header) - 1] = 0;
  int D = 0, Q = 0, L[3] = {0, 0, 0};
  unsigned stored = 0;
  if (status == MLBM_OK) {
    const char* dims = strstr(header, "\"dimD\": ");
    const char* lengths = strstr(header, "\"global_length\": [");
    const char* it = strstr(header, "\"iteration\": ");
    if (!strstr(header, "metalbm_b200 checkpoint 1") || !dims || !lengths || !it || sscanf(dims, "\"dimD\": %d, \"dimQ\": %d", &D, &Q) != 2 ||
        sscanf(lengths, "\"global_length\": [%d, %d, %d]", &L[0], &L[1], &L[2]) != 3 || sscanf(it, "\"iteration\": %u", &stored) != 1)
      status = fail(MLBM_ERR_INVALID, "%s: not a metalbm_b200 checkpoint header", path);
  }
  if (status == MLBM_OK) {
    const int* G = ctx->config.global_length;
    if (D != ctx->D || Q != ctx->Q || L[0] != G[0] || L[1] != (ctx->D >= 2 ? G[1] : 1) || L[2] != (ctx->D == 3 ? G[2] : 1))
      status = fail(MLBM_ERR_INVALID, "%s holds D%dQ%d %dx%dx%d, the context is D%dQ%d %dx%dx%d", path, D, Q, L[0], L[1], L[2], ctx->D, ctx->Q, G[0],
                    ctx->D >= 2 ? G[1] : 1, ctx->D == 3 ? G[2] : 1);
  }
  const size_t es = ctx->elementSize;
  std::vector<double> slab((size_t)c.localElements);
  std::vector<char> local(status == MLBM_OK ? (size_t)c.localElements * ctx->Q * es : 0);
  for (int q = 0; q < ctx->Q && status == MLBM_OK; ++q) {
    const off_t offset = (off_t)kCheckpointHeaderBytes + (off_t)sizeof(double) * ((off_t)q * c.globalElements + (off_t)ctx->config.rank * c.localElements);
    const size_t bytes = sizeof(double) * (size_t)c.localElements;
    size_t done = 0;
    while (done < bytes) {
      const ssize_t step = pread(fd, reinterpret_cast<char*>(slab.data()) + done, bytes - done, offset + (off_t)done);
      if (step <= 0) { status = fail(MLBM_ERR_INVALID, "%s: data set distribution%d is truncated", path, q); break; }
      done += (size_t)step;
    }
    char* target = local.data() + (size_t)q * c.localElements * es;
    if (es == 8) memcpy(target, slab.data(), bytes);
    else for (long long i = 0; i < c.localElements; ++i) reinterpret_cast<float*>(target)[i] = (float)slab[(size_t)i];
  }
  close(fd);
  if (status != MLBM_OK) return status;
  if (iteration) *iteration = stored;
  return mlbm_upload_distribution(ctx, local.data(), (size_t)c.localElements, (size_t)c.paddedY, (size_t)c.paddedZ);
}

// a field [components][LX][paddedY][paddedZ] on the host <-> dense [components][LX][NM][NR] on the device
static int copyField(mlbm_ctx* ctx, void* host, void* device, int components, size_t componentStride, size_t paddedY,
                     size_t paddedZ, bool upload) {
  const size_t es = ctx->elementSize;
  const bool threeD = ctx->D == 3;
  const size_t hostPitch = threeD ? paddedZ : paddedY;
  if (hostPitch < (size_t)ctx->NR || (threeD && paddedY < (size_t)ctx->NM)) return fail(MLBM_ERR_INVALID, "padded lengths smaller than the local lengths");
  const bool uniform = !threeD || paddedY == (size_t)ctx->NM;
  for (int c = 0; c < components; ++c) {
    const int chunks = uniform ? 1 : ctx->LX;
    const size_t rows = uniform ? (size_t)ctx->LX * ctx->NM : (size_t)ctx->NM;
    for (int chunk = 0; chunk < chunks; ++chunk) {
      char* d = static_cast<char*>(device) + ((size_t)c * ctx->fieldStride + (size_t)chunk * ctx->plane) * es;
      char* h = static_cast<char*>(host) + ((size_t)c * componentStride + (size_t)chunk * paddedY * paddedZ) * es;
      if (upload) MLBM_CUDA(cudaMemcpy2DAsync(d, ctx->NR * es, h, hostPitch * es, ctx->NR * es, rows, cudaMemcpyHostToDevice, ctx->computeStream));
      else MLBM_CUDA(cudaMemcpy2DAsync(h, hostPitch * es, d, ctx->NR * es, ctx->NR * es, rows, cudaMemcpyDeviceToHost, ctx->computeStream));
    }
  }
  return MLBM_OK;
}

int mlbm_init_equilibrium(mlbm_ctx* ctx, const void* density, const void* velocity, size_t componentStride,
                          size_t paddedY, size_t paddedZ) {
  if (!ctx || !density || !velocity) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (int status = ensureFields(ctx)) return status;
  if (int status = copyField(ctx, const_cast<void*>(density), ctx->density, 1, componentStride, paddedY, paddedZ, true)) return status;
  if (int status = copyField(ctx, const_cast<void*>(velocity), ctx->velocity, ctx->D, componentStride, paddedY, paddedZ, true)) return status;
  void* target = ctx->populations[ctx->current];
  if (ctx->config.dtype == MLBM_F64) {
    if (ctx->config.equilibrium == MLBM_EXACT)
      launchInitEquilibrium<kExact, double>(ctx->config.lattice, ctx->computeStream, static_cast<double*>(target), static_cast<const double*>(ctx->density), static_cast<const double*>(ctx->velocity), ctx->stride, ctx->interior, ctx->fieldStride, ctx->nodes);
    else
      launchInitEquilibrium<kTruncationMa3, double>(ctx->config.lattice, ctx->computeStream, static_cast<double*>(target), static_cast<const double*>(ctx->density), static_cast<const double*>(ctx->velocity), ctx->stride, ctx->interior, ctx->fieldStride, ctx->nodes);
  } else {
    if (ctx->config.equilibrium == MLBM_EXACT)
      launchInitEquilibrium<kExact, float>(ctx->config.lattice, ctx->computeStream, static_cast<float*>(target), static_cast<const float*>(ctx->density), static_cast<const float*>(ctx->velocity), ctx->stride, ctx->interior, ctx->fieldStride, ctx->nodes);
    else
      launchInitEquilibrium<kTruncationMa3, float>(ctx->config.lattice, ctx->computeStream, static_cast<float*>(target), static_cast<const float*>(ctx->density), static_cast<const float*>(ctx->velocity), ctx->stride, ctx->interior, ctx->fieldStride, ctx->nodes);
  }
  MLBM_CUDA(cudaGetLastError());
  ctx->launches += 1;
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  ctx->halosValid = false;
  if (ctx->shell && shellForceIsTimeDependent(ctx->shell)) ctx->forceStale = true;  // fieldList holds the initial fields (Initialize.h)
  return MLBM_OK;
}

int mlbm_init_synthetic(mlbm_ctx* ctx, double densityAmplitude, double velocityAmplitude) {
  if (!ctx) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  void* target = ctx->populations[ctx->current];
  const int xOffset = ctx->config.rank * ctx->LX;  // gSD::sOffset (Domain.h:155-162)
  const int* global = ctx->config.global_length;
  const bool exact = ctx->config.equilibrium == MLBM_EXACT;
  if (ctx->config.dtype == MLBM_F64) {
    if (exact) launchInitSynthetic<kExact, double>(ctx->config.lattice, ctx->computeStream, static_cast<double*>(target), ctx->stride, ctx->plane, ctx->interior, ctx->nodes, ctx->NR, xOffset, global, densityAmplitude, velocityAmplitude);
    else launchInitSynthetic<kTruncationMa3, double>(ctx->config.lattice, ctx->computeStream, static_cast<double*>(target), ctx->stride, ctx->plane, ctx->interior, ctx->nodes, ctx->NR, xOffset, global, densityAmplitude, velocityAmplitude);
  } else {
    if (exact) launchInitSynthetic<kExact, float>(ctx->config.lattice, ctx->computeStream, static_cast<float*>(target), ctx->stride, ctx->plane, ctx->interior, ctx->nodes, ctx->NR, xOffset, global, densityAmplitude, velocityAmplitude);
    else launchInitSynthetic<kTruncationMa3, float>(ctx->config.lattice, ctx->computeStream, static_cast<float*>(target), ctx->stride, ctx->plane, ctx->interior, ctx->nodes, ctx->NR, xOffset, global, densityAmplitude, velocityAmplitude);
  }
  MLBM_CUDA(cudaGetLastError());
  ctx->launches += 1;
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  ctx->halosValid = false;
  return MLBM_OK;
}

int mlbm_perturb_distribution(mlbm_ctx* ctx, double eps, uint64_t seed) {
  if (!ctx) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  const unsigned grid = (unsigned)((ctx->nodes + 255) / 256);
  const long long offset = (long long)ctx->config.rank * ctx->nodes;
  void* target = ctx->populations[ctx->current];
  if (ctx->config.dtype == MLBM_F64)
    perturbKernel<double><<<grid, 256, 0, ctx->computeStream>>>(static_cast<double*>(target), ctx->stride, ctx->interior, ctx->nodes, offset, ctx->Q, eps, seed);
  else
    perturbKernel<float><<<grid, 256, 0, ctx->computeStream>>>(static_cast<float*>(target), ctx->stride, ctx->interior, ctx->nodes, offset, ctx->Q, eps, seed);
  MLBM_CUDA(cudaGetLastError());
  ctx->launches += 1;
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  ctx->halosValid = false;
  return MLBM_OK;
}

int mlbm_set_alpha(mlbm_ctx* ctx, const void* host, size_t paddedY, size_t paddedZ) {
  if (!ctx || !host) return fail(MLBM_ERR_INVALID, "null argument");
  if (!ctx->alpha) return fail(MLBM_ERR_STATE, "the BGK alpha field is the constant 2 (Collision.h:121)");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (int status = copyField(ctx, const_cast<void*>(host), ctx->alpha, 1, 0, paddedY, paddedZ, true)) return status;
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  return MLBM_OK;
}

int mlbm_set_force_field(mlbm_ctx* ctx, const void* host, size_t componentStride, size_t paddedY, size_t paddedZ) {
  if (!ctx || !host) return fail(MLBM_ERR_INVALID, "null argument");
  if (ctx->config.force != MLBM_FORCE_FIELD) return fail(MLBM_ERR_STATE, "the context was not created with MLBM_FORCE_FIELD");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (int status = copyField(ctx, const_cast<void*>(host), ctx->force, ctx->D, componentStride, paddedY, paddedZ, true)) return status;
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  return MLBM_OK;
}

int mlbm_step(mlbm_ctx* ctx, unsigned iteration, int isStored) {
  (void)iteration;  // Force::update is a no-op for the time-independent forces of this path (Force.h:51-54)
  if (!ctx) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (int status = enqueueStep(ctx, isStored ? (isStored & 3 ? isStored & 3 : 1) : 0, true, ctx->profiling)) return status;
  // the reference's iterate returns after cudaDeviceSynchronize (Algorithm.h:355)
  if (int status = synchronizeOrDiagnose(ctx, ctx->computeStream)) return status;
  float communication = 0.f, computation = 0.f;
  MLBM_CUDA(cudaEventElapsedTime(&communication, ctx->timeStart, ctx->timeMid));
  MLBM_CUDA(cudaEventElapsedTime(&computation, ctx->timeMid, ctx->timeStop));
  ctx->lastCommunication = communication * 1e-3;
  ctx->lastComputation = computation * 1e-3;
  return MLBM_OK;
}

int mlbm_run_async_stored(mlbm_ctx* ctx, unsigned firstIteration, unsigned count, unsigned storeEvery, int storedMode) {
  if (!ctx) return fail(MLBM_ERR_INVALID, "null argument");
  if (storedMode != 1 && storedMode != 2) return fail(MLBM_ERR_INVALID, "stored mode %d (1 = fields + all observables, 2 = energy / mass / Mach only)", storedMode);
  MLBM_CUDA(cudaSetDevice(ctx->device));
  const bool profile = ctx->profiling;
  for (unsigned i = 0; i < count; ++i) {
    const unsigned iteration = firstIteration + i;
    const int isStored = (storeEvery && iteration % storeEvery == 0) ? storedMode : 0;
    if (int status = enqueueStep(ctx, isStored, false, profile)) return status;
  }
  return MLBM_OK;
}

int mlbm_run_async(mlbm_ctx* ctx, unsigned firstIteration, unsigned count, unsigned storeEvery) {
  return mlbm_run_async_stored(ctx, firstIteration, count, storeEvery, 1);
}

int mlbm_sync(mlbm_ctx* ctx) {
  if (!ctx) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (int status = synchronizeOrDiagnose(ctx, ctx->computeStream)) return status;
  if (int status = synchronizeOrDiagnose(ctx, ctx->commStream)) return status;
  if (int status = synchronizeOrDiagnose(ctx, ctx->analysisStream)) return status;
  return MLBM_OK;
}

int mlbm_download_fields(mlbm_ctx* ctx, void* density, void* velocity, void* alpha, void* force, size_t componentStride,
                         size_t paddedY, size_t paddedZ) {
  if (!ctx) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if ((density || velocity || force) && !ctx->fieldsStored) return fail(MLBM_ERR_STATE, "no stored step yet (Algorithm::isStored was never set)");
  if (density) { if (int status = copyField(ctx, density, ctx->density, 1, componentStride, paddedY, paddedZ, false)) return status; }
  if (velocity) { if (int status = copyField(ctx, velocity, ctx->velocity, ctx->D, componentStride, paddedY, paddedZ, false)) return status; }
  if (force) { if (int status = copyField(ctx, force, ctx->force, ctx->D, componentStride, paddedY, paddedZ, false)) return status; }
  if (alpha) {
    if (ctx->alpha) {
      if (int status = copyField(ctx, alpha, ctx->alpha, 1, componentStride, paddedY, paddedZ, false)) return status;
    } else {
      // BGK: alpha == 2 everywhere (Collision.h:121, Algorithm.h:105-106)
      const size_t hostPitch = ctx->D == 3 ? paddedZ : paddedY;
      for (int x = 0; x < ctx->LX; ++x)
        for (int m = 0; m < ctx->NM; ++m)
          for (int r = 0; r < ctx->NR; ++r) {
            const size_t index = ctx->D == 3 ? ((size_t)x * paddedY + m) * paddedZ + r : (size_t)x * hostPitch + r;
            if (ctx->config.dtype == MLBM_F64) static_cast<double*>(alpha)[index] = 2.0;
            else static_cast<float*>(alpha)[index] = 2.0f;
          }
    }
  }
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  return MLBM_OK;
}

int mlbm_alloc_pinned(size_t bytes, void** out) {
  if (!out) return fail(MLBM_ERR_INVALID, "null argument");
  *out = nullptr;
  cudaError_t error = cudaMallocHost(out, bytes ? bytes : 1);
  if (error != cudaSuccess) return fail(MLBM_ERR_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(error));
  return MLBM_OK;
}

int mlbm_free_pinned(void* pointer) {
  if (!pointer) return MLBM_OK;
  cudaError_t error = cudaFreeHost(pointer);
  if (error != cudaSuccess) return fail(MLBM_ERR_CUDA, "cudaFreeHost: %s", cudaGetErrorString(error));
  return MLBM_OK;
}

int mlbm_timers(mlbm_ctx* ctx, double* communicationSeconds, double* computationSeconds) {
  if (!ctx) return fail(MLBM_ERR_INVALID, "null argument");
  if (communicationSeconds) *communicationSeconds = ctx->lastCommunication;
  if (computationSeconds) *computationSeconds = ctx->lastComputation;
  return MLBM_OK;
}

int mlbm_device_distribution(mlbm_ctx* ctx, mlbm_device_layout* out) {
  if (!ctx || !out) return fail(MLBM_ERR_INVALID, "null argument");
  out->populations = ctx->populations[ctx->current];
  out->component_stride = (size_t)ctx->stride;
  out->plane = (size_t)ctx->plane;
  out->row = (size_t)ctx->NR;
  out->halo_x = ctx->H;
  out->local_length[0] = ctx->LX;
  out->local_length[1] = ctx->D == 3 ? ctx->NM : ctx->NR;
  out->local_length[2] = ctx->D == 3 ? ctx->NR : 1;
  return MLBM_OK;
}

int mlbm_launch_count(mlbm_ctx* ctx, uint64_t* launches) {
  if (!ctx || !launches) return fail(MLBM_ERR_INVALID, "null argument");
  *launches = ctx->launches;
  return MLBM_OK;
}

int mlbm_stream(mlbm_ctx* ctx, void** cudaStream) {
  if (!ctx || !cudaStream) return fail(MLBM_ERR_INVALID, "null argument");
  *cudaStream = ctx->computeStream;
  return MLBM_OK;
}

int mlbm_kernel_time(mlbm_ctx* ctx, double* averageMs, uint64_t* launches) {
  if (!ctx) return fail(MLBM_ERR_INVALID, "null argument");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->profiling) {
    // first call switches per-launch event timing on
    ctx->profiling = true;
    if (averageMs) *averageMs = 0.0;
    if (launches) *launches = 0;
    return MLBM_OK;
  }
  MLBM_CUDA(cudaStreamSynchronize(ctx->computeStream));
  if (int status = collectProfile(ctx)) return status;
  if (averageMs) *averageMs = ctx->profileLaunches ? ctx->profileMs / (double)ctx->profileLaunches : 0.0;
  if (launches) *launches = ctx->profileLaunches;
  ctx->profileMs = 0.0;
  ctx->profileLaunches = 0;
  return MLBM_OK;
}

int mlbm_mark(mlbm_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot >= 8) return fail(MLBM_ERR_INVALID, "bad mark slot");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->marks[slot]) MLBM_CUDA(cudaEventCreate(&ctx->marks[slot]));
  MLBM_CUDA(cudaEventRecord(ctx->marks[slot], ctx->computeStream));
  return MLBM_OK;
}

int mlbm_elapsed(mlbm_ctx* ctx, int from, int to, double* milliseconds) {
  if (!ctx || !milliseconds || from < 0 || from >= 8 || to < 0 || to >= 8 || !ctx->marks[from] || !ctx->marks[to])
    return fail(MLBM_ERR_INVALID, "bad mark slot");
  MLBM_CUDA(cudaSetDevice(ctx->device));
  MLBM_CUDA(cudaEventSynchronize(ctx->marks[to]));
  float ms = 0.f;
  MLBM_CUDA(cudaEventElapsedTime(&ms, ctx->marks[from], ctx->marks[to]));
  *milliseconds = ms;
  return MLBM_OK;
}

}  // extern "C"
