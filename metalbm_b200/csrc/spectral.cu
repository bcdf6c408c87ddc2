// spectral.cu -- the reference's total enstrophy (spectral vorticity) on x-slabs, as a device-side analysis.
//
// What the reference does on every stored step (Routine.h:129-132, Transformer.h:118-295, Analysis.h:68-98):
//   1. in-place r2c FFT of the stored hydrodynamic velocity (unnormalised, FFTW conventions);
//   2. vorticity spectrum  w^ = i k x u^  with INTEGER wave numbers  k_d = i <= N_d/2 ? i : i - N_d  (the Nyquist
//      index keeps +N_d/2); in 2-D only the real part is formed and the imaginary part is forced to zero
//      (Transformer.h:151-165);
//   3. c2r FFT back, divided by the global volume V (BackwardFFT::execute, Transformer.h:101-108), then divided by V
//      AGAIN (Curl::normalize, :179-185, :284-294);
//   4. Z = sum_x sum_d 0.5 w_d(x)^2, MPI-summed, divided by V (Analysis.h:85-93, :30).
//
// Only the scalar is needed on the step path, so step 3 is replaced by Parseval's identity: with W(x) the
// unnormalised c2r output, sum_x W(x)^2 = V sum_k |w^_k|^2 over the FULL spectrum that the c2r transform implies.
// For a half spectrum stored along the last axis (index j = 0 .. N/2) that is
//     sum_k = sum_{0 < j < N/2} 2 |w^|^2  +  sum_{j in {0, N/2}} |sym w^|^2 ,
// where the c2r transform drops the non-Hermitian part of the two self-conjugate planes: sym w^(k) =
// (w^(k) + conj w^(-k)) / 2.  Because u^ is the transform of a real field, sym w^ is again i k' x u^ with every
// Nyquist component of k' replaced by 0 (the +N/2 of the plane and of its mirror image cancel).  The identity,
// including the Nyquist bookkeeping and the 2-D real-part quirk, is pinned against the literal algorithm
// (oracle/oracle.py: spectral_enstrophy, itself pinned to the reference's Curl on golden vectors) in
// tests/test_parity_gpu.py::test_spectral_enstrophy_*.
//
// Distributed transform: batched FFT over the local (m, r) planes, one all-to-all (NCCL send/recv, grouped) that
// re-slabs the half spectrum from x-slabs to contiguous column chunks, batched FFT along x.  cuFFT does the 1-D
// and 2-D transforms (library code, off the hot path: this runs on stored steps only); it is resolved with
// dlopen so that the library still loads where cuFFT is absent.
#include "spectral.h"

#include <cufft.h>
#include <dlfcn.h>

#include <algorithm>
#include <string>

namespace mlbm {

namespace {

struct CufftApi {
  cufftResult (*Create)(cufftHandle*);
  cufftResult (*MakePlanMany64)(cufftHandle, int, long long*, long long*, long long, long long, long long*, long long,
                                long long, cufftType, long long, size_t*);
  cufftResult (*SetStream)(cufftHandle, cudaStream_t);
  cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal*, cufftDoubleComplex*);
  cufftResult (*ExecZ2Z)(cufftHandle, cufftDoubleComplex*, cufftDoubleComplex*, int);
  cufftResult (*Destroy)(cufftHandle);
};

const CufftApi* loadCufft(std::string* error) {
  static CufftApi api;
  static bool tried = false, ok = false;
  static std::string message;
  if (!tried) {
    tried = true;
    void* handle = nullptr;
    for (const char* name : {"libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so.12"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      message = std::string("cannot load libcufft (needed for the spectral enstrophy): ") + dlerror();
    } else {
      ok = true;
      auto resolve = [&](const char* name) -> void* {
        void* symbol = dlsym(handle, name);
        if (!symbol) { ok = false; message = std::string("libcufft lacks ") + name; }
        return symbol;
      };
      api.Create = reinterpret_cast<decltype(api.Create)>(resolve("cufftCreate"));
      api.MakePlanMany64 = reinterpret_cast<decltype(api.MakePlanMany64)>(resolve("cufftMakePlanMany64"));
      api.SetStream = reinterpret_cast<decltype(api.SetStream)>(resolve("cufftSetStream"));
      api.ExecD2Z = reinterpret_cast<decltype(api.ExecD2Z)>(resolve("cufftExecD2Z"));
      api.ExecZ2Z = reinterpret_cast<decltype(api.ExecZ2Z)>(resolve("cufftExecZ2Z"));
      api.Destroy = reinterpret_cast<decltype(api.Destroy)>(resolve("cufftDestroy"));
    }
  }
  if (!ok) { if (error) *error = message; return nullptr; }
  return &api;
}

constexpr int kBlock = 256;

__global__ void widenKernel(const float* __restrict__ in, double* __restrict__ out, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (double)in[i];
}

// x-slab half spectrum [LX][columns] -> per-peer contiguous blocks [peer][LX][columns of that peer]
__global__ void packColumnsKernel(const double2* __restrict__ local, double2* __restrict__ packed, int LX, long long columns,
                                  long long chunk) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)LX * columns) return;
  const long long x = i / columns, column = i % columns;
  const long long peer = column / chunk, c = column % chunk;
  const long long first = peer * chunk;
  const long long width = columns - first < chunk ? columns - first : chunk;
  packed[(long long)LX * first + x * width + c] = local[i];
}

// sum over this rank's part of the spectrum of  weight * |k' x u^|^2  (see the file header); one partial per block
template <int D>
__global__ void vorticityNormKernel(const double2* __restrict__ ux, const double2* __restrict__ uy, const double2* __restrict__ uz,
                                    int NX, int NM, int NR, long long myColumns, long long firstColumn,
                                    double* __restrict__ blockSums) {
  double total = 0.0;
  // fixed grid, grid-stride: the partition (and with it the summation order) does not depend on timing
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)NX * myColumns; i += (long long)gridDim.x * blockDim.x) {
    double value = 0.0;
    const int NRc = NR / 2 + 1;
    const int ix = (int)(i / myColumns);
    const long long column = firstColumn + i % myColumns;
    const int im = (int)(column / NRc), ir = (int)(column % NRc);
    const bool selfConjugate = ir == 0 || (NR % 2 == 0 && ir == NR / 2);
    auto waveNumber = [&](int index, int n) {
      if (selfConjugate && n % 2 == 0 && index == n / 2) return 0.0;
      return (double)(index <= n / 2 ? index : index - n);
    };
    const double weight = selfConjugate ? 1.0 : 2.0;
    const double kx = waveNumber(ix, NX);
    if (D == 2) {
      // Transformer.h:151-165: Re w^ = -kx Im(uy^) + ky Im(ux^), Im w^ = 0
      const double ky = waveNumber(ir, NR);
      const double w = -kx * uy[i].y + ky * ux[i].y;
      value = weight * w * w;
    } else {
      const double ky = waveNumber(im, NM), kz = waveNumber(ir, NR);
      const double2 a = ux[i], b = uy[i], c = uz[i];
      // |i k x u^|^2 = |k x u^|^2, component by component (Transformer.h:214-270)
      const double wxr = ky * c.x - kz * b.x, wxi = ky * c.y - kz * b.y;
      const double wyr = kz * a.x - kx * c.x, wyi = kz * a.y - kx * c.y;
      const double wzr = kx * b.x - ky * a.x, wzi = kx * b.y - ky * a.y;
      value = weight * (wxr * wxr + wxi * wxi + wyr * wyr + wyi * wyi + wzr * wzr + wzi * wzi);
    }
    total += value;
  }
  double value = total;
  for (int offset = 16; offset > 0; offset >>= 1) value += __shfl_xor_sync(0xffffffffu, value, offset);
  __shared__ double scratch[kBlock / 32];
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = value;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
    for (int w = 0; w < kBlock / 32; ++w) sum += scratch[w];
    blockSums[blockIdx.x] = sum;
  }
}

// deterministic final sum; scale = 1 / (2 V^3) so that the caller's division by V gives Z = sum / (2 V^4)
__global__ void finishKernel(const double* __restrict__ blockSums, long long count, double invVolume, double* __restrict__ out) {
  __shared__ double scratch[kBlock];
  double sum = 0.0;
  for (long long i = threadIdx.x; i < count; i += blockDim.x) sum += blockSums[i];
  scratch[threadIdx.x] = sum;
  __syncthreads();
  for (int width = kBlock / 2; width > 0; width >>= 1) {
    if ((int)threadIdx.x < width) scratch[threadIdx.x] += scratch[threadIdx.x + width];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = scratch[0] * invVolume * invVolume * invVolume * 0.5;
}

// PowerSpectra::operator() (Analysis.h:148-168) over this rank's part of the half spectrum: sum_d |a^_d|^2, halved where the
// wave number of the last (halved) dimension is 0, into bin floor(|k|) when that is below `bins`.  Bit-reproducible: a block
// is ONE warp with private bins in shared memory; per round the 32 lanes publish (bin, value) and lane l adds, in lane order,
// the values of the bins it owns (bin % 32 == l) -- no atomics, a fixed summation order; then one partial row per block
// (summed in block order by binSumKernel).
constexpr int kBinBlock = 32;
template <int D>
__global__ void __launch_bounds__(kBinBlock) powerBinKernel(const double2* __restrict__ ax, const double2* __restrict__ ay, const double2* __restrict__ az, int NX, int NM,
                               int NR, long long myColumns, long long firstColumn, int bins, double* __restrict__ blockBins) {
  extern __shared__ double shared[];
  __shared__ int keys[kBinBlock];
  __shared__ double values[kBinBlock];
  for (int b = threadIdx.x; b < bins; b += blockDim.x) shared[b] = 0.0;
  __syncthreads();
  const int NRc = NR / 2 + 1;
  const long long total = (long long)NX * myColumns, stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += stride) {   // block-uniform trip count
    const long long i = base + threadIdx.x;
    int key = -1;
    double value = 0.0;
    if (i < total) {
      const int ix = (int)(i / myColumns);
      const long long column = firstColumn + i % myColumns;
      const int im = (int)(column / NRc), ir = (int)(column % NRc);
      const long long kx = ix <= NX / 2 ? ix : ix - NX, ky = D == 3 ? (im <= NM / 2 ? im : im - NM) : 0, kr = ir;  // AnalysisList.h:141-149
      const unsigned kNorm = (unsigned)sqrt((double)(kx * kx + ky * ky + kr * kr));
      if (kNorm < (unsigned)bins) {
        const double2 a = ax[i], b = ay[i];
        double energy = a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
        if (D == 3) { const double2 c = az[i]; energy += c.x * c.x + c.y * c.y; }
        key = (int)kNorm;
        value = (ir == 0 ? 0.5 : 1.0) * energy;
      }
    }
    keys[threadIdx.x] = key;
    values[threadIdx.x] = value;
    __syncthreads();
    for (int source = 0; source < kBinBlock; ++source) {
      const int k = keys[source];
      if (k >= 0 && (k % kBinBlock) == (int)threadIdx.x) shared[k] += values[source];
    }
    __syncthreads();
  }
  for (int b = threadIdx.x; b < bins; b += blockDim.x) blockBins[(long long)blockIdx.x * bins + b] = shared[b];
}

__global__ void binSumKernel(const double* __restrict__ blockBins, int blocks, int bins, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bins) return;
  double total = 0.0;
  for (int block = 0; block < blocks; ++block) total += blockBins[(long long)block * bins + b];
  out[b] = total;
}

}  // namespace

class SpectralEnstrophy {
 public:
  SpectralGeometry g;
  const CufftApi* fft = nullptr;
  const NcclApi* nccl = nullptr;
  ncclComm_t comm = nullptr;
  int NX = 0, NRc = 0;
  long long columns = 0, chunk = 0, myColumns = 0, firstColumn = 0;
  cufftHandle planPlanes = 0, planX = 0;
  bool havePlanes = false, haveX = false;
  double* realStage = nullptr;     // FP32 storage only: one velocity component widened to double
  double2* local = nullptr;        // nranks > 1: [LX][columns] half spectrum of one component
  double2* packed = nullptr;       // nranks > 1: the same, grouped by destination rank
  double2* spectrum[3] = {nullptr, nullptr, nullptr};  // [NX][myColumns] per component
  double* blockSums = nullptr;
  long long blocks = 0;
  double* blockBins = nullptr;     // power spectra: [blocks][bins] partial bins, made on first use
  int blockBinsCapacity = 0;

  ~SpectralEnstrophy() {
    if (havePlanes) fft->Destroy(planPlanes);
    if (haveX) fft->Destroy(planX);
    for (void* pointer : {(void*)realStage, (void*)local, (void*)packed, (void*)spectrum[0], (void*)spectrum[1],
                          (void*)spectrum[2], (void*)blockSums, (void*)blockBins})
      if (pointer) cudaFree(pointer);
  }
};

static bool check(cudaError_t status, const char* what, std::string* error) {
  if (status == cudaSuccess) return true;
  if (error) *error = std::string(what) + ": " + cudaGetErrorString(status);
  return false;
}

static bool checkFft(cufftResult status, const char* what, std::string* error) {
  if (status == CUFFT_SUCCESS) return true;
  if (error) *error = std::string(what) + ": cuFFT error " + std::to_string((int)status);
  return false;
}

SpectralEnstrophy* spectralCreate(const SpectralGeometry& geometry, const NcclApi* nccl, ncclComm_t comm, std::string* error) {
  const CufftApi* fft = loadCufft(error);
  if (!fft) return nullptr;
  if (geometry.nranks > 1 && (!nccl || !comm)) {
    if (error) *error = "spectral enstrophy on several ranks needs the NCCL communicator (mlbm_comm_init)";
    return nullptr;
  }
  SpectralEnstrophy* s = new SpectralEnstrophy();
  s->g = geometry;
  s->fft = fft;
  s->nccl = nccl;
  s->comm = comm;
  s->NX = geometry.LX * geometry.nranks;
  s->NRc = geometry.NR / 2 + 1;
  s->columns = (long long)geometry.NM * s->NRc;
  s->chunk = (s->columns + geometry.nranks - 1) / geometry.nranks;
  s->firstColumn = std::min(s->columns, s->chunk * geometry.rank);
  s->myColumns = std::min(s->chunk, s->columns - s->firstColumn);
  const long long localNodes = (long long)geometry.LX * geometry.NM * geometry.NR;
  const long long localSpectrum = (long long)geometry.LX * s->columns;
  const long long mySpectrum = (long long)s->NX * s->myColumns;
  s->blocks = std::min<long long>((mySpectrum + kBlock - 1) / kBlock, 148 * 8);

  auto failed = [&]() { delete s; return (SpectralEnstrophy*)nullptr; };
  if (geometry.elementSize == 4 && !check(cudaMalloc(&s->realStage, sizeof(double) * localNodes), "cudaMalloc", error)) return failed();
  if (geometry.nranks > 1) {
    if (!check(cudaMalloc(&s->local, sizeof(double2) * localSpectrum), "cudaMalloc", error)) return failed();
    if (!check(cudaMalloc(&s->packed, sizeof(double2) * localSpectrum), "cudaMalloc", error)) return failed();
  }
  for (int d = 0; d < geometry.D; ++d)
    if (!check(cudaMalloc(&s->spectrum[d], sizeof(double2) * std::max(1LL, mySpectrum)), "cudaMalloc", error)) return failed();
  if (!check(cudaMalloc(&s->blockSums, sizeof(double) * std::max(1LL, s->blocks)), "cudaMalloc", error)) return failed();

  // real -> half-complex transform of every local x plane: 2-D over (m, r) in 3-D, 1-D over r in 2-D
  size_t workSize = 0;
  if (!checkFft(fft->Create(&s->planPlanes), "cufftCreate", error)) return failed();
  s->havePlanes = true;
  if (geometry.D == 3) {
    long long n[2] = {geometry.NM, geometry.NR};
    if (!checkFft(fft->MakePlanMany64(s->planPlanes, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, geometry.LX, &workSize),
                  "cufftMakePlanMany64(D2Z planes)", error)) return failed();
  } else {
    long long n[1] = {geometry.NR};
    if (!checkFft(fft->MakePlanMany64(s->planPlanes, 1, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, geometry.LX, &workSize),
                  "cufftMakePlanMany64(D2Z rows)", error)) return failed();
  }
  // complex transform along x of every column this rank owns: element (x, c) at x * myColumns + c
  if (s->myColumns > 0) {
    if (!checkFft(fft->Create(&s->planX), "cufftCreate", error)) return failed();
    s->haveX = true;
    long long n[1] = {s->NX};
    long long embed[1] = {s->NX};
    if (!checkFft(fft->MakePlanMany64(s->planX, 1, n, embed, s->myColumns, 1, embed, s->myColumns, 1, CUFFT_Z2Z, s->myColumns, &workSize),
                  "cufftMakePlanMany64(Z2Z x)", error)) return failed();
  }
  return s;
}

void spectralDestroy(SpectralEnstrophy* plan) { delete plan; }

// forward transform of the D components of a dense field into s->spectrum[d] ([NX][myColumns], this rank's columns)
static int transformComponents(SpectralEnstrophy* s, const void* velocity, long long fieldStride, cudaStream_t stream,
                               unsigned long long& count, std::string* error) {
  const SpectralGeometry& g = s->g;
  const long long localNodes = (long long)g.LX * g.NM * g.NR;
  const long long localSpectrum = (long long)g.LX * s->columns;
  if (!checkFft(s->fft->SetStream(s->planPlanes, stream), "cufftSetStream", error)) return -1;
  if (s->haveX && !checkFft(s->fft->SetStream(s->planX, stream), "cufftSetStream", error)) return -1;

  for (int d = 0; d < g.D; ++d) {
    double* real;
    if (g.elementSize == 4) {
      widenKernel<<<(unsigned)((localNodes + kBlock - 1) / kBlock), kBlock, 0, stream>>>(
          static_cast<const float*>(velocity) + d * fieldStride, s->realStage, localNodes);
      real = s->realStage;
      ++count;
    } else {
      real = const_cast<double*>(static_cast<const double*>(velocity)) + d * fieldStride;
    }
    double2* target = g.nranks > 1 ? s->local : s->spectrum[d];
    if (!checkFft(s->fft->ExecD2Z(s->planPlanes, real, reinterpret_cast<cufftDoubleComplex*>(target)), "cufftExecD2Z", error)) return -1;
    ++count;
    if (g.nranks > 1) {
      packColumnsKernel<<<(unsigned)((localSpectrum + kBlock - 1) / kBlock), kBlock, 0, stream>>>(s->local, s->packed, g.LX, s->columns, s->chunk);
      ++count;
      // all-to-all: block `peer` of the packed spectrum goes to rank `peer`, whose x planes land at rows peer * LX
      ncclResult_t result = s->nccl->GroupStart();
      for (int peer = 0; peer < g.nranks && result == ncclSuccess; ++peer) {
        const long long first = std::min(s->columns, s->chunk * peer);
        const long long width = std::min(s->chunk, s->columns - first);
        if (width > 0) result = s->nccl->Send(s->packed + (long long)g.LX * first, (size_t)(2 * g.LX * width), ncclDouble, peer, s->comm, stream);
        if (result == ncclSuccess && s->myColumns > 0)
          result = s->nccl->Recv(s->spectrum[d] + (long long)peer * g.LX * s->myColumns, (size_t)(2 * g.LX * s->myColumns), ncclDouble, peer, s->comm, stream);
      }
      const ncclResult_t end = s->nccl->GroupEnd();
      if (result == ncclSuccess) result = end;
      if (result != ncclSuccess) {
        if (error) *error = std::string("spectral all-to-all: ") + s->nccl->GetErrorString(result);
        return -1;
      }
      ++count;
    }
    if (s->haveX) {
      cufftDoubleComplex* data = reinterpret_cast<cufftDoubleComplex*>(s->spectrum[d]);
      if (!checkFft(s->fft->ExecZ2Z(s->planX, data, data, CUFFT_FORWARD), "cufftExecZ2Z", error)) return -1;
      ++count;
    }
  }
  return 0;
}

int spectralEnqueue(SpectralEnstrophy* s, const void* velocity, long long fieldStride, double* out, cudaStream_t stream,
                    unsigned long long* launches, std::string* error) {
  const SpectralGeometry& g = s->g;
  unsigned long long count = 0;
  if (transformComponents(s, velocity, fieldStride, stream, count, error)) return -1;
  if (s->blocks > 0 && s->myColumns > 0) {
    if (g.D == 3)
      vorticityNormKernel<3><<<(unsigned)s->blocks, kBlock, 0, stream>>>(s->spectrum[0], s->spectrum[1], s->spectrum[2], s->NX, g.NM, g.NR,
                                                                      s->myColumns, s->firstColumn, s->blockSums);
    else
      vorticityNormKernel<2><<<(unsigned)s->blocks, kBlock, 0, stream>>>(s->spectrum[0], s->spectrum[1], nullptr, s->NX, g.NM, g.NR,
                                                                      s->myColumns, s->firstColumn, s->blockSums);
    ++count;
  }
  double volume = (double)s->NX * g.NM * g.NR;
  finishKernel<<<1, kBlock, 0, stream>>>(s->blockSums, s->myColumns > 0 ? s->blocks : 0, 1.0 / volume, out);
  ++count;
  if (!check(cudaGetLastError(), "spectral enstrophy kernels", error)) return -1;
  if (launches) *launches += count;
  return 0;
}

int spectralPowerSpectrum(SpectralEnstrophy* s, const void* field, long long fieldStride, int bins, double* out, cudaStream_t stream,
                          unsigned long long* launches, std::string* error) {
  const SpectralGeometry& g = s->g;
  unsigned long long count = 0;
  if (bins <= 0) return 0;
  if ((size_t)bins * sizeof(double) > 48 * 1024) {
    if (error) *error = "power spectra: more than 6144 wave-number bins";
    return -1;
  }
  const int blocks = (int)std::max<long long>(1, s->blocks);
  if (s->blockBinsCapacity < blocks * bins) {
    if (s->blockBins) cudaFree(s->blockBins);
    s->blockBins = nullptr;
    s->blockBinsCapacity = 0;
    if (!check(cudaMalloc(&s->blockBins, sizeof(double) * (size_t)blocks * bins), "cudaMalloc", error)) return -1;
    s->blockBinsCapacity = blocks * bins;
  }
  if (transformComponents(s, field, fieldStride, stream, count, error)) return -1;
  if (s->myColumns > 0) {
    if (g.D == 3)
      powerBinKernel<3><<<(unsigned)blocks, kBinBlock, bins * sizeof(double), stream>>>(s->spectrum[0], s->spectrum[1], s->spectrum[2], s->NX, g.NM, g.NR,
                                                                                    s->myColumns, s->firstColumn, bins, s->blockBins);
    else
      powerBinKernel<2><<<(unsigned)blocks, kBinBlock, bins * sizeof(double), stream>>>(s->spectrum[0], s->spectrum[1], nullptr, s->NX, g.NM, g.NR,
                                                                                    s->myColumns, s->firstColumn, bins, s->blockBins);
    binSumKernel<<<(unsigned)((bins + 127) / 128), 128, 0, stream>>>(s->blockBins, blocks, bins, out);
    count += 2;
  } else if (!check(cudaMemsetAsync(out, 0, sizeof(double) * bins, stream), "cudaMemsetAsync", error)) {
    return -1;
  }
  if (!check(cudaGetLastError(), "power spectra kernels", error)) return -1;
  if (launches) *launches += count;
  return 0;
}

}  // namespace mlbm
