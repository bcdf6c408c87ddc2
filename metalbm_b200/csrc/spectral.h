// spectral.h -- total enstrophy of the stored velocity field with the reference's SPECTRAL vorticity
// (Curl, Transformer.h:118-295; Routine.h:129-132; TotalEnstrophy, Analysis.h:68-98), on x-slabs.
#pragma once

#include <cuda_runtime.h>

#include <string>

#include "nccl_loader.h"

namespace mlbm {

struct SpectralGeometry {
  int D;            // 2 or 3
  int LX, NM, NR;   // local slab extents on the kernel axes (x, m, r); NM == 1 in 2-D
  int rank, nranks; // x-slab index and count: global NX = LX * nranks
  int elementSize;  // 8 (velocity stored as double) or 4 (float)
};

class SpectralEnstrophy;

// Both return nullptr / non-zero and fill *error on failure (cuFFT missing, out of memory, ...).
SpectralEnstrophy* spectralCreate(const SpectralGeometry& geometry, const NcclApi* nccl, ncclComm_t comm, std::string* error);
void spectralDestroy(SpectralEnstrophy* plan);

// Enqueues on `stream`: *out = this rank's share of  sum_x sum_d 0.5 * vorticity_d(x)^2  (the quantity
// TotalEnstrophy accumulates before AnalysisScalar::normalize), so that the sum over ranks divided by the
// global volume is the reference's total enstrophy.  `velocity` is the dense device field
// [D][LX][NM][NR] of the last stored step, components `fieldStride` elements apart.
// `launches` is incremented by the number of kernels / library calls enqueued.
int spectralEnqueue(SpectralEnstrophy* plan, const void* velocity, long long fieldStride, double* out, cudaStream_t stream,
                    unsigned long long* launches, std::string* error);

// Enqueues on `stream`: out[0 .. bins) = this rank's share of the reference's power spectrum of a dense field of D components
// (PowerSpectra, Analysis.h:122-177, as driven by SpectralAnalysisList::writeAnalyses, AnalysisList.h:132-170): per stored
// half-spectrum mode of the unnormalised transform, sum_d |a^_d|^2 (halved where the last wave number is 0) into bin
// floor(|k|).  The sum over ranks is the reference's spectrum before AnalysisSpectral::normalize.
int spectralPowerSpectrum(SpectralEnstrophy* plan, const void* field, long long fieldStride, int bins, double* out,
                          cudaStream_t stream, unsigned long long* launches, std::string* error);

}  // namespace mlbm
