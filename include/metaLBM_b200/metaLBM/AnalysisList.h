// metaLBM/AnalysisList.h (B200 drop-in) -- `ScalarAnalysisList<T, architecture>` (AnalysisList.h:26-97): total energy
// and total enstrophy of the stored step, summed over ranks, normalised by the global volume and appended by rank 0
// to the reference's ASCII table `../output/<prefix>/observables_<startIteration>.dat` ("iteration total_energy
// total_enstrophy", Writer.h:140-190) in the reference's exact format.  The reference loops over the host field arrays
// (Analysis.h:53-61, 85-93) and calls MPI_Reduce; here the sums were already formed on the device by the step
// kernel (warp shuffles + block partials) and reduced across GPUs by NCCL -- writeAnalyses only fetches them.
// Enstrophy is the reference's spectral one (Curl, Transformer.h:118-295), evaluated on the device (csrc/spectral.cu).
// The two observables the reference does not have (max Mach number, total mass) go to a second table next to it,
// `observables_b200_<startIteration>.dat`, so that the reference's file keeps its three columns.
#pragma once

#include <string>
#include <vector>

#include "Algorithm.h"
#include "Writer.h"

namespace lbm {

template <class T, Architecture architecture>
class ScalarAnalysisList {
  using Algorithm_t = Algorithm<T, algorithmT, architecture, memoryL, partitionningT, communicationT, overlappingT>;
  Algorithm_t& algorithm;
  const unsigned int startIteration;

 public:
  ScalarAnalysisWriter_ scalarAnalysisWriter;
  ScalarAnalysisWriter_ extraAnalysisWriter;
  T totalEnergy = (T)0, totalEnstrophy = (T)0, maxMach = (T)0, totalMass = (T)0;

  ScalarAnalysisList(Algorithm_t& algorithm_in, const unsigned int scalarAnalysisStep_in, const unsigned int startIteration_in)
      : algorithm(algorithm_in), startIteration(startIteration_in),
        scalarAnalysisWriter(prefix, "observables", startIteration_in, scalarAnalysisStep_in),
        extraAnalysisWriter(prefix, "observables_b200", startIteration_in, scalarAnalysisStep_in) {
    if (MPIInit::rank[d::X] == 0) {
      scalarAnalysisWriter.writeHeader("iteration total_energy total_enstrophy");  // AnalysisList.h:88-93
      extraAnalysisWriter.writeHeader("iteration max_mach total_mass");
    }
  }

  inline bool getIsAnalyzed(const unsigned int iteration) { return scalarAnalysisWriter.getIsAnalyzed(iteration); }

  inline void writeAnalyses(const unsigned int iteration) {
    if (!getIsAnalyzed(iteration) || iteration == startIteration) return;
    double out[4];
    algorithm.getObservables(out);
    totalEnergy = (T)out[0]; totalEnstrophy = (T)out[1]; maxMach = (T)out[2]; totalMass = (T)out[3];
    if (MPIInit::rank[d::X] == 0) {
      T scalarList[] = {totalEnergy, totalEnstrophy};  // AnalysisList.h:64-69
      scalarAnalysisWriter.openFile(iteration);
      scalarAnalysisWriter.template writeAnalysis<2>(iteration, scalarList);
      scalarAnalysisWriter.closeFile();
      T extraList[] = {maxMach, totalMass};
      extraAnalysisWriter.openFile(iteration);
      extraAnalysisWriter.template writeAnalysis<2>(iteration, extraList);
      extraAnalysisWriter.closeFile();
    }
  }
};

// `SpectralAnalysisList<T, architecture>` (AnalysisList.h:99-202): energy spectrum of the stored velocity and forcing
// spectrum of the force array, appended by rank 0 to `../output/<prefix>/spectra_<startIteration>.dat`.  The reference runs
// two in-place FFTW-MPI transforms per field on the host; here the device transforms of csrc/spectral.cu and a binning
// kernel do the work (mlbm_power_spectra) and the fields are left untouched.
template <class T, Architecture architecture>
class SpectralAnalysisList {
  using Algorithm_t = Algorithm<T, algorithmT, architecture, memoryL, partitionningT, communicationT, overlappingT>;
  Algorithm_t& algorithm;
  const unsigned int startIteration;

 public:
  SpectralAnalysisWriter_ spectralAnalysisWriter;
  std::vector<double> energySpectra, forcingSpectra;   // gFD::maxWaveNumber() bins each (FourierDomain.h:77-79)

  SpectralAnalysisList(Algorithm_t& algorithm_in, const unsigned int spectralAnalysisStep_in, const unsigned int startIteration_in)
      : algorithm(algorithm_in), startIteration(startIteration_in),
        spectralAnalysisWriter(prefix, "spectra", startIteration_in, spectralAnalysisStep_in) {
    if (MPIInit::rank[d::X] == 0 && spectralAnalysisStep_in != 0)
      spectralAnalysisWriter.writeHeader("iteration wavenumber energy_spectra forcing_spectra");   // AnalysisList.h:194-199
  }

  inline bool getIsAnalyzed(const unsigned int iteration) { return spectralAnalysisWriter.getIsAnalyzed(iteration); }

  inline void writeAnalyses(const unsigned int iteration) {
    const int bins = algorithm.getPowerSpectra(nullptr, nullptr, 0);
    energySpectra.assign((size_t)bins, 0.0);
    forcingSpectra.assign((size_t)bins, 0.0);
    algorithm.getPowerSpectra(energySpectra.data(), forcingSpectra.data(), bins);
    if (MPIInit::rank[d::X] == 0) {
      std::vector<T> energy(energySpectra.begin(), energySpectra.end()), forcing(forcingSpectra.begin(), forcingSpectra.end());
      T* spectraList[2] = {energy.data(), forcing.data()};   // AnalysisList.h:163-168
      spectralAnalysisWriter.openFile(iteration);
      spectralAnalysisWriter.template writeAnalysis<2>(iteration, (unsigned int)bins, spectraList);
      spectralAnalysisWriter.closeFile();
    }
  }
};

}  // namespace lbm
