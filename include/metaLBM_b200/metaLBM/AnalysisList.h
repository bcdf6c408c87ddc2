// metaLBM/AnalysisList.h (B200 drop-in) -- `ScalarAnalysisList<T, architecture>` (AnalysisList.h:26-97): total energy
// and total enstrophy of the stored step, summed over ranks, normalised by the global volume and appended by rank 0
// to an ASCII table "iteration total_energy total_enstrophy".  The reference loops over the host field arrays
// (Analysis.h:53-61, 85-93) and calls MPI_Reduce; here the sums were already formed on the device by the step
// kernel (warp shuffles + block partials) and reduced across GPUs by NCCL -- writeAnalyses only fetches them.
// Enstrophy is the reference's spectral one (Curl, Transformer.h:118-295), evaluated on the device (csrc/spectral.cu).
#pragma once

#include <fstream>
#include <iomanip>
#include <string>

#include "Algorithm.h"

namespace lbm {

template <class T, Architecture architecture>
class ScalarAnalysisList {
  using Algorithm_t = Algorithm<T, algorithmT, architecture, memoryL, partitionningT, communicationT, overlappingT>;
  Algorithm_t& algorithm;
  const unsigned int analysisStep, startIteration;
  const std::string fileName;

 public:
  T totalEnergy = (T)0, totalEnstrophy = (T)0, maxMach = (T)0, totalMass = (T)0;

  ScalarAnalysisList(Algorithm_t& algorithm_in, const unsigned int scalarAnalysisStep_in, const unsigned int startIteration_in,
                     const std::string& fileName_in = std::string("observables_") + prefix + ".dat")
      : algorithm(algorithm_in), analysisStep(scalarAnalysisStep_in), startIteration(startIteration_in), fileName(fileName_in) {
    if (MPIInit::rank[d::X] == 0) std::ofstream(fileName, std::ios::trunc) << "iteration total_energy total_enstrophy max_mach total_mass\n";
  }

  inline bool getIsAnalyzed(const unsigned int iteration) { return analysisStep && (iteration % analysisStep) == 0; }

  inline void writeAnalyses(const unsigned int iteration) {
    if (!getIsAnalyzed(iteration) || iteration == startIteration) return;
    double out[4];
    algorithm.getObservables(out);
    totalEnergy = (T)out[0]; totalEnstrophy = (T)out[1]; maxMach = (T)out[2]; totalMass = (T)out[3];
    if (MPIInit::rank[d::X] == 0)
      std::ofstream(fileName, std::ios::app) << std::setprecision(17) << iteration << " " << out[0] << " " << out[1] << " " << out[2]
                                             << " " << out[3] << "\n";
  }
};

}  // namespace lbm
