// metaLBM/Routine.h (B200 drop-in) -- `Routine<T, algorithmT, Architecture::GPU, memoryL, partitionningT,
// communicationT, overlappingT>` with the reference's constructor-and-compute() shape (Routine.h:23-241): builds the
// field list, the equilibrium distribution, the communication and the algorithm, then runs the time loop
//     algorithm.unpack; for it: algorithm.isStored = ...; algorithm.iterate(...); analyses      (Routine.h:90-154)
// The HDF5 / XDMF field writers and the performance table of the reference are CALLERS of the hot path and out of scope
// (SURVEY.md section 2); the scalar observables and spectra are kept because they are device reductions of the step itself,
// and the checkpoint of the distribution (backUpStep, Routine.h:212-216; restart from startIteration, Initialize.h:119-124)
// because it is the data format either side of the path.
#pragma once

#include <chrono>
#include <iostream>

#include "Algorithm.h"
#include "AnalysisList.h"

namespace lbm {

namespace b200 {
inline bool isMultiple(const unsigned int iteration, const unsigned int step) { return step != 0 && iteration % step == 0; }
}  // namespace b200

template <class T, AlgorithmType algorithmType, Architecture architecture, MemoryLayout memoryLayout,
          PartitionningType partitionningType, CommunicationType communicationType, Overlapping overlapping>
class Routine {
 protected:
  using Clock = std::chrono::high_resolution_clock;
  using Algorithm_ = Algorithm<T, algorithmType, architecture, memoryLayout, partitionningType, communicationType, overlapping>;

  Communication_ communication;
  Stream<architecture> defaultStream, bulkStream, leftStream, rightStream;
  Event<architecture> leftEvent, rightEvent;
  FieldList<T, architecture> fieldList;
  Distribution<T, architecture> distribution;
  Algorithm_ algorithm;
  ScalarAnalysisList<T, architecture> scalarAnalysisList;
  SpectralAnalysisList<T, architecture> spectralAnalysisList;   // Routine.h:52, 78-79
  DistributionWriter_ distributionWriter;                        // Routine.h:45, 67
  double computationTime = 0, communicationTime = 0, totalTime = 0;

 public:
  Routine()
      : communication(), defaultStream(true), bulkStream(false), leftStream(false), rightStream(false),
        fieldList(defaultStream),
        distribution(initDistribution<T, architecture>(fieldList.density, fieldList.velocity, defaultStream)),
        algorithm(fieldList, distribution, communication),
        scalarAnalysisList(algorithm, scalarAnalysisStep, startIteration),
        spectralAnalysisList(algorithm, spectralAnalysisStep, startIteration), distributionWriter(prefix) {}

  FieldList<T, architecture>& getFieldList() { return fieldList; }
  Distribution<T, architecture>& getDistribution() { return distribution; }
  Algorithm_& getAlgorithm() { return algorithm; }
  ScalarAnalysisList<T, architecture>& getScalarAnalysisList() { return scalarAnalysisList; }
  SpectralAnalysisList<T, architecture>& getSpectralAnalysisList() { return spectralAnalysisList; }

  void compute() {
    algorithm.unpack(defaultStream);
    const auto t0 = Clock::now();
    for (unsigned int iteration = startIteration + 1; iteration <= endIteration; ++iteration) {
      algorithm.isStored = scalarAnalysisList.getIsAnalyzed(iteration) || spectralAnalysisList.getIsAnalyzed(iteration) ||
                           b200::isMultiple(iteration, writeStep);   // Routine.h:122-124
      algorithm.iterate(iteration, defaultStream, bulkStream, leftStream, rightStream, leftEvent, rightEvent);
      if (algorithm.isStored) scalarAnalysisList.writeAnalyses(iteration);
      if (spectralAnalysisList.getIsAnalyzed(iteration)) spectralAnalysisList.writeAnalyses(iteration);   // Routine.h:227-229
      communicationTime += algorithm.getCommunicationTime();
      computationTime += algorithm.getComputationTime();
      if (distributionWriter.getIsBackedUp(iteration)) {   // Routine.h:212-216
        algorithm.pack(defaultStream);
        distributionWriter.openFile(iteration);
        distributionWriter.writeDistribution(distribution);
        distributionWriter.closeFile();
      }
    }
    defaultStream.synchronize();
    totalTime = std::chrono::duration<double>(Clock::now() - t0).count();
    if (MPIInit::rank[d::X] == 0) printOutputs();
  }

 protected:
  void printOutputs() {
    const double nodes = (double)gSD::sVolume() * (double)(endIteration - startIteration);
    std::cout << "Total time         : " << totalTime << " s\n"
              << "Computatation time : " << computationTime << " s\n"
              << "Communication time : " << communicationTime << " s\n"
              << "MLUPS              : " << nodes / totalTime * 1e-6 << "\n";
  }
};

}  // namespace lbm
