// metaLBM/Algorithm.h (B200 drop-in) -- `Algorithm<T, AlgorithmType::Pull, Architecture::GPU, MemoryLayout::SoA,
// PartitionningType::OneD, CommunicationType::MPI, Overlapping::{Off, On}>` with the members Routine::compute uses
// (Algorithm.h:31-452, Routine.h:90-154): the constructor (fieldList, distribution, communication), `isStored`,
// `iterate(iteration, 4 streams, 2 events)`, `pack(stream)`, `unpack(stream)`, `getCommunicationTime()`,
// `getComputationTime()`.  Every call forwards to the C-ABI; the fused node update (Algorithm::operator(),
// Algorithm.h:97-126) is the CUDA kernel behind mlbm_step.  Overlapping::On is the real thing here (in the reference
// snapshot it does not compile).
#pragma once

#include "Collision.h"
#include "Communication.h"
#include "Context.h"
#include "Distribution.h"
#include "Event.h"
#include "FieldList.h"
#include "Stream.h"

namespace lbm {

template <class T, AlgorithmType algorithmType, Architecture architecture, MemoryLayout memoryLayout,
          PartitionningType partitionningType, CommunicationType communicationType, Overlapping overlapping>
class Algorithm {
  static_assert(architecture == Architecture::GPU, "metalbm_b200 provides Architecture::GPU only: there is no CPU fallback");
  static_assert(algorithmType == AlgorithmType::Pull && memoryLayout == MemoryLayout::SoA && partitionningType == PartitionningType::OneD,
                "metalbm_b200 implements the Pull / SoA / OneD algorithm (Algorithm.h:300-452)");
  static_assert(overlapping == overlappingT, "the context is configured from the global overlappingT");

 public:
  using Communication_t = Communication<T, L::Type, AlgorithmType::Pull, memoryLayout, PartitionningType::OneD, communicationType, L::dimD>;

  T* densityPtr;
  T* velocityPtr;
  T* forcePtr;
  T* alphaPtr;
  T* distributionPtr;

 protected:
  mlbm_ctx* context;
  Collision_<architecture> collision;
  Communication_t communication;
  double dtComputation, dtCommunication;
  static constexpr size_t pY() { return lSD::pLength()[d::Y]; }
  static constexpr size_t pZ() { return lSD::pLength()[d::Z]; }

 public:
  bool isStored;

  Algorithm(FieldList<T, architecture>& fieldList_in, Distribution<T, architecture>& distribution_in,
            Communication_t& communication_in)
      : densityPtr(fieldList_in.density.getData(FFTWInit::numberElements)),
        velocityPtr(fieldList_in.velocity.getData(FFTWInit::numberElements)),
        forcePtr(fieldList_in.force.getData(FFTWInit::numberElements)),
        alphaPtr(fieldList_in.alpha.getData(FFTWInit::numberElements)),
        distributionPtr(distribution_in.getData(FFTWInit::numberElements)),
        context(b200::Context::get()),
        collision(relaxationTime, fieldList_in, forceAmplitude, forceWaveLength, forcekMin, forcekMax),
        communication(communication_in),
        dtComputation(0), dtCommunication(0), isStored(false) {}

  // Algorithm.h:326-358 (Off) / :392-447 (On)
  void iterate(const unsigned int iteration, Stream<architecture>&, Stream<architecture>&, Stream<architecture>&,
               Stream<architecture>&, Event<architecture>&, Event<architecture>&) {
    collision.update(iteration, FFTWInit::numberElements);
    LBM_B200_CALL(mlbm_step(context, iteration, isStored ? 1 : 0));
    LBM_B200_CALL(mlbm_timers(context, &dtCommunication, &dtComputation));
    if (isStored) {
      // Algorithm::storeFields (Algorithm.h:150-194): the reference's kernel writes the pinned host fields directly
      LBM_B200_CALL(mlbm_download_fields(context, densityPtr, velocityPtr, alphaPtr, forcePtr, FFTWInit::numberElements, pY(), pZ()));
    }
  }

  double getCommunicationTime() { return dtCommunication; }
  double getComputationTime() { return dtComputation; }

  // Algorithm.h:132-139: device -> local padded host array (before checkpoints)
  void pack(const Stream<architecture>&) {
    LBM_B200_CALL(mlbm_download_distribution(context, distributionPtr, FFTWInit::numberElements, pY(), pZ()));
  }
  // Algorithm.h:141-147: local padded host array -> device
  void unpack(const Stream<architecture>&) {
    LBM_B200_CALL(mlbm_upload_distribution(context, distributionPtr, FFTWInit::numberElements, pY(), pZ()));
    if (alphaPtr) LBM_B200_CALL(setAlphaIfEntropic());
  }

  // B200 additions (not in the reference): on-line observables of the last stored step, reduced over ranks
  //   out = {total energy, total enstrophy, max Mach, total mass}   (AnalysisList.h:55-73 as device reductions)
  void getObservables(double out[4]) { LBM_B200_CALL(mlbm_observables(context, out)); }
  //   energy / forcing spectra of the last stored step (SpectralAnalysisList, AnalysisList.h:132-170); returns the number of bins
  int getPowerSpectra(double* energySpectrum, double* forcingSpectrum, int capacity) {
    int count = 0;
    LBM_B200_CALL(mlbm_power_spectra(context, energySpectrum, forcingSpectrum, capacity, &count));
    return count;
  }
  mlbm_ctx* getContext() { return context; }

 private:
  int setAlphaIfEntropic() {
    return collisionT == CollisionType::BGK ? 0 : mlbm_set_alpha(context, alphaPtr, pY(), pZ());
  }
};

}  // namespace lbm
