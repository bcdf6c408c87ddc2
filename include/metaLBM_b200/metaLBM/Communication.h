// metaLBM/Communication.h (B200 drop-in) -- `Communication<T, latticeType, AlgorithmType::Pull, memoryLayout,
// PartitionningType::OneD, CommunicationType::MPI, Dimension>` and the alias `Communication_`
// (Communication.h:26-222, 494-500, 619-620).
//
// The x-slab halo exchange (Communication.h:134-180) is part of mlbm_step: boundary planes first, then
// ncclSend/ncclRecv (or peer copies) of the faceQ crossing populations on a high-priority stream while the bulk
// kernel runs.  communicateHalos() therefore has nothing left to do on the host and is a documented no-op;
// reduce() is the reference's sum over the local interior followed by a sum over ranks.
#pragma once

#include "Computation.h"
#include "Context.h"

namespace lbm {

template <class T, LatticeType latticeType, AlgorithmType algorithmType, MemoryLayout memoryLayout,
          PartitionningType partitionningType, CommunicationType communicationType, unsigned int Dimension>
class Communication {
  static_assert(algorithmType == AlgorithmType::Pull && partitionningType == PartitionningType::OneD,
                "metalbm_b200: the 1-D x-slab pull communication is the implemented one (2-D/3-D partitions are TODO stubs in the reference, Communication.h:183-206)");

 public:
  Communication() {}

  // Communication.h:494-500 -- performed inside mlbm_step, overlapped with the bulk kernel
  void communicateHalos(T*) {}

  // Communication.h:76-89: element-wise sum over ranks (available on every rank here)
  void reduce(T* localSumPtr, unsigned int numberComponents) {
    double buffer[16];
    for (unsigned int offset = 0; offset < numberComponents; offset += 16) {
      const unsigned int count = numberComponents - offset < 16 ? numberComponents - offset : 16;
      for (unsigned int i = 0; i < count; ++i) buffer[i] = (double)localSumPtr[offset + i];
      LBM_B200_CALL(mlbm_reduce_sum(b200::Context::get(), buffer, (int)count));
      for (unsigned int i = 0; i < count; ++i) localSumPtr[offset + i] = (T)buffer[i];
    }
  }

  // Communication.h:91-101: sum of a local scalar field over the interior, then over ranks
  T reduce(T* localPtr) {
    T localSum = (T)0;
    Computation<Architecture::CPU, L::dimD>(lSD::sStart(), lSD::sEnd()).Do([&](const Position& iP) { localSum += localPtr[lSD::getIndex(iP)]; });
    reduce(&localSum, 1);
    return localSum;
  }
};

typedef Communication<dataT, latticeT, algorithmT, memoryL, partitionningT, communicationT, L::dimD> Communication_;

}  // namespace lbm
