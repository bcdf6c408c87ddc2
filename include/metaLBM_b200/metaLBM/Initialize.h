// metaLBM/Initialize.h (B200 drop-in) -- initDensity / initVelocity / initForce / initAlpha / initDistribution
// (Initialize.h:19-148).  The field initialisers fill host arrays; initDistribution reads the checkpoint of startIteration
// when that is not 0 (Initialize.h:119-124) and otherwise evaluates f = feq(rho, u) ON THE
// DEVICE (mlbm_init_equilibrium) and brings the result back into the Distribution's local array, so the host copy
// is what the reference would hold before Algorithm::unpack.
#pragma once

#include <iostream>

#include "Context.h"
#include "Distribution.h"
#include "Field.h"
#include "Writer.h"

namespace lbm {

template <class T, Architecture architecture>
Field<T, 1, architecture, true> initDensity(const Stream<architecture>& stream) {
  Field<T, 1, architecture, true> densityFieldR("density", (T)initDensityValue, stream);
  switch (initDensityT) {
    case InitDensityType::Homogeneous: break;
    case InitDensityType::Peak:  // 3 rho_0 at (0.4, 0.3, 0.2) (L - 1) of rank 0 (Initialize.h:30-46)
      if (MPIInit::rank[d::X] == 0) {
        Position center = {{static_cast<unsigned int>((lSD::sLength()[d::X] - 1) * (T)0.4),
                            static_cast<unsigned int>((lSD::sLength()[d::Y] - 1) * (T)0.3),
                            static_cast<unsigned int>((lSD::sLength()[d::Z] - 1) * (T)0.2)}};
        densityFieldR.setValue(center, (T)3.0 * (T)initDensityValue, FFTWInit::numberElements);
      }
      break;
    default: std::cout << "Wrong type of density initialization.";
  }
  return densityFieldR;
}

template <class T, Architecture architecture>
Field<T, L::dimD, architecture, true> initVelocity(const Stream<architecture>& stream) {
  MathVector<T, L::dimD> projected = {};
  for (int iD = 0; iD < L::dimD; ++iD) projected[iD] = (T)initVelocityVector[iD];
  Field<T, L::dimD, architecture, true> velocityFieldR("velocity", projected, stream);
  if (initVelocityT != InitVelocityType::Homogeneous) std::cout << "Wrong type of velocity initialization.";
  return velocityFieldR;
}

template <class T, Architecture architecture>
Field<T, L::dimD, architecture, writeForce> initForce(const Stream<architecture>& stream) {
  return Field<T, L::dimD, architecture, writeForce>("force", (T)0, stream);
}

template <class T, Architecture architecture>
Field<T, 1, architecture, writeAlpha> initAlpha(const Stream<architecture>& stream) {
  return Field<T, 1, architecture, writeAlpha>("alpha", (T)2, stream);
}

template <class T, Architecture architecture>
Distribution<T, architecture> initDistribution(Field<T, 1, architecture, true>& densityField,
                                               Field<T, L::dimD, architecture, true>& velocityField,
                                               const Stream<architecture>&) {
  static_assert(architecture == Architecture::GPU, "metalbm_b200 provides the Architecture::GPU path only (no CPU fallback)");
  Distribution<T, architecture> distributionR;
  mlbm_ctx* context = b200::Context::get();
  const size_t n = FFTWInit::numberElements, pY = lSD::pLength()[d::Y], pZ = lSD::pLength()[d::Z];
  if (startIteration == 0) {
    LBM_B200_CALL(mlbm_init_equilibrium(context, densityField.getData(n), velocityField.getData(n), n, pY, pZ));
    LBM_B200_CALL(mlbm_download_distribution(context, distributionR.getData(n), n, pY, pZ));
  } else {
    // restart (Initialize.h:119-124): the checkpoint DistributionWriter left at startIteration
    DistributionReader_ distributionReader(prefix);
    distributionReader.openFile(startIteration);
    distributionReader.readDistribution(distributionR);
    distributionReader.closeFile();
  }
  return distributionR;
}

}  // namespace lbm
