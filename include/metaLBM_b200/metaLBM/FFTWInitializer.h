// metaLBM/FFTWInitializer.h (B200 drop-in) -- keeps `FFTWInit::numberElements`, the component stride of every
// local field (FFTWInitializer.h:13-41).  For the slab decomposition FFTW-MPI returns exactly the padded local
// volume, which is what this header uses; no FFTW is involved on the step path.
#pragma once

#include "Domain.h"

namespace lbm {

template <int numThreadsAtCompileTime>
struct FFTWInitializer {
  static unsigned int numberElements;
  FFTWInitializer() { numberElements = lSD::pVolume(); }
  ~FFTWInitializer() {}
};

using FFTWInit = FFTWInitializer<numThreads>;
template <> unsigned int FFTWInit::numberElements = lSD::pVolume();

}  // namespace lbm
