// The reference's src/main.cu, unchanged except Architecture::CPU -> Architecture::GPU (src/main.cu:20 runs the CPU
// algorithm from the "GPU" main; this one runs the B200 path), compiled against the drop-in headers:
//   g++ -std=c++14 -DNPROCS=1 -DNTHREADS=1 -DGLOBAL_LENGTH_X=128 -DGLOBAL_LENGTH_Y=128 -DGLOBAL_LENGTH_Z=1
//       -DLBM_POSTFIX='"demo"' -include examples/Input_d2q9_kolmogorov.in -I include/metaLBM_b200
//       examples/main_gpu.cpp -L metalbm_b200 -lmetalbm_b200 -Wl,-rpath,$PWD/metalbm_b200
// (`-include <Input.in>` plays the role of the reference's `#include "Input.in"` first line, which resolves to a
// git-ignored copy made by its CMake, src/CMakeLists.txt:45-60.)
#include "metaLBM/Computation.cuh"
#include "metaLBM/Event.cuh"
#include "metaLBM/CUDAInitializer.h"
#include "metaLBM/MPIInitializer.h"
#include "metaLBM/FFTWInitializer.h"
#include "metaLBM/Commons.h"
#include "metaLBM/MathVector.h"
#include "metaLBM/Routine.h"

int main(int argc, char* argv[]) {
  using namespace lbm;
  LBM_INSTRUMENT_ON("main", 0)

  auto mpiLauncher = MPIInitializer<numProcs>{argc, argv};
  auto cudaLauncher = CUDAInitializer{};
  auto fftwLauncher = FFTWInitializer<numThreads>{};

  Routine<dataT, algorithmT, Architecture::GPU, memoryL, partitionningT, communicationT, overlappingT> routine;

  routine.compute();
}
