// Parity harness for the C++ drop-in layer: drives the reference's object model (FieldList, Distribution,
// Communication_, Algorithm_ -- the members of Routine, Routine.h:36-56) by hand so that arbitrary populations can be
// injected and read back.  Used by tests/test_cpp_shim.py; compiled with -include <Input.in> like examples/main_gpu.cpp.
//   shim_check <populations_in.bin> <steps> <populations_out.bin> <fields_out.bin> [<moments_out.bin>]
// populations_*.bin: raw dataT [dimQ][lx][ly][lz] of this rank's slab (interior only, z fastest).
// fields_out.bin   : density [lx][ly][lz], velocity [dimD][lx][ly][lz], alpha [lx][ly][lz], then 4 observables.
// moments_out.bin  : density, velocity [dimD], hydrodynamic velocity [dimD] of the state AFTER the last step, computed on the
//                    host with Collision::calculateMoments / getHydrodynamicVelocity over Distribution::getHaloDataPreviousHost().
#include <cstdio>
#include <vector>

#include "metaLBM/CUDAInitializer.h"
#include "metaLBM/MPIInitializer.h"
#include "metaLBM/FFTWInitializer.h"
#include "metaLBM/Algorithm.h"

using namespace lbm;

template <class Function>
static void forEachNode(Function function) {
  size_t n = 0;
  for (unsigned int x = 0; x < lSD::sLength()[d::X]; ++x)
    for (unsigned int y = 0; y < lSD::sLength()[d::Y]; ++y)
      for (unsigned int z = 0; z < lSD::sLength()[d::Z]; ++z) function(n++, lSD::getIndex(Position{{x, y, z}}));
}

int main(int argc, char* argv[]) {
  if (argc < 5) return 2;
  auto mpiLauncher = MPIInitializer<numProcs>{argc, argv};
  auto cudaLauncher = CUDAInitializer{};
  auto fftwLauncher = FFTWInitializer<numThreads>{};
  (void)cudaLauncher; (void)fftwLauncher;

  using Algorithm_ = Algorithm<dataT, algorithmT, Architecture::GPU, memoryL, partitionningT, communicationT, overlappingT>;
  Communication_ communication;
  Stream<Architecture::GPU> defaultStream(true), bulkStream(false), leftStream(false), rightStream(false);
  Event<Architecture::GPU> leftEvent, rightEvent;
  FieldList<dataT, Architecture::GPU> fieldList(defaultStream);
  Distribution<dataT, Architecture::GPU> distribution =
      initDistribution<dataT, Architecture::GPU>(fieldList.density, fieldList.velocity, defaultStream);
  Algorithm_ algorithm(fieldList, distribution, communication);

  const size_t volume = lSD::sVolume();
  std::vector<dataT> buffer(volume * L::dimQ);
  FILE* in = std::fopen(argv[1], "rb");
  if (!in || std::fread(buffer.data(), sizeof(dataT), buffer.size(), in) != buffer.size()) return 3;
  std::fclose(in);
  for (int iQ = 0; iQ < L::dimQ; ++iQ) {
    dataT* component = distribution.getData(FFTWInit::numberElements, iQ);
    forEachNode([&](size_t n, unsigned int index) { component[index] = buffer[iQ * volume + n]; });
  }

  const unsigned int steps = (unsigned int)std::atoi(argv[2]);
  algorithm.unpack(defaultStream);
  for (unsigned int iteration = 1; iteration <= steps; ++iteration) {
    algorithm.isStored = iteration == steps;
    algorithm.iterate(iteration, defaultStream, bulkStream, leftStream, rightStream, leftEvent, rightEvent);
  }
  double observables[4];
  algorithm.getObservables(observables);
  algorithm.pack(defaultStream);

  for (int iQ = 0; iQ < L::dimQ; ++iQ) {
    const dataT* component = distribution.getData(FFTWInit::numberElements, iQ);
    forEachNode([&](size_t n, unsigned int index) { buffer[iQ * volume + n] = component[index]; });
  }
  FILE* out = std::fopen(argv[3], "wb");
  std::fwrite(buffer.data(), sizeof(dataT), buffer.size(), out);
  std::fclose(out);

  std::vector<dataT> fields(volume * (2 + L::dimD));
  forEachNode([&](size_t n, unsigned int index) {
    fields[n] = fieldList.density.getData(FFTWInit::numberElements)[index];
    for (int iD = 0; iD < L::dimD; ++iD) fields[(1 + iD) * volume + n] = fieldList.velocity.getData(FFTWInit::numberElements, iD)[index];
    fields[(1 + L::dimD) * volume + n] = fieldList.alpha.getData(FFTWInit::numberElements)[index];
  });
  out = std::fopen(argv[4], "wb");
  std::fwrite(fields.data(), sizeof(dataT), fields.size(), out);
  std::fwrite(observables, sizeof(double), 4, out);
  std::fclose(out);
  if (argc > 5) {
    // the host-callable per-node surface (Moment.h:14-47, Collision.h:60-93) over a host copy of the halo-space buffer the
    // NEXT step would read: density, velocity and hydrodynamic velocity of the populations pulled to every interior node
    const dataT* halo = distribution.getHaloDataPreviousHost();
    Collision_<Architecture::GPU> collision(relaxationTime, fieldList, forceAmplitude, forceWaveLength, forcekMin, forcekMax);
    std::vector<dataT> moments(volume * (1 + 2 * L::dimD));
    size_t n = 0;
    for (unsigned int x = 0; x < lSD::sLength()[d::X]; ++x)
      for (unsigned int y = 0; y < lSD::sLength()[d::Y]; ++y)
        for (unsigned int z = 0; z < lSD::sLength()[d::Z]; ++z, ++n) {
          const Position iP = Position{{x + L::halo()[d::X], y + L::halo()[d::Y], z + L::halo()[d::Z]}};
          collision.calculateMoments(halo, iP);
          collision.setForce(fieldList.force.getData(FFTWInit::numberElements), iP, gSD::sOffset(MPIInit::rank), FFTWInit::numberElements);
          dataT density;
          Moment_::calculateDensity(halo, iP, density);
          if (density != collision.getDensity()) return 4;
          moments[n] = collision.getDensity();
          for (int iD = 0; iD < L::dimD; ++iD) {
            moments[(1 + iD) * volume + n] = collision.getVelocity()[iD];
            moments[(1 + L::dimD + iD) * volume + n] = collision.getHydrodynamicVelocity()[iD];
          }
        }
    out = std::fopen(argv[5], "wb");
    std::fwrite(moments.data(), sizeof(dataT), moments.size(), out);
    std::fclose(out);
  }
  const double mass = communication.reduce(fieldList.density.getData(FFTWInit::numberElements));
  std::printf("ok rank %d mass %.17g comm %.3e s comp %.3e s\n", MPIInit::rank[d::X], mass, algorithm.getCommunicationTime(),
              algorithm.getComputationTime());
  return 0;
}
