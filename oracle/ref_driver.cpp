// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference CPU implementation (/root/reference/include/metaLBM,
// compiled where it lies by oracle/refbuild.py with the stub headers in oracle/shim/)
// through its own step API, the way Routine::compute does (Routine.h:90-154):
//
//   unpack -> for it: isStored = ...; iterate(it, streams, events) -> pack
//
// but with arbitrary initial populations read from a file and with populations,
// fields and scalar observables dumped as raw float64 for the parity tests.
// Everything physical is done by reference code: Algorithm::iterate (Algorithm.h:326-358),
// TotalEnergy / TotalEnstrophy (Analysis.h:33-98), Curl (Transformer.h:118-295),
// Communication::reduce (Communication.h:76-89), SpectralAnalysisList (AnalysisList.h:99-202).
//
// usage: ref_driver <populations.bin | -> <output prefix> <steps> <store every> [observables 0|1] [dump 0|1] [warm-up steps]
//   populations.bin : float64 [Q][GX][GY][GZ] global interior populations ("-" = the
//                     reference's own equilibrium initialisation, Initialize.h:91-118)
#include "Input.in"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "metaLBM/Commons.h"
#include "metaLBM/MPIInitializer.h"
#include "metaLBM/FFTWInitializer.h"
#include "metaLBM/MathVector.h"
#include "metaLBM/Routine.h"

using namespace lbm;

static void writeRaw(const std::string& name, const std::vector<double>& data) {
  FILE* file = fopen(name.c_str(), "wb");
  if (!file) { fprintf(stderr, "ref_driver: cannot open %s\n", name.c_str()); exit(2); }
  fwrite(data.data(), sizeof(double), data.size(), file);
  fclose(file);
}

int main(int argc, char* argv[]) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <populations.bin|-> <prefix> <steps> <store every> [observables]\n", argv[0]);
    return 1;
  }
  const std::string inputName = argv[1];
  const std::string outputPrefix = argv[2];
  const int steps = atoi(argv[3]);
  const int storeEvery = atoi(argv[4]);
  const bool withObservables = argc > 5 ? atoi(argv[5]) != 0 : true;
  const bool withDump = argc > 6 ? atoi(argv[6]) != 0 : true;  // timing runs skip the dump and the stored step
  const int warmup = argc > 7 ? atoi(argv[7]) : 0;             // leading steps left out of the timers

  auto mpiLauncher = MPIInitializer<numProcs>{argc, argv};
  auto fftwLauncher = FFTWInitializer<numThreads>{};
  const int rank = MPIInit::rank[d::X];
  const unsigned int numberElements = FFTWInit::numberElements;

  constexpr Architecture arch = Architecture::CPU;
  Communication_ communication;
  Stream<arch> defaultStream(true), bulkStream(false), leftStream(false), rightStream(false);
  Event<arch> leftEvent, rightEvent;
  FieldWriter_ fieldWriter(prefix);
  FieldList<dataT, arch> fieldList(fieldWriter, defaultStream);
  Distribution<dataT, arch> distribution =
      initDistribution<dataT, arch>(fieldList.density, fieldList.velocity, defaultStream);

  const Position local = lSD::sLength();
  const Position global = gSD::sLength();
  const size_t localVolume = (size_t)local[d::X] * local[d::Y] * local[d::Z];
  const size_t globalVolume = (size_t)global[d::X] * global[d::Y] * global[d::Z];

  if (inputName != "-") {
    FILE* file = fopen(inputName.c_str(), "rb");
    if (!file) { fprintf(stderr, "ref_driver: cannot open %s\n", inputName.c_str()); return 2; }
    std::vector<double> slab(localVolume);
    for (int iQ = 0; iQ < L::dimQ; ++iQ) {
      const size_t begin = (size_t)iQ * globalVolume + (size_t)rank * localVolume;
      fseek(file, (long)(begin * sizeof(double)), SEEK_SET);
      if (fread(slab.data(), sizeof(double), localVolume, file) != localVolume) {
        fprintf(stderr, "ref_driver: short read on %s\n", inputName.c_str());
        return 2;
      }
      dataT* component = distribution.getData(numberElements, iQ);
      size_t i = 0;
      for (unsigned int x = 0; x < local[d::X]; ++x)
        for (unsigned int y = 0; y < local[d::Y]; ++y)
          for (unsigned int z = 0; z < local[d::Z]; ++z)
            component[lSD::getIndex(Position({x, y, z}))] = slab[i++];
    }
    fclose(file);
  }

  const Position fourierOffset = gFD::offset(MPIInit::rank);
  const auto lengths = Cast<unsigned int, ptrdiff_t, 3>::Do(gSD::sLength());
  Curl<double, Architecture::CPU, PartitionningType::OneD, L::dimD, L::dimD> curlVelocity(
      fieldList.velocity.getData(numberElements), fieldList.vorticity.getData(numberElements),
      lengths.data(), fourierOffset);
  TotalEnergy<dataT> totalEnergy(fieldList.density.getData(numberElements),
                                 fieldList.velocity.getData(numberElements));
  TotalEnstrophy<dataT> totalEnstrophy(fieldList.vorticity.getData(numberElements));
  Computation<Architecture::CPU, L::dimD> computationLocal(lSD::sStart(), lSD::sEnd());

  Algorithm<dataT, algorithmT, arch, memoryL, partitionningT, communicationT, overlappingT>
      algorithm(fieldList, distribution, communication);

  algorithm.unpack(defaultStream);

  std::string observables;
  double computationTime = 0.0, communicationTime = 0.0;
  const auto wallStart = std::chrono::high_resolution_clock::now();
  for (int iteration = 1; iteration <= steps; ++iteration) {
    algorithm.isStored = (storeEvery > 0 && iteration % storeEvery == 0) || (withDump && iteration == steps);
    algorithm.iterate(iteration, defaultStream, bulkStream, leftStream, rightStream,
                      leftEvent, rightEvent);
    if (iteration > warmup) {
      computationTime += algorithm.getComputationTime();
      communicationTime += algorithm.getCommunicationTime();
    }

    if (algorithm.isStored && withObservables) {
      // Routine.h:129-132 then ScalarAnalysisList::writeAnalyses (AnalysisList.h:55-73)
      // (the stub FFT is single-rank: with NPROCS > 1 the curl is skipped and enstrophy is 0)
      if (numProcs == 1) {
        curlVelocity.executeSpace();
        curlVelocity.normalize();
      }
      totalEnergy.reset();
      totalEnstrophy.reset();
      computationLocal.Do([&] LBM_HOST(const Position& iP) {
        totalEnergy(iP);
        totalEnstrophy(iP);
      });
      totalEnergy.normalize();
      totalEnstrophy.normalize();
      communication.reduce(&(totalEnergy.scalar), 1);
      communication.reduce(&(totalEnstrophy.scalar), 1);
      if (rank == 0) {
        char line[256];
        snprintf(line, sizeof(line), "obs %d %.17g %.17g\n", iteration, totalEnergy.scalar,
                 totalEnstrophy.scalar);
        observables += line;
      }
    }
  }
  const auto wallEnd = std::chrono::high_resolution_clock::now();

  algorithm.pack(defaultStream);

  // dump this rank's slab: f[Q], density, velocity[D], alpha, force[D], vorticity[2D-3]
  const int numberFields = L::dimQ + 1 + L::dimD + 1 + L::dimD + (2 * L::dimD - 3);
  std::vector<double> out;
  if (withDump) out.reserve((size_t)numberFields * localVolume);
  auto gather = [&](dataT* component) {
    if (!withDump) return;
    for (unsigned int x = 0; x < local[d::X]; ++x)
      for (unsigned int y = 0; y < local[d::Y]; ++y)
        for (unsigned int z = 0; z < local[d::Z]; ++z)
          out.push_back(component[lSD::getIndex(Position({x, y, z}))]);
  };
  for (int iQ = 0; iQ < L::dimQ; ++iQ) gather(distribution.getData(numberElements, iQ));
  gather(fieldList.density.getData(numberElements));
  for (int iD = 0; iD < L::dimD; ++iD) gather(fieldList.velocity.getData(numberElements, iD));
  gather(fieldList.alpha.getData(numberElements));
  for (int iD = 0; iD < L::dimD; ++iD) gather(fieldList.force.getData(numberElements, iD));
  for (int iD = 0; iD < 2 * L::dimD - 3; ++iD) gather(fieldList.vorticity.getData(numberElements, iD));
  if (withDump) writeRaw(outputPrefix + ".r" + std::to_string(rank) + ".bin", out);

  // SpectralAnalysisList::writeAnalyses (AnalysisList.h:132-170) on the fields of the last stored step: energy spectrum of
  // the stored velocity, forcing spectrum of the force array.  Done after the dump: its transforms run in place on the
  // fields (forward, then backward / V).  The writer's file goes to ../output/ (Writer.h:37), which need not exist.
  std::string spectra;
  if (withDump && withObservables && numProcs == 1) {
    SpectralAnalysisList<dataT, arch> spectralAnalysisList(fieldList, communication, 1, 0);
    spectralAnalysisList.writeAnalyses(steps);
    for (unsigned int k = 0; k < gFD::maxWaveNumber(); ++k) {
      char line[256];
      snprintf(line, sizeof(line), "spec %u %.17g %.17g\n", k, spectralAnalysisList.energySpectra.spectra[k],
               spectralAnalysisList.forcingSpectra.spectra[k]);
      spectra += line;
    }
  }

  communication.reduce(&computationTime, 1);
  communication.reduce(&communicationTime, 1);
  if (rank == 0) {
    FILE* file = fopen((outputPrefix + ".txt").c_str(), "w");
    fprintf(file, "lattice D%dQ%d\nranks %d\nlocal %u %u %u\nglobal %u %u %u\nsteps %d\n", L::dimD,
            L::dimQ, numProcs, local[d::X], local[d::Y], local[d::Z], global[d::X], global[d::Y],
            global[d::Z], steps);
    fprintf(file, "%s", observables.c_str());
    fprintf(file, "%s", spectra.c_str());
    // per-rank averages of the reference's own timers (Algorithm.h:340-357)
    fprintf(file, "time_computation %.9g\ntime_communication %.9g\ntime_wall %.9g\n",
            computationTime / numProcs, communicationTime / numProcs,
            std::chrono::duration<double>(wallEnd - wallStart).count());
    fclose(file);
  }
  return 0;
}
