// metaLBM/Writer.h (B200 drop-in) -- only what sits right behind the hot path: the two ASCII tables the analysis lists
// append to, `../output/<prefix>/observables_<startIteration>.dat` and `spectra_<startIteration>.dat`.  The method names
// are the ones ScalarAnalysisList / SpectralAnalysisList call in the reference (AnalysisList.h:41-93, 163-199); the file
// format (one line per record: unsigned iteration, then the values with 16 significant digits, every field followed by ONE
// blank, also the last) is pinned byte for byte against the reference's own writer in
// tests/test_cpp_shim.py::test_observables_file_format_equals_the_reference.  One deviation: the output directory is created
// when it is missing.  The HDF5 / XDMF field and checkpoint writers are out of scope.
#pragma once

#include <sys/stat.h>

#include <cstdio>
#include <string>

#include "Context.h"
#include "Options.h"

namespace lbm {
namespace b200 {

// iteration is a multiple of period; a period of 0 means never (the reference divides by it unguarded)
inline bool isDue(unsigned int iteration, unsigned int period) { return period != 0 && iteration % period == 0; }

// an append-only text table; every call opens and closes the file, like the reference's writers do
class AnalysisTable {
  std::string folder, path;
  unsigned int period;

  std::FILE* open(const char* mode) const {
    std::FILE* file = std::fopen(path.c_str(), mode);
    if (!file) {  // first use: ../output/ and ../output/<prefix>/ may not exist yet
      ::mkdir("../output", 0777);
      ::mkdir(folder.c_str(), 0777);
      file = std::fopen(path.c_str(), mode);
    }
    if (!file) std::printf("Could not open file %s\n", path.c_str());
    return file;
  }

 public:
  AnalysisTable(const std::string& prefix_in, const std::string& name, unsigned int startIteration, unsigned int period_in)
      : folder("../output/" + prefix_in), path(folder + "/" + name + "_" + std::to_string(startIteration) + ".dat"), period(period_in) {}

  bool due(unsigned int iteration) const { return isDue(iteration, period); }
  void header(const std::string& text) const {
    if (std::FILE* file = open("w")) { std::fprintf(file, "%s\n", text.c_str()); std::fclose(file); }
  }
  // "<key_0> <key_1> ... <value_0> ... <value_n-1> \n"
  template <class T>
  void record(const unsigned int* keys, unsigned int keyCount, T* const* columns, unsigned int columnCount, unsigned int row) const {
    std::FILE* file = open("a");
    if (!file) return;
    for (unsigned int k = 0; k < keyCount; ++k) std::fprintf(file, "%u ", keys[k]);
    for (unsigned int c = 0; c < columnCount; ++c) std::fprintf(file, "%.16g ", (double)columns[c][row]);
    std::fprintf(file, "\n");
    std::fclose(file);
  }
};

}  // namespace b200

template <class T, InputOutputFormat inputOutputFormat>
class ScalarAnalysisWriter {
  static_assert(inputOutputFormat == InputOutputFormat::ascii, "metalbm_b200 writes the analyses as ascii (the reference's only instantiation)");
  b200::AnalysisTable table;

 public:
  ScalarAnalysisWriter(const std::string& folder, const std::string& name, unsigned int startIteration, unsigned int step)
      : table(folder, name, startIteration, step) {}
  bool getIsAnalyzed(unsigned int iteration) { return table.due(iteration); }
  void writeHeader(const std::string& header) { table.header(header); }
  void openFile(unsigned int) {}
  void closeFile() {}
  template <unsigned int Count>
  void writeAnalysis(unsigned int iteration, T* data) {
    T* columns[Count];
    for (unsigned int c = 0; c < Count; ++c) columns[c] = data + c;
    table.record(&iteration, 1, columns, Count, 0);
  }
};
typedef ScalarAnalysisWriter<dataT, InputOutputFormat::ascii> ScalarAnalysisWriter_;

template <class T, InputOutputFormat inputOutputFormat>
class SpectralAnalysisWriter {
  static_assert(inputOutputFormat == InputOutputFormat::ascii, "metalbm_b200 writes the analyses as ascii (the reference's only instantiation)");
  b200::AnalysisTable table;

 public:
  SpectralAnalysisWriter(const std::string& folder, const std::string& name, unsigned int startIteration, unsigned int step)
      : table(folder, name, startIteration, step) {}
  bool getIsAnalyzed(unsigned int iteration) { return table.due(iteration); }
  void writeHeader(const std::string& header) { table.header(header); }
  void openFile(unsigned int) {}
  void closeFile() {}
  // one line per wave number: "iteration wavenumber energy_spectra forcing_spectra "
  template <unsigned int Count>
  void writeAnalysis(unsigned int iteration, unsigned int maxWaveNumber, T* data[Count]) {
    for (unsigned int k = 0; k < maxWaveNumber; ++k) {
      const unsigned int keys[2] = {iteration, k};
      table.record(keys, 2, data, Count, k);
    }
  }
};
typedef SpectralAnalysisWriter<dataT, InputOutputFormat::ascii> SpectralAnalysisWriter_;

// DistributionWriter<T, InputOutput::HDF5> (Writer.h:400-445) and DistributionReader<T, InputOutput::HDF5> (Reader.h:119-157):
// the checkpoint `../output/<prefix>/distribution-<iteration>` -- dimQ data sets "distribution<iQ>" of the padded global box,
// every rank its hyperslab.  Same call sequence as the reference (getIsBackedUp / openFile / writeDistribution / closeFile;
// openFile / readDistribution / closeFile), same data-set layout; the container is the flat one of mlbm_checkpoint_write
// (extension .mlbm; tools/checkpoint_to_hdf5.py converts to and from the reference's .h5) because this build has no HDF5.
// writeDistribution saves the DEVICE state, which is what Algorithm::pack has just copied into `distribution`
// (Routine.h:212-216); readDistribution loads it to the device AND into the host array.
namespace b200 {
inline std::string checkpointPath(const std::string& filePrefix, unsigned int iteration, bool create) {
  const std::string folder = "../output/" + filePrefix;
  if (create) {
    ::mkdir("../output", 0777);
    ::mkdir(folder.c_str(), 0777);
  }
  return folder + "/distribution-" + std::to_string(iteration) + ".mlbm";
}
}  // namespace b200

template <class T, InputOutput inputOutput>
class DistributionWriter {};

template <class T>
class DistributionWriter<T, InputOutput::HDF5> {
  const std::string filePrefix;
  std::string path;
  unsigned int iteration = 0;

 public:
  DistributionWriter(const std::string& filePrefix_in) : filePrefix(filePrefix_in) {}
  bool getIsBackedUp(const unsigned int iteration_in) { return b200::isDue(iteration_in, backUpStep); }  // Writer.h:412-414
  void openFile(const unsigned int iteration_in) {
    iteration = iteration_in;
    path = b200::checkpointPath(filePrefix, iteration_in, true);   // every rank: mkdir is idempotent, nobody waits for rank 0
  }
  template <class Distribution_>
  void writeDistribution(Distribution_&) { LBM_B200_CALL(mlbm_checkpoint_write(b200::Context::get(), path.c_str(), iteration)); }
  void closeFile() {}
};
typedef DistributionWriter<dataT, InputOutput::HDF5> DistributionWriter_;

template <class T, InputOutput inputOutput>
class DistributionReader {};

template <class T>
class DistributionReader<T, InputOutput::HDF5> {
  const std::string filePrefix;
  std::string path;

 public:
  DistributionReader(const std::string& filePrefix_in) : filePrefix(filePrefix_in) {}
  void openFile(const unsigned int iteration) { path = b200::checkpointPath(filePrefix, iteration, false); }
  template <class Distribution_>
  void readDistribution(Distribution_& distribution) {
    LBM_B200_CALL(mlbm_checkpoint_read(b200::Context::get(), path.c_str(), nullptr));
    LBM_B200_CALL(mlbm_download_distribution(b200::Context::get(), distribution.getData(FFTWInit::numberElements), FFTWInit::numberElements,
                                             lSD::pLength()[d::Y], lSD::pLength()[d::Z]));
  }
  void closeFile() {}
};
typedef DistributionReader<dataT, InputOutput::HDF5> DistributionReader_;

}  // namespace lbm
