// metaLBM/Writer.h (B200 drop-in) -- the part of the reference's writers that sits right behind the hot path: the ASCII
// table of the scalar analyses, `../output/<prefix>/observables_<startIteration>.dat` (Writer.h:22-105 Writer<ascii>,
// :140-190 ScalarAnalysisWriter; opened by ScalarAnalysisList, AnalysisList.h:41).  Same file name, header line, column
// order, precision (16 significant digits, default float format) and the same trailing blank before every newline, so
// that post-processing scripts written against the reference keep working; pinned byte for byte against the reference's
// own class in tests/test_cpp_shim.py::test_observables_file_format_equals_the_reference.
// The only deviation: the output directory is created when it is missing (the reference prints "Could not open file").
// The HDF5 / XDMF field and checkpoint writers (Writer.h:252-560) are out of scope.
#pragma once

#include <sys/stat.h>

#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "Options.h"

namespace lbm {

template <class T, InputOutput inputOutput, InputOutputFormat inputOutputFormat>
class Writer {};

template <class T>
class Writer<T, InputOutput::Generic, InputOutputFormat::ascii> {
 protected:
  const std::string writeFolder;
  const std::string writerFolder;
  const std::string fileExtension;
  const std::string filePrefix;
  std::ofstream file;

  Writer(const std::string& writerFolder_in, const std::string& filePrefix_in, const std::string& fileExtension_in)
      : writeFolder("../output/"), writerFolder(writerFolder_in + "/"), fileExtension(fileExtension_in), filePrefix(filePrefix_in) {}

  inline std::string getFileName(const std::string& postfix = "") {
    return writeFolder + writerFolder + filePrefix + postfix + fileExtension;
  }

  inline void makeFolders() {
    ::mkdir(writeFolder.c_str(), 0777);
    ::mkdir((writeFolder + writerFolder).c_str(), 0777);
  }

  inline void open(const std::string& fileName, std::ios_base::openmode mode) {
    file.open(fileName, std::ofstream::out | mode);
    if (!file) {
      makeFolders();
      file.clear();
      file.open(fileName, std::ofstream::out | mode);
    }
    file.precision(16);
    if (!file) std::cout << "Could not open file " << fileName << std::endl;
  }

  inline void openAndAppend(const std::string& fileName) { open(fileName, std::ofstream::app); }
  inline void openAndTruncate(const std::string& fileName) { open(fileName, std::ofstream::trunc); }

  template <class U>
  inline void write(const U data) { file << data; }
};

template <class T, InputOutputFormat inputOutputFormat>
class ScalarAnalysisWriter : public Writer<T, InputOutput::Generic, inputOutputFormat> {
  static_assert(inputOutputFormat == InputOutputFormat::ascii, "metalbm_b200 writes the scalar analyses as ascii (the reference's only instantiation, Writer.h:562-563)");
  using Base = Writer<T, InputOutput::Generic, inputOutputFormat>;
  unsigned int startIteration;
  unsigned int analysisStep;

 public:
  ScalarAnalysisWriter(const std::string& writerFolder_in, const std::string& filePrefix_in, const unsigned int startIteration_in,
                       const unsigned int analysisStep_in)
      : Base(writerFolder_in, filePrefix_in, ".dat"), startIteration(startIteration_in), analysisStep(analysisStep_in) {}

  // the reference divides by analysisStep unguarded (Writer.h:160-162); 0 means "never" here instead of a division by zero
  inline bool getIsAnalyzed(const unsigned int iteration) { return analysisStep != 0 && (iteration % analysisStep) == 0; }

  inline std::string fileName() { return Base::getFileName("_" + std::to_string(startIteration)); }
  inline void openFile(const unsigned int) { Base::openAndAppend(fileName()); }
  inline void closeFile() { Base::file.close(); }

  template <unsigned int NumberScalarAnalyses>
  void writeAnalysis(const unsigned int iteration, T* data) {
    Base::write(iteration);
    Base::file << " ";
    for (unsigned int iS = 0; iS < NumberScalarAnalyses; ++iS) {
      Base::write(data[iS]);
      Base::file << " ";
    }
    Base::file << std::endl;
  }

  void writeHeader(const std::string& header) {
    Base::openAndTruncate(fileName());
    Base::file << header << std::endl;
    closeFile();
  }
};

typedef ScalarAnalysisWriter<dataT, InputOutputFormat::ascii> ScalarAnalysisWriter_;

// SpectralAnalysisWriter (Writer.h:193-251): `../output/<prefix>/spectra_<startIteration>.dat`, one line per wave number and
// analysed iteration: "iteration wavenumber energy_spectra forcing_spectra " (the same trailing blank).
template <class T, InputOutputFormat inputOutputFormat>
class SpectralAnalysisWriter : public Writer<T, InputOutput::Generic, inputOutputFormat> {
  static_assert(inputOutputFormat == InputOutputFormat::ascii, "metalbm_b200 writes the spectral analyses as ascii (Writer.h:564-565)");
  using Base = Writer<T, InputOutput::Generic, inputOutputFormat>;
  unsigned int startIteration;
  unsigned int analysisStep;

 public:
  SpectralAnalysisWriter(const std::string& writerFolder_in, const std::string& filePrefix_in, const unsigned int startIteration_in,
                         const unsigned int analysisStep_in)
      : Base(writerFolder_in, filePrefix_in, ".dat"), startIteration(startIteration_in), analysisStep(analysisStep_in) {}

  // the reference divides by analysisStep unguarded (Writer.h:211-213); 0 means "never" here
  inline bool getIsAnalyzed(const unsigned int iteration) { return analysisStep != 0 && (iteration % analysisStep) == 0; }

  inline std::string fileName() { return Base::getFileName("_" + std::to_string(startIteration)); }
  inline void openFile(const unsigned int) { Base::openAndAppend(fileName()); }
  inline void closeFile() { Base::file.close(); }

  template <unsigned int NumberSpectralAnalyses>
  void writeAnalysis(const unsigned int iteration, const unsigned int maxWaveNumber, T* data[NumberSpectralAnalyses]) {
    for (unsigned int kNorm = 0; kNorm < maxWaveNumber; ++kNorm) {   // Writer.h:226-238
      Base::write(iteration);
      Base::file << " ";
      Base::write(kNorm);
      Base::file << " ";
      for (unsigned int iS = 0; iS < NumberSpectralAnalyses; ++iS) {
        Base::write(data[iS][kNorm]);
        Base::file << " ";
      }
      Base::file << std::endl;
    }
  }

  void writeHeader(const std::string& header) {
    Base::openAndTruncate(fileName());
    Base::file << header << std::endl;
    closeFile();
  }
};

typedef SpectralAnalysisWriter<dataT, InputOutputFormat::ascii> SpectralAnalysisWriter_;

}  // namespace lbm
