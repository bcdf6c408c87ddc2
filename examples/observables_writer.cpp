// Writes an observables table through ScalarAnalysisWriter_ the way ScalarAnalysisList does (AnalysisList.h:41, 64-69,
// 88-93).  Compiled twice by tests/test_cpp_shim.py: against the drop-in headers (include/metaLBM_b200) and, where the
// reference tree is present, against the reference's own Writer.h -- the two files must be identical byte for byte.
#ifdef OBSERVABLES_WRITER_REFERENCE
#include "Input.in"
#include "metaLBM/Commons.h"
#include "metaLBM/MPIInitializer.h"
#include "metaLBM/FFTWInitializer.h"
#include "metaLBM/MathVector.h"
#endif
#include "metaLBM/Writer.h"

int main() {
  using namespace lbm;
  ScalarAnalysisWriter_ writer(prefix, "observables", 7, 5);
  writer.writeHeader("iteration total_energy total_enstrophy");
  dataT rows[][2] = {{1.0 / 3.0, 2.0 / 3.0e9}, {0.0, -0.0}, {123456789.123456789, 1e-300}, {5.622066758466514e-09, 4.27299627548299e-08},
                     {1e22, 1.7976931348623157e308}, {0.1, 100.0}};
  unsigned int iteration = 5;
  for (auto& row : rows) {
    if (!writer.getIsAnalyzed(iteration)) return 2;
    writer.openFile(iteration);
    writer.writeAnalysis<2>(iteration, row);
    writer.closeFile();
    iteration += 5;
  }
  return writer.getIsAnalyzed(7) ? 3 : 0;
}
