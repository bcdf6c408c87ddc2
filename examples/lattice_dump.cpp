// Prints the lattice descriptor the drop-in headers expose (runs without a GPU); tests/test_cpp_shim.py compares it
// with the oracle's tables, i.e. with the reference's Lattice.h.
#include <cstdio>

#include "metaLBM/Lattice.h"
#include "metaLBM/Domain.h"

int main() {
  using namespace lbm;
  std::printf("%d %d %d %d\n", L::dimD, L::dimQ, L::dimH, L::faceQ);
  for (int iQ = 0; iQ < L::dimQ; ++iQ) {
    for (int iD = 0; iD < 3; ++iD) std::printf("%d ", iD < L::dimD ? (int)L::celerity()[iQ][iD] : 0);
    std::printf("%.17g\n", (double)L::weight()[iQ]);
  }
  for (unsigned int i = 0; i < L::iQ_Bottom().size(); ++i) std::printf("%u ", L::iQ_Bottom()[i]);
  std::printf("| ");
  for (unsigned int i = 0; i < L::iQ_Top().size(); ++i) std::printf("%u ", L::iQ_Top()[i]);
  std::printf("| ");
  for (unsigned int i = 0; i < L::iQ_Front().size(); ++i) std::printf("%u ", L::iQ_Front()[i]);
  std::printf("| ");
  for (unsigned int i = 0; i < L::iQ_Back().size(); ++i) std::printf("%u ", L::iQ_Back()[i]);
  std::printf("\n%u %u %u %u %u\n", lSD::pLength()[d::X], lSD::pLength()[d::Y], lSD::pLength()[d::Z], lSD::pVolume(), hSD::volume());
  // unsigned celerities wrap like the reference's uiL (Lattice.h:806): -1 -> 2^32 - 1
  std::printf("%u\n", uiL::celerity()[1][0]);
  return 0;
}
