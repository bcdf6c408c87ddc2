// tests/emu/step_emu.cpp -- TEST INFRASTRUCTURE ONLY (never part of the product).
//
// Compiles the SOURCE of the fused step kernels (a copy of metalbm_b200/csrc/step_kernel.cuh in which the single
// `extern __shared__` declaration is redirected to the emulator's shared-memory buffer, made by tests/emu/build.py) for
// the host and runs one launch of fusedStepKernel<...> under tests/emu/cuda_emu.h.  The caller (tests/test_kernel_logic.py)
// lays the arrays out exactly as csrc/context.cu does and compares the result with the oracle.
#include "step_kernel_emu.cuh"

using namespace mlbm;

struct EmuLaunch {
  int lattice, collision, equilibrium, scheme, f32;
  const void* prev; void* next; void* alpha; void* density; void* velocity; void* force; double* partials;
  const double* forceTable[3]; int forceAxis[3];
  long long stride, plane, fieldStride;
  int LX, NM, NR, x0, planeCount, planeStep, planesPerBlock;
  void* peerLow; void* peerHigh;
  int wrapX, isStored, hydroShift, hasForce;
  double beta, guoFactor;
};

template <class L, int COLLISION, int EQ, int SCHEME, typename StoreT>
static int run(const EmuLaunch& e) {
  StepParams p;
  std::memset(&p, 0, sizeof(p));
  p.prev = e.prev; p.next = e.next; p.alpha = e.alpha; p.density = e.density; p.velocity = e.velocity; p.force = e.force;
  p.partials = e.partials;
  for (int d = 0; d < 3; ++d) { p.forceTable[d] = e.forceTable[d]; p.forceAxis[d] = e.forceAxis[d]; }
  p.stride = e.stride; p.plane = e.plane; p.fieldStride = e.fieldStride;
  p.LX = e.LX; p.NM = e.NM; p.NR = e.NR; p.x0 = e.x0; p.planeStep = e.planeStep; p.planeCount = e.planeCount;
  p.planesPerBlock = e.planesPerBlock; p.peerLow = e.peerLow; p.peerHigh = e.peerHigh;
  p.wrapX = e.wrapX; p.isStored = e.isStored; p.hydroShift = e.hydroShift; p.hasForce = e.hasForce;
  p.beta = e.beta; p.guoFactor = e.guoFactor;
  const int gridR = (e.NR + kStepBlock - 1) / kStepBlock;
  const dim3 grid((unsigned)gridR, (unsigned)e.NM, (unsigned)((e.planeCount + e.planesPerBlock - 1) / e.planesPerBlock));
  const size_t shared = COLLISION != kBGK ? (size_t)entropicSharedBytes(L::Q, logTableInShared(L::Q)) : 0;
  cuda_emu::launch<StepParams>(fusedStepKernel<L, COLLISION, EQ, SCHEME, StoreT>, grid, kStepBlock, shared, p);
  return 0;
}

template <class L, int COLLISION, int EQ, typename StoreT>
static int runScheme(const EmuLaunch& e) {
  switch (e.scheme) {
    case kSchemeNone: return run<L, COLLISION, EQ, kSchemeNone, StoreT>(e);
    case kSchemeGuo: return run<L, COLLISION, EQ, kSchemeGuo, StoreT>(e);
    case kSchemeEDM: return run<L, COLLISION, EQ, kSchemeEDM, StoreT>(e);
    default: return -1;
  }
}

template <class L, bool HAS_EXACT, typename StoreT>
static int runLattice(const EmuLaunch& e) {
  if (e.equilibrium == kTruncationMa3) {
    return e.collision == kBGK ? runScheme<L, kBGK, kTruncationMa3, StoreT>(e)
         : e.collision == kELBM ? runScheme<L, kELBM, kTruncationMa3, StoreT>(e) : runScheme<L, kELBMForcing, kTruncationMa3, StoreT>(e);
  }
  if constexpr (HAS_EXACT) {
    if (e.equilibrium == kExact)
      return e.collision == kBGK ? runScheme<L, kBGK, kExact, StoreT>(e)
           : e.collision == kELBM ? runScheme<L, kELBM, kExact, StoreT>(e) : runScheme<L, kELBMForcing, kExact, StoreT>(e);
  }
  return -1;
}

template <typename StoreT>
static int runType(const EmuLaunch& e) {
  switch (e.lattice) {
    case kD2Q5: return runLattice<Lattice<kD2Q5>, false, StoreT>(e);
    case kD2Q9: return runLattice<Lattice<kD2Q9>, true, StoreT>(e);
    case kD3Q15: return runLattice<Lattice<kD3Q15>, false, StoreT>(e);
    case kD3Q19: return runLattice<Lattice<kD3Q19>, false, StoreT>(e);
    case kD3Q27: return runLattice<Lattice<kD3Q27>, true, StoreT>(e);
    case kD2Q13: return runLattice<Lattice<kD2Q13>, false, StoreT>(e);
    case kD2Q17: return runLattice<Lattice<kD2Q17>, false, StoreT>(e);
    case kD2Q21: return runLattice<Lattice<kD2Q21>, false, StoreT>(e);
    case kD3Q33: return runLattice<Lattice<kD3Q33>, false, StoreT>(e);
    default: return -1;
  }
}

#define EMU_API extern "C" __attribute__((visibility("default")))

EMU_API int emu_fused_step(const EmuLaunch* launch) {
  if (launch->f32) return launch->lattice == kD2Q9 ? runLattice<Lattice<kD2Q9>, true, float>(*launch) : -1;
  return runType<double>(*launch);
}

EMU_API int emu_entropic_shared_bytes(int Q) { return entropicSharedBytes(Q, logTableInShared(Q)); }
EMU_API int emu_step_block(void) { return kStepBlock; }
EMU_API int emu_observable_slots(void) { return kObservableSlots; }
