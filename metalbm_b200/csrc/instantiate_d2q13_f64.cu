// instantiate_d2q13.cu -- explicit instantiations of the fused step kernel for D2Q13 / f64.
#include "step_kernel.cuh"

namespace mlbm {

template <int COLLISION, int EQ, int SCHEME>
static StepKernel pick() {
  return fusedStepKernel<Lattice<kD2Q13>, COLLISION, EQ, SCHEME, double>;
}

template <int COLLISION, int EQ>
static StepKernel pickScheme(int scheme) {
  switch (scheme) {
    case kSchemeNone: return pick<COLLISION, EQ, kSchemeNone>();
    case kSchemeGuo: return pick<COLLISION, EQ, kSchemeGuo>();
    case kSchemeEDM: return pick<COLLISION, EQ, kSchemeEDM>();
    default: return nullptr;
  }
}

template <int COLLISION>
static StepKernel pickEquilibrium(int equilibrium, int scheme) {
  if (equilibrium == kTruncationMa3) return pickScheme<COLLISION, kTruncationMa3>(scheme);
  return nullptr;
}

StepKernel lookupStepKernel_d2q13_f64(int collision, int equilibrium, int scheme) {
  if (collision == kBGK) return pickEquilibrium<kBGK>(equilibrium, scheme);
  if (collision == kELBM) return pickEquilibrium<kELBM>(equilibrium, scheme);
  if (collision == kELBMForcing) return pickEquilibrium<kELBMForcing>(equilibrium, scheme);
  return nullptr;
}

}  // namespace mlbm
