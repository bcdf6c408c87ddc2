// instantiate_d2q21.cu -- explicit instantiations of the fused step kernel for D2Q21 / f32.
#include "step_kernel.cuh"

namespace mlbm {

template <int COLLISION, int EQ, int SCHEME>
static StepKernel pick() {
  return fusedStepKernel<Lattice<kD2Q21>, COLLISION, EQ, SCHEME, float>;
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

StepKernel lookupStepKernel_d2q21_f32(int collision, int equilibrium, int scheme) {
  if (collision == kBGK) return pickEquilibrium<kBGK>(equilibrium, scheme);
  if (collision == kELBM) return pickEquilibrium<kELBM>(equilibrium, scheme);
  if (collision == kELBMForcing) return pickEquilibrium<kELBMForcing>(equilibrium, scheme);
  return nullptr;
}

}  // namespace mlbm
