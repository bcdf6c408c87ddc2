// metaLBM/Collision.h (B200 drop-in) -- the physics selectors of the reference as template names:
//   Collision<T, CollisionType, Architecture>       Collision.h:20-909   (BGK :103-180, ELBM :182-376, ForcedNR_ELBM :684-724)
//   Equilibrium<T, LatticeType, EquilibriumType>    Equilibrium.h:14-126
//   ForcingScheme<T, ForcingSchemeType>             ForcingScheme.h:17-198
//   Force<T, ForceType, Architecture>               Force.h:23-292
//   Moment<T>                                        Moment.h:14-47
//   Boundary<T, BoundaryType::Periodic, ...>         Boundary.h:45-102
// In the reference these are per-node functors that the generic kernel calls; in the B200 build the per-node work
// is one fused CUDA kernel specialised on the same choices (metalbm_b200/csrc/step_kernel.cuh), so on the host side
// each name is a descriptor: it says which kernel specialisation the C-ABI selects (`abi`) and carries the
// parameters the reference constructor takes.  The `_` aliases follow Collision.h:913-914 / Force.h / Moment.h.
#pragma once

#include "Context.h"

namespace lbm {

template <class T, LatticeType latticeType, EquilibriumType equilibriumType>
class Equilibrium {
 public:
  static constexpr int abi = b200::abiEquilibrium(equilibriumType);
  static_assert(abi >= 0, "metalbm_b200: unsupported EquilibriumType");
  static_assert(equilibriumType != EquilibriumType::Exact || latticeType == LatticeType::D2Q9 || latticeType == LatticeType::D3Q27,
                "the exact equilibrium exists for D2Q9 and D3Q27 (Equilibrium.h:36-126)");
};
typedef Equilibrium<dataT, latticeT, equilibriumT> Equilibrium_;

template <class T, ForcingSchemeType forcingSchemeType>
class ForcingScheme {
 public:
  static constexpr int abi = b200::abiScheme(forcingSchemeType);
  static_assert(abi >= 0, "metalbm_b200: unsupported ForcingSchemeType");
  const T tau;  // Guo's prefactor uses the INPUT relaxation time, also under ELBM (ForcingScheme.h:115, Collision.h:44)
  ForcingScheme(const T& tau_in) : tau(tau_in) {}
};
typedef ForcingScheme<dataT, forcingSchemeT> ForcingScheme_;

template <class T, ForceType forceType, Architecture architecture>
class Force {
 public:
  static constexpr int abi = b200::abiForce(forceType);
  static_assert(abi >= 0, "metalbm_b200: unsupported ForceType");
  const MathVector<T, 3> amplitude, waveLength;
  Force(const MathVector<T, 3>& amplitude_in, const MathVector<T, 3>& waveLength_in, const unsigned int = 0, const unsigned int = 0)
      : amplitude(amplitude_in), waveLength(waveLength_in) {}
  // time-independent forces: nothing to do per iteration (Force.h:51-54)
  void update(const unsigned int, const unsigned int) {}
};
template <Architecture architecture> using Force_ = Force<dataT, forceT, architecture>;

template <class T>
class Moment {};
typedef Moment<dataT> Moment_;

template <class T, CollisionType collisionType, Architecture architecture>
class Collision {
 public:
  static constexpr int abi = b200::abiCollision(collisionType);
  static_assert(abi >= 0, "metalbm_b200: unsupported CollisionType");
  const T tau;
  Force_<architecture> force;
  ForcingScheme_ forcingScheme;

  template <class FieldList_>
  Collision(const T tau_in, FieldList_&, const MathVector<T, 3>& amplitude_in, const MathVector<T, 3>& waveLength_in,
            const unsigned int kMin_in = 0, const unsigned int kMax_in = 0)
      : tau(tau_in), force(amplitude_in, waveLength_in, kMin_in, kMax_in), forcingScheme(tau_in) {}

  void update(const unsigned int iteration, const unsigned int numberElements) { force.update(iteration, numberElements); }
};
template <Architecture architecture> using Collision_ = Collision<dataT, collisionT, architecture>;


// Periodic boundaries are index arithmetic inside the fused kernel (no y/z halo cells, no boundary launches);
// the name is kept for code that spells the reference's type (Boundary.h:45-102).
template <class T, BoundaryType boundaryType, AlgorithmType algorithmType>
class Boundary {
  static_assert(boundaryType == BoundaryType::Periodic || boundaryType == BoundaryType::Generic,
                "metalbm_b200: only periodic boundaries are on the path (the other types are empty stubs in the reference, Boundary.h:373-391)");
};

}  // namespace lbm
