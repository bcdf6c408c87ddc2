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
// The per-node members host code can call in the reference (LBM_HOST) are kept callable on the HOST, over a host copy
// of the halo-space distribution (Distribution::getHaloDataPreviousHost): Moment<T>::calculateDensity / calculateVelocity
// (Moment.h:14-47), Force::setForce (Force.h:39-48, 154-159, 208-215, 262-267), ForcingScheme::calculateHydrodynamicVelocity
// (ForcingScheme.h:26-33, 50-57) and Collision::calculateMoments / setForce / getDensity / getVelocity / getForce /
// getHydrodynamicVelocity (Collision.h:60-93).  They are inspection tools (post-processing, tests): the step itself never
// runs on the host.
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
  // the velocity the fields store: u + F / (2 rho) (ForcingScheme.h:26-33); the scheme None stores u (:50-57)
  MathVector<T, L::dimD> calculateHydrodynamicVelocity(const MathVector<T, L::dimD>& force, const T& density,
                                                       const MathVector<T, L::dimD>& velocity) const {
    if (forcingSchemeType == ForcingSchemeType::None) return velocity;
    MathVector<T, L::dimD> shifted = velocity;
    const T half = (T)0.5 / density;
    for (int iD = 0; iD < L::dimD; ++iD) shifted[iD] += half * force[iD];
    return shifted;
  }
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
  // the force at LOCAL interior position iP (the caller passes iP - L::halo(), Collision.h:86): the analytic profiles of
  // Force.h:154-159 (Constant), :208-215 (Sinusoidal), :262-267 (Kolmogorov: x component only, profile along y); every
  // other type reads component iD of the force array (Force.h:39-48)
  void setForce(const T* forcePtr, const Position& iP, MathVector<T, L::dimD>& force, const unsigned int numberElements) const {
    for (int iD = 0; iD < L::dimD; ++iD) force[iD] = (T)0;
    if (forceType == ForceType::None) return;
    if (forceType == ForceType::Constant) {
      for (int iD = 0; iD < L::dimD; ++iD) force[iD] = amplitude[iD];
    } else if (forceType == ForceType::Sinusoidal) {
      for (int iD = 0; iD < L::dimD; ++iD) force[iD] = amplitude[iD] * std::sin(iP[iD] * 2 * M_PI / waveLength[iD]);
    } else if (forceType == ForceType::Kolmogorov) {
      force[d::X] = amplitude[d::X] * std::sin(iP[d::Y] * 2 * M_PI / waveLength[d::X]);
    } else if (forcePtr) {
      for (int iD = 0; iD < L::dimD; ++iD) force[iD] = forcePtr[(size_t)iD * numberElements + lSD::getIndex(iP)];
    }
  }
};
template <Architecture architecture> using Force_ = Force<dataT, forceT, architecture>;

// Moment.h:14-47 on a HOST array in the reference's halo space (hSD, Domain.h:173-283; Distribution::getHaloDataPreviousHost):
// the moments of the populations PULLED to iP, f_q(iP - c_q), summed in increasing iQ like the reference; iP is a halo-space
// position (interior node + L::halo())
template <class T>
class Moment {
  static unsigned int upstream(const Position& iP, const int iQ) {
    Position from = iP;
    for (int iD = 0; iD < L::dimD; ++iD) from[iD] = (unsigned int)((int)iP[iD] - (int)L::celerity()[iQ][iD]);
    return hSD::getIndex(from, (unsigned int)iQ);
  }

 public:
  static void calculateDensity(const T* haloDistributionPtr, const Position& iP, T& density) {
    T sum = haloDistributionPtr[upstream(iP, 0)];
    for (int iQ = 1; iQ < L::dimQ; ++iQ) sum += haloDistributionPtr[upstream(iP, iQ)];
    density = sum;
  }
  static void calculateVelocity(const T* haloDistributionPtr, const Position& iP, const T density, MathVector<T, L::dimD>& velocity) {
    for (int iD = 0; iD < L::dimD; ++iD) velocity[iD] = L::celerity()[0][iD] * haloDistributionPtr[upstream(iP, 0)];
    for (int iQ = 1; iQ < L::dimQ; ++iQ) {
      const T population = haloDistributionPtr[upstream(iP, iQ)];
      for (int iD = 0; iD < L::dimD; ++iD) velocity[iD] += L::celerity()[iQ][iD] * population;
    }
    velocity /= density;
  }
};
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

  // ---- the per-node state of Collision<T, GenericSRT> (Collision.h:28-33, 60-93), on the host ----
  const T& getDensity() { return density; }
  const MathVector<T, L::dimD>& getVelocity() { return velocity; }
  const MathVector<T, L::dimD>& getForce() { return nodeForce; }
  // Collision.h:72-79: density, velocity and velocity^2 of the populations pulled to the halo-space position iP
  void calculateMoments(const T* haloDistributionPreviousPtr, const Position& iP) {
    Moment_::calculateDensity(haloDistributionPreviousPtr, iP, density);
    Moment_::calculateVelocity(haloDistributionPreviousPtr, iP, density, velocity);
    velocity2 = velocity.norm2();
  }
  // Collision.h:81-88: the force at the node, evaluated at the LOCAL position iP - L::halo()
  void setForce(const T* forcePtr, const Position& iP, const Position&, const unsigned int numberElements) {
    Position local = iP;
    for (int iD = 0; iD < 3; ++iD) local[iD] = iP[iD] - L::halo()[iD];
    force.setForce(forcePtr, local, nodeForce, numberElements);
  }
  // Collision.h:90-93
  const MathVector<T, L::dimD> getHydrodynamicVelocity() { return forcingScheme.calculateHydrodynamicVelocity(nodeForce, density, velocity); }

 private:
  T density = (T)0, velocity2 = (T)0;
  MathVector<T, L::dimD> velocity = {}, nodeForce = {};
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
