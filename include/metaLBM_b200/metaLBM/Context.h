// metaLBM/Context.h (B200 drop-in, no reference counterpart) -- the process-wide mlbm_ctx behind the template API.
//
// Every physics choice of the reference is a compile-time global of Input.in (latticeT, collisionT, ...; read by the
// headers as global names, e.g. Lattice.h:804-806, Collision.h:913-914) and one process drives one GPU
// (CUDAInitializer.h:23-26), so exactly one context per process is the faithful mapping.  It is created on first
// use from those globals and MPIInit::rank, and destroyed at exit.
#pragma once

#include <cstdlib>
#include <vector>

#include "Commons.h"
#include "Domain.h"
#include "FFTWInitializer.h"
#include "Lattice.h"
#include "MPIInitializer.h"
#include "Options.h"

namespace lbm {
namespace b200 {

constexpr int abiCollision(CollisionType c) {
  return c == CollisionType::BGK ? (int)MLBM_BGK : c == CollisionType::ELBM ? (int)MLBM_ELBM
       : c == CollisionType::ForcedNR_ELBM ? (int)MLBM_FORCED_NR_ELBM
       : c == CollisionType::Approached_ELBM ? (int)MLBM_APPROACHED_ELBM
       : c == CollisionType::Malaspinas_ELBM ? (int)MLBM_MALASPINAS_ELBM
       : c == CollisionType::Essentially1_ELBM ? (int)MLBM_ESSENTIALLY1_ELBM
       : c == CollisionType::Essentially2_ELBM ? (int)MLBM_ESSENTIALLY2_ELBM
       : c == CollisionType::ForcedBNR_ELBM ? (int)MLBM_FORCED_BNR_ELBM
       : c == CollisionType::ForcedNR_ELBM_Forcing ? (int)MLBM_FORCED_NR_ELBM_FORCING : -1;
}
constexpr int abiEquilibrium(EquilibriumType e) {
  return e == EquilibriumType::TruncationMa3 ? (int)MLBM_TRUNCATION_MA3 : e == EquilibriumType::Exact ? (int)MLBM_EXACT : -1;
}
constexpr int abiScheme(ForcingSchemeType s) {
  return s == ForcingSchemeType::None ? (int)MLBM_SCHEME_NONE : s == ForcingSchemeType::Guo ? (int)MLBM_GUO
       : s == ForcingSchemeType::ShanChen ? (int)MLBM_SHAN_CHEN
       : s == ForcingSchemeType::ExactDifferenceMethod ? (int)MLBM_EXACT_DIFFERENCE : -1;
}
constexpr int abiForce(ForceType f) {
  return f == ForceType::None ? (int)MLBM_FORCE_NONE : f == ForceType::Constant ? (int)MLBM_FORCE_CONSTANT
       : f == ForceType::Sinusoidal ? (int)MLBM_FORCE_SINUSOIDAL : f == ForceType::Kolmogorov ? (int)MLBM_FORCE_KOLMOGOROV
       : (f == ForceType::ConstantShell && L::dimD == 2) ? (int)MLBM_FORCE_CONSTANT_SHELL
       : (f == ForceType::EnergyRemoval && L::dimD == 2) ? (int)MLBM_FORCE_ENERGY_REMOVAL
       : (f == ForceType::Turbulent2D && L::dimD == 2) ? (int)MLBM_FORCE_TURBULENT_2D : -1;
}

static_assert(abiCollision(collisionT) >= 0, "metalbm_b200: collisionT must be BGK, ELBM or one of the entropic variants that behave like ELBM in the reference (Approached_, Malaspinas_, Essentially1_, Essentially2_, ForcedNR_, ForcedBNR_ELBM) or ForcedNR_ELBM_Forcing");
static_assert(abiEquilibrium(equilibriumT) >= 0, "metalbm_b200: equilibriumT must be TruncationMa3 or Exact");
static_assert(equilibriumT != EquilibriumType::Exact || latticeT == LatticeType::D2Q9 || latticeT == LatticeType::D3Q27,
              "metalbm_b200: the exact equilibrium exists for D2Q9 and D3Q27 only (Equilibrium.h:36-126)");
static_assert(abiScheme(forcingSchemeT) >= 0, "metalbm_b200: forcingSchemeT must be None, Guo, ShanChen or ExactDifferenceMethod");
static_assert(abiForce(forceT) >= 0, "metalbm_b200: forceT must be None, Constant, Sinusoidal, Kolmogorov or, on 2-D lattices, ConstantShell, EnergyRemoval, Turbulent2D (their 3-D variants go through the C-ABI's MLBM_FORCE_FIELD)");
static_assert(algorithmT == AlgorithmType::Pull && memoryL == MemoryLayout::SoA && partitionningT == PartitionningType::OneD,
              "metalbm_b200 implements the Pull / SoA / OneD step (Algorithm.h:300-452)");
static_assert(sizeof(dataT) == 8 || sizeof(dataT) == 4, "dataT must be double or float");

class Context {
  mlbm_ctx* handle = nullptr;
  Context() {
    mlbm_config config = {};
    config.abi_version = MLBM_ABI_VERSION;
    config.lattice = L::abi;
    config.collision = abiCollision(collisionT);
    config.equilibrium = abiEquilibrium(equilibriumT);
    config.forcing_scheme = abiScheme(forcingSchemeT);
    config.force = abiForce(forceT);
    config.dtype = sizeof(dataT) == 8 ? MLBM_F64 : MLBM_F32;
    config.overlap = overlappingT == Overlapping::On ? MLBM_OVERLAP_ON : MLBM_OVERLAP_OFF;
    for (int iD = 0; iD < 3; ++iD) {
      config.global_length[iD] = globalLengthInt[iD];
      config.force_amplitude[iD] = (double)forceAmplitude[iD];
      config.force_wavelength[iD] = (double)forceWaveLength[iD];
    }
    config.rank = MPIInit::rank[d::X];
    config.nranks = numProcs;
    config.device = -1;  // rank % device count, CUDAInitializer.h:23-26
    config.tau = (double)relaxationTime;
    config.force_k_min = (int)forcekMin;  // Algorithm.h:90-91
    config.force_k_max = (int)forcekMax;
    config.removal_k_min = (int)removalForcekMin;  // Force.h:586-587
    config.removal_k_max = (int)removalForcekMax;
    for (int iD = 0; iD < 3; ++iD) config.removal_amplitude[iD] = (double)removalForceAmplitude[iD];
    LBM_B200_CALL(mlbm_create(&config, &handle));
    if (numProcs > 1) {
      unsigned char id[128] = {0};
      if (MPIInit::rank[d::X] == 0) LBM_B200_CALL(mlbm_comm_unique_id(id));
      MPIInit::broadcastFromRoot(id, sizeof(id));
      LBM_B200_CALL(mlbm_comm_init(handle, id));
      // Overlapping::On: the boundary kernel stores into the neighbours' halo planes over NVLink (CUDA IPC mappings);
      // MLBM_PEER_HALOS=0 keeps the NCCL send/recv exchange
      const char* peerHalos = std::getenv("MLBM_PEER_HALOS");
      // (the multi-speed lattices carry dimH > 1 halo planes per side: they keep the overlapped NCCL exchange,
      // mlbm_comm_peer_attach is built for one plane)
      if (overlappingT == Overlapping::On && L::dimH == 1 && !(peerHalos && peerHalos[0] == '0')) {
        unsigned char mine[MLBM_PEER_HANDLE_BYTES];
        std::vector<unsigned char> all((size_t)MLBM_PEER_HANDLE_BYTES * numProcs);
        LBM_B200_CALL(mlbm_comm_peer_export(handle, mine));
        MPIInit::allGather(mine, all.data(), sizeof(mine), id, sizeof(id));
        LBM_B200_CALL(mlbm_comm_peer_attach(handle, all.data() + (size_t)MLBM_PEER_HANDLE_BYTES * MPIInit::rankLeft,
                                            all.data() + (size_t)MLBM_PEER_HANDLE_BYTES * MPIInit::rankRight));
      }
    }
  }
  ~Context() { mlbm_destroy(handle); }
  static bool& alive() { static bool flag = false; return flag; }

 public:
  Context(const Context&) = delete;
  static mlbm_ctx* get() {
    static Context instance;
    alive() = true;
    return instance.handle;
  }
  static bool exists() { return alive(); }
};

inline void synchronizeContextIfAny() {
  if (Context::exists()) LBM_B200_CALL(mlbm_sync(Context::get()));
}

}  // namespace b200
}  // namespace lbm
