// shell_force.h -- the reference's spectral body forces on 2-D lattices (Force.h:296-616), rebuilt without a distributed
// transform: ConstantShell (a fixed solenoidal field on a shell of wave numbers), EnergyRemoval (minus the band-passed
// momentum of the last stored fields) and Turbulent2D (their sum).
#pragma once

#include <cuda_runtime.h>

#include <string>

#include "nccl_loader.h"

namespace mlbm {

struct ShellForceGeometry {
  int LX, NR;        // local slab extents (2-D: x planes of NR nodes)
  int rank, nranks;  // x-slab index and count: global NX = LX * nranks
  int globalX, globalY;
  int elementSize;   // 8 (fields stored as double) or 4 (float)
};

struct ShellForceSpec {
  bool injection;              // ConstantShell part (Force.h:296-420): psi^ = injectionAmplitude on the shell
  double injectionAmplitude;   // forceAmplitude[0]
  int injectionKMin, injectionKMax;
  bool removal;                // EnergyRemoval part (Force.h:423-561): F^_d = -removalAmplitude[d] * (rho u_d)^ on the shell
  double removalAmplitude[2];
  int removalKMin, removalKMax;
};

class ShellForce;

// nullptr and *error on failure (out of memory)
ShellForce* shellForceCreate(const ShellForceGeometry& geometry, const ShellForceSpec& spec, std::string* error);
void shellForceDestroy(ShellForce* plan);
bool shellForceIsTimeDependent(const ShellForce* plan);

// Enqueues on `stream`: force = the part that does not depend on the fields (the injection array, or zero): the state
// before any field exists (Collision.h:51-54 with fieldList at rest).
int shellForceInitial(ShellForce* plan, void* force, long long fieldStride, cudaStream_t stream, unsigned long long* launches,
                      std::string* error);

// Enqueues on `stream` the (re)computation of the force field `force` ([2] components, `fieldStride` elements apart) from
// the stored fields `density` / `velocity` -- Force::update (Force.h:552-558, 605-609).  The removal part projects the
// momentum onto the shell's modes (block partial sums, a deterministic second stage, one all-reduce over the ranks when
// nranks > 1) and synthesises the band-passed field; the injection part was synthesised once at creation.
int shellForceUpdate(ShellForce* plan, const void* density, const void* velocity, void* force, long long fieldStride,
                     const NcclApi* nccl, ncclComm_t comm, cudaStream_t stream, unsigned long long* launches, std::string* error);

}  // namespace mlbm
