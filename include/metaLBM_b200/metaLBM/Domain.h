// metaLBM/Domain.h (B200 drop-in) -- the index spaces of the reference (Domain.h:10-351) under their alias
// names: lSD (local, FFTW-padded), gSD (global), hSD (halo).  z is the fastest index, x the slowest;
// SoA component stride for local/global fields is the padded volume, for halo space the halo volume.
#pragma once

#include <cstddef>

#include "Lattice.h"
#include "MathVector.h"
#include "Options.h"

namespace lbm {

constexpr int globalLengthInt[3] = {globalLengthX, L::dimD > 1 ? globalLengthY : 1, L::dimD > 2 ? globalLengthZ : 1};
constexpr unsigned int globalLengthUInt[3] = {globalLengthX, L::dimD > 1 ? globalLengthY : 1, L::dimD > 2 ? globalLengthZ : 1};
constexpr ptrdiff_t globalLengthPtrdiff_t[3] = {globalLengthX, L::dimD > 1 ? globalLengthY : 1, L::dimD > 2 ? globalLengthZ : 1};

constexpr Position localLength = {{(unsigned int)(globalLengthX / numProcs), L::dimD > 1 ? (unsigned int)globalLengthY : 1u,
                                   L::dimD > 2 ? (unsigned int)globalLengthZ : 1u}};

template <DomainType domainType, PartitionningType partitionningType, MemoryLayout memoryLayout, unsigned int NumberComponents>
struct Domain {};

// lSD: the local padded space (Domain.h:42-92).  The last used dimension is padded to 2 (N/2 + 1) for the in-place
// real FFT of the reference's analysis layer (ProjectPadRealAndLeave1, MathVector.h:330-344).
template <unsigned int NumberComponents>
struct Domain<DomainType::LocalSpace, PartitionningType::Generic, MemoryLayout::Generic, NumberComponents> {
  static constexpr Position pStart() { return Position{{0, 0, 0}}; }
  static constexpr Position pEnd() {
    return Position{{L::dimD == 1 ? 2 * (localLength[d::X] / 2 + 1) : localLength[d::X],
                     L::dimD == 2 ? 2 * (localLength[d::Y] / 2 + 1) : localLength[d::Y],
                     L::dimD == 3 ? 2 * (localLength[d::Z] / 2 + 1) : localLength[d::Z]}};
  }
  static constexpr Position pLength() { return pEnd(); }
  static constexpr unsigned int pVolume() { return pLength()[d::X] * pLength()[d::Y] * pLength()[d::Z]; }
  static constexpr Position sStart() { return Position{{0, 0, 0}}; }
  static constexpr Position sEnd() { return localLength; }
  static constexpr Position sLength() { return sEnd(); }
  static constexpr unsigned int sVolume() { return sLength()[d::X] * sLength()[d::Y] * sLength()[d::Z]; }
  static constexpr unsigned int getIndex(const Position& iP) {
    return pLength()[d::Z] * (pLength()[d::Y] * iP[d::X] + iP[d::Y]) + iP[d::Z];
  }
  static constexpr unsigned int getIndex(const Position& iP, const unsigned int iC) { return iC * pVolume() + getIndex(iP); }
};

// gSD: the global space (Domain.h:94-171)
template <PartitionningType partitionningType, unsigned int NumberComponents>
struct Domain<DomainType::GlobalSpace, partitionningType, MemoryLayout::Generic, NumberComponents>
    : public Domain<DomainType::LocalSpace, PartitionningType::Generic, MemoryLayout::Generic, NumberComponents> {
  static constexpr Position sStart() { return Position{{0, 0, 0}}; }
  static constexpr Position sEnd() { return Position{{globalLengthUInt[0], globalLengthUInt[1], globalLengthUInt[2]}}; }
  static constexpr Position sLength() { return sEnd(); }
  static constexpr unsigned int sVolume() { return sLength()[d::X] * sLength()[d::Y] * sLength()[d::Z]; }
  // offset of the slab of `rank` in the global space (Domain.h:155-162)
  static Position sOffset(const MathVector<int, 3>& rank) {
    return Position{{(unsigned int)rank[d::X] * localLength[d::X], 0u, 0u}};
  }
};

// hSD: the halo space the reference streams in (Domain.h:173-283).  The B200 layout keeps x halo planes only
// (periodic images in y/z are reached by index arithmetic), see mlbm_device_layout; these functions describe the
// REFERENCE halo space so that code computing sizes / indices with them keeps compiling.
template <unsigned int NumberComponents>
struct Domain<DomainType::HaloSpace, PartitionningType::Generic, MemoryLayout::SoA, NumberComponents>
    : public Domain<DomainType::LocalSpace, PartitionningType::Generic, MemoryLayout::Generic, NumberComponents> {
  static constexpr Position start() { return Position{{0, 0, 0}}; }
  static constexpr Position end() {
    return Position{{localLength[d::X] + 2 * L::halo()[d::X], localLength[d::Y] + 2 * L::halo()[d::Y],
                     localLength[d::Z] + 2 * L::halo()[d::Z]}};
  }
  static constexpr Position length() { return end(); }
  static constexpr unsigned int volume() { return length()[d::X] * length()[d::Y] * length()[d::Z]; }
  static constexpr unsigned int getIndex(const Position& iP) {
    return length()[d::Z] * (length()[d::Y] * iP[d::X] + iP[d::Y]) + iP[d::Z];
  }
  static constexpr unsigned int getIndex(const Position& iP, const unsigned int iC) { return iC * volume() + getIndex(iP); }
  static constexpr unsigned int getIndexLocal(const Position& iP) {
    using Local = Domain<DomainType::LocalSpace, PartitionningType::Generic, MemoryLayout::Generic, NumberComponents>;
    return Local::getIndex(Position{{iP[d::X] - L::halo()[d::X], iP[d::Y] - L::halo()[d::Y], iP[d::Z] - L::halo()[d::Z]}});
  }
};

using gSD = Domain<DomainType::GlobalSpace, PartitionningType::Generic, MemoryLayout::Generic, 1>;
using lSD = Domain<DomainType::LocalSpace, PartitionningType::Generic, MemoryLayout::Generic, 1>;
using hSD = Domain<DomainType::HaloSpace, PartitionningType::Generic, memoryL, L::dimQ>;

}  // namespace lbm
