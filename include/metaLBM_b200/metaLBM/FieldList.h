// metaLBM/FieldList.h (B200 drop-in) -- `FieldList<T, architecture>` (FieldList.h:22-87) with the fields of the step
// path: density, velocity, force, alpha (+ vorticity for the analysis layer).  The kinetic-diagnostic fields
// (T2..fNonEq8, `writeKinetics`) are research output outside the hot path.
#pragma once

#include "Field.h"
#include "Initialize.h"

namespace lbm {

template <class T, Architecture architecture>
class FieldList {
 public:
  Field<T, 1, architecture, true> density;
  Field<T, L::dimD, architecture, true> velocity;
  Field<T, L::dimD, architecture, writeForce> force;
  Field<T, 1, architecture, writeAlpha> alpha;
  Field<T, 2 * L::dimD - 3, architecture, writeVorticity> vorticity;

  FieldList(const Stream<architecture>& stream_in)
      : density(initDensity<T, architecture>(stream_in)),
        velocity(initVelocity<T, architecture>(stream_in)),
        force(initForce<T, architecture>(stream_in)),
        alpha(initAlpha<T, architecture>(stream_in)),
        vorticity("vorticity") {}

  // the reference signature FieldList(FieldWriter_&, const Stream&) (FieldList.h:44-45); the writer is the caller's
  template <class FieldWriter>
  FieldList(FieldWriter&, const Stream<architecture>& stream_in) : FieldList(stream_in) {}
};

}  // namespace lbm
