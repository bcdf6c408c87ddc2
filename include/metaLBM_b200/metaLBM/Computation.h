// metaLBM/Computation.h (B200 drop-in) -- the reference's executor applies a functor to every Position of a
// box (Computation.h:10-127 CPU loops, Computation.cuh:12-141 generic kernels).  On the B200 path the per-node
// physics is not a host functor any more (it is the fused CUDA kernel behind mlbm_step), so only the CPU
// executor is kept, for host-side loops such as field initialisation and the scalar analyses.
#pragma once

#include "Commons.h"
#include "MathVector.h"
#include "Options.h"
#include "Stream.h"

namespace lbm {

template <Architecture architecture, unsigned int Dimension>
class Computation {
 protected:
  const Position start, end;

 public:
  Computation(const Position& start_in, const Position& end_in, const MathVector<unsigned int, 3>& = {{0, 1, 2}})
      : start(start_in), end(end_in) {}

  template <typename Callback, typename... Arguments>
  void Do(const Stream<architecture>&, Callback function, const Arguments... arguments) { Do(function, arguments...); }

  template <typename Callback, typename... Arguments>
  void Do(Callback function, const Arguments... arguments) {
    Position iP = {{0, 0, 0}};
    for (unsigned int x = start[d::X]; x < (Dimension > 0 ? end[d::X] : start[d::X] + 1); ++x)
      for (unsigned int y = start[d::Y]; y < (Dimension > 1 ? end[d::Y] : start[d::Y] + 1); ++y)
        for (unsigned int z = start[d::Z]; z < (Dimension > 2 ? end[d::Z] : start[d::Z] + 1); ++z) {
          iP[d::X] = x; iP[d::Y] = y; iP[d::Z] = z;
          function(iP, arguments...);
        }
  }

  void synchronize() {}
};

}  // namespace lbm
