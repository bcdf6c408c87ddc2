// metaLBM/Event.h (B200 drop-in) -- `Event<Architecture>` (Event.h:9-37, Event.cuh:9-26).  The boundary-done /
// exchange-done hand-offs that Algorithm.h:418-432 sketches with leftEvent / rightEvent happen inside mlbm_step
// on the context's own CUDA events; these objects keep the iterate() signature intact.
#pragma once

#include "Stream.h"

namespace lbm {

template <Architecture architecture>
class Event {
 public:
  Event() {}
  void synchronize() {}
  void record(Stream<architecture>&) {}
  void wait(Stream<architecture>&) {}
};

}  // namespace lbm
