// metaLBM/Stream.h (B200 drop-in) -- `Stream<Architecture>` (Stream.h:9-30, Stream.cuh:9-30).  The CUDA streams
// of the B200 path (one compute stream, one high-priority communication stream, joined by events) are owned by
// the mlbm_ctx; a Stream<Architecture::GPU> is a handle whose synchronize() drains them.
#pragma once

#include "Commons.h"
#include "Options.h"

namespace lbm {

namespace b200 { inline void synchronizeContextIfAny(); }

template <Architecture architecture>
class Stream {};

template <>
class Stream<Architecture::CPU> {
 public:
  Stream(bool isDefault_in = true) { (void)isDefault_in; }
  void synchronize() {}
};

template <>
class Stream<Architecture::GPU> {
 public:
  Stream(bool isDefault_in = true) { (void)isDefault_in; }
  void synchronize() { b200::synchronizeContextIfAny(); }
};

}  // namespace lbm
