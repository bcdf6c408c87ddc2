// metaLBM/CUDAInitializer.h (B200 drop-in) -- the reference binds rank -> device `localRank % deviceCount`
// (CUDAInitializer.h:11-41).  The C-ABI applies the same rule inside mlbm_create (config.device = -1), so
// this RAII object only keeps the spelling `auto cudaLauncher = CUDAInitializer{};` of src/main.cu:16 valid.
#pragma once

namespace lbm {
struct CUDAInitializer {
  CUDAInitializer() {}
  ~CUDAInitializer() {}
};
}  // namespace lbm
