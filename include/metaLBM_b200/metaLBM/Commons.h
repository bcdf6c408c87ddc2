// metaLBM/Commons.h (B200 drop-in) -- same macro spellings as the reference's Commons.h:6-103.
// The host side of this layer is plain C++; everything that runs on the device lives behind the C-ABI
// (include/metalbm_b200.h, libmetalbm_b200.so), so the LBM_* decorators expand to nothing here.
#pragma once

#include <cstdio>
#include <cstdlib>

#include "../../metalbm_b200.h"

#define LBM_HOST
#define LBM_DEVICE
#define LBM_SHARED
#define LBM_CONSTANT
#define LBM_GLOBAL
#define LBM_INLINE inline
#define LBM_INSTRUMENT_ON(name, colorID)
#define LBM_INSTRUMENT_OFF(name, colorID)

// The reference prints "[file:line] ... failed" and exit(-1)s on any CUDA / MPI error (Commons.h:6-28);
// the C-ABI returns a status instead, and this macro re-applies the reference's convention.
#define LBM_B200_CALL(call)                                                                        \
  do {                                                                                             \
    const int lbm_b200_status_ = (call);                                                           \
    if (lbm_b200_status_ != 0) {                                                                   \
      std::fprintf(stderr, "[%s:%d] metalbm_b200 failed with %d: %s\n", __FILE__, __LINE__,        \
                   lbm_b200_status_, mlbm_last_error());                                           \
      std::exit(-1);                                                                               \
    }                                                                                              \
  } while (0)
#define LBM_CUDA_CALL(call) LBM_B200_CALL(call)
#define LBM_MPI_CALL(call) (call)
