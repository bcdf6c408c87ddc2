// metaLBM/Distribution.h (B200 drop-in) -- `Distribution<T, architecture>` (Distribution.h:15-43): the local padded
// SoA array of dimQ components (what initDistribution fills, Algorithm::unpack uploads, Algorithm::pack refreshes
// and the checkpoint writer dumps).  The ping-pong halo pair lives on the device inside the mlbm_ctx:
// getHaloDataPrevious() returns the DEVICE pointer of the buffer the next step reads, in the layout described by
// mlbm_device_layout -- x halo planes only, no y/z halo cells -- for callers that interoperate on the device.
#pragma once

#include <vector>

#include "Context.h"
#include "Field.h"

namespace lbm {

template <class T, Architecture architecture>
class Distribution : public Field<T, L::dimQ, architecture, true> {
 private:
  using Base = Field<T, L::dimQ, architecture, true>;

 public:
  using Base::fieldName;
  using Base::getData;

  Distribution() : Base("distribution") {}

  T* getHaloDataPrevious() {
    mlbm_device_layout layout;
    LBM_B200_CALL(mlbm_device_distribution(b200::Context::get(), &layout));
    return static_cast<T*>(layout.populations);
  }
  // The same buffer as a HOST array in the reference's halo space (hSD::getIndex(iP, iQ), Domain.h:272-276), every halo
  // cell holding what the halo exchange and the periodic boundaries deliver before the node update: what the host-callable
  // per-node functions (Moment<T>, Collision::calculateMoments) read.  Refreshed by every call; an inspection path.
  const T* getHaloDataPreviousHost() {
    haloMirror.resize((size_t)hSD::volume() * L::dimQ);
    LBM_B200_CALL(mlbm_download_halo_distribution(b200::Context::get(), haloMirror.data(), haloMirror.size()));
    return haloMirror.data();
  }
  mlbm_device_layout getHaloLayout() {
    mlbm_device_layout layout;
    LBM_B200_CALL(mlbm_device_distribution(b200::Context::get(), &layout));
    return layout;
  }
 private:
  std::vector<T> haloMirror;
};

}  // namespace lbm
