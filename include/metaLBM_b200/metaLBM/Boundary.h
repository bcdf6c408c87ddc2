// metaLBM/Boundary.h (B200 drop-in): see Collision.h, which holds all per-node physics descriptors of the fused kernel.
#pragma once
#include "Collision.h"
