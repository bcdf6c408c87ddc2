// tests/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY: what `#include <cuda_runtime.h>` resolves to when the
// kernel source is compiled for the host emulator (tests/emu/cuda_emu.h).
#pragma once
#include "../cuda_emu.h"
