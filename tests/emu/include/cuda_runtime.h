// tests/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY: what `#include <cuda_runtime.h>` resolves to when the
// kernel source is compiled for the host emulator (tests/emu/cuda_emu.h); with MLBM_EMU_HOST also the host side of the
// runtime API (tests/emu/include/cuda_host_emu.h), for the emulated build of the whole library.
#pragma once
#include "../cuda_emu.h"
#ifdef MLBM_EMU_HOST
#include "cuda_host_emu.h"
#endif
