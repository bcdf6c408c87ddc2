// metaLBM/DynamicArray.cuh (B200 drop-in): the reference keeps its CUDA specialisations in .cuh twins (src/main.cu:2-3);
// here the device side lives behind the C-ABI, so the .cuh name simply forwards.
#pragma once
#include "DynamicArray.h"
