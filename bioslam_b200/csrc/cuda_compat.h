// Build switch: real CUDA (nvcc, sm_100a) is the product; -DB200_USE_EMU (tests/cudaemu, g++) is a CPU debugging
// harness for the same sources and is never shipped nor loaded by the package.
#pragma once
#ifdef B200_USE_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define B200_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char b200_dyn_smem_[]; \
    type* name = reinterpret_cast<type*>(b200_dyn_smem_)
#endif
