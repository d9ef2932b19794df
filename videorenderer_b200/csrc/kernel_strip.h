// kernel_strip.h -- launcher of the persistent warp-specialised 2x kernel (kernel_strip.cuh); used by LaunchFused.
#pragma once
#include <cuda_runtime.h>
#include "kernels.h"

namespace b200vr {

// bi-planar source, bilinear chroma, 16-byte aligned planes / pitches (TMA), 4 or 6 taps
bool        StripEligible(const FusedPlan& p);
// gT / gD / gG / exc: device threshold table of the gamma + UNORM10 quantiser, as separate arrays and as 16-byte entries
// (nullptr: evaluate the pow directly)
struct GammaEntry;
cudaError_t LaunchStrip2x(const FusedPlan& p, const float* gT, const float* gD, const GammaEntry* gG, const uint32_t exc[2], cudaStream_t st);
// debug: enabled = 0 keeps every 2x plan on the tile kernel (k_fused2x); cfg selects the warp-role split of the headline variant
void        SetStripKernel(int enabled, int cfg);
// debug: the kernel's watchdog record on the current device (all zero = every mbarrier wait completed); returns 0 / -1
int         StripStatus(unsigned int out[8], int clear);

}  // namespace b200vr
