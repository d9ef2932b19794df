// dither32f.cuh -- the 32x32 ordered-dither table as floats d = k/1024 (exact in fp32), a compile-time constant in device
// memory: no init kernel, no ordering between streams or instances (Source/DX11VideoProcessor.cpp:1414-1440 uploads the
// same values once as a texture; Shaders/d3d11/ps_final_pass.hlsl:23-31 samples it).
#pragma once
#include <stdint.h>
#include "dither32.h"

namespace b200vr {

struct DitherTableF { float v[1024]; };

constexpr DitherTableF MakeDitherTableF(float scale)
{
    constexpr uint16_t k[1024] = { MPCVR_DITHER32_VALUES };
    DitherTableF t{};
    for (int i = 0; i < 1024; i++) t.v[i] = (float)k[i] * scale;
    return t;
}

}  // namespace b200vr
