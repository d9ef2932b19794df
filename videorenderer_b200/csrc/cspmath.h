// cspmath.h -- host colour math of the frame path: YUV->RGB transform (3x3 + offset) and
// primaries conversion.  Same results, bit for bit, as the reference's mpv-derived csputils
// (mp_get_csp_matrix Source/csputils.cpp:392-509, GetColorspaceGamutConversionMatrix :549-557),
// as driven by CDX11VideoProcessor::SetShaderConvertColorParams (Source/DX11VideoProcessor.cpp:813-887).
#pragma once
#include "fmt_table.h"

namespace b200vr {

struct ColorTransform {        // PS_COLOR_TRANSFORM, Source/Shaders.h:25-30
    float m[3][3];             // rows cm_r, cm_g, cm_b
    float c[3];                // cm_c
    bool  enable;              // m_PSConvColorData.bEnable, Source/DX11VideoProcessor.cpp:849-853
};

struct ProcAmp { float brightness = 0, contrast = 1, hue = 0, saturation = 1; };   // DXVA2FixedToFloat values

enum Chromaticities { CHR_BT601_525, CHR_BT601_625, CHR_BT709, CHR_BT2020, CHR_BT470M, CHR_DCI_P3 };

bool BuildColorTransform(const ExtendedFormat& srcExFmt, const FmtConvParams& fmt, const ProcAmp& amp, ColorTransform& out);
void BuildGamutMatrix(Chromaticities from, Chromaticities to, float out[3][3]);

}  // namespace b200vr
