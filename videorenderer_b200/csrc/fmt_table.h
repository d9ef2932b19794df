// fmt_table.h -- input format table and extended-format defaults of the frame path.
// Mirrors the columns of the reference's s_FmtConvMapping that the shader chain depends on
// (Source/Helper.cpp:295-359; FmtConvParams_t Source/Helper.h:143-165) and the
// DXVA2_ExtendedFormat handling (SpecifyExtendedFormat Source/Helper.cpp:1169-1211).
#pragma once
#include <cstdint>
#include "../../include/mpcvr_b200.h"

namespace b200vr {

enum ColorSystem { CS_YUV = 0, CS_RGB = 1, CS_GRAY = 2 };           // ColorSystem_t, Source/Helper.h:129-133

// how the kernels address the source ("plane views" of DX11PlaneConfig_t, Source/Helper.cpp:295-307)
enum SrcLayout {
    LAYOUT_UNSUPPORTED = 0,
    LAYOUT_BIPLANAR,     // R8|R16 luma + R8G8|R16G16 interleaved chroma   (NV12, P01x, P21x)
    LAYOUT_PLANAR3,      // three R8|R16 planes                            (YV12.., YUV4xxP*)
    LAYOUT_PACKED422,    // R8G8B8A8|R16G16B16A16 macro-pixel              (YUY2, UYVY, Y210, Y216)
    LAYOUT_PACKED444,    // one texel per pixel                            (AYUV, Y410, Y416)
    LAYOUT_RGB,          // one RGB(A) pixel per packsize bytes, sample layout (RGB24 ... b64a)
    LAYOUT_GRAY,         // one R8|R16 plane                               (Y8, Y10, Y16)
};

// packed_kind of LAYOUT_RGB: the SAMPLE byte layout.  What the reference's texture would hold after its CPU repack
// (Source/Helper.cpp:446-787) is noted on the right; fetch_rgb() returns exactly those texel values.
enum RgbKind {
    RGBK_BGRX32 = 0,     // B G R X bytes                      -> B8G8R8X8 / B8G8R8A8          (CopyPlaneAsIs)
    RGBK_BGR24,          // B G R bytes                        -> B8G8R8X8                     (CopyFrameRGB24 :446-482)
    RGBK_R210,           // big-endian x2 R10 G10 B10          -> R10G10B10A2                  (CopyFrameR210 :769-787)
    RGBK_RGB48,          // R G B little-endian words          -> R16G16B16A16                 (CopyFrameRGB48 :541-565)
    RGBK_BGR48,          // B G R words                        -> R16G16B16A16                 (CopyFrameBGR48 :600-640)
    RGBK_BGRA64,         // B G R A words                      -> R16G16B16A16                 (CopyFrameBGRA64 :642-658)
    RGBK_B64A,           // big-endian A R G B words           -> R16G16B16A16                 (CopyFrameB64A :660-677)
};

struct FmtConvParams {
    mpcvr_cformat cformat;
    const char*   str;
    SrcLayout     layout;
    int           bytes;          // bytes per stored component (Y410: 4 = one packed word)
    int           Packsize;       // Source/Helper.cpp:309-359 column
    int           PitchCoeff;
    ColorSystem   CSType;
    int           Subsampling;    // 420 / 422 / 444 / 400
    int           CDepth;
    int           div_chroma_w, div_chroma_h;
    int           lsb_shift;      // CopyPlane10to16 (Source/Helper.cpp:789-803): stored code << 6
    bool          v_before_u;     // YV12/YV16/YV24 plane order (Source/Shaders.cpp:159-165)
    int           packed_kind;    // packed422: 0 Y0 U Y1 V, 1 U Y0 V Y1, 2 v210 words; packed444: 0 AYUV, 1 Y410, 2 Y416; rgb: RgbKind
};

const FmtConvParams& GetFmtConvParams(int cformat);                  // Source/Helper.cpp:365-369
bool FormatSupported(const FmtConvParams& p);

// DXVA2_ExtendedFormat bit fields (dxva2api.h) -- values as used by Source/Helper.cpp:949-1004
struct ExtendedFormat {
    uint32_t value = 0;
    uint32_t SampleFormat() const { return value & 0xFF; }
    uint32_t ChromaSubsampling() const { return (value >> 8) & 0xF; }
    uint32_t NominalRange() const { return (value >> 12) & 0x7; }
    uint32_t TransferMatrix() const { return (value >> 15) & 0x7; }
    uint32_t Lighting() const { return (value >> 18) & 0xF; }
    uint32_t Primaries() const { return (value >> 22) & 0x1F; }
    uint32_t TransferFunction() const { return (value >> 27) & 0x1F; }
    static uint32_t Make(uint32_t sample, uint32_t chroma, uint32_t range, uint32_t matrix, uint32_t light, uint32_t prim, uint32_t trc) {
        return sample | (chroma << 8) | (range << 12) | (matrix << 15) | (light << 18) | (prim << 22) | (trc << 27);
    }
};
enum { ChromaSub_Unknown = 0, ChromaSub_MPEG1 = 1, ChromaSub_MPEG2 = 5, ChromaSub_DV_PAL = 6, ChromaSub_Cosited = 7 };
enum { Range_Unknown = 0, Range_0_255 = 1, Range_16_235 = 2 };
enum { Matrix_Unknown = 0, Matrix_BT709 = 1, Matrix_BT601 = 2, Matrix_SMPTE240M = 3, Matrix_BT2020_10 = 4, Matrix_YCgCo = 7 };
enum { Prim_Unknown = 0, Prim_BT709 = 2, Prim_BT470_2_SysM = 3, Prim_BT470_2_SysBG = 4, Prim_SMPTE170M = 5,
       Prim_SMPTE240M = 6, Prim_BT2020 = 9, Prim_DCI_P3 = 11 };
enum { Trc_Unknown = 0, Trc_10 = 1, Trc_18 = 2, Trc_20 = 3, Trc_22 = 4, Trc_709 = 5, Trc_240M = 6, Trc_sRGB = 7,
       Trc_28 = 8, Trc_2020_const = 12, Trc_2020 = 13, Trc_26 = 14, Trc_2084 = 15, Trc_HLG = 16 };
enum { Light_Unknown = 0, Light_dim = 3 };

ExtendedFormat SpecifyExtendedFormat(ExtendedFormat ex, const FmtConvParams& fmt, unsigned width, unsigned height);

}  // namespace b200vr
