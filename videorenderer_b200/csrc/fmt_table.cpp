// fmt_table.cpp -- see fmt_table.h.  Reference: Source/Helper.cpp:283-359, :1169-1211.
#include "fmt_table.h"

namespace b200vr {

#define U LAYOUT_UNSUPPORTED
static const FmtConvParams s_table[MPCVR_CF_COUNT] = {
    // cformat            str          layout           bytes Pack PitchC CSType   Sub  CDepth dw dh shift v<u kind
    {MPCVR_CF_NONE,      "unknown",   U,                0,   0, 0, CS_YUV,    0,  0, 1, 1, 0, false, 0},
    {MPCVR_CF_NV12,      "NV12",      LAYOUT_BIPLANAR,  1,   1, 3, CS_YUV,  420,  8, 2, 2, 0, false, 0},
    {MPCVR_CF_P010,      "P010",      LAYOUT_BIPLANAR,  2,   2, 3, CS_YUV,  420, 16, 2, 2, 0, false, 0},
    {MPCVR_CF_P016,      "P016",      LAYOUT_BIPLANAR,  2,   2, 3, CS_YUV,  420, 16, 2, 2, 0, false, 0},
    {MPCVR_CF_YUY2,      "YUY2",      LAYOUT_PACKED422, 1,   2, 2, CS_YUV,  422,  8, 2, 1, 0, false, 0},
    {MPCVR_CF_UYVY,      "UYVY",      LAYOUT_PACKED422, 1,   2, 2, CS_YUV,  422,  8, 2, 1, 0, false, 1},
    {MPCVR_CF_P210,      "P210",      LAYOUT_BIPLANAR,  2,   2, 4, CS_YUV,  422, 16, 2, 1, 0, false, 0},
    {MPCVR_CF_P216,      "P216",      LAYOUT_BIPLANAR,  2,   2, 4, CS_YUV,  422, 16, 2, 1, 0, false, 0},
    {MPCVR_CF_Y210,      "Y210",      LAYOUT_PACKED422, 2,   4, 2, CS_YUV,  422, 10, 2, 1, 0, false, 0},
    {MPCVR_CF_Y216,      "Y216",      LAYOUT_PACKED422, 2,   4, 2, CS_YUV,  422, 16, 2, 1, 0, false, 0},
    {MPCVR_CF_V210,      "v210",      LAYOUT_PACKED422, 2,   2, 2, CS_YUV,  422, 10, 2, 1, 0, false, 2},   // read in place (no CopyFrameV210 repack pass)
    {MPCVR_CF_AYUV,      "AYUV",      LAYOUT_PACKED444, 1,   4, 2, CS_YUV,  444,  8, 1, 1, 0, false, 0},
    {MPCVR_CF_Y410,      "Y410",      LAYOUT_PACKED444, 4,   4, 2, CS_YUV,  444, 10, 1, 1, 0, false, 1},
    {MPCVR_CF_Y416,      "Y416",      LAYOUT_PACKED444, 2,   8, 2, CS_YUV,  444, 16, 1, 1, 0, false, 2},
    {MPCVR_CF_YV12,      "YV12",      LAYOUT_PLANAR3,   1,   1, 3, CS_YUV,  420,  8, 2, 2, 0, true,  0},
    {MPCVR_CF_YV16,      "YV16",      LAYOUT_PLANAR3,   1,   1, 4, CS_YUV,  422,  8, 2, 1, 0, true,  0},
    {MPCVR_CF_YV24,      "YV24",      LAYOUT_PLANAR3,   1,   1, 6, CS_YUV,  444,  8, 1, 1, 0, true,  0},
    {MPCVR_CF_YUV420P8,  "YUV420P8",  LAYOUT_PLANAR3,   1,   1, 3, CS_YUV,  420,  8, 2, 2, 0, false, 0},
    {MPCVR_CF_YUV422P8,  "YUV422P8",  LAYOUT_PLANAR3,   1,   1, 4, CS_YUV,  422,  8, 2, 1, 0, false, 0},
    {MPCVR_CF_YUV444P8,  "YUV444P8",  LAYOUT_PLANAR3,   1,   1, 6, CS_YUV,  444,  8, 1, 1, 0, false, 0},
    {MPCVR_CF_YUV420P10, "YUV420P10", LAYOUT_PLANAR3,   2,   2, 3, CS_YUV,  420, 10, 2, 2, 6, false, 0},
    {MPCVR_CF_YUV420P16, "YUV420P16", LAYOUT_PLANAR3,   2,   2, 3, CS_YUV,  420, 16, 2, 2, 0, false, 0},
    {MPCVR_CF_YUV422P10, "YUV422P10", LAYOUT_PLANAR3,   2,   2, 4, CS_YUV,  422, 10, 2, 1, 6, false, 0},
    {MPCVR_CF_YUV422P16, "YUV422P16", LAYOUT_PLANAR3,   2,   2, 4, CS_YUV,  422, 16, 2, 1, 0, false, 0},
    {MPCVR_CF_YUV444P10, "YUV444P10", LAYOUT_PLANAR3,   2,   2, 6, CS_YUV,  444, 10, 1, 1, 6, false, 0},
    {MPCVR_CF_YUV444P16, "YUV444P16", LAYOUT_PLANAR3,   2,   2, 6, CS_YUV,  444, 16, 1, 1, 0, false, 0},
    // RGB / gray inputs (Source/Helper.cpp:341-358).  The reference repacks these samples on the CPU into a DXGI texture
    // (CopyFrame*, Source/Helper.cpp:446-803); here the kernels read the sample layout directly -- packed_kind selects it.
    {MPCVR_CF_GBRP8,     "GBRP8",     LAYOUT_PLANAR3,   1,   1, 6, CS_RGB,  444,  8, 1, 1, 0, false, 0},
    {MPCVR_CF_GBRP10,    "GBRP10",    LAYOUT_PLANAR3,   2,   2, 6, CS_RGB,  444, 10, 1, 1, 6, false, 0},
    {MPCVR_CF_GBRP16,    "GBRP16",    LAYOUT_PLANAR3,   2,   2, 6, CS_RGB,  444, 16, 1, 1, 0, false, 0},
    {MPCVR_CF_RGB24,     "RGB24",     LAYOUT_RGB,       1,   3, 2, CS_RGB,  444,  8, 1, 1, 0, false, RGBK_BGR24},
    {MPCVR_CF_XRGB32,    "RGB32",     LAYOUT_RGB,       1,   4, 2, CS_RGB,  444,  8, 1, 1, 0, false, RGBK_BGRX32},
    {MPCVR_CF_ARGB32,    "ARGB32",    LAYOUT_RGB,       1,   4, 2, CS_RGB,  444,  8, 1, 1, 0, false, RGBK_BGRX32},
    {MPCVR_CF_r210,      "r210",      LAYOUT_RGB,       4,   4, 2, CS_RGB,  444, 10, 1, 1, 0, false, RGBK_R210},
    {MPCVR_CF_RGB48,     "RGB48",     LAYOUT_RGB,       2,   6, 2, CS_RGB,  444, 16, 1, 1, 0, false, RGBK_RGB48},
    {MPCVR_CF_BGR48,     "BGR48",     LAYOUT_RGB,       2,   6, 2, CS_RGB,  444, 16, 1, 1, 0, false, RGBK_BGR48},
    {MPCVR_CF_BGRA64,    "BGRA64",    LAYOUT_RGB,       2,   8, 2, CS_RGB,  444, 16, 1, 1, 0, false, RGBK_BGRA64},
    {MPCVR_CF_B64A,      "b64a",      LAYOUT_RGB,       2,   8, 2, CS_RGB,  444, 16, 1, 1, 0, false, RGBK_B64A},
    {MPCVR_CF_Y8,        "Y8",        LAYOUT_GRAY,      1,   1, 2, CS_GRAY, 400,  8, 1, 1, 0, false, 0},
    {MPCVR_CF_Y10,       "Y10",       LAYOUT_GRAY,      2,   2, 2, CS_GRAY, 400, 10, 1, 1, 6, false, 0},
    {MPCVR_CF_Y16,       "Y16",       LAYOUT_GRAY,      2,   2, 2, CS_GRAY, 400, 16, 1, 1, 0, false, 0},
};
#undef U

const FmtConvParams& GetFmtConvParams(int cformat)
{
    if (cformat < 0 || cformat >= MPCVR_CF_COUNT) return s_table[0];
    return s_table[cformat];
}

bool FormatSupported(const FmtConvParams& p) { return p.layout != LAYOUT_UNSUPPORTED; }

ExtendedFormat SpecifyExtendedFormat(ExtendedFormat ex, const FmtConvParams& fmt, unsigned width, unsigned height)
{
    if (fmt.CSType == CS_RGB) { ex.value = 0u; return ex; }
    if (fmt.CSType != CS_YUV) return ex;
    uint32_t chroma = ex.ChromaSubsampling(), range = ex.NominalRange(), matrix = ex.TransferMatrix(),
             light = ex.Lighting(), prim = ex.Primaries(), trc = ex.TransferFunction();
    if (fmt.Subsampling != 420)               chroma = ChromaSub_Unknown;
    else if (chroma == ChromaSub_Unknown)     chroma = ChromaSub_MPEG2;
    if (range == Range_Unknown)               range = Range_16_235;
    if (matrix == Matrix_Unknown)             matrix = (width <= 1024 && height <= 576) ? Matrix_BT601 : Matrix_BT709;
    if (light == Light_Unknown)               light = Light_dim;
    if (prim == Prim_Unknown)                 prim = Prim_BT709;
    if (trc == Trc_Unknown)                   trc = Trc_709;
    ex.value = ExtendedFormat::Make(ex.SampleFormat(), chroma, range, matrix, light, prim, trc);
    return ex;
}

}  // namespace b200vr
