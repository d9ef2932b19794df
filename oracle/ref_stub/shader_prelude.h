// oracle/ref_stub/shader_prelude.h -- TEST INFRASTRUCTURE ONLY.
// Windows-type prelude that lets the reference's own format table, extended-format defaults and convert-shader GENERATOR
// compile where they lie: oracle/Makefile pipes this file, the reference's lines (sed -n, nothing is copied into the repo):
//   Source/Helper.h:86-165      ColorFormat_t ... FmtConvParams_t
//   Source/Helper.cpp:283-369   plane configs, s_FmtConvMapping, GetFmtConvParams(ColorFormat_t)
//   Source/Helper.cpp:949-1004  set_colorspace
//   Source/Helper.cpp:1169-1211 SpecifyExtendedFormat
//   Source/IVideoRenderer.h:47-52, Source/Shaders.h:54-58   CHROMA_* and SHADER_CONVERT_* enums
//   Source/Shaders.cpp:66-79, 82-589, 593-930   ShaderGetPixels, the Dolby Vision reshape text, GetShaderConvertColor
// and ref_stub/shader_wrap.inc into g++.  CompileShader is replaced by a stub that keeps the generated HLSL text, so the
// tests can compare the op sequence and the embedded literals the REFERENCE emits with the oracle's branch choice.
// The Windows SDK types below (dxva2api.h / mfobjects.h / d3d9types.h / dxgiformat.h) carry their documented values.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <format>
#include "csputils.h"                      // the reference's (sed-patched enum type, see Makefile)

typedef unsigned int UINT; typedef unsigned long DWORD; typedef long HRESULT; typedef void* LPVOID; typedef const char* LPCSTR;
typedef unsigned char BYTE; typedef long LONG;
struct RECT { LONG left, top, right, bottom; };
struct ID3DBlob; struct D3D_SHADER_MACRO { LPCSTR Name; LPCSTR Definition; };
struct CMediaType;
#define S_OK ((HRESULT)0)
#define E_FAIL ((HRESULT)0x80004005L)
#define ASSERT(x) ((void)0)
#define DLog(...) ((void)0)

enum D3DFORMAT { D3DFMT_UNKNOWN = 0, D3DFMT_A8R8G8B8 = 21, D3DFMT_X8R8G8B8 = 22, D3DFMT_A2B10G10R10 = 31, D3DFMT_G16R16 = 34, D3DFMT_A16B16G16R16 = 36,
                 D3DFMT_L8 = 50, D3DFMT_A8L8 = 51, D3DFMT_L16 = 81,
                 D3DFMT_NV12 = 0x3231564E, D3DFMT_YV12 = 0x32315659, D3DFMT_YUY2 = 0x32595559, D3DFMT_UYVY = 0x59565955, D3DFMT_P010 = 0x30313050,
                 D3DFMT_P016 = 0x36313050, D3DFMT_P210 = 0x30313250, D3DFMT_P216 = 0x36313250, D3DFMT_Y410 = 0x30313459, D3DFMT_Y416 = 0x36313459,
                 D3DFMT_PLANAR = 0xFFFF };
enum DXGI_FORMAT { DXGI_FORMAT_UNKNOWN = 0, DXGI_FORMAT_R16G16B16A16_UNORM = 11, DXGI_FORMAT_R10G10B10A2_UNORM = 24, DXGI_FORMAT_R8G8B8A8_UNORM = 28,
                   DXGI_FORMAT_R16G16_UNORM = 35, DXGI_FORMAT_R8G8_UNORM = 49, DXGI_FORMAT_R16_UNORM = 56, DXGI_FORMAT_R8_UNORM = 61,
                   DXGI_FORMAT_B8G8R8A8_UNORM = 87, DXGI_FORMAT_B8G8R8X8_UNORM = 88, DXGI_FORMAT_AYUV = 100, DXGI_FORMAT_Y410 = 101, DXGI_FORMAT_Y416 = 102,
                   DXGI_FORMAT_NV12 = 103, DXGI_FORMAT_P010 = 104, DXGI_FORMAT_P016 = 105, DXGI_FORMAT_YUY2 = 107, DXGI_FORMAT_Y210 = 108, DXGI_FORMAT_Y216 = 109,
                   DXGI_FORMAT_PLANAR = 0xFFFF };

// dxva2api.h
typedef struct _DXVA2_ExtendedFormat {
    union {
        struct { UINT SampleFormat : 8; UINT VideoChromaSubsampling : 4; UINT NominalRange : 3; UINT VideoTransferMatrix : 3;
                 UINT VideoLighting : 4; UINT VideoPrimaries : 5; UINT VideoTransferFunction : 5; };
        UINT value;
    };
} DXVA2_ExtendedFormat;
enum { DXVA2_VideoChromaSubsampling_Unknown = 0, DXVA2_VideoChromaSubsampling_MPEG1 = 1, DXVA2_VideoChromaSubsampling_MPEG2 = 5,
       DXVA2_VideoChromaSubsampling_DV_PAL = 6, DXVA2_VideoChromaSubsampling_Cosited = 7 };
enum { DXVA2_NominalRange_Unknown = 0, DXVA2_NominalRange_0_255 = 1, DXVA2_NominalRange_16_235 = 2, DXVA2_NominalRange_48_208 = 3 };
enum { DXVA2_VideoTransferMatrix_Unknown = 0, DXVA2_VideoTransferMatrix_BT709 = 1, DXVA2_VideoTransferMatrix_BT601 = 2, DXVA2_VideoTransferMatrix_SMPTE240M = 3,
       MFVideoTransferMatrix_BT2020_10 = 4, MFVideoTransferMatrix_BT2020_12 = 5 };
#define VIDEOTRANSFERMATRIX_YCgCo 7        /* Source/Utils/MediaTypes.h:47 */
enum { DXVA2_VideoLighting_Unknown = 0, DXVA2_VideoLighting_bright = 1, DXVA2_VideoLighting_office = 2, DXVA2_VideoLighting_dim = 3, DXVA2_VideoLighting_dark = 4 };
enum { DXVA2_VideoPrimaries_Unknown = 0, DXVA2_VideoPrimaries_reserved = 1, DXVA2_VideoPrimaries_BT709 = 2, DXVA2_VideoPrimaries_BT470_2_SysM = 3,
       DXVA2_VideoPrimaries_BT470_2_SysBG = 4, DXVA2_VideoPrimaries_SMPTE170M = 5, DXVA2_VideoPrimaries_SMPTE240M = 6, DXVA2_VideoPrimaries_EBU3213 = 7,
       DXVA2_VideoPrimaries_SMPTE_C = 8, MFVideoPrimaries_BT2020 = 9, MFVideoPrimaries_XYZ = 10, MFVideoPrimaries_DCI_P3 = 11, MFVideoPrimaries_ACES = 12 };
enum { DXVA2_VideoTransFunc_Unknown = 0, DXVA2_VideoTransFunc_10 = 1, DXVA2_VideoTransFunc_18 = 2, DXVA2_VideoTransFunc_20 = 3, DXVA2_VideoTransFunc_22 = 4,
       DXVA2_VideoTransFunc_709 = 5, DXVA2_VideoTransFunc_240M = 6, DXVA2_VideoTransFunc_sRGB = 7, DXVA2_VideoTransFunc_28 = 8,
       MFVideoTransFunc_Log_100 = 9, MFVideoTransFunc_Log_316 = 10, MFVideoTransFunc_709_sym = 11, MFVideoTransFunc_2020_const = 12, MFVideoTransFunc_2020 = 13,
       MFVideoTransFunc_26 = 14, MFVideoTransFunc_2084 = 15, MFVideoTransFunc_HLG = 16, MFVideoTransFunc_10_rel = 17 };

// resource ids (Source/resource.h:10-12) and the resource loader: the HLSL include files are represented by one marker line each
#define IDF_HLSL_ST2084 501
#define IDF_HLSL_HLG 502
#define IDF_HLSL_HDR_TONE_MAPPING 503
static HRESULT GetDataFromResource(LPVOID& data, DWORD& size, UINT resid)
{
    static const char st[] = "//#include st2084.hlsl", hl[] = "//#include hlg.hlsl", tm[] = "//#include hdr_tone_mapping.hlsl";
    const char* p = resid == IDF_HLSL_ST2084 ? st : (resid == IDF_HLSL_HLG ? hl : (resid == IDF_HLSL_HDR_TONE_MAPPING ? tm : nullptr));
    if (!p) return E_FAIL;
    data = (LPVOID)p; size = (DWORD)strlen(p);
    return S_OK;
}

// Include/IMediaSideData.h:154-230 (the fields GetShaderConvertColor reads)
#pragma pack(push, 1)
#define LAV_DOVI_MAX_PIECES 8
struct MediaSideDataDOVIMetadata {
    struct { uint8_t rpu_type; uint16_t rpu_format; uint8_t vdr_rpu_profile, vdr_rpu_level, chroma_resampling_explicit_filter_flag, coef_data_type, coef_log2_denom,
             vdr_rpu_normalized_idc, bl_video_full_range_flag, bl_bit_depth, el_bit_depth, vdr_bit_depth, spatial_resampling_filter_flag,
             el_spatial_resampling_filter_flag, disable_residual_flag; } Header;
    struct {
        uint8_t vdr_rpu_id, mapping_color_space, mapping_chroma_format_idc;
        struct { uint8_t num_pivots; uint16_t pivots[LAV_DOVI_MAX_PIECES + 1]; uint8_t mapping_idc[LAV_DOVI_MAX_PIECES]; uint8_t poly_order[LAV_DOVI_MAX_PIECES];
                 int64_t poly_coef[LAV_DOVI_MAX_PIECES][3]; uint8_t mmr_order[LAV_DOVI_MAX_PIECES]; int64_t mmr_constant[LAV_DOVI_MAX_PIECES];
                 int64_t mmr_coef[LAV_DOVI_MAX_PIECES][3][7]; } curves[3];
    } Mapping;
    struct { double ycc_to_rgb_matrix[9]; double ycc_to_rgb_offset[3]; double rgb_to_lms_matrix[9]; } ColorMetadata;
};
#pragma pack(pop)

static std::string g_lastShader;
static HRESULT CompileShader(const std::string& srcCode, const D3D_SHADER_MACRO*, LPCSTR, ID3DBlob**) { g_lastShader = srcCode; return S_OK; }
