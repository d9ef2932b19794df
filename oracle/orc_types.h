/*
 * oracle/orc_types.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Shared plain-C types of the CPU oracle: a restatement of the per-frame
 * shader chain of MPC Video Renderer (reference paths are relative to
 * /root/reference).  Nothing under videorenderer_b200/ may include, link or
 * call anything in oracle/; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * Enum values are numerically identical to the reference's:
 *   ColorFormat_t          Source/Helper.h:86-127
 *   TEXFMT_/CHROMA_/UPSCALE_/DOWNSCALE_   Source/IVideoRenderer.h:25-72
 *   DXVA2_ExtendedFormat   dxva2api.h bit layout, MFVideo* extensions as used
 *                          in Source/Helper.cpp:949-1004
 */
#ifndef ORC_TYPES_H
#define ORC_TYPES_H
#include <stdint.h>

enum { /* Source/Helper.h:86-127 */
    ORC_CF_NONE = 0, ORC_CF_NV12, ORC_CF_P010, ORC_CF_P016, ORC_CF_YUY2, ORC_CF_UYVY,
    ORC_CF_P210, ORC_CF_P216, ORC_CF_Y210, ORC_CF_Y216, ORC_CF_V210, ORC_CF_AYUV,
    ORC_CF_Y410, ORC_CF_Y416, ORC_CF_YV12, ORC_CF_YV16, ORC_CF_YV24,
    ORC_CF_YUV420P8, ORC_CF_YUV422P8, ORC_CF_YUV444P8, ORC_CF_YUV420P10, ORC_CF_YUV420P16,
    ORC_CF_YUV422P10, ORC_CF_YUV422P16, ORC_CF_YUV444P10, ORC_CF_YUV444P16,
    ORC_CF_GBRP8, ORC_CF_GBRP10, ORC_CF_GBRP16, ORC_CF_RGB24, ORC_CF_XRGB32, ORC_CF_ARGB32,
    ORC_CF_r210, ORC_CF_RGB48, ORC_CF_BGR48, ORC_CF_BGRA64, ORC_CF_B64A,
    ORC_CF_Y8, ORC_CF_Y10, ORC_CF_Y16, ORC_CF_COUNT
};

enum { ORC_TEXFMT_AUTOINT = 0, ORC_TEXFMT_8INT = 8, ORC_TEXFMT_10INT = 10, ORC_TEXFMT_16FLOAT = 16 };
enum { ORC_CHROMA_Nearest = 0, ORC_CHROMA_Bilinear, ORC_CHROMA_CatmullRom };
enum { ORC_UPSCALE_Nearest = 0, ORC_UPSCALE_Mitchell, ORC_UPSCALE_CatmullRom, ORC_UPSCALE_Lanczos2,
       ORC_UPSCALE_Lanczos3, ORC_UPSCALE_Jinc2 };
enum { ORC_DOWNSCALE_Box = 0, ORC_DOWNSCALE_Bilinear, ORC_DOWNSCALE_Hamming, ORC_DOWNSCALE_Bicubic,
       ORC_DOWNSCALE_BicubicSharp, ORC_DOWNSCALE_Lanczos };

/* surface formats (DXGI names in the reference): m_InternalTexFmt / m_SwapChainFmt */
enum { ORC_SURF_BGRA8 = 0, ORC_SURF_RGB10A2 = 1, ORC_SURF_RGBA16F = 2,
       ORC_SURF_RAW = 3 /* oracle-internal: values as sampled from the source texture (no render-target write) */ };

/* shader variant: which twin of the HLSL files is restated */
enum { ORC_VARIANT_D3D11 = 0, ORC_VARIANT_D3D9 = 1 };

/* arithmetic mode */
enum { ORC_MODE_SPEC = 0,    /* the frozen fp32 contract the CUDA path must match bit for bit   */
       ORC_MODE_LITERAL = 1, /* written-order fp32, no contraction, libm exp2f/log2f, true divides */
       ORC_MODE_F64 = 2 };   /* float64 twin of LITERAL, arbitrates                             */

/* DXVA2_ExtendedFormat bit fields (dxva2api.h): value =
 *  SampleFormat:8 | ChromaSubsampling:4 | NominalRange:3 | TransferMatrix:3 |
 *  Lighting:4 | Primaries:5 | TransferFunction:5 */
#define ORC_EXFMT(sample, chroma, range, matrix, light, prim, trc) \
    ((uint32_t)(sample) | ((uint32_t)(chroma) << 8) | ((uint32_t)(range) << 12) | \
     ((uint32_t)(matrix) << 15) | ((uint32_t)(light) << 18) | ((uint32_t)(prim) << 22) | \
     ((uint32_t)(trc) << 27))
#define ORC_EXFMT_CHROMA(v) (((v) >> 8) & 0xF)
#define ORC_EXFMT_RANGE(v)  (((v) >> 12) & 0x7)
#define ORC_EXFMT_MATRIX(v) (((v) >> 15) & 0x7)
#define ORC_EXFMT_LIGHT(v)  (((v) >> 18) & 0xF)
#define ORC_EXFMT_PRIM(v)   (((v) >> 22) & 0x1F)
#define ORC_EXFMT_TRC(v)    (((v) >> 27) & 0x1F)

enum { ORC_CHROMA_SUB_UNKNOWN = 0, ORC_CHROMA_SUB_MPEG1 = 1, ORC_CHROMA_SUB_MPEG2 = 5,
       ORC_CHROMA_SUB_DV_PAL = 6, ORC_CHROMA_SUB_COSITED = 7 };
enum { ORC_RANGE_UNKNOWN = 0, ORC_RANGE_0_255 = 1, ORC_RANGE_16_235 = 2, ORC_RANGE_48_208 = 3 };
enum { ORC_MATRIX_UNKNOWN = 0, ORC_MATRIX_BT709 = 1, ORC_MATRIX_BT601 = 2, ORC_MATRIX_SMPTE240M = 3,
       ORC_MATRIX_BT2020_10 = 4, ORC_MATRIX_BT2020_12 = 5, ORC_MATRIX_YCgCo = 7 };
enum { ORC_PRIM_UNKNOWN = 0, ORC_PRIM_BT709 = 2, ORC_PRIM_BT470_2_SysM = 3, ORC_PRIM_BT470_2_SysBG = 4,
       ORC_PRIM_SMPTE170M = 5, ORC_PRIM_SMPTE240M = 6, ORC_PRIM_EBU3213 = 7, ORC_PRIM_SMPTE_C = 8,
       ORC_PRIM_BT2020 = 9, ORC_PRIM_XYZ = 10, ORC_PRIM_DCI_P3 = 11 };
enum { ORC_TRC_UNKNOWN = 0, ORC_TRC_10 = 1, ORC_TRC_18 = 2, ORC_TRC_20 = 3, ORC_TRC_22 = 4, ORC_TRC_709 = 5,
       ORC_TRC_240M = 6, ORC_TRC_sRGB = 7, ORC_TRC_28 = 8, ORC_TRC_Log_100 = 9, ORC_TRC_Log_316 = 10,
       ORC_TRC_709_sym = 11, ORC_TRC_2020_const = 12, ORC_TRC_2020 = 13, ORC_TRC_26 = 14,
       ORC_TRC_2084 = 15, ORC_TRC_HLG = 16 };
enum { ORC_LIGHT_UNKNOWN = 0, ORC_LIGHT_dim = 3 };

typedef struct orc_config {
    int32_t  cformat;            /* ORC_CF_*                                         */
    int32_t  width, height;      /* m_srcWidth / m_srcHeight (source texture size)   */
    int32_t  src_rect[4];        /* m_srcRect  L,T,R,B                               */
    uint32_t exfmt;              /* m_srcExFmt.value (defaults already applied, or 0 bits = unknown) */
    int32_t  dst_rect[4];        /* m_videoRect L,T,R,B in window coordinates        */
    int32_t  win_w, win_h;       /* m_windowRect size == back buffer size            */
    int32_t  backbuffer_fmt;     /* ORC_SURF_* : m_SwapChainFmt (RGBA16F = extension)*/
    /* Settings_t subset, Source/IVideoRenderer.h:104-186 */
    int32_t  iTexFormat, iChromaScaling, iUpscaling, iDownscaling;
    int32_t  bInterpolateAt50pct, bUseDither, bConvertToSdr, iSDRDisplayNits, bDeintBlend;
    /* DXVA2 ProcAmp values already in float (DXVA2FixedToFloat) */
    float    brightness, contrast, hue, saturation;
    int32_t  variant;            /* ORC_VARIANT_*                                    */
    /* HDR output.  hdr_out = m_bHdrPassthroughSupport && (m_bHdrPassthrough || m_bHdrLocalToneMapping):
     * picks the convert type (UpdateConvertColorShader, DX11VideoProcessor.cpp:2948-2950).  tonemap = the
     * ps_hdr10_tonemap post-scale step runs (:3359-3367) with the raw arguments of SetHDR10ShaderParams (:907-921). */
    int32_t  hdr_out, tonemap;
    float    tm_min_nits, tm_max_nits, tm_max_cll, tm_max_fall, tm_display_nits;
    int32_t  tm_type;
    /* SetRotation / SetFlip (Source/VideoProcessor.h:205-210): 0/90/180/270 clockwise, horizontal flip of the source */
    int32_t  rotation, flip;
    /* path-B correction shader drawn as a post-scale step on RGB that an upstream fixed-function converter produced
     * (m_pPSCorrection, DX11VideoProcessor.cpp:1893-1931, :3354-3357): ORC_CORR_* */
    int32_t  correction;
} orc_config;

enum { ORC_CORR_NONE = 0, ORC_CORR_CONVERT_PQ_TO_SDR = 1, ORC_CORR_FIXCONVERT_PQ_TO_SDR = 2, ORC_CORR_FIXCONVERT_HLG_TO_SDR = 3,
       ORC_CORR_FIX_BT2020 = 4, ORC_CORR_CONVERT_HLG_TO_PQ = 5, ORC_CORR_FIX_YCGCO = 6 };

/* Dolby Vision reshaping data of one frame: the fields of MediaSideDataDOVIMetadata (Include/IMediaSideData.h:166-243)
 * that the convert shader consumes (Source/Shaders.cpp:531-589, :703-859; Source/DX11VideoProcessor.cpp:817-834, :990-1141). */
typedef struct orc_dovi_curve {
    uint8_t  num_pivots;                 /* [2, 9] */
    uint16_t pivots[9];
    uint8_t  mapping_idc[8];             /* 0 polynomial, 1 mmr */
    uint8_t  poly_order[8];
    int64_t  poly_coef[8][3];
    uint8_t  mmr_order[8];               /* [1, 3] */
    int64_t  mmr_constant[8];
    int64_t  mmr_coef[8][3][7];
} orc_dovi_curve;
typedef struct orc_dovi_ext {           /* one slot of MediaSideDataDOVIMetadata::Extensions (Include/IMediaSideData.h:245-300) */
    uint8_t  level;                      /* 0 = unused slot; 2 = Level2 {target_max_pq, trim_slope, trim_offset, trim_power,
                                            trim_chroma_weight, trim_saturation_gain, ms_weight}; other levels are carried, not used */
    uint16_t v[7];
} orc_dovi_ext;
typedef struct orc_dovi {
    uint8_t  coef_log2_denom, bl_bit_depth;
    orc_dovi_curve curves[3];
    double   ycc_to_rgb_matrix[9], ycc_to_rgb_offset[3], rgb_to_lms_matrix[9];
    uint16_t source_min_pq, source_max_pq;
    orc_dovi_ext ext[32];
} orc_dovi;

/* test hook (orc_test_trace_*): one record per draw of the resize stage, written by the pass functions themselves.
 * cu / cv = the texture coordinate (in texels) the pass computed for the first and the last output index that drives u / v;
 * u_from_x = 1 when the output column drives u (upright and half-turn quads), 0 when the output row does (quarter turns). */
typedef struct orc_trace_entry {
    int32_t type;            /* 0 separable shader upright (resize_pass), 1 oriented draw (oriented_pass), 2 Jinc2, 4 point copy, 5 point-sampled stretch,
                                6 ConvertColorPass, 7 correction shader, 8 HDR10 tone-map shader, 9 FinalPass */
    int32_t in_id, out_id;   /* the reference's textures: 0 m_TexSrcVideo, 1 m_TexResize, 2 the back buffer, 3 m_TexConvertOutput, 4 / 5 m_TexsPostScale[0 / 1] */
    int32_t in_w, in_h, in_rect[4], out_rect[4], rot, flip;
    int32_t in_axis;         /* texture axis the shader filters: 0 X shader, 1 Y shader, -1 none, 2 both (Jinc2) */
    int32_t kind, filter;    /* 1 interpolation / 2 convolution and the scaler index */
    int32_t u_from_x;
    float   scale;           /* the `scale` constant the convolution shader divides by (TextureResizeShader :351-354), 0 when unused */
    float   cu[2], cv[2];
} orc_trace_entry;

/* optional taps into the pass chain for stage-wise parity tests */
typedef struct orc_debug {
    float*    convert_f32;   /* [srcRectH][srcRectW][3]  convert pass output before the RT write    */
    void*     convert_surf;  /* packed m_TexConvertOutput in its DXGI memory format                 */
    int32_t   convert_surf_pitch;
    uint16_t* resize_x_f16;  /* [h1][win_w][4] m_TexResize (RGBA16F) if the two-pass path ran       */
    void*     postscale_surf;/* packed m_TexsPostScale[0] if a final pass ran                        */
    int32_t   postscale_pitch;
    int32_t   out_internal_fmt, out_final_pass, out_resize_x, out_resize_y; /* decisions taken       */
} orc_debug;

#endif
