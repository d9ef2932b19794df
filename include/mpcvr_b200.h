/*
 * mpcvr_b200.h -- C ABI of the B200-native frame path (libmpcvr_b200.so).
 *
 * Drop-in boundary for the per-frame pixel path of MPC Video Renderer: every entry point
 * replaces one member of the reference's CVideoProcessor / IVideoRenderer / ISubRender11 /
 * ID3D11DecoderConfiguration surface (reference paths relative to /root/reference).
 * Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Object model: one `mpcvr` instance == one CVideoProcessor (one stream, one GPU, one CUDA
 * stream).  Instances are externally synchronised like the reference's (all frame and
 * settings calls arrive under m_RendererLock, Source/VideoRenderer.cpp:422,444,558,586);
 * distinct instances are fully independent.
 *
 * Errors: HRESULT values as int32_t (S_OK 0, S_FALSE 1 = frame skipped,
 * Source/DX11VideoProcessor.cpp:2179); BOOL-returning reference methods return 1/0.
 * CUDA failures map to MPCVR_E_FAIL with text available from mpcvr_last_error().
 * There is no CPU fallback: without a usable CUDA device frame calls fail.
 */
#ifndef MPCVR_B200_H
#define MPCVR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCVR_ABI_VERSION 1

/* HRESULT codes used by the reference on this path */
#define MPCVR_S_OK          ((int32_t)0)
#define MPCVR_S_FALSE       ((int32_t)1)
#define MPCVR_E_NOTIMPL     ((int32_t)0x80004001)
#define MPCVR_E_POINTER     ((int32_t)0x80004003)
#define MPCVR_E_ABORT       ((int32_t)0x80004004)
#define MPCVR_E_FAIL        ((int32_t)0x80004005)
#define MPCVR_E_UNEXPECTED  ((int32_t)0x8000FFFF)
#define MPCVR_E_OUTOFMEMORY ((int32_t)0x8007000E)
#define MPCVR_E_INVALIDARG  ((int32_t)0x80070057)

/* ColorFormat_t, Source/Helper.h:86-127 (same numeric values) */
typedef enum mpcvr_cformat {
    MPCVR_CF_NONE = 0, MPCVR_CF_NV12, MPCVR_CF_P010, MPCVR_CF_P016, MPCVR_CF_YUY2, MPCVR_CF_UYVY,
    MPCVR_CF_P210, MPCVR_CF_P216, MPCVR_CF_Y210, MPCVR_CF_Y216, MPCVR_CF_V210, MPCVR_CF_AYUV,
    MPCVR_CF_Y410, MPCVR_CF_Y416, MPCVR_CF_YV12, MPCVR_CF_YV16, MPCVR_CF_YV24,
    MPCVR_CF_YUV420P8, MPCVR_CF_YUV422P8, MPCVR_CF_YUV444P8, MPCVR_CF_YUV420P10, MPCVR_CF_YUV420P16,
    MPCVR_CF_YUV422P10, MPCVR_CF_YUV422P16, MPCVR_CF_YUV444P10, MPCVR_CF_YUV444P16,
    MPCVR_CF_GBRP8, MPCVR_CF_GBRP10, MPCVR_CF_GBRP16, MPCVR_CF_RGB24, MPCVR_CF_XRGB32, MPCVR_CF_ARGB32,
    MPCVR_CF_r210, MPCVR_CF_RGB48, MPCVR_CF_BGR48, MPCVR_CF_BGRA64, MPCVR_CF_B64A,
    MPCVR_CF_Y8, MPCVR_CF_Y10, MPCVR_CF_Y16, MPCVR_CF_COUNT
} mpcvr_cformat;

/* Source/IVideoRenderer.h:25-95 */
enum { MPCVR_TEXFMT_AUTOINT = 0, MPCVR_TEXFMT_8INT = 8, MPCVR_TEXFMT_10INT = 10, MPCVR_TEXFMT_16FLOAT = 16 };
enum { MPCVR_CHROMA_Nearest = 0, MPCVR_CHROMA_Bilinear, MPCVR_CHROMA_CatmullRom, MPCVR_CHROMA_COUNT };
enum { MPCVR_UPSCALE_Nearest = 0, MPCVR_UPSCALE_Mitchell, MPCVR_UPSCALE_CatmullRom, MPCVR_UPSCALE_Lanczos2,
       MPCVR_UPSCALE_Lanczos3, MPCVR_UPSCALE_Jinc2, MPCVR_UPSCALE_COUNT };
enum { MPCVR_DOWNSCALE_Box = 0, MPCVR_DOWNSCALE_Bilinear, MPCVR_DOWNSCALE_Hamming, MPCVR_DOWNSCALE_Bicubic,
       MPCVR_DOWNSCALE_BicubicSharp, MPCVR_DOWNSCALE_Lanczos, MPCVR_DOWNSCALE_COUNT };
#define MPCVR_SDR_NITS_DEF 125
#define MPCVR_SDR_NITS_MIN 25
#define MPCVR_SDR_NITS_MAX 400

/* Settings_t, Source/IVideoRenderer.h:104-186, field for field (bool -> int32_t).
 * Only the fields SURVEY 8(a24) marks act on this path; the rest are stored and returned unchanged. */
typedef struct mpcvr_settings {
    int32_t bUseD3D11;
    int32_t bShowStats;
    int32_t iResizeStats;
    int32_t iTexFormat;              /* acts: MPCVR_TEXFMT_* */
    int32_t VPFmts_bNV12, VPFmts_bP01x, VPFmts_bYUY2, VPFmts_bOther;
    int32_t iVPDeinterlacing;
    int32_t bDeintDouble;
    int32_t bVPScaling;
    int32_t iVPSuperRes;
    int32_t bVPRTXVideoHDR;
    int32_t iChromaScaling;          /* acts */
    int32_t iUpscaling;              /* acts */
    int32_t iDownscaling;            /* acts */
    int32_t bInterpolateAt50pct;     /* acts */
    int32_t bUseDither;              /* acts */
    int32_t bDeintBlend;             /* acts (interlaced 4:2:0 only) */
    int32_t iSwapEffect;
    int32_t bExclusiveFS;
    int32_t bVBlankBeforePresent;
    int32_t bAdjustPresentTime;
    int32_t bReinitByDisplay;
    int32_t bHdrPreferDoVi;
    int32_t bHdrPassthrough;         /* effective only after mpcvr_set_display_hdr(vp, 1) */
    int32_t iHdrToggleDisplay;
    int32_t iHdrOsdBrightness;
    int32_t bConvertToSdr;           /* acts */
    int32_t iSDRDisplayNits;         /* acts */
    int32_t bHdrLocalToneMapping;
    int32_t iHdrLocalToneMappingType;
    int32_t iHdrDisplayMaxNits;
} mpcvr_settings;

/* Settings_t::SetDefault(), Source/IVideoRenderer.h:140-185 */
void mpcvr_settings_default(mpcvr_settings* s);

typedef struct mpcvr_rect { int32_t left, top, right, bottom; } mpcvr_rect;

/* Back-buffer ("swap chain") formats: m_SwapChainFmt, Source/DX11VideoProcessor.cpp:1478.
 * RGBA16F is an extension (BASELINE config 4 asks for the post-scale fp16 surface). */
typedef enum mpcvr_surface_format {
    MPCVR_SURF_BGRA8 = 0,     /* DXGI_FORMAT_B8G8R8A8_UNORM: bytes B,G,R,A                */
    MPCVR_SURF_RGB10A2 = 1,   /* DXGI_FORMAT_R10G10B10A2_UNORM: R bits 0-9, G 10-19, B 20-29, A 30-31 */
    MPCVR_SURF_RGBA16F = 2,   /* DXGI_FORMAT_R16G16B16A16_FLOAT                           */
    MPCVR_SURF_RGBA16 = 3,    /* DXGI_FORMAT_R16G16B16A16_UNORM: internal only (16-bit RGB sources read without a convert
                                 pass, Source/DX11VideoProcessor.cpp:3320-3322); not a back-buffer format */
    MPCVR_SURF_RGB8 = 4       /* extension: packed 3 bytes per pixel R,G,B -- the codes of the B8G8R8A8 back buffer without the alpha
                                 byte (every decision that looks at the swap-chain format treats it as B8G8R8A8); for callers
                                 that read frames back over PCIe (25 % fewer bytes).  Not a D3D format. */
} mpcvr_surface_format;

/* What CMediaType / VIDEOINFOHEADER2 carry into VerifyMediaType / InitMediaType
 * (Source/DX11VideoProcessor.cpp:1742-1959). */
typedef struct mpcvr_media_type {
    int32_t    cformat;           /* mpcvr_cformat (GetColorFormat, Source/Helper.cpp:263-281)        */
    int32_t    width;             /* biWidth  : buffer width in pixels (pitch = width*Packsize rule)  */
    int32_t    height;            /* |biHeight|                                                        */
    int32_t    frame_width;       /* VR_Extradata.FrameWidth  (0 = same as width), Source/Helper.h:57-65 */
    int32_t    frame_height;      /* VR_Extradata.FrameHeight (0 = same as height)                    */
    mpcvr_rect src_rect;          /* rcSource; all zero = whole frame                                 */
    uint32_t   aspect_x, aspect_y;/* dwPictAspectRatioX/Y                                             */
    uint32_t   ext_format;        /* DXVA2_ExtendedFormat.value taken from dwControlFlags (0 = none)  */
    int32_t    interlaced;        /* dwInterlaceFlags & AMINTERLACE_IsInterlaced                      */
    int64_t    avg_time_per_frame;/* REFERENCE_TIME, 100 ns units                                     */
    int32_t    bottom_up;         /* biCompression == BI_RGB && biHeight > 0: rows are stored last-to-first and the
                                     pitch is negated (Source/DX11VideoProcessor.cpp:1801-1803)                   */
    int32_t    reserved;
} mpcvr_media_type;

/* IMediaSideData blobs (Include/IMediaSideData.h:35-60); layouts identical (pack 1). */
#pragma pack(push, 1)
typedef struct mpcvr_side_data_hdr {          /* MediaSideDataHDR */
    double display_primaries_x[3];
    double display_primaries_y[3];
    double white_point_x, white_point_y;
    double max_display_mastering_luminance, min_display_mastering_luminance;
} mpcvr_side_data_hdr;
typedef struct mpcvr_side_data_hdr_cll {      /* MediaSideDataHDRContentLightLevel */
    uint32_t MaxCLL, MaxFALL;
} mpcvr_side_data_hdr_cll;
#pragma pack(pop)

/* Dolby Vision reshaping data of one frame: the fields of MediaSideDataDOVIMetadata (IID_MediaSideDataDOVIMetadataV2,
 * Include/IMediaSideData.h:166-243) that the convert shader consumes -- Header.coef_log2_denom / bl_bit_depth,
 * Mapping.curves[3], ColorMetadata.ycc_to_rgb_matrix / ycc_to_rgb_offset / rgb_to_lms_matrix -- in natural (unpacked)
 * C layout.  The adapter copies them field by field.  Accepted as CheckDoviMetadata does
 * (Source/VideoProcessor.cpp:273-297): num_pivots in [2, 9], mapping_idc 0 (polynomial) or 1 (MMR). */
typedef struct mpcvr_dovi_curve {
    uint8_t  num_pivots;
    uint16_t pivots[9];
    uint8_t  mapping_idc[8];
    uint8_t  poly_order[8];
    int64_t  poly_coef[8][3];
    uint8_t  mmr_order[8];
    int64_t  mmr_constant[8];
    int64_t  mmr_coef[8][3][7];
} mpcvr_dovi_curve;
typedef struct mpcvr_dovi_ext {      /* one slot of MediaSideDataDOVIMetadata::Extensions (Include/IMediaSideData.h:245-300) */
    uint8_t  level;                   /* 0 = unused; 1 = Level1 {min_pq, max_pq, avg_pq}; 3 = Level3 {min, max, avg pq offsets};
                                         2 = Level2: v = {target_max_pq, trim_slope, trim_offset, trim_power,
                                         trim_chroma_weight, trim_saturation_gain, ms_weight} -> the L2 trims of the to-SDR chain
                                         (Source/DX11VideoProcessor.cpp:2383-2466, Source/Shaders.cpp:787-795);
                                         6 = Level6 {max_luminance, min_luminance, max_cll, max_fall}: static HDR10 metadata (:2477-2490);
                                         other levels are carried, not used */
    uint16_t v[7];
} mpcvr_dovi_ext;
typedef struct mpcvr_dovi_metadata {
    uint8_t  coef_log2_denom, bl_bit_depth;
    mpcvr_dovi_curve curves[3];
    double   ycc_to_rgb_matrix[9], ycc_to_rgb_offset[3], rgb_to_lms_matrix[9];
    uint16_t source_min_pq, source_max_pq;     /* ColorMetadata */
    mpcvr_dovi_ext ext[32];
    /* the header / NLQ fields CheckDoviMetadata reads to turn away profile 4 and profile 7 FEL streams (Source/VideoProcessor.cpp:273-284,
     * Source/VideoProcessor.h:249-263); a base-layer-only stream (profiles 5, 8) has disable_residual_flag = 1 */
    uint8_t  el_spatial_resampling_filter_flag, disable_residual_flag, vdr_bit_depth, nlq_method_idc;
    struct { uint16_t nlq_offset; uint64_t vdr_in_max, linear_deadzone_slope, linear_deadzone_threshold; } nlq[3];
} mpcvr_dovi_metadata;

#define MPCVR_SAMPLE_DEVICE_MEMORY 0x1u  /* planes are CUDA device pointers (IMediaSampleD3D11 analogue)   */

/* What IMediaSample hands to ProcessSample/CopySample (Source/DX11VideoProcessor.cpp:2143-2597).
 * Host sample: `data` = IMediaSample::GetPointer(), planes follow each other as in
 * MemCopyToTexSrcVideo (:1213-1252): Y, then UV (or U, V) at data + pitch*height ...; pitch per
 * InitMediaType (:1789-1803) unless `pitch` != 0.  Device sample: planes[]/pitches[] explicit. */
typedef struct mpcvr_sample {
    uint32_t    flags;
    const void* data;             /* host: contiguous frame buffer                                    */
    int64_t     size;             /* host: bytes available at data (checked like :2579)               */
    int32_t     pitch;            /* host: 0 = m_srcPitch                                             */
    const void* planes[3];        /* device: plane base pointers                                      */
    int32_t     pitches[3];       /* device: bytes                                                    */
    void*       ready_event;      /* device: optional cudaEvent_t the decoder recorded                */
    int64_t     rt_start, rt_end; /* REFERENCE_TIME                                                   */
    const mpcvr_side_data_hdr*     hdr;      /* optional IID_MediaSideDataHDR                         */
    const mpcvr_side_data_hdr_cll* hdr_cll;  /* optional IID_MediaSideDataHDRContentLightLevel        */
    const mpcvr_dovi_metadata*     dovi;     /* optional Dolby Vision reshaping data (YUV types only; used when
                                                bHdrPreferDoVi or the stream is not tagged PQ/HLG, :2271)        */
} mpcvr_sample;

/* DXVA2_ProcAmpValues in float (DXVA2FixedToFloat), SetProcAmpValues Source/DX11VideoProcessor.cpp:4506-4537 */
#define MPCVR_PROCAMP_BRIGHTNESS 0x1u
#define MPCVR_PROCAMP_CONTRAST   0x2u
#define MPCVR_PROCAMP_HUE        0x4u
#define MPCVR_PROCAMP_SATURATION 0x8u
typedef struct mpcvr_procamp { float brightness, contrast, hue, saturation; } mpcvr_procamp;

typedef struct mpcvr mpcvr;   /* opaque: the CVideoProcessor instance */

/* --- lifetime: CMpcVideoRenderer ctor creates the processor, Source/VideoRenderer.cpp:284-303;
 *     CVideoProcessor::Init(hwnd,...) Source/VideoProcessor.h:173 ---------------------------------
 * cuda_device >= 0: device ordinal; the instance creates its own stream unless `stream` != NULL.
 * cuda_device == -1: planning-only instance (media-type negotiation and parameter queries work,
 * every frame call fails with MPCVR_E_UNEXPECTED) -- lets hosts negotiate formats off-GPU. */
int32_t mpcvr_create(const mpcvr_settings* settings, int32_t cuda_device, void* stream, mpcvr** out);
void    mpcvr_destroy(mpcvr* vp);
int32_t mpcvr_abi_version(void);
const char* mpcvr_last_error(const mpcvr* vp);

/* --- media type: VerifyMediaType / InitMediaType / GetAlignmentSize, Source/VideoProcessor.h:178-181 */
int32_t mpcvr_verify_media_type(mpcvr* vp, const mpcvr_media_type* mt);              /* BOOL */
int32_t mpcvr_init_media_type(mpcvr* vp, const mpcvr_media_type* mt);                /* BOOL */
int32_t mpcvr_get_alignment_size(mpcvr* vp, const mpcvr_media_type* mt, int32_t* cx, int32_t* cy); /* BOOL; Source/DX11VideoProcessor.cpp:2055-2141 */

/* --- frames: ProcessSample / Render / FillBlack / Flush / Reset, Source/VideoProcessor.h:183-189 */
int32_t mpcvr_process_sample(mpcvr* vp, const mpcvr_sample* sample);                 /* CopySample + Render(1, rtStart) */
int32_t mpcvr_render(mpcvr* vp, int32_t field, int64_t frame_start_time);            /* re-run Process() on the current source */
int32_t mpcvr_fill_black(mpcvr* vp);
void    mpcvr_flush(mpcvr* vp);
int32_t mpcvr_reset(mpcvr* vp, int32_t display_mode_change);
int32_t mpcvr_wait_idle(mpcvr* vp);      /* cudaStreamSynchronize: the "Present" completion point */

/* --- geometry: SetVideoRect / SetWindowRect / SetRotation / SetFlip, Source/VideoProcessor.h:195-210 */
void    mpcvr_get_source_rect(mpcvr* vp, mpcvr_rect* r);
void    mpcvr_get_video_rect(mpcvr* vp, mpcvr_rect* r);
void    mpcvr_set_video_rect(mpcvr* vp, const mpcvr_rect* r);
int32_t mpcvr_set_window_rect(mpcvr* vp, const mpcvr_rect* r);
int32_t mpcvr_set_rotation(mpcvr* vp, int32_t degrees);       /* 0 / 90 / 180 / 270 clockwise (Source/DX11VideoProcessor.cpp:4052-4058); else E_INVALIDARG */
int32_t mpcvr_set_flip(mpcvr* vp, int32_t flip);              /* horizontal flip of the source (Source/VideoProcessor.h:210) */
/* m_bHdrPassthroughSupport (Source/DX11VideoProcessor.cpp:1476): the reference reads it from the DXGI output; here the
 * caller states whether the presentation target is an HDR10 (PQ / BT.2020) display.  With it set, bHdrPassthrough /
 * bHdrLocalToneMapping pick SHADER_CONVERT_NONE / _TO_PQ (:2948-2950) and the ps_hdr10_tonemap post-scale step
 * (:3359-3367) fed from mpcvr_sample.hdr / .hdr_cll (:2233-2268, :2715-2725). */
int32_t mpcvr_set_display_hdr(mpcvr* vp, int32_t hdr10_supported);
/* Path-B correction shaders for RGB-input callers.  In the reference these post-process what the vendor's
 * D3D11 VideoProcessor blit produced (a YCbCr->RGB conversion that knows neither BT.2020 nor PQ/HLG): m_pPSCorrection is
 * picked from the stream's colour tags (Source/DX11VideoProcessor.cpp:1893-1931) and drawn after the resize (:3354-3357).
 * Here the caller states that its RGB samples (RGB32 / r210 / RGB48 ...) come from such a fixed-function converter; the
 * ext_format of the media type then describes the ORIGINAL stream and selects ps_convert_pq_to_sdr (pq_supported = the
 * converter handled PQ, ID3D11VideoProcessor IsPqSupported) / ps_fixconvert_pq_to_sdr / ps_fixconvert_hlg_to_sdr /
 * ps_convert_hlg_to_pq / ps_fix_bt2020 as the reference does.  present = 0 restores the plain RGB path. */
int32_t mpcvr_set_upstream_converter(mpcvr* vp, int32_t present, int32_t pq_supported);
void    mpcvr_get_video_size(mpcvr* vp, int32_t* w, int32_t* h);                      /* GetVideoSize */
void    mpcvr_get_aspect_ratio(mpcvr* vp, int32_t* ax, int32_t* ay);                  /* GetAspectRatio */

/* --- settings: IVideoRenderer::GetSettings/SetSettings -> Configure, Source/IVideoRenderer.h:188-197,
 *     Source/DX11VideoProcessor.cpp:3800-4050 */
void    mpcvr_get_settings(mpcvr* vp, mpcvr_settings* s);
void    mpcvr_configure(mpcvr* vp, const mpcvr_settings* s);
int32_t mpcvr_get_active(mpcvr* vp);                                                  /* IVideoRenderer::GetActive */
int32_t mpcvr_set_procamp(mpcvr* vp, uint32_t flags, const mpcvr_procamp* values);
int32_t mpcvr_get_procamp(mpcvr* vp, uint32_t flags, mpcvr_procamp* values);

/* --- output.  The reference presents a DXGI swap chain back buffer (m_SwapChainFmt, window sized).
 * Here the back buffer is a device buffer owned by the instance; the caller may also bind its own. */
int32_t mpcvr_set_backbuffer_format(mpcvr* vp, int32_t fmt);                          /* b10BitOutput analogue, :1478 */
int32_t mpcvr_bind_backbuffer(mpcvr* vp, void* device_ptr, size_t pitch);             /* NULL = internal buffer */
int32_t mpcvr_get_backbuffer(mpcvr* vp, void** device_ptr, size_t* pitch, int32_t* w, int32_t* h, int32_t* fmt);
int32_t mpcvr_read_backbuffer(mpcvr* vp, void* host_dst, size_t dst_pitch);           /* D2H copy of the presented frame */
int32_t mpcvr_read_backbuffer_async(mpcvr* vp, void* pinned_host_dst, size_t dst_pitch); /* same, enqueued on the instance's stream; complete after mpcvr_wait_idle */
/* GetCurentImage: runs Process() at source-rect size into B8G8R8X8, top-down, row pitch = dst_pitch
 * (Source/DX11VideoProcessor.cpp:3493-3608; BITMAPINFOHEADER omitted, pixels only). */
int32_t mpcvr_get_current_image(mpcvr* vp, void* host_dst, size_t dst_pitch, int32_t* w, int32_t* h);
/* GetDisplayedImage (:3610): copy of what was last presented, as 32-bit BGRX regardless of back-buffer format */
int32_t mpcvr_get_displayed_image(mpcvr* vp, void* host_dst, size_t dst_pitch, int32_t* w, int32_t* h);
int32_t mpcvr_get_vp_info(mpcvr* vp, char* dst, size_t dst_size);                     /* GetVPInfo */

/* --- ISubRender11: Include/ISubRender11.h:30-47, invoked as at Source/DX11VideoProcessor.cpp:3262-3282.
 * Render11(rtStart, rtStop, avgTimePerFrame, croppedVideoRect, originalVideoRect, viewportRect,
 *          videoStretchFactor, xOffsetInPixels, flags) gains the CUDA stream and the target surface in
 * place of the D3D11 device context state; it runs on the frame's stream after the pixel kernels and
 * before the frame is signalled complete. SetDevice11(dev) becomes set_device(user, cuda_device|-1). */
typedef int32_t (*mpcvr_sub_render_fn)(void* user, int64_t rt_start, int64_t rt_stop, int64_t avg_time_per_frame,
                                       mpcvr_rect cropped_video_rect, mpcvr_rect original_video_rect,
                                       mpcvr_rect viewport_rect, double video_stretch_factor,
                                       int32_t x_offset_in_pixels, uint32_t flags,
                                       void* cuda_stream, void* device_target, size_t target_pitch,
                                       int32_t target_fmt);
typedef int32_t (*mpcvr_sub_set_device_fn)(void* user, int32_t cuda_device);
int32_t mpcvr_set_sub_callback(mpcvr* vp, mpcvr_sub_render_fn render, mpcvr_sub_set_device_fn set_device, void* user);

/* --- ID3D11DecoderConfiguration: Include/ID3DVideoMemoryConfiguration.h:20-29,
 *     impl Source/VideoRendererInputPin.cpp:186-202 -> CDX11VideoProcessor::SetDevice (:1254) */
int32_t mpcvr_activate_device_decoding(mpcvr* vp, int32_t cuda_device, void* decode_stream, uint32_t flags);
int32_t mpcvr_get_device_index(mpcvr* vp);                                            /* GetD3D11AdapterIndex */

/* --- measurement helpers (device events on the instance's stream) */
int32_t mpcvr_get_stream(mpcvr* vp, void** cuda_stream);
int32_t mpcvr_get_last_frame_kernels(mpcvr* vp, int32_t* launches);   /* kernels launched by the last Render() */

#ifdef __cplusplus
}
#endif
#endif /* MPCVR_B200_H */
