// frame_plan.h -- plain structs handed from the host-side processor to the CUDA kernels.
// One FramePlan describes what CDX11VideoProcessor::Process (Source/DX11VideoProcessor.cpp:3285-3424)
// would do for the current media type / rects / settings.
#pragma once
#include <cstdint>

namespace b200vr {

enum Transfer : int32_t {           // which body GetShaderConvertColor emits, Source/Shaders.cpp:861-923
    TRANSFER_NONE = 0,              // matrix only
    TRANSFER_PQ_TO_SDR = 1,         // :870-884
    TRANSFER_HLG_TO_SDR = 2,        // :862-868 then :870-884
    TRANSFER_SDR_2020 = 3,          // :892-915
    TRANSFER_HLG_TO_PQ = 4,         // :886-891 (SHADER_CONVERT_TO_PQ, HDR output)
};

struct TonemapParams {              // HDRParamsConstantBuffer, Shaders/d3d11/ps_hdr10_tonemap.hlsl:12-21 (after SetHDR10ShaderParams' clamps)
    float min_nits, max_nits, max_cll, max_fall, display;
    int32_t sel;                    // 1 ACES, 2 Reinhard, 3 Habel, 4 Moebius, 5 BT.2390, 6 ST 2094-10
    // DolbyConstants (cbuffer b1 of ps_hdr10_tonemap.hlsl:23-32), bound when the frame carries Dolby Vision L2 trims
    int32_t l2_enabled;
    float   l2_chroma_weight, l2_saturation_gain, l2_slope, l2_offset, l2_power;
};

struct SrcDesc {                    // the source "texture views" (Source/DX11Helper.h:87-170)
    const uint8_t* pl[3];           // Y | Y,UV | Y,U,V (U/V already ordered) | packed
    int32_t pitch[3];
    int32_t W, H;                   // luma size in pixels
    int32_t cw, ch;                 // chroma plane size in texels
    int32_t layout;                 // SrcLayout
    int32_t bytes;                  // per component
    int32_t lsb_shift;
    int32_t packed_kind;
    int32_t subsampling;            // 420/422/444
    int32_t chroma_mode;            // MPCVR_CHROMA_*
    int32_t siting;                 // ChromaSub_*
    int32_t deint_blend;
    int32_t rect_l, rect_t, rect_w, rect_h;   // m_srcRect
};

struct ColorParams {
    float cm[3][3];
    float cc[3];
    float gamut[3][3];
    float lum_scale;                // 10000 / iSDRDisplayNits
    float sdr_gamma;                // toLinear exponent of the BT.2020 SDR branch
    int32_t transfer;               // Transfer
    // compile-time constants of the HLSL (fxc folds them in fp32); evaluated once on the host in the same fp32 order
    float inv_m1, inv_m2;           // 1.0/ST2084_m1, 1.0/ST2084_m2                      (st2084.hlsl:11,13)
    float inv_gamma22;              // 1.0/2.2                                           (Source/Shaders.cpp:921)
    float hable_cb, hable_de, hable_df, hable_ef, hable_rdiv;   // C*B, D*E, D*F, E/F, 1/hable(4.8)   (hdr_tone_mapping.hlsl:3-12)
    float hlg_ra;                   // 1/B67_a                                           (hlg.hlsl:9)
    float r1000;                    // 1/1000 (LinearToST2084 divider)                   (Source/Shaders.cpp:866)
};

struct DoviParams {                 // PS_DOVI_CURVE x3 (Source/Shaders.cpp:715-733) + the LMS matrix literal (:823-840)
    float   pivots[3][7];
    float   coeffs[3][8][4];
    float   mmr[3][48][4];
    int32_t methods[3], mmr_single[3], min_order[3], max_order[3];
    float   mat[3][3];
    int32_t has_mmr;
    // DolbyConstants (Source/Shaders.cpp:775-786; SetDolbyVisionDynamicParams Source/DX11VideoProcessor.cpp:953-960)
    int32_t l2_enabled;
    float   l2_chroma_weight, l2_saturation_gain, l2_slope, l2_offset, l2_power;
};

enum Correction : int32_t {         // m_pPSCorrection (Source/DX11VideoProcessor.cpp:1893-1931); the files under Shaders/d3d11/
    CORR_NONE = 0, CORR_CONVERT_PQ_TO_SDR = 1, CORR_FIXCONVERT_PQ_TO_SDR = 2, CORR_FIXCONVERT_HLG_TO_SDR = 3,
    CORR_FIX_BT2020 = 4, CORR_CONVERT_HLG_TO_PQ = 5, CORR_FIX_YCGCO = 6,
};
struct CorrectionParams {           // compile-time matrices of the correction shaders, folded on the host
    float fix2020[3][3];            // mul(ycbcr2020nc_rgb, rgb_ycbcr709)           (conv_matrix.hlsl, ps_fix_bt2020.hlsl:10)
    float fixycgco[3][3];           // mul(ycgco_rgb, rgb_ycbcr709)                 (ps_fix_ycgco.hlsl:6)
    float gamut[3][3];              // convert_matrix_2020_to_709                   (colorspace_gamut_conversion.hlsl:86-88)
    int32_t mode;
};

struct SurfDesc {                   // a render target / shader resource of the pass chain
    void*   ptr;
    int32_t pitch;                  // bytes
    int32_t w, h;
    int32_t fmt;                    // mpcvr_surface_format
};

enum ScalerKind : int32_t { SCALER_NONE = 0, SCALER_INTERP = 1, SCALER_CONV = 2, SCALER_NEAREST = 3 };

struct AxisTable {                  // per output index i of one resize axis (TextureResizeShader, :332-377)
    const int32_t* base;            // first source texel of the tap window
    const float*   w;               // [dst_len][ntaps] weights (zero padded)
    const float*   wsum;            // convolution: sum of weights (divide), else unused
    int32_t ntaps;
    int32_t kind;                   // ScalerKind
    int32_t dup_q1;                 // D3D11 ps_interpolation_lanczos3.hlsl:33-34: tap 1 re-reads tap 0's texel
    int32_t dst0, dst_len;          // destination range on this axis
    int32_t src0, src_len;          // source rect on this axis (for the point-sampled case)
};

}  // namespace b200vr
