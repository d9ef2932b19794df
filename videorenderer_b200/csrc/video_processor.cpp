// video_processor.cpp -- see video_processor.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include "video_processor.h"

namespace b200vr {

int PlaneLoadBytes(const FmtConvParams& f, int plane);

static inline int AlignUp(int v, int a) { return (v + a - 1) / a * a; }
static inline int SurfBpp(int fmt) { return (fmt == MPCVR_SURF_RGBA16F || fmt == MPCVR_SURF_RGBA16) ? 8 : (fmt == MPCVR_SURF_RGB8 ? 3 : 4); }

// ------------------------------------------------------------------------------------------------
// Tex2D / DeviceAxis
// ------------------------------------------------------------------------------------------------
void Tex2D::Release()
{
    if (ptr && owned) cudaFree(ptr);
    ptr = nullptr; pitch = 0; w = h = 0; bytes = 0; owned = true;
}

cudaError_t Tex2D::CheckCreate(int f, int width, int height)
{
    if (ptr && owned && fmt == f && w == width && h == height) return cudaSuccess;
    Release();
    if (width <= 0 || height <= 0) return cudaErrorInvalidValue;
    const size_t p = (size_t)AlignUp(width * SurfBpp(f), 256);
    cudaError_t e = cudaMalloc(&ptr, p * height);
    if (e != cudaSuccess) { ptr = nullptr; return e; }
    pitch = p; w = width; h = height; fmt = f; bytes = p * height; owned = true;
    return cudaSuccess;
}

AxisTable DeviceAxis::table() const
{
    AxisTable t;
    t.base = base.ptr; t.w = w.ptr; t.wsum = wsum.ptr;
    t.ntaps = host.ntaps; t.kind = host.kind; t.dup_q1 = host.dup_q1 ? 1 : 0;
    t.dst0 = host.dst0; t.dst_len = host.dst_len; t.src0 = host.src0; t.src_len = host.src_len;
    return t;
}

cudaError_t DeviceAxis::Upload(cudaStream_t st)
{
    cudaError_t e = base.Upload(host.base, st);
    if (e == cudaSuccess) e = w.Upload(host.w, st);
    if (e == cudaSuccess) e = wsum.Upload(host.wsum, st);
    return e;
}

// ------------------------------------------------------------------------------------------------
CB200VideoProcessor::CB200VideoProcessor() { mpcvr_settings_default(&m_settings); ApplySettings(m_settings); }

CB200VideoProcessor::~CB200VideoProcessor()
{
    if (m_device >= 0) {
        cudaSetDevice(m_device);
        if (m_stream) cudaStreamSynchronize(m_stream);
        for (auto& t : m_TexSrcVideo) t.Release();
        m_TexConvertOutput.Release(); m_TexSrcUnpacked.Release(); m_TexResize.Release(); m_TexsPostScale.Release(); m_TexsPostScale2.Release(); m_BackBuffer.Release(); m_TexImage.Release();
        m_dbgConvertF32.Release(); m_doviDev.Release();
        m_axisX.Release(); m_axisY.Release(); m_pointX.Release(); m_pointY.Release(); m_pointRX.Release(); m_pointRY.Release(); m_oFilt.Release(); m_oU.Release(); m_oV.Release();
        if (m_ownStream && m_stream) cudaStreamDestroy(m_stream);
    }
}

int32_t CB200VideoProcessor::Fail(int32_t hr, const char* what, cudaError_t e)
{
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    else snprintf(buf, sizeof(buf), "%s", what);
    m_lastError = buf;
    return hr;
}

bool CB200VideoProcessor::Cuda(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    Fail(MPCVR_E_FAIL, what, e);
    return false;
}

void CB200VideoProcessor::ApplySettings(const mpcvr_settings& s)
{
    m_settings = s;
    m_bUseD3D11 = s.bUseD3D11 != 0;
    m_iTexFormat = s.iTexFormat;
    m_iChromaScaling = s.iChromaScaling;
    m_iUpscaling = s.iUpscaling;
    m_iDownscaling = s.iDownscaling;
    m_bInterpolateAt50pct = s.bInterpolateAt50pct != 0;
    m_bUseDither = s.bUseDither != 0;
    m_bDeintBlend = s.bDeintBlend != 0;
    m_bConvertToSdr = s.bConvertToSdr != 0;
    m_iSDRDisplayNits = s.iSDRDisplayNits;
    if (m_iSDRDisplayNits < MPCVR_SDR_NITS_MIN || m_iSDRDisplayNits > MPCVR_SDR_NITS_MAX) m_iSDRDisplayNits = MPCVR_SDR_NITS_DEF;
    m_bHdrPreferDoVi = s.bHdrPreferDoVi != 0;
    m_bHdrPassthrough = s.bHdrPassthrough != 0;
    m_bHdrLocalToneMapping = s.bHdrLocalToneMapping != 0;
    m_iHdrLocalToneMappingType = s.iHdrLocalToneMappingType;
    m_iHdrDisplayMaxNits = s.iHdrDisplayMaxNits;
}

int32_t CB200VideoProcessor::Init(const mpcvr_settings* s, int device, void* stream)
{
    if (s) ApplySettings(*s);
    m_device = device;
    if (device < 0) { m_device = -1; return MPCVR_S_OK; }         // planning-only instance
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device >= count) {
        m_device = -1;
        return Fail(MPCVR_E_FAIL, "mpcvr_create: no such CUDA device (this library has no CPU path)", e);
    }
    if (!Cuda(cudaSetDevice(device), "cudaSetDevice")) { m_device = -1; return MPCVR_E_FAIL; }
    if (stream) { m_stream = (cudaStream_t)stream; m_ownStream = false; }
    else {
        if (!Cuda(cudaStreamCreateWithFlags(&m_stream, cudaStreamNonBlocking), "cudaStreamCreate")) { m_device = -1; return MPCVR_E_FAIL; }
        m_ownStream = true;
    }
    if (m_subSetDevice) m_subSetDevice(m_subUser, m_device);
    return MPCVR_S_OK;
}

// ------------------------------------------------------------------------------------------------
// media type
// ------------------------------------------------------------------------------------------------
int32_t CB200VideoProcessor::VerifyMediaType(const mpcvr_media_type* mt)
{
    if (!mt) return 0;
    const FmtConvParams& p = GetFmtConvParams(mt->cformat);
    if (p.cformat == MPCVR_CF_NONE || !FormatSupported(p)) return 0;      // VerifyMediaType: known format + a usable VP
    if (mt->width <= 0 || mt->height <= 0) return 0;
    return 1;
}

int32_t CB200VideoProcessor::InitMediaType(const mpcvr_media_type* mt)
{
    if (!VerifyMediaType(mt)) return 0;
    const FmtConvParams FmtParams = GetFmtConvParams(mt->cformat);

    m_decExFmt.value = 0;
    m_Dovi.bValid = false; m_Dovi.bHasMMR = false;                          // m_Dovi = {} on a new media type
    m_srcRect = CRect(mt->src_rect.left, mt->src_rect.top, mt->src_rect.right, mt->src_rect.bottom);
    m_srcAspectRatioX = mt->aspect_x; m_srcAspectRatioY = mt->aspect_y;
    m_upstreamExFmt.value = (FmtParams.CSType == CS_RGB) ? mt->ext_format : 0u;   // used only with mpcvr_set_upstream_converter
    if (FmtParams.CSType == CS_YUV && mt->ext_format) {
        // m_decExFmt.value = dwControlFlags; SampleFormat = AMCONTROL_USED | AMCONTROL_COLORINFO_PRESENT  (:1765-1768)
        m_decExFmt.value = (mt->ext_format & ~0xFFu) | 0x3u;
    }
    m_bInterlaced = mt->interlaced != 0;
    m_rtAvgTimePerFrame = mt->avg_time_per_frame;

    const unsigned biWidth = (unsigned)mt->width, biHeight = (unsigned)std::abs(mt->height);
    m_srcLines = biHeight * FmtParams.PitchCoeff / 2;                       // :1789-1803
    m_srcPitch = biWidth * FmtParams.Packsize;
    switch (FmtParams.cformat) {
    case MPCVR_CF_Y8: case MPCVR_CF_NV12: case MPCVR_CF_RGB24: case MPCVR_CF_BGR48:
        m_srcPitch = AlignUp(m_srcPitch, 4); break;
    case MPCVR_CF_V210:
        m_srcPitch = AlignUp((biWidth + 5) / 6 * 16, 128); break;
    default: break;
    }
    if (mt->bottom_up) m_srcPitch = -m_srcPitch;                            // BI_RGB && biHeight > 0, :1801-1803
    unsigned origW = biWidth, origH = biHeight;                             // VR_Extradata, :1805-1813
    if (mt->frame_width > 0 && mt->frame_height != 0) { origW = (unsigned)mt->frame_width; origH = (unsigned)std::abs(mt->frame_height); }
    if (origW > biWidth || origH > biHeight) return 0;                      // VR_Extradata names the frame INSIDE the (pitch-aligned) buffer

    if (m_srcRect.IsRectNull()) m_srcRect = CRect(0, 0, (int)origW, (int)origH);
    if (m_srcRect.left < 0 || m_srcRect.top < 0 || m_srcRect.right > (int)origW || m_srcRect.bottom > (int)origH || m_srcRect.IsRectEmpty())
        return 0;
    m_srcRectWidth = m_srcRect.Width(); m_srcRectHeight = m_srcRect.Height();

    m_srcExFmt = SpecifyExtendedFormat(m_decExFmt, FmtParams, m_srcRectWidth, m_srcRectHeight);

    const unsigned g = std::gcd(m_srcRectWidth, m_srcRectHeight);           // :1842-1856
    const unsigned srcFrameARX = m_srcRectWidth / g, srcFrameARY = m_srcRectHeight / g;
    if (!m_srcAspectRatioX || !m_srcAspectRatioY) { m_srcAspectRatioX = srcFrameARX; m_srcAspectRatioY = srcFrameARY; m_srcAnamorphic = false; }
    else {
        const unsigned ag = std::gcd(m_srcAspectRatioX, m_srcAspectRatioY);
        m_srcAspectRatioX /= ag; m_srcAspectRatioY /= ag;
        m_srcAnamorphic = (srcFrameARX != m_srcAspectRatioX || srcFrameARY != m_srcAspectRatioY);
    }

    UpdateTexParams(FmtParams.CDepth);

    // InitializeTexVP, :2018-2047: the source texture is origW x origH
    m_srcWidth = origW; m_srcHeight = origH; m_srcParams = FmtParams;
    m_haveFrame = false;
    if (m_device >= 0) {
        cudaSetDevice(m_device);
        for (auto& t : m_TexSrcVideo) t.Release();
        // plane buffers (bytes addressed as R8 "textures" of the right byte width)
        int nplanes = 1, pw[3] = {0, 0, 0}, ph[3] = {0, 0, 0};
        const int bpc = FmtParams.bytes;
        switch (FmtParams.layout) {
        case LAYOUT_BIPLANAR:
            nplanes = 2; pw[0] = origW * bpc; ph[0] = origH;
            pw[1] = (origW / FmtParams.div_chroma_w) * 2 * bpc; ph[1] = origH / FmtParams.div_chroma_h; break;
        case LAYOUT_PLANAR3:
            nplanes = 3; pw[0] = origW * bpc; ph[0] = origH;
            pw[1] = pw[2] = (origW / FmtParams.div_chroma_w) * bpc; ph[1] = ph[2] = origH / FmtParams.div_chroma_h; break;
        case LAYOUT_PACKED422:
            pw[0] = (FmtParams.packed_kind == 2) ? std::abs(m_srcPitch) : (int)(origW / 2) * 4 * bpc; ph[0] = origH; break;   // v210: whole 16-byte groups
        default:               pw[0] = origW * FmtParams.Packsize; ph[0] = origH; break;
        }
        for (int i = 0; i < nplanes; i++) {
            Tex2D& t = m_TexSrcVideo[i];
            const size_t p = (size_t)AlignUp(pw[i], 256);
            if (!Cuda(cudaMalloc(&t.ptr, p * ph[i]), "cudaMalloc(m_TexSrcVideo)")) return 0;
            t.pitch = p; t.w = pw[i]; t.h = ph[i]; t.bytes = p * ph[i]; t.owned = true;
            cudaMemsetAsync(t.ptr, 0, t.bytes, m_stream);
        }
    }
    SetShaderConvertColorParams();
    SetShaderLuminanceParams();
    UpdateCorrection();
    m_bInit = true;
    m_planDirty = true;
    if (m_videoRect.IsRectEmpty()) m_videoRect = CRect(0, 0, (int)m_srcRectWidth, (int)m_srcRectHeight);
    if (m_windowRect.IsRectEmpty()) m_windowRect = m_videoRect;
    UpdateRenderRect();
    if (UpdateTexures() != MPCVR_S_OK) return 0;
    if (UpdatePostScaleTexures() != MPCVR_S_OK) return 0;
    return 1;
}

int32_t CB200VideoProcessor::GetAlignmentSize(const mpcvr_media_type* mt, int32_t* cx, int32_t* cy)
{
    // The reference returns RowPitch/Packsize of a mapped probe texture so that the upload is one memcpy
    // (:2055-2141).  Here the preferred row pitch is a multiple of 256 bytes (coalesced 128-bit loads, TMA
    // tensor-map stride granularity).
    if (!VerifyMediaType(mt) || !cx || !cy) return 0;
    const FmtConvParams& p = GetFmtConvParams(mt->cformat);
    *cx = AlignUp(mt->width * p.Packsize, 256) / p.Packsize;
    *cy = std::abs(mt->height);
    return 1;
}

void CB200VideoProcessor::UpdateTexParams(int cdepth)
{
    switch (m_iTexFormat) {
    case MPCVR_TEXFMT_8INT:    m_InternalTexFmt = MPCVR_SURF_BGRA8; break;
    case MPCVR_TEXFMT_10INT:   m_InternalTexFmt = MPCVR_SURF_RGB10A2; break;
    case MPCVR_TEXFMT_16FLOAT: m_InternalTexFmt = MPCVR_SURF_RGBA16F; break;
    case MPCVR_TEXFMT_AUTOINT:
    default:                   m_InternalTexFmt = (cdepth > 8) ? MPCVR_SURF_RGB10A2 : MPCVR_SURF_BGRA8; break;
    }
}

int CB200VideoProcessor::Transfer() const
{
    // GetShaderConvertColor, Source/Shaders.cpp:613-616, :861-915; convertType per UpdateConvertColorShader (:2948-2950)
    // with no HDR display: SHADER_CONVERT_TO_SDR iff m_bConvertToSdr.
    // HDR output: SHADER_CONVERT_TO_PQ for HLG, SHADER_CONVERT_NONE otherwise (:2948-2950)
    const uint32_t trc = m_srcExFmt.TransferFunction();
    const bool toSdr = m_bConvertToSdr && !HdrOut();
    // bConvertHDRtoSDR also holds for Dolby Vision (the reshaped signal is PQ); bApplyHLG excludes it (Shaders.cpp:614-615)
    if (toSdr && (trc == Trc_2084 || m_Dovi.bValid)) return TRANSFER_PQ_TO_SDR;
    if (toSdr && trc == Trc_HLG)  return TRANSFER_HLG_TO_SDR;
    if (HdrOut() && trc == Trc_HLG && !m_Dovi.bValid) return TRANSFER_HLG_TO_PQ;
    if (m_srcExFmt.Primaries() == Prim_BT2020) {
        switch (trc) {
        case Trc_18: case Trc_20: case Trc_HLG: case Trc_22: case Trc_709: case Trc_240M: case Trc_sRGB: case Trc_28: case Trc_26:
            return TRANSFER_SDR_2020;
        default: break;   // incl. TransFunc_10: the generated HLSL does not compile ("\\nothing", :895) -> matrix-only fallback
        }
    }
    return TRANSFER_NONE;
}

void CB200VideoProcessor::SetShaderConvertColorParams()
{
    BuildColorTransform(m_srcExFmt, m_srcParams, m_ProcAmp, m_PSConvColorData);
    if (m_Dovi.bValid) {                                                    // :817-834
        const float brightness = m_ProcAmp.brightness / 255, contrast = m_ProcAmp.contrast;
        float (&m)[3][3] = m_PSConvColorData.m; float (&c)[3] = m_PSConvColorData.c;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = (float)m_Dovi.msd.ycc_to_rgb_matrix[i * 3 + j] * contrast;
        for (int i = 0; i < 3; i++) {
            c[i] = brightness;
            for (int j = 0; j < 3; j++) c[i] = (float)((double)c[i] - (double)m[i][j] * m_Dovi.msd.ycc_to_rgb_offset[j]);
        }
        m_PSConvColorData.enable = true;
    }
    {
        float (&m)[3][3] = m_PSConvColorData.m;
        if (m_srcParams.CSType == CS_RGB && m_srcParams.layout == LAYOUT_PLANAR3) {          // GBRP: planes are G, B, R (:863-867)
            for (int i = 0; i < 3; i++) { std::swap(m[i][0], m[i][1]); std::swap(m[i][1], m[i][2]); }
            m_PSConvColorData.enable = true;                                                  // :848
        } else if (m_srcParams.CSType == CS_GRAY) {                                          // :868-873
            m[1][0] = m[1][1]; m[1][1] = 0;
            m[2][0] = m[2][2]; m[2][2] = 0;
        }
    }
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) m_color.cm[i][j] = m_PSConvColorData.m[i][j]; m_color.cc[i] = m_PSConvColorData.c[i]; }
    BuildGamutMatrix(CHR_BT2020, CHR_BT709, m_color.gamut);               // Source/Shaders.cpp:634-645
    m_color.transfer = Transfer();
    float g = 1.0f;
    switch (m_srcExFmt.TransferFunction()) {                                // Source/Shaders.cpp:894-905
    case Trc_18: g = 1.8f; break; case Trc_20: g = 2.0f; break; case Trc_28: g = 2.8f; break; case Trc_26: g = 2.6f; break;
    default: g = 2.2f; break;
    }
    m_color.sdr_gamma = g;
    // shader literals, folded in fp32 like fxc does (host code is built with -ffp-contract=off; fmaf = the shader's mad)
    m_color.inv_m1 = 1.0f / 0.1593017578125f; m_color.inv_m2 = 1.0f / 78.84375f; m_color.inv_gamma22 = 1.0f / 2.2f;
    {
        const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
        m_color.hable_cb = C * B; m_color.hable_de = D * E; m_color.hable_df = D * F; m_color.hable_ef = E / F;
        const float x = 4.8f;
        const float hdiv = std::fmaf(x, std::fmaf(A, x, m_color.hable_cb), m_color.hable_de) / std::fmaf(x, std::fmaf(A, x, B), m_color.hable_df) - m_color.hable_ef;
        m_color.hable_rdiv = 1.0f / hdiv;
    }
    m_color.hlg_ra = 1.0f / 0.17883277f; m_color.r1000 = 1.0f / 1000.0f;
    m_planDirty = true;
}

void CB200VideoProcessor::SetShaderLuminanceParams()
{
    m_color.lum_scale = 10000.0f / m_iSDRDisplayNits;                       // :891
}

void CB200VideoProcessor::UpdateRenderRect() { m_renderRect.IntersectRect(m_videoRect, m_windowRect); }

int32_t CB200VideoProcessor::UpdateTexures()
{
    if (!m_srcWidth || !m_srcHeight) return MPCVR_S_OK;
    m_planDirty = true;
    if (m_device < 0) return MPCVR_S_OK;
    cudaSetDevice(m_device);
    if (!Cuda(m_TexConvertOutput.CheckCreate(m_InternalTexFmt, m_srcRectWidth, m_srcRectHeight), "m_TexConvertOutput")) return MPCVR_E_FAIL;
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::UpdatePostScaleTexures()
{
    const bool needDither =
        ((m_SwapChainFmt == MPCVR_SURF_BGRA8 || m_SwapChainFmt == MPCVR_SURF_RGB8) && m_InternalTexFmt != MPCVR_SURF_BGRA8)
        || (m_SwapChainFmt == MPCVR_SURF_RGB10A2 && m_InternalTexFmt == MPCVR_SURF_RGBA16F);
    m_bFinalPass = m_bUseDither && needDither;
    m_planDirty = true;
    if (m_device < 0 || m_windowRect.IsRectEmpty()) return MPCVR_S_OK;
    cudaSetDevice(m_device);
    if (m_bFinalPass) {
        if (!Cuda(m_TexsPostScale.CheckCreate(m_InternalTexFmt, m_windowRect.Width(), m_windowRect.Height()), "m_TexsPostScale")) return MPCVR_E_FAIL;
    } else m_TexsPostScale.Release();
    if (!Cuda(m_BackBuffer.CheckCreate(m_SwapChainFmt, m_windowRect.Width(), m_windowRect.Height()), "back buffer")) return MPCVR_E_FAIL;
    return MPCVR_S_OK;
}

void CB200VideoProcessor::SetVideoRect(const CRect& r)
{
    m_videoRect = r;
    UpdateRenderRect();
    UpdateTexures();
}

int32_t CB200VideoProcessor::SetWindowRect(const CRect& r)
{
    m_windowRect = r;
    UpdateRenderRect();
    return UpdatePostScaleTexures();
}

int32_t CB200VideoProcessor::SetRotation(int v)
{
    if (v != 0 && v != 90 && v != 180 && v != 270) return Fail(MPCVR_E_INVALIDARG, "rotation must be 0, 90, 180 or 270");
    m_iRotation = v; m_planDirty = true;                                    // Source/DX11VideoProcessor.cpp:4052-4058
    return MPCVR_S_OK;
}
int32_t CB200VideoProcessor::SetFlip(bool v)
{
    m_bFlip = v; m_planDirty = true;                                        // Source/VideoProcessor.h:210
    return MPCVR_S_OK;
}

bool CB200VideoProcessor::CheckDoviMetadata(const mpcvr_dovi_metadata* p) const
{
    if (p->bl_bit_depth < 8 || p->bl_bit_depth > 16 || p->coef_log2_denom > 31) return false;
    // "ignore Profile 4" (Source/VideoProcessor.cpp:275-279)
    if (p->el_spatial_resampling_filter_flag && !p->disable_residual_flag && p->vdr_bit_depth != 12) return false;
    // "Ignore Profile 7 Full Enhancement Layer" (SourceIsDoviFel, Source/VideoProcessor.h:249-263)
    if (!p->disable_residual_flag && p->nlq_method_idc == 0) {
        for (int i = 0; i < 3; i++)
            if (p->nlq[i].nlq_offset != 0 || p->nlq[i].vdr_in_max != (1ULL << p->coef_log2_denom) ||
                p->nlq[i].linear_deadzone_slope != 0 || p->nlq[i].linear_deadzone_threshold != 0) return false;
    }
    for (const auto& curve : p->curves) {
        if (curve.num_pivots < 2 || curve.num_pivots > 9) return false;
        for (int i = 0; i < (int)curve.num_pivots - 1; i++) {
            if (curve.mapping_idc[i] > 1) return false;                     // 0 polynomial, 1 mmr
            if (curve.mapping_idc[i] == 1 && (curve.mmr_order[i] < 1 || curve.mmr_order[i] > 3)) return false;
        }
    }
    return true;
}

int32_t CB200VideoProcessor::SetShaderDoviCurves()
{
    // SetShaderDoviCurves / SetShaderDoviCurvesPoly (:990-1141) and the matrix literal of GetShaderConvertColor
    // (Source/Shaders.cpp:823-840: dovi_lms2rgb x rgb_to_lms through mul_matrix3x3, Source/csputils.cpp:531-538)
    const mpcvr_dovi_metadata& dv = m_Dovi.msd;
    DoviParams& P = m_doviHost;
    memset(&P, 0, sizeof(P));
    bool any_mmr = false;
    for (const auto& cv : dv.curves) for (int i = 0; i < cv.num_pivots - 1; i++) if (cv.mapping_idc[i] == 1) any_mmr = true;
    m_Dovi.bHasMMR = any_mmr; P.has_mmr = any_mmr ? 1 : 0;
    const float scale = 1.0f / (float)((1 << dv.bl_bit_depth) - 1);
    const float scale_coef = 1.0f / (float)(1u << dv.coef_log2_denom);
    for (int c = 0; c < 3; c++) {
        const mpcvr_dovi_curve& cv = dv.curves[c];
        bool has_poly = false, has_mmr = false, mmr_single = true;
        int mmr_idx = 0, min_order = 3, max_order = 1;
        for (int i = 0; i < cv.num_pivots - 1; i++) {
            float* co = P.coeffs[c][i];
            if (cv.mapping_idc[i] == 0) {
                has_poly = true;
                co[0] = scale_coef * (float)cv.poly_coef[i][0];
                co[1] = cv.poly_order[i] >= 1 ? scale_coef * (float)cv.poly_coef[i][1] : 0.0f;
                co[2] = cv.poly_order[i] >= 2 ? scale_coef * (float)cv.poly_coef[i][2] : 0.0f;
                co[3] = 0.0f;                                                // order 0 signals polynomial
            } else {
                min_order = std::min<int>(min_order, cv.mmr_order[i]); max_order = std::max<int>(max_order, cv.mmr_order[i]);
                mmr_single = !has_mmr; has_mmr = true;
                co[0] = scale_coef * (float)cv.mmr_constant[i]; co[1] = (float)mmr_idx; co[3] = (float)cv.mmr_order[i];
                for (int j = 0; j < cv.mmr_order[i]; j++) {                  // weights per order as two packed float4s
                    for (int k = 0; k < 3; k++) P.mmr[c][mmr_idx][k] = scale_coef * (float)cv.mmr_coef[i][j][k];
                    mmr_idx++;
                    for (int k = 0; k < 4; k++) P.mmr[c][mmr_idx][k] = scale_coef * (float)cv.mmr_coef[i][j][3 + k];
                    mmr_idx++;
                }
            }
        }
        const int n = cv.num_pivots - 2;
        for (int i = 0; i < 7; i++) P.pivots[c][i] = i < n ? scale * (float)cv.pivots[i + 1] : 1e9f;
        P.methods[c] = (has_poly ? 1 : 0) | (has_mmr ? 2 : 0);
        if (has_mmr) { P.mmr_single[c] = mmr_single; P.min_order[c] = min_order; P.max_order[c] = max_order; }
    }
    static const float lms2rgb[3][3] = { { 3.06441879f, -2.16597676f,  0.10155818f}, {-0.65612108f,  1.78554118f, -0.12943749f},
                                         { 0.01736321f, -0.04725154f,  1.03004253f} };
    float lin[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) lin[i][j] = (float)dv.rgb_to_lms_matrix[i * 3 + j];
    for (int r = 0; r < 3; r++)
        for (int i = 0; i < 3; i++) {
            float v = lms2rgb[r][0] * lin[0][i]; v = v + lms2rgb[r][1] * lin[1][i]; v = v + lms2rgb[r][2] * lin[2][i];
            P.mat[r][i] = v;
        }
    // Level 2 trims for this display (:2383-2466) -> DolbyConstants (:953-960).  Host arithmetic as written there (libm powf);
    // std::lerp spelled out so that the value does not depend on the C++ library at hand.
    {
        auto lerpf = [](float a, float b, float t) {
            if ((a <= 0 && b >= 0) || (a >= 0 && b <= 0)) return t * b + (1 - t) * a;
            if (t == 1) return b;
            const float x = a + t * (b - a);
            return ((t > 1) == (b > a)) ? (b < x ? x : b) : (b > x ? x : b);
        };
        const float PQ_M1 = 2610.f / (4096.f * 4.f), PQ_M2 = 2523.f / 4096.f * 128.f, PQ_C1 = 3424.f / 4096.f,
                    PQ_C2 = 2413.f / 4096.f * 32.f, PQ_C3 = 2392.f / 4096.f * 32.f;
        float y = (float)m_iHdrDisplayMaxNits; y /= 10000.0f; y = fmaxf(y, 0.0f); y = powf(y, PQ_M1);
        y = (PQ_C1 + PQ_C2 * y) / (1.0f + PQ_C3 * y);
        const float display_pq = powf(y, PQ_M2);
        int lower = -1, upper = -1; bool present = false; float dl = 1.0f, du = 1.0f;
        for (int i = 0; i < 32; i++) {
            if (dv.ext[i].level != 2) continue;
            present = true;
            const float t = dv.ext[i].v[0] / 4095.0f;
            if (t <= display_pq) { const float d = display_pq - t; if (d < dl) { dl = d; lower = i; } }
            else                 { const float d = t - display_pq; if (d < du) { du = d; upper = i; } }
        }
        if (present) {
            float tv[5] = { 1.0f, 0.0f, 1.0f, 0.0f, 0.0f };                 // slope, offset, power, chroma weight, saturation gain
            auto clamp01 = [](float w) { return w < 0.0f ? 0.0f : (w > 1.0f ? 1.0f : w); };
            if (lower != -1 && upper != -1) {
                const float lp = dv.ext[lower].v[0] / 4095.0f, up = dv.ext[upper].v[0] / 4095.0f;
                const float w = clamp01((up != lp) ? (display_pq - lp) / (up - lp) : 0.0f);
                for (int k = 0; k < 5; k++) tv[k] = lerpf((float)dv.ext[lower].v[1 + k], (float)dv.ext[upper].v[1 + k], w);
            } else if (lower != -1) {
                const float mp = dv.source_max_pq / 4095.0f, lp = dv.ext[lower].v[0] / 4095.0f;
                const float w = clamp01((mp > lp) ? (display_pq - lp) / (mp - lp) : 0.0f);
                for (int k = 0; k < 5; k++) tv[k] = lerpf((float)dv.ext[lower].v[1 + k], 2048.0f, w);
            } else if (upper != -1) {
                for (int k = 0; k < 5; k++) tv[k] = (float)dv.ext[upper].v[1 + k];
            }
            m_doviL2Present = true;                                       // m_DoviExtensionMetadata.L2.present = true, :2463 -- never cleared per frame
            m_doviL2[0] = tv[0] / 4096.0f + 0.5f; m_doviL2[1] = tv[1] / 4096.0f - 0.5f; m_doviL2[2] = tv[2] / 4096.0f + 0.5f;
            m_doviL2[3] = tv[3] / 4096.0f - 0.5f; m_doviL2[4] = tv[4] / 4096.0f - 0.5f;
        }
        if (m_doviL2Present) {                                           // SetDolbyVisionDynamicParams, :953-960: the last trims seen
            P.l2_enabled = 1;
            P.l2_slope = m_doviL2[0]; P.l2_offset = m_doviL2[1]; P.l2_power = m_doviL2[2];
            P.l2_chroma_weight = m_doviL2[3]; P.l2_saturation_gain = m_doviL2[4];
        }
    }
    // Level 1 (+ Level 3 offsets) -> nits (:2347-2381): feeds SetHDR10ShaderParams when local tone mapping is on
    for (int i = 0; i < 32; i++) {                                           // the first Level 1 block; none: the previous values stay (sticky)
        if (dv.ext[i].level != 1) continue;
        uint16_t v[3] = { dv.ext[i].v[0], dv.ext[i].v[1], dv.ext[i].v[2] };
        for (int k = 0; k < 32; k++)
            if (dv.ext[k].level == 3) { for (int j = 0; j < 3; j++) v[j] = (uint16_t)(v[j] + dv.ext[k].v[j] - 2048); break; }
        const float PQ_M1 = 2610.f / (4096.f * 4.f), PQ_M2 = 2523.f / 4096.f * 128.f, PQ_C1 = 3424.f / 4096.f,
                    PQ_C2 = 2413.f / 4096.f * 32.f, PQ_C3 = 2392.f / 4096.f * 32.f;
        for (int j = 0; j < 3; j++) {
            float x = v[j] / 4095.f;
            x = powf(x, 1.0f / PQ_M2);
            x = fmaxf(x - PQ_C1, 0.0f) / (PQ_C2 - PQ_C3 * x);
            x = powf(x, 1.0f / PQ_M1);
            m_doviL1Nits[j] = (float)(uint16_t)(unsigned int)(x * 10000.0f);
        }
        m_doviL1Present = true;
        break;
    }
    if (m_device >= 0) {
        cudaSetDevice(m_device);
        if (!m_doviDev.ptr) { if (!Cuda(cudaMalloc(&m_doviDev.ptr, sizeof(DoviParams)), "cudaMalloc(dovi)")) return MPCVR_E_FAIL; m_doviDev.n = 1; }
        if (!Cuda(cudaMemcpyAsync(m_doviDev.ptr, &P, sizeof(P), cudaMemcpyHostToDevice, m_stream), "H2D dovi")) return MPCVR_E_FAIL;
    }
    m_planDirty = true;
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::SetUpstreamConverter(bool present, bool pqSupported)
{
    m_upstreamVP = present; m_upstreamPq = pqSupported;
    UpdateCorrection();
    m_planDirty = true;
    return MPCVR_S_OK;
}

void CB200VideoProcessor::UpdateCorrection()
{
    // shader choice of InitMediaType, :1893-1931, on the tags of the original stream
    int mode = CORR_NONE;
    if (m_upstreamVP && m_srcParams.CSType == CS_RGB && m_upstreamExFmt.value) {
        const uint32_t trc = m_upstreamExFmt.TransferFunction(), prim = m_upstreamExFmt.Primaries();
        const bool bTransFunc22 = (trc == Trc_22 || trc == Trc_709 || trc == Trc_240M);
        if (!m_bUseD3D11) {                                   // the D3D9 twin's rule, Source/DX9VideoProcessor.cpp:1225-1238
            if (trc == Trc_2084 && m_bConvertToSdr) mode = CORR_FIXCONVERT_PQ_TO_SDR;
            else if (trc == Trc_HLG && m_bConvertToSdr) mode = CORR_FIXCONVERT_HLG_TO_SDR;
            else if (prim == Prim_BT2020) mode = CORR_FIX_BT2020;
        }
        else if (trc == Trc_2084 && !HdrOut() && m_bConvertToSdr) mode = m_upstreamPq ? CORR_CONVERT_PQ_TO_SDR : CORR_FIXCONVERT_PQ_TO_SDR;
        else if (trc == Trc_HLG) {
            if (HdrOut()) mode = CORR_CONVERT_HLG_TO_PQ;
            else if (m_bConvertToSdr) mode = CORR_FIXCONVERT_HLG_TO_SDR;
            else if (prim == Prim_BT2020) mode = CORR_FIX_BT2020;                 // HLG compatible with SDR
        }
        else if ((bTransFunc22 || !m_bHdrPassthroughSupport) && prim == Prim_BT2020) mode = CORR_FIX_BT2020;
    }
    if (mode != CORR_NONE && m_corr.gamut[0][0] == 0.0f) {
        // `static const` matrices of the shaders: fxc folds them at compile time in double precision (warning X4122 in
        // colorspace_gamut_conversion.hlsl); evaluated in double in the written order, rounded once to fp32
        static const double rgb_ycbcr709[3][3] = { { 0.2126, 0.7152, 0.0722 }, { -0.114572, -0.385428, 0.5 }, { 0.5, -0.454153, -0.045847 } };
        static const double ycbcr2020nc_rgb[3][3] = { { 1.0, 0.0, 1.4746 }, { 1.0, -0.164553, -0.571353 }, { 1.0, 1.8814, 0.0 } };
        static const double ycgco_rgb[3][3] = { { 1.0, -1.0, 1.0 }, { 1.0, 1.0, 0.0 }, { 1.0, -1.0, -1.0 } };
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double a = 0, b = 0;
                for (int k = 0; k < 3; k++) { a += ycbcr2020nc_rgb[i][k] * rgb_ycbcr709[k][j]; b += ycgco_rgb[i][k] * rgb_ycbcr709[k][j]; }
                m_corr.fix2020[i][j] = (float)a; m_corr.fixycgco[i][j] = (float)b;
            }
        // Colorspace_Gamut_Conversion_2020_to_709 (colorspace_gamut_conversion.hlsl:1-93, from zimg)
        auto det2 = [](double a00, double a01, double a10, double a11) { return a00 * a11 - a01 * a10; };
        auto inverse = [&](const double X[3][3], double I[3][3]) {
            double det = 0;
            det += X[0][0] * det2(X[1][1], X[1][2], X[2][1], X[2][2]);
            det -= X[0][1] * det2(X[1][0], X[1][2], X[2][0], X[2][2]);
            det += X[0][2] * det2(X[1][0], X[1][1], X[2][0], X[2][1]);
            I[0][0] = det2(X[1][1], X[1][2], X[2][1], X[2][2]) / det; I[0][1] = det2(X[0][2], X[0][1], X[2][2], X[2][1]) / det; I[0][2] = det2(X[0][1], X[0][2], X[1][1], X[1][2]) / det;
            I[1][0] = det2(X[1][2], X[1][0], X[2][2], X[2][0]) / det; I[1][1] = det2(X[0][0], X[0][2], X[2][0], X[2][2]) / det; I[1][2] = det2(X[0][2], X[0][0], X[1][2], X[1][0]) / det;
            I[2][0] = det2(X[1][0], X[1][1], X[2][0], X[2][1]) / det; I[2][1] = det2(X[0][1], X[0][0], X[2][1], X[2][0]) / det; I[2][2] = det2(X[0][0], X[0][1], X[1][0], X[1][1]) / det;
        };
        static const double P709[3][2] = { { 0.640, 0.330 }, { 0.300, 0.600 }, { 0.150, 0.060 } };
        static const double P2020[3][2] = { { 0.708, 0.292 }, { 0.170, 0.797 }, { 0.131, 0.046 } };
        const double wx = 0.3127, wy = 0.3290;
        double M[2][3][3];
        for (int s2 = 0; s2 < 2; s2++) {
            const double (*P)[2] = s2 ? P2020 : P709;
            double X[3][3], I[3][3], S[3];
            for (int c = 0; c < 3; c++) { X[0][c] = P[c][0] / P[c][1]; X[1][c] = 1.0; X[2][c] = (1.0 - P[c][0] - P[c][1]) / P[c][1]; }
            const double W[3] = { wx / wy, 1.0, (1.0 - wx - wy) / wy };
            inverse(X, I);
            for (int r = 0; r < 3; r++) S[r] = I[r][0] * W[0] + I[r][1] * W[1] + I[r][2] * W[2];
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) M[s2][r][c] = X[r][c] * S[c];
        }
        double I709[3][3];
        inverse(M[0], I709);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) { double a = 0; for (int k = 0; k < 3; k++) a += I709[r][k] * M[1][k][c]; m_corr.gamut[r][c] = (float)a; }
    }
    if (mode != m_corr.mode) { m_corr.mode = mode; m_planDirty = true; }
}

int32_t CB200VideoProcessor::CorrectionPass(const Tex2D& tex, const Tex2D& rt, const CRect& rect)
{
    m_lastLaunches++;
    return Cuda(LaunchCorrectionPass(tex.desc(), rt.desc(), m_color, m_corr, rect.left, rect.top, rect.right, rect.bottom, m_stream), "correction shader") ? MPCVR_S_OK : MPCVR_E_FAIL;
}

int32_t CB200VideoProcessor::SetDisplayHdr(bool supported)
{
    // m_bHdrPassthroughSupport: in the reference it is read from the DXGI output (HDR-capable display + OS HDR mode)
    if (supported == m_bHdrPassthroughSupport) return MPCVR_S_OK;
    m_bHdrPassthroughSupport = supported;
    m_hdrColorSpaceActive = false; m_bTonemap = false; m_lastHdr10 = Hdr10();
    if (m_bInit) { SetShaderConvertColorParams(); UpdateCorrection(); }
    m_planDirty = true;
    return MPCVR_S_OK;
}

void CB200VideoProcessor::SetHDR10ShaderParams(float mn, float mx, float cll, float fall, float disp, int type)
{
    if (mn <= 0.f) mn = 0.f;                                                  // :911-916
    if (mx <= 10.f) mx = 1000.f;
    if (cll <= 10.f) cll = mx;
    if (fall <= 1.f) fall = cll;
    if (disp < 100.f || disp > 10000.f) disp = 1000.f;
    if (type < 1 || type > 6) type = 1;
    m_tm.min_nits = mn; m_tm.max_nits = mx; m_tm.max_cll = cll; m_tm.max_fall = fall; m_tm.display = disp; m_tm.sel = type;
    m_bTonemap = true;                                                        // m_pPSHDR10ToneMapping created, :924-929
}

void CB200VideoProcessor::UpdateHdrState()
{
    // The part of Render() under "if (m_pDXGISwapChain4)" (:2626-2727); that swap chain exists when
    // bHdrOutput = m_bHdrPassthroughSupport && (m_bHdrPassthrough || m_bHdrLocalToneMapping) && SourceIsHDR() (:1476).
    if (!(HdrOut() && (SourceIsHDR10orHLG() || m_Dovi.bValid))) { m_hdrColorSpaceActive = false; m_bTonemap = false; return; }   // SourceIsHDR(), :1476
    if (m_Dovi.bValid) {                                                      // :2627-2648: the RPU's static metadata tops up the stream's
        if (!m_hdr10.bValid && m_lastHdr10.bValid) { m_hdr10 = m_lastHdr10; m_hdr10.bValid = true; }
        if (m_hdr10.bValid) {
            if (m_DoviMaxMasteringLuminance > m_hdr10.MaxMasteringLuminance) m_hdr10.MaxMasteringLuminance = m_DoviMaxMasteringLuminance;
            if (m_DoviMinMasteringLuminance && m_DoviMinMasteringLuminance != m_hdr10.MinMasteringLuminance) m_hdr10.MinMasteringLuminance = m_DoviMinMasteringLuminance;
            if (m_DoviMaxContentLightLevel && m_DoviMaxContentLightLevel != m_hdr10.MaxCLL) m_hdr10.MaxCLL = (uint16_t)m_DoviMaxContentLightLevel;
            if (m_DoviMaxFrameAverageLightLevel && m_DoviMaxFrameAverageLightLevel != m_hdr10.MaxFALL) m_hdr10.MaxFALL = (uint16_t)m_DoviMaxFrameAverageLightLevel;
        }
    }
    if (!m_hdrColorSpaceActive) {                                             // SetColorSpace1(G2084/P2020) succeeded, :2651-2690
        m_hdrColorSpaceActive = true;
        if (m_hdr10.bValid) m_lastHdr10 = m_hdr10;
        else if (!m_lastHdr10.bValid) {                                       // Display-P3 defaults, :2673-2692
            m_lastHdr10 = Hdr10(); m_lastHdr10.bValid = true;
            m_lastHdr10.MaxMasteringLuminance = m_DoviMaxMasteringLuminance ? m_DoviMaxMasteringLuminance : 1000;
            m_lastHdr10.MinMasteringLuminance = m_DoviMinMasteringLuminance ? m_DoviMinMasteringLuminance : 50;
            // the reference writes the RPU's MaxCLL / MaxFALL into m_hdr10 here, which is not valid at this point (:2687-2692): kept
            if (m_DoviMaxContentLightLevel) m_hdr10.MaxCLL = (uint16_t)m_DoviMaxContentLightLevel;
            if (m_DoviMaxFrameAverageLightLevel) m_hdr10.MaxFALL = (uint16_t)m_DoviMaxFrameAverageLightLevel;
        }
    } else if (m_hdr10.bValid) {                                              // :2703-2712
        if (m_hdr10.MaxMasteringLuminance != m_lastHdr10.MaxMasteringLuminance || m_hdr10.MinMasteringLuminance != m_lastHdr10.MinMasteringLuminance ||
            m_hdr10.MaxCLL != m_lastHdr10.MaxCLL || m_hdr10.MaxFALL != m_lastHdr10.MaxFALL || !m_lastHdr10.bValid) m_lastHdr10 = m_hdr10;
    }
    m_tm.l2_enabled = 0;
    if (m_bHdrLocalToneMapping && m_Dovi.bValid && m_doviL1Present) {         // :2716-2720: Dolby Vision L1 drives the curve, BT.2390 becomes ST 2094-10
        SetHDR10ShaderParams(m_doviL1Nits[0], m_doviL1Nits[1], m_doviL1Nits[1], m_doviL1Nits[2], (float)m_iHdrDisplayMaxNits,
                             m_iHdrLocalToneMappingType == 5 ? 6 : m_iHdrLocalToneMappingType);
    } else if (m_bHdrLocalToneMapping && m_lastHdr10.bValid) {                // :2721-2725
        SetHDR10ShaderParams((float)m_lastHdr10.MinMasteringLuminance, (float)m_lastHdr10.MaxMasteringLuminance, (float)m_lastHdr10.MaxCLL,
                             (float)m_lastHdr10.MaxFALL, (float)m_iHdrDisplayMaxNits, m_iHdrLocalToneMappingType);
    } else m_bTonemap = false;
    if (m_bTonemap && m_Dovi.bValid && m_doviHost.l2_enabled) {               // cbuffer b1 = m_pDoViDynamicConstants, :3362-3364
        m_tm.l2_enabled = 1; m_tm.l2_chroma_weight = m_doviHost.l2_chroma_weight; m_tm.l2_saturation_gain = m_doviHost.l2_saturation_gain;
        m_tm.l2_slope = m_doviHost.l2_slope; m_tm.l2_offset = m_doviHost.l2_offset; m_tm.l2_power = m_doviHost.l2_power;
    }
}

void CB200VideoProcessor::GetVideoSize(int32_t& w, int32_t& h) const
{
    const bool q = (m_iRotation == 90 || m_iRotation == 270);               // CVideoProcessor::GetVideoSize, Source/VideoProcessor.cpp:30-39
    w = q ? m_srcRectHeight : m_srcRectWidth; h = q ? m_srcRectWidth : m_srcRectHeight;
}
void CB200VideoProcessor::GetAspectRatio(int32_t& ax, int32_t& ay) const
{
    const bool q = (m_iRotation == 90 || m_iRotation == 270);               // Source/VideoProcessor.cpp:41-50
    ax = q ? m_srcAspectRatioY : m_srcAspectRatioX; ay = q ? m_srcAspectRatioX : m_srcAspectRatioY;
}

void CB200VideoProcessor::Configure(const mpcvr_settings& s)
{
    // diff and rebuild the minimum, like Source/DX11VideoProcessor.cpp:3800-4050
    const int oldTex = m_iTexFormat; const bool oldSdr = m_bConvertToSdr; const int oldNits = m_iSDRDisplayNits;
    const bool oldHdrOut = HdrOut(); const int oldHdrNits = m_iHdrDisplayMaxNits; const bool oldPreferDoVi = m_bHdrPreferDoVi;
    ApplySettings(s);
    if (!m_bInit) return;
    if (oldPreferDoVi != m_bHdrPreferDoVi && m_Dovi.bValid && !m_bHdrPreferDoVi && SourceIsHDR10orHLG()) {   // :3891-3898
        m_Dovi.bValid = false; m_Dovi.bHasMMR = false; m_Dovi.msd = mpcvr_dovi_metadata();
        m_doviL1Present = false; m_doviL2Present = false;
        SetShaderConvertColorParams();
    }
    // the convert shader is rebuilt on a chroma-scaling / blend change (:3985-3999): the current frame re-renders with the new fetch
    m_src.chroma_mode = m_iChromaScaling;
    m_src.deint_blend = (m_bDeintBlend && m_bInterlaced && m_srcParams.Subsampling == 420) ? 1 : 0;
    // the Dolby Vision Level-2 trims are picked for iHdrDisplayMaxNits (:2383-2466): recompute them when the display changes
    if (m_Dovi.bValid && oldHdrNits != m_iHdrDisplayMaxNits) SetShaderDoviCurves();
    if (oldTex != m_iTexFormat) { UpdateTexParams(m_srcParams.CDepth); UpdateTexures(); }
    if (oldSdr != m_bConvertToSdr || oldHdrOut != HdrOut()) SetShaderConvertColorParams();
    if (!HdrOut()) { m_hdrColorSpaceActive = false; m_bTonemap = false; }
    UpdateCorrection();
    if (oldNits != m_iSDRDisplayNits) SetShaderLuminanceParams();
    UpdatePostScaleTexures();
    m_planDirty = true;
}

int32_t CB200VideoProcessor::SetProcAmpValues(uint32_t flags, const mpcvr_procamp* v)
{
    if (!v) return MPCVR_E_POINTER;
    // ranges of SetDefaultDXVA2ProcAmpRanges: brightness -100..100, contrast 0..2, hue -180..180, saturation 0..2
    auto clampf = [](float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); };
    if (flags & MPCVR_PROCAMP_BRIGHTNESS) m_ProcAmp.brightness = clampf(v->brightness, -100.f, 100.f);
    if (flags & MPCVR_PROCAMP_CONTRAST)   m_ProcAmp.contrast   = clampf(v->contrast, 0.f, 2.f);
    if (flags & MPCVR_PROCAMP_HUE)        m_ProcAmp.hue        = clampf(v->hue, -180.f, 180.f);
    if (flags & MPCVR_PROCAMP_SATURATION) m_ProcAmp.saturation = clampf(v->saturation, 0.f, 2.f);
    if (m_bInit) SetShaderConvertColorParams();
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::GetProcAmpValues(uint32_t flags, mpcvr_procamp* v) const
{
    if (!v) return MPCVR_E_POINTER;
    (void)flags;
    v->brightness = m_ProcAmp.brightness; v->contrast = m_ProcAmp.contrast; v->hue = m_ProcAmp.hue; v->saturation = m_ProcAmp.saturation;
    return MPCVR_S_OK;
}

// ------------------------------------------------------------------------------------------------
// output surfaces
// ------------------------------------------------------------------------------------------------
int32_t CB200VideoProcessor::SetBackbufferFormat(int fmt)
{
    if (fmt != MPCVR_SURF_BGRA8 && fmt != MPCVR_SURF_RGB10A2 && fmt != MPCVR_SURF_RGBA16F && fmt != MPCVR_SURF_RGB8) return MPCVR_E_INVALIDARG;
    m_SwapChainFmt = fmt;
    m_UserBackBuffer = Tex2D();
    return UpdatePostScaleTexures();
}

int32_t CB200VideoProcessor::BindBackbuffer(void* dev, size_t pitch)
{
    if (!dev) { m_UserBackBuffer = Tex2D(); return MPCVR_S_OK; }
    if (m_windowRect.IsRectEmpty()) return Fail(MPCVR_E_UNEXPECTED, "bind_backbuffer: window rect not set");
    if (pitch < (size_t)m_windowRect.Width() * SurfBpp(m_SwapChainFmt) || (pitch & 15)) return Fail(MPCVR_E_INVALIDARG, "bind_backbuffer: pitch too small or not 16-byte aligned");
    m_UserBackBuffer.ptr = dev; m_UserBackBuffer.pitch = pitch; m_UserBackBuffer.w = m_windowRect.Width(); m_UserBackBuffer.h = m_windowRect.Height();
    m_UserBackBuffer.fmt = m_SwapChainFmt; m_UserBackBuffer.owned = false;
    m_planDirty = true;
    return MPCVR_S_OK;
}

Tex2D& CB200VideoProcessor::CurrentTarget() { return m_UserBackBuffer.ptr ? m_UserBackBuffer : m_BackBuffer; }

int32_t CB200VideoProcessor::GetBackbuffer(void** dev, size_t* pitch, int32_t* w, int32_t* h, int32_t* fmt)
{
    Tex2D& t = CurrentTarget();
    if (!t.ptr) return Fail(MPCVR_E_UNEXPECTED, "no back buffer yet");
    if (dev) *dev = t.ptr; if (pitch) *pitch = t.pitch; if (w) *w = t.w; if (h) *h = t.h; if (fmt) *fmt = t.fmt;
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::ReadBackbuffer(void* host, size_t pitch, bool sync)
{
    if (m_device < 0) return Fail(MPCVR_E_UNEXPECTED, "no CUDA device (planning-only instance)");
    Tex2D& t = CurrentTarget();
    if (!t.ptr || !host) return Fail(MPCVR_E_POINTER, "read_backbuffer: nothing to read");
    cudaSetDevice(m_device);
    if (!Cuda(cudaMemcpy2DAsync(host, pitch, t.ptr, t.pitch, (size_t)t.w * SurfBpp(t.fmt), t.h, cudaMemcpyDeviceToHost, m_stream), "D2H back buffer")) return MPCVR_E_FAIL;
    if (sync && !Cuda(cudaStreamSynchronize(m_stream), "sync")) return MPCVR_E_FAIL;
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::WaitIdle()
{
    if (m_device < 0) return MPCVR_S_OK;
    cudaSetDevice(m_device);
    return Cuda(cudaStreamSynchronize(m_stream), "cudaStreamSynchronize") ? MPCVR_S_OK : MPCVR_E_FAIL;
}

// ------------------------------------------------------------------------------------------------
// frames
// ------------------------------------------------------------------------------------------------
int32_t CB200VideoProcessor::MemCopyToTexSrcVideo(const uint8_t* data, int pitch)
{
    // plane walk of MemCopyToTexSrcVideo, :1213-1252
    const FmtConvParams& f = m_srcParams;
    const int H = (int)m_srcHeight;
    auto copy = [&](Tex2D& t, const uint8_t* src, int spitch, int lines) {
        return cudaMemcpy2DAsync(t.ptr, t.pitch, src, spitch, (size_t)std::min(t.w, spitch), lines, cudaMemcpyHostToDevice, m_stream);
    };
    if (f.layout == LAYOUT_BIPLANAR || f.layout == LAYOUT_PLANAR3) {
        if (!Cuda(copy(m_TexSrcVideo[0], data, pitch, H), "H2D luma")) return MPCVR_E_FAIL;
        const int chromaH = H / f.div_chroma_h;
        const int chromaPitch = (f.layout == LAYOUT_PLANAR3) ? pitch / f.div_chroma_w : pitch;
        const uint8_t* p = data + (size_t)pitch * H;
        if (!Cuda(copy(m_TexSrcVideo[1], p, chromaPitch, chromaH), "H2D chroma")) return MPCVR_E_FAIL;
        if (f.layout == LAYOUT_PLANAR3) {
            p += (size_t)chromaPitch * chromaH;
            if (!Cuda(copy(m_TexSrcVideo[2], p, chromaPitch, chromaH), "H2D chroma 2")) return MPCVR_E_FAIL;
        }
    } else {
        if (!Cuda(copy(m_TexSrcVideo[0], data, pitch, H), "H2D packed")) return MPCVR_E_FAIL;
    }
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::CopySample(const mpcvr_sample* s)
{
    const FmtConvParams& f = m_srcParams;
    SrcDesc& S = m_src;
    memset(&S, 0, sizeof(S));
    S.W = m_srcWidth; S.H = m_srcHeight;
    S.cw = m_srcWidth / f.div_chroma_w; S.ch = m_srcHeight / f.div_chroma_h;
    S.layout = f.layout; S.bytes = f.bytes; S.lsb_shift = f.lsb_shift; S.packed_kind = f.packed_kind; S.subsampling = f.Subsampling;
    S.chroma_mode = m_iChromaScaling; S.siting = m_srcExFmt.ChromaSubsampling();
    S.deint_blend = (m_bDeintBlend && m_bInterlaced && f.Subsampling == 420) ? 1 : 0;   // :2965, :3075
    S.rect_l = m_srcRect.left; S.rect_t = m_srcRect.top; S.rect_w = m_srcRectWidth; S.rect_h = m_srcRectHeight;

    m_hdr10 = Hdr10();                                                        // :2231
    if (SourceIsHDR10orHLG() && (m_bHdrPassthrough || m_bHdrLocalToneMapping)) {   // IMediaSideData HDR blobs, :2233-2268
        if (const mpcvr_side_data_hdr* hdr = s->hdr) {
            const double* px = hdr->display_primaries_x; const double* py = hdr->display_primaries_y;
            if (px[0] > 0. && px[1] > 0. && px[2] > 0. && py[0] > 0. && py[1] > 0. && py[2] > 0. && hdr->white_point_x > 0. && hdr->white_point_y > 0.
                && hdr->max_display_mastering_luminance > 0. && hdr->min_display_mastering_luminance > 0.) {
                m_hdr10.bValid = true;
                m_hdr10.MaxMasteringLuminance = (uint32_t)std::lround(hdr->max_display_mastering_luminance);
                m_hdr10.MinMasteringLuminance = (uint32_t)std::lround(hdr->min_display_mastering_luminance * 10000.0);
            }
        }
        if (const mpcvr_side_data_hdr_cll* cll = s->hdr_cll) { m_hdr10.MaxCLL = (uint16_t)cll->MaxCLL; m_hdr10.MaxFALL = (uint16_t)cll->MaxFALL; }
    }

    // Dolby Vision side data (:2271-2330): taken for YUV types when preferred or when the stream is not tagged PQ/HLG
    {
        const bool take = s->dovi && f.CSType == CS_YUV && (m_bHdrPreferDoVi || !SourceIsHDR10orHLG()) && CheckDoviMetadata(s->dovi);
        if (take) {
            const bool changed = !m_Dovi.bValid || memcmp(&m_Dovi.msd, s->dovi, sizeof(m_Dovi.msd)) != 0;
            // bMasteringLuminanceChanged compares with whatever m_Dovi.msd held before (zeros at first), :2300-2301
            const bool masteringChanged = m_Dovi.msd.source_max_pq != s->dovi->source_max_pq || m_Dovi.msd.source_min_pq != s->dovi->source_min_pq;
            if (masteringChanged) {                                       // :2472-2490
                auto PqToLinearNits = [](float x) {
                    const float PQ_M1 = 2610.f / (4096.f * 4.f), PQ_M2 = 2523.f / 4096.f * 128.f, PQ_C1 = 3424.f / 4096.f,
                                PQ_C2 = 2413.f / 4096.f * 32.f, PQ_C3 = 2392.f / 4096.f * 32.f;
                    x = powf(x, 1.0f / PQ_M2);
                    x = fmaxf(x - PQ_C1, 0.0f) / (PQ_C2 - PQ_C3 * x);
                    x = powf(x, 1.0f / PQ_M1);
                    return x * 10000.0f;
                };
                m_DoviMaxMasteringLuminance = (uint32_t)PqToLinearNits(s->dovi->source_max_pq / 4095.f);
                m_DoviMinMasteringLuminance = (uint32_t)(PqToLinearNits(s->dovi->source_min_pq / 4095.f) * 10000.0f);
                for (int i = 0; i < 32; i++) {
                    if (s->dovi->ext[i].level != 6) continue;
                    m_DoviMaxMasteringLuminance = s->dovi->ext[i].v[0]; m_DoviMinMasteringLuminance = s->dovi->ext[i].v[1];
                    m_DoviMaxContentLightLevel = s->dovi->ext[i].v[2]; m_DoviMaxFrameAverageLightLevel = s->dovi->ext[i].v[3];
                    break;
                }
            }
            if (changed) {
                m_Dovi.msd = *s->dovi; m_Dovi.bValid = true;
                const int32_t hr = SetShaderDoviCurves();
                if (hr != MPCVR_S_OK) return hr;
                SetShaderConvertColorParams();
            }
        } else if (m_Dovi.bValid && !s->dovi) {
            // the reference keeps the last metadata until the media type changes; a sample without it keeps using it (:2271-2330)
        }
    }

    if (s->flags & MPCVR_SAMPLE_DEVICE_MEMORY) {
        // IMediaSampleD3D11 path (:2528-2570): the decoder's planes are read in place
        const int np = f.layout == LAYOUT_BIPLANAR ? 2 : (f.layout == LAYOUT_PLANAR3 ? 3 : 1);
        for (int i = 0; i < np; i++) {
            if (!s->planes[i] || s->pitches[i] == 0 || (s->pitches[i] < 0 && np > 1)) return Fail(MPCVR_E_POINTER, "device sample: missing plane");
            // the kernels read a plane with typed loads of its texel (a D3D11 texture is always that aligned; a raw pointer need not be)
            const int a = PlaneLoadBytes(f, i);
            if (((uintptr_t)s->planes[i] % a) || (s->pitches[i] % a)) return Fail(MPCVR_E_INVALIDARG, "device sample: plane address / pitch not aligned to the texel size");
            S.pl[i] = (const uint8_t*)s->planes[i]; S.pitch[i] = s->pitches[i];
        }
        if (s->ready_event && !Cuda(cudaStreamWaitEvent(m_stream, (cudaEvent_t)s->ready_event, 0), "cudaStreamWaitEvent")) return MPCVR_E_FAIL;
    } else {
        if (!s->data) return Fail(MPCVR_E_POINTER, "host sample: null data");
        const int pitch = s->pitch ? s->pitch : m_srcPitch;
        if (pitch == 0) return Fail(MPCVR_E_INVALIDARG, "host sample: zero pitch");
        if (pitch < 0 && (f.layout == LAYOUT_BIPLANAR || f.layout == LAYOUT_PLANAR3)) return Fail(MPCVR_E_NOTIMPL, "bottom-up (negative pitch) samples are single-plane only (:1243-1249)");
        if (s->size > 0 && s->size < (int64_t)std::abs(pitch) * (int64_t)m_srcLines) return Fail(MPCVR_E_FAIL, "host sample smaller than pitch*lines (:2579)");
        // a bottom-up DIB is uploaded as it lies in memory; the kernels walk it from its last row with a negative pitch,
        // which is what "src = srcData + srcPitch * (1 - lines)" (:1246) does on the CPU
        // bottom-up DIB inside a taller buffer (VR_Extradata): the frame's rows are the LAST ones in memory, "src = srcData + srcPitch * (1 - lines)"
        // with lines = m_srcLines (:1246)
        const uint8_t* first = (const uint8_t*)s->data + (pitch < 0 ? (size_t)std::abs(pitch) * (m_srcLines > m_srcHeight ? m_srcLines - m_srcHeight : 0) : 0);
        const int32_t hr = MemCopyToTexSrcVideo(first, std::abs(pitch));
        if (hr != MPCVR_S_OK) return hr;
        for (int i = 0; i < 3; i++) { S.pl[i] = (const uint8_t*)m_TexSrcVideo[i].ptr; S.pitch[i] = (int32_t)m_TexSrcVideo[i].pitch; }
        if (pitch < 0) { S.pl[0] += (size_t)S.pitch[0] * (m_srcHeight - 1); S.pitch[0] = -S.pitch[0]; }
    }
    if (f.layout == LAYOUT_PLANAR3 && f.v_before_u) { std::swap(S.pl[1], S.pl[2]); std::swap(S.pitch[1], S.pitch[2]); }   // texV:t1, texU:t2
    m_rtStart = s->rt_start; m_rtEnd = s->rt_end;
    m_haveFrame = true;
    return MPCVR_S_OK;
}

// width in bytes of the loads the kernels issue on plane `plane` of a sample (pixel_ops.cuh fetch_*)
int PlaneLoadBytes(const FmtConvParams& f, int plane)
{
    switch (f.layout) {
    case LAYOUT_BIPLANAR:  return plane == 0 ? f.bytes : 2 * f.bytes;                         // R8|R16 ; R8G8|R16G16 texels
    case LAYOUT_PLANAR3:   return f.bytes;
    case LAYOUT_PACKED422: return f.packed_kind == 2 ? 4 : f.bytes;                           // v210 words ; byte / 16-bit components
    case LAYOUT_PACKED444: return f.packed_kind == 2 ? 8 : 4;                                 // Y416 ; AYUV, Y410
    case LAYOUT_RGB:
        switch (f.packed_kind) { case RGBK_BGR24: return 1; case RGBK_RGB48: case RGBK_BGR48: return 2;
                                 case RGBK_BGRA64: case RGBK_B64A: return 8; default: return 4; }
    default:               return f.bytes;
    }
}

int32_t CB200VideoProcessor::ProcessSample(const mpcvr_sample* s)
{
    if (m_device < 0) return Fail(MPCVR_E_UNEXPECTED, "no CUDA device: this library has no CPU path (planning-only instance)");
    if (!m_bInit) return Fail(MPCVR_E_UNEXPECTED, "ProcessSample before InitMediaType");
    if (!s) return MPCVR_E_POINTER;
    cudaSetDevice(m_device);
    int32_t hr = CopySample(s);
    if (hr != MPCVR_S_OK) return hr;
    return Render(1, s->rt_start);
}

int32_t CB200VideoProcessor::Render(int field, int64_t frameStartTime)
{
    (void)field; (void)frameStartTime;
    if (m_device < 0) return Fail(MPCVR_E_UNEXPECTED, "no CUDA device: this library has no CPU path (planning-only instance)");
    if (!m_bInit || !m_haveFrame) return Fail(MPCVR_E_UNEXPECTED, "Render without a frame");
    if (m_windowRect.IsRectEmpty() || m_videoRect.IsRectEmpty()) return MPCVR_S_FALSE;     // nothing visible: frame skipped
    cudaSetDevice(m_device);
    m_lastLaunches = 0;
    Tex2D& target = CurrentTarget();
    if (!target.ptr) { int32_t hr = UpdatePostScaleTexures(); if (hr != MPCVR_S_OK) return hr; }
    UpdateHdrState();
    int32_t hr = Process(CurrentTarget(), m_srcRect, m_videoRect, false);
    if (hr != MPCVR_S_OK) return hr;
    return DrawSubtitles(CurrentTarget());
}

int32_t CB200VideoProcessor::DrawSubtitles(const Tex2D& rt)
{
    if (!m_subRender) return MPCVR_S_OK;
    // arguments as at Source/DX11VideoProcessor.cpp:3262-3282
    mpcvr_rect cropped = {m_videoRect.left, m_videoRect.top, m_videoRect.right, m_videoRect.bottom};
    mpcvr_rect original = cropped;
    mpcvr_rect viewport = {0, 0, m_windowRect.Width(), m_windowRect.Height()};
    m_subRender(m_subUser, m_rtStart, m_rtEnd, m_rtAvgTimePerFrame, cropped, original, viewport, 1.0, 0, 0,
                (void*)m_stream, rt.ptr, rt.pitch, rt.fmt);
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::FillBlack()
{
    if (m_device < 0) return Fail(MPCVR_E_UNEXPECTED, "no CUDA device");
    Tex2D& t = CurrentTarget();
    if (!t.ptr) return MPCVR_S_OK;
    cudaSetDevice(m_device);
    return Cuda(LaunchClear(t.desc(), m_stream), "clear") ? MPCVR_S_OK : MPCVR_E_FAIL;
}

void CB200VideoProcessor::Flush()
{
    m_haveFrame = false; m_rtStart = 0;
    m_doviL1Present = false; m_doviL2Present = false;                       // m_DoviExtensionMetadata = {}, Source/DX11VideoProcessor.cpp:4082
}

int32_t CB200VideoProcessor::Reset(bool) { m_planDirty = true; return MPCVR_S_OK; }

// ------------------------------------------------------------------------------------------------
// plan: everything Process() needs that depends only on media type / rects / settings
// ------------------------------------------------------------------------------------------------
int32_t CB200VideoProcessor::UpdatePlan()
{
    if (!m_planDirty) return MPCVR_S_OK;
    // after ConvertColorPass rSrc is the whole convert output (:3319); without one the chain samples the source
    // texture itself and rSrc stays m_srcRect (:3320-3322)
    const CRect rSrc = m_PSConvColorData.enable ? CRect(0, 0, (int)m_srcRectWidth, (int)m_srcRectHeight) : m_srcRect;
    const CRect& dst = m_videoRect;
    const int rot = m_iRotation; const bool rotated = (rot == 90 || rot == 270);
    m_oriented = (rot != 0 || m_bFlip);
    // ResizeShaderPass, :3106-3127: a quarter turn swaps the source dimensions the scaler choice looks at
    const int w1 = rotated ? rSrc.Height() : rSrc.Width(), h1 = rotated ? rSrc.Width() : rSrc.Height(), w2 = dst.Width(), h2 = dst.Height();
    int kx = ChooseScaler(w1, w2, m_bInterpolateAt50pct), ky = ChooseScaler(h1, h2, m_bInterpolateAt50pct);
    // UPSCALE_Nearest has no shader (m_pShaderUpscaleX == nullptr, :94-95, :2919): that axis is point sampled
    if (kx == SCALER_INTERP && m_iUpscaling == MPCVR_UPSCALE_Nearest) kx = SCALER_NEAREST;
    if (ky == SCALER_INTERP && m_iUpscaling == MPCVR_UPSCALE_Nearest) ky = SCALER_NEAREST;
    const int fx = kx == SCALER_CONV ? m_iDownscaling : m_iUpscaling, fy = ky == SCALER_CONV ? m_iDownscaling : m_iUpscaling;
    const bool d3d9 = !m_bUseD3D11;
    // Jinc2 has no per-axis table: it is one 2-D draw (LaunchJinc2Pass); the slot keeps a point map so the bookkeeping stays uniform
    m_kx = kx; m_ky = ky;
    m_jincX = (kx == SCALER_INTERP && m_iUpscaling == MPCVR_UPSCALE_Jinc2);
    m_jincY = (ky == SCALER_INTERP && m_iUpscaling == MPCVR_UPSCALE_Jinc2);
    if (!BuildAxisTable(m_jincX ? SCALER_NEAREST : kx, fx, d3d9, rSrc.left, w1, dst.left, w2, m_axisX.host)) return Fail(MPCVR_E_NOTIMPL, "X scaler not available");
    // two shader passes: the Y pass reads m_TexResize, whose rows start at 0 (resize rect (dst.left, 0, dst.right, h1), :3150-3160)
    const bool twoPass = (kx == SCALER_INTERP || kx == SCALER_CONV) && (ky == SCALER_INTERP || ky == SCALER_CONV);
    if (!BuildAxisTable(m_jincY ? SCALER_NEAREST : ky, fy, d3d9, twoPass ? 0 : rSrc.top, h1, dst.top, h2, m_axisY.host)) return Fail(MPCVR_E_NOTIMPL, "Y scaler not available");
    // point-sampled maps: other axis of a single-axis resize, and plain copies
    BuildAxisTable(SCALER_NEAREST, 0, d3d9, rSrc.left, w1, dst.left, w2, m_pointX.host);
    BuildAxisTable(SCALER_NEAREST, 0, d3d9, rSrc.top, h1, dst.top, h2, m_pointY.host);
    // two-pass: X pass keeps rows (h1 -> h1), Y pass keeps columns (w2 -> w2)
    BuildAxisTable(SCALER_NEAREST, 0, d3d9, rSrc.top, h1, 0, h1, m_pointRY.host);
    BuildAxisTable(SCALER_NEAREST, 0, d3d9, dst.left, w2, dst.left, w2, m_pointRX.host);
    m_oInAxis = -1; m_oOutAxis = 0; m_oTwoPass = false;
    if (m_oriented) {
        // The first draw of the source carries the orientation (:3137-3181).  Output pixel (x, y) samples the texture at
        //   rot 0: u<-x, v<-y   rot 90: u<-y, v<-x reversed   rot 180: both reversed   rot 270: u<-y reversed, v<-x;  flip inverts u.
        // For a quarter turn the X-direction resizer is the Y shader (:3111-3119); when both directions pick the same shader
        // object (both up or both down) the reference takes its "one pass resize" branch (:3135-3141): one direction is
        // filtered, the other merely point sampled.  Kept as is.
        const bool shaderX = (kx == SCALER_INTERP || kx == SCALER_CONV), shaderY = (ky == SCALER_INTERP || ky == SCALER_CONV);
        const bool sameShader = rotated && shaderX && shaderY && kx == ky;
        m_oTwoPass = shaderX && shaderY && !sameShader;
        const int ow = w2, oh = m_oTwoPass ? h1 : h2;                      // first-draw target: m_TexResize rows when two passes follow
        const int oy0 = m_oTwoPass ? 0 : dst.top;
        const bool uFromX = !rotated;
        const int uLen = uFromX ? ow : oh, vLen = uFromX ? oh : ow;
        const bool uRev = ((rot == 180 || rot == 270) != m_bFlip), vRev = (rot == 90 || rot == 180);
        BuildAxisTable(SCALER_NEAREST, 0, d3d9, rSrc.left, rSrc.Width(), uFromX ? dst.left : oy0, uLen, m_oU.host, uRev);
        BuildAxisTable(SCALER_NEAREST, 0, d3d9, rSrc.top, rSrc.Height(), uFromX ? oy0 : dst.left, vLen, m_oV.host, vRev);
        int kind = SCALER_NONE, filt = 0;
        // a Jinc2 direction needs no table (the 2-D shader takes the turned quad as it is, LaunchJinc2Pass' orient); the only table an
        // oriented Jinc2 plan may need is the convolution of a shrinking X direction, whose draw then carries the orientation
        if (shaderX && !m_jincX)                   { m_oInAxis = rotated ? 1 : 0; m_oOutAxis = 0; kind = kx; filt = fx; }
        else if (!shaderX && shaderY && !m_jincY)  { m_oInAxis = rotated ? 0 : 1; m_oOutAxis = 1; kind = ky; filt = fy; }
        if (m_oInAxis >= 0) {
            // shader constant: scale.x = srcRect.W / dstRect.W, scale.y = srcRect.H / dstRect.H (:351-354), turned or not
            const float qscale = m_oInAxis == 0 ? (float)rSrc.Width() / (float)ow : (float)rSrc.Height() / (float)oh;
            const bool ok = m_oInAxis == 0
                ? BuildAxisTable(kind, filt, d3d9, rSrc.left, rSrc.Width(), m_oOutAxis == 0 ? dst.left : oy0, uLen, m_oFilt.host, uRev, qscale)
                : BuildAxisTable(kind, filt, d3d9, rSrc.top, rSrc.Height(), m_oOutAxis == 0 ? dst.left : oy0, vLen, m_oFilt.host, vRev, qscale);
            if (!ok) return Fail(MPCVR_E_NOTIMPL, "scaler not available for the rotated draw (Jinc2 is a 'next' row)");
        }
        if (m_device >= 0) {
            cudaSetDevice(m_device);
            if (!Cuda(m_oU.Upload(m_stream), "upload") || !Cuda(m_oV.Upload(m_stream), "upload") ||
                (m_oInAxis >= 0 && !Cuda(m_oFilt.Upload(m_stream), "upload"))) return MPCVR_E_FAIL;
        }
    }
    if (m_device >= 0) {
        cudaSetDevice(m_device);
        if (!Cuda(m_axisX.Upload(m_stream), "upload X table") || !Cuda(m_axisY.Upload(m_stream), "upload Y table") ||
            !Cuda(m_pointX.Upload(m_stream), "upload") || !Cuda(m_pointY.Upload(m_stream), "upload") ||
            !Cuda(m_pointRX.Upload(m_stream), "upload") || !Cuda(m_pointRY.Upload(m_stream), "upload")) return MPCVR_E_FAIL;
    }
    m_planDirty = false;
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::ConvertColorPass(const Tex2D& rt)
{
    float4* dbg = nullptr;
    if (m_captureF32) {
        const size_t n = (size_t)rt.w * rt.h;
        if (m_dbgConvertF32.n < n) { m_dbgConvertF32.Release(); if (!Cuda(cudaMalloc(&m_dbgConvertF32.ptr, n * sizeof(float4)), "dbg alloc")) return MPCVR_E_FAIL; m_dbgConvertF32.n = n; }
        dbg = m_dbgConvertF32.ptr;
    }
    m_lastLaunches++;
    return Cuda(LaunchConvertPass(m_src, m_color, rt.desc(), dbg, m_stream, m_Dovi.bValid ? m_doviDev.ptr : nullptr), "ConvertColorPass") ? MPCVR_S_OK : MPCVR_E_FAIL;
}

int32_t CB200VideoProcessor::ResizeShaderPass(const Tex2D& tex, const Tex2D& rt, const CRect& srcRect, const CRect& dstRect)
{
    const int kx = m_kx, ky = m_ky;
    const bool shaderX = (kx == SCALER_INTERP || kx == SCALER_CONV), shaderY = (ky == SCALER_INTERP || ky == SCALER_CONV);
    const int h1 = m_axisY.host.src_len;
    if (m_jincX || m_jincY) {
        // resizerX == resizerY (or only one of them): one 2-D draw (:3135-3141, :3169-3177); if the other direction shrinks, the
        // convolution shader takes that pass and m_TexResize sits in between (:3143-3167).  The FIRST draw carries rotation / flip.
        const int rot = m_iRotation;
        const int orient = m_oriented ? (((rot == 90 || rot == 270) ? 1 : 0) | ((((rot == 180 || rot == 270) != m_bFlip)) ? 2 : 0) | ((rot == 90 || rot == 180) ? 4 : 0)) : 0;
        if (kx == SCALER_CONV || ky == SCALER_CONV) {
            if (!Cuda(m_TexResize.CheckCreate(MPCVR_SURF_RGBA16F, rt.w, h1), "m_TexResize")) return MPCVR_E_FAIL;
            m_lastLaunches += 2;
            cudaError_t e1;
            if (m_jincX) e1 = LaunchJinc2Pass(tex.desc(), m_TexResize.desc(), srcRect.left, srcRect.top, srcRect.Width(), srcRect.Height(), dstRect.left, 0, dstRect.Width(), h1, m_stream, orient);
            else if (m_oriented) {                 // turned convolution of the X direction (see the oriented branch below)
                const int32_t* other = (m_oInAxis == 0) ? m_oV.base.ptr : m_oU.base.ptr;
                e1 = LaunchResizePass(tex.desc(), m_TexResize.desc(), m_oFilt.table(), 0, other, 0, h1, m_stream, m_oInAxis);
            } else e1 = LaunchResizePass(tex.desc(), m_TexResize.desc(), m_axisX.table(), 0, m_pointRY.base.ptr, 0, h1, m_stream);
            if (!Cuda(e1, "resize X")) return MPCVR_E_FAIL;
            const cudaError_t e2 = m_jincY
                ? LaunchJinc2Pass(m_TexResize.desc(), rt.desc(), dstRect.left, 0, dstRect.Width(), h1, dstRect.left, dstRect.top, dstRect.Width(), dstRect.Height(), m_stream)
                : LaunchResizePass(m_TexResize.desc(), rt.desc(), m_axisY.table(), 1, m_pointRX.base.ptr, dstRect.left, dstRect.Width(), m_stream);
            if (!Cuda(e2, "resize Y")) return MPCVR_E_FAIL;
        } else {
            m_lastLaunches++;
            if (!Cuda(LaunchJinc2Pass(tex.desc(), rt.desc(), srcRect.left, srcRect.top, srcRect.Width(), srcRect.Height(),
                                      dstRect.left, dstRect.top, dstRect.Width(), dstRect.Height(), m_stream, orient), "Jinc2")) return MPCVR_E_FAIL;
        }
        return MPCVR_S_OK;
    }
    if (m_oriented) {
        const bool rotated = (m_iRotation == 90 || m_iRotation == 270);
        const Tex2D* first = &rt;
        int oy0 = dstRect.top, oh = dstRect.Height();
        if (m_oTwoPass) {
            if (!Cuda(m_TexResize.CheckCreate(MPCVR_SURF_RGBA16F, rt.w, h1), "m_TexResize")) return MPCVR_E_FAIL;
            first = &m_TexResize; oy0 = 0; oh = h1;
        }
        m_lastLaunches++;
        if (m_oInAxis < 0) {
            if (!Cuda(LaunchCopyPass(tex.desc(), first->desc(), m_oU.base.ptr, m_oV.base.ptr, dstRect.left, oy0, dstRect.Width(), oh, m_stream, rotated ? 1 : 0), "oriented copy")) return MPCVR_E_FAIL;
        } else {
            // the other output axis is point sampled along the other texture axis
            const int32_t* other = (m_oInAxis == 0) ? m_oV.base.ptr : m_oU.base.ptr;
            const int other0 = m_oOutAxis == 0 ? oy0 : dstRect.left, otherLen = m_oOutAxis == 0 ? oh : dstRect.Width();
            if (!Cuda(LaunchResizePass(tex.desc(), first->desc(), m_oFilt.table(), m_oOutAxis, other, other0, otherLen, m_stream, m_oInAxis), "oriented resize")) return MPCVR_E_FAIL;
        }
        if (m_oTwoPass) {
            m_lastLaunches++;
            if (!Cuda(LaunchResizePass(m_TexResize.desc(), rt.desc(), m_axisY.table(), 1, m_pointRX.base.ptr, dstRect.left, dstRect.Width(), m_stream), "resize Y")) return MPCVR_E_FAIL;
        }
        return MPCVR_S_OK;
    }
    if (shaderX && shaderY) {
        // two passes through m_TexResize: RGBA16F, width = render-target width, height = h1 (:3143-3167)
        if (!Cuda(m_TexResize.CheckCreate(MPCVR_SURF_RGBA16F, rt.w, h1), "m_TexResize")) return MPCVR_E_FAIL;
        m_lastLaunches += 2;
        if (!Cuda(LaunchResizePass(tex.desc(), m_TexResize.desc(), m_axisX.table(), 0, m_pointRY.base.ptr, 0, h1, m_stream), "resize X")) return MPCVR_E_FAIL;
        if (!Cuda(LaunchResizePass(m_TexResize.desc(), rt.desc(), m_axisY.table(), 1, m_pointRX.base.ptr, dstRect.left, dstRect.Width(), m_stream), "resize Y")) return MPCVR_E_FAIL;
    } else if (shaderX) {
        m_lastLaunches++;
        if (!Cuda(LaunchResizePass(tex.desc(), rt.desc(), m_axisX.table(), 0, m_pointY.base.ptr, dstRect.top, dstRect.Height(), m_stream), "resize X")) return MPCVR_E_FAIL;
    } else if (shaderY) {
        m_lastLaunches++;
        if (!Cuda(LaunchResizePass(tex.desc(), rt.desc(), m_axisY.table(), 1, m_pointX.base.ptr, dstRect.left, dstRect.Width(), m_stream), "resize Y")) return MPCVR_E_FAIL;
    } else {
        m_lastLaunches++;
        if (!Cuda(LaunchCopyPass(tex.desc(), rt.desc(), m_pointX.base.ptr, m_pointY.base.ptr, dstRect.left, dstRect.top, dstRect.Width(), dstRect.Height(), m_stream), "copy")) return MPCVR_E_FAIL;
    }
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::FinalPass(const Tex2D& tex, const Tex2D& rt, const CRect& rect)
{
    m_lastLaunches++;
    const float Q = (rt.fmt == MPCVR_SURF_RGB10A2) ? 1023.0f : 255.0f;
    return Cuda(LaunchFinalPass(tex.desc(), rt.desc(), rect.left, rect.top, rect.right, rect.bottom, Q, m_stream), "FinalPass") ? MPCVR_S_OK : MPCVR_E_FAIL;
}

int32_t CB200VideoProcessor::TonemapPass(const Tex2D& tex, const Tex2D& rt, const CRect& rect)
{
    m_lastLaunches++;
    return Cuda(LaunchTonemapPass(tex.desc(), rt.desc(), m_color, m_tm, rect.left, rect.top, rect.right, rect.bottom, m_stream), "HDR10 tone-map") ? MPCVR_S_OK : MPCVR_E_FAIL;
}

int32_t CB200VideoProcessor::Process(const Tex2D& target, const CRect& srcRect, const CRect& dstRect, bool skipConvert)
{
    (void)srcRect;
    int32_t hr = UpdatePlan();
    if (hr != MPCVR_S_OK) return hr;
    m_bDitherUsed = false;

    const bool covers = dstRect.left <= 0 && dstRect.top <= 0 && dstRect.right >= target.w && dstRect.bottom >= target.h;

    // ---- fused family
    if (!skipConvert && m_forcePath != MPCVR_PATH_MULTIPASS) {
        FusedPlan fp; const char* why = nullptr;
        // PATH_AUTO takes the specialised fused kernels (1:1, 2x); the generic table-driven fused kernel is currently slower than
        // the multipass chain on compute-bound plans (profiles/r01e_bench_all_configs.jsonl), so it runs only on request
        if (PlanFused(target, dstRect, fp, &why) && (m_forcePath == MPCVR_PATH_FUSED || FusedKind(fp) != 3)) {
            if (!covers) { m_lastLaunches++; if (!Cuda(LaunchClear(target.desc(), m_stream), "clear")) return MPCVR_E_FAIL; }
            int launches = 0;
            const cudaError_t fe = LaunchFused(fp, m_stream, &launches);
            if (fe != cudaErrorNotSupported || m_forcePath == MPCVR_PATH_FUSED) {      // "not supported" at launch time: the multipass chain renders the frame
                if (!Cuda(fe, "fused frame kernel")) return MPCVR_E_FAIL;
                m_lastLaunches += launches;
                m_path = MPCVR_PATH_FUSED; m_bDitherUsed = fp.final_pass != 0;
                return MPCVR_S_OK;
            }
            cudaGetLastError();
        }
        if (m_forcePath == MPCVR_PATH_FUSED) return Fail(MPCVR_E_NOTIMPL, why ? why : "plan has no fused form");
    }
    m_path = MPCVR_PATH_MULTIPASS;

    // ---- multipass family: the reference's passes, one kernel each
    if (!covers) { m_lastLaunches++; if (!Cuda(LaunchClear(target.desc(), m_stream), "clear")) return MPCVR_E_FAIL; }   // ClearRenderTargetView, :2617
    const Tex2D* pInputTexture = &m_TexConvertOutput;
    CRect rSrc(0, 0, m_TexConvertOutput.w, m_TexConvertOutput.h);
    if (!m_PSConvColorData.enable && !skipConvert) {
        // "pInputTexture = &m_TexSrcVideo" (:3320-3322): RGB source, default ProcAmp -> no ConvertColorPass.  The sample stays in
        // its own byte layout here, so one unpack kernel writes the texture the reference's CPU repack would have produced
        // (B8G8R8X8 / R10G10B10A2 / R16G16B16A16, Source/Helper.cpp:345-354) -- an exact, code-preserving copy -- and the
        // later passes sample that with rSrc = m_srcRect.
        const int fmt = m_srcParams.bytes == 1 ? MPCVR_SURF_BGRA8 : (m_srcParams.packed_kind == RGBK_R210 ? MPCVR_SURF_RGB10A2 : MPCVR_SURF_RGBA16);
        if (!Cuda(m_TexSrcUnpacked.CheckCreate(fmt, (int)m_srcWidth, (int)m_srcHeight), "m_TexSrcUnpacked")) return MPCVR_E_FAIL;
        SrcDesc whole = m_src; whole.rect_l = 0; whole.rect_t = 0; whole.rect_w = (int32_t)m_srcWidth; whole.rect_h = (int32_t)m_srcHeight;
        ColorParams ident = m_color; ident.transfer = TRANSFER_NONE;
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) ident.cm[i][j] = (i == j) ? 1.0f : 0.0f; ident.cc[i] = 0.0f; }
        m_lastLaunches++;
        if (!Cuda(LaunchConvertPass(whole, ident, m_TexSrcUnpacked.desc(), nullptr, m_stream), "source unpack")) return MPCVR_E_FAIL;
        pInputTexture = &m_TexSrcUnpacked; rSrc = m_srcRect;
    } else if (!skipConvert) { hr = ConvertColorPass(m_TexConvertOutput); if (hr != MPCVR_S_OK) return hr; }

    const bool correction = m_corr.mode != CORR_NONE;
    const int numSteps = (correction ? 1 : 0) + (m_bTonemap ? 1 : 0) + (m_bFinalPass ? 1 : 0);   // GetPostScaleSteps, :779-795
    if (numSteps) {                                                           // :3325-3416
        if (!Cuda(m_TexsPostScale.CheckCreate(m_InternalTexFmt, target.w, target.h), "m_TexsPostScale")) return MPCVR_E_FAIL;
        if (numSteps > 1 && !Cuda(m_TexsPostScale2.CheckCreate(m_InternalTexFmt, target.w, target.h), "m_TexsPostScale[1]")) return MPCVR_E_FAIL;
        const Tex2D* pTex = &m_TexsPostScale;
        CRect rect; rect.IntersectRect(dstRect, CRect(0, 0, pTex->w, pTex->h));
        if (rSrc != dstRect || m_iRotation != 0) { hr = ResizeShaderPass(*pInputTexture, *pTex, rSrc, dstRect); if (hr != MPCVR_S_OK) return hr; }
        else pTex = pInputTexture;                                            // "Hmm", :3351 (a flip alone is lost here, as in the reference)
        // StepSetting(), :3330-3339: ping-pong through the two-texture ring, the last step writes the back buffer
        int stepsLeft = numSteps;
        const Tex2D* ring[2] = { &m_TexsPostScale, &m_TexsPostScale2 };
        int next = (pTex == &m_TexsPostScale) ? 1 : 0;
        auto stepTarget = [&]() -> const Tex2D& { const Tex2D& rt = (--stepsLeft == 0) ? target : *ring[next]; next ^= 1; return rt; };
        if (correction) {                                                     // :3354-3357
            const Tex2D& rt = stepTarget();
            hr = CorrectionPass(*pTex, rt, rect); if (hr != MPCVR_S_OK) return hr;
            pTex = &rt;
        }
        if (m_bTonemap) {                                                     // :3359-3367
            const Tex2D& rt = stepTarget();
            hr = TonemapPass(*pTex, rt, rect); if (hr != MPCVR_S_OK) return hr;
            pTex = &rt;
        }
        if (m_bFinalPass) { hr = FinalPass(*pTex, target, rect); m_bDitherUsed = true; }
        return hr;
    }
    return ResizeShaderPass(*pInputTexture, target, rSrc, dstRect);           // :3418
}

// ------------------------------------------------------------------------------------------------
// fused planning
// ------------------------------------------------------------------------------------------------
bool CB200VideoProcessor::PlanFused(const Tex2D& target, const CRect& dstRect, FusedPlan& fp, const char** why)
{
    memset(&fp, 0, sizeof(fp));
    if (!m_PSConvColorData.enable) { if (why) *why = "no convert pass on this plan (RGB source, default ProcAmp): multipass only"; return false; }
    if (m_bTonemap) { if (why) *why = "HDR10 tone-map post-scale step: multipass only"; return false; }
    if (m_oriented) { if (why) *why = "rotation / flip: multipass only"; return false; }
    if (m_corr.mode != CORR_NONE) { if (why) *why = "correction shader step: multipass only"; return false; }
    if (m_Dovi.bValid) { if (why) *why = "Dolby Vision reshaping: multipass only"; return false; }
    if (m_jincX || m_jincY) { if (why) *why = "Jinc2: multipass only"; return false; }
    fp.src = m_src; fp.color = m_color;
    fp.ax = m_axisX.table(); fp.ay = m_axisY.table();
    fp.out = target.desc();
    fp.internal_fmt = m_InternalTexFmt;
    fp.final_pass = m_bFinalPass ? 1 : 0;
    const int kx = m_axisX.host.kind, ky = m_axisY.host.kind;
    fp.two_pass = ((kx == SCALER_INTERP || kx == SCALER_CONV) && (ky == SCALER_INTERP || ky == SCALER_CONV)) ? 1 : 0;
    fp.Q = (target.fmt == MPCVR_SURF_RGB10A2) ? 1023.0f : 255.0f;
    (void)dstRect;
    fp.hbase_x = m_axisX.host.base.data(); fp.hbase_y = m_axisY.host.base.data();
    // the register-blocked 2x kernel assumes periodic tables: output i = 2j+ph reads [j + b0 + ph, ...) with one of two weight vectors
    auto periodic = [&](const HostAxisTable& T, float* wE, float* wO, int& b0) {
        const int nt = T.ntaps;
        if (T.kind != SCALER_INTERP || T.dst_len < 2 || T.src_len * 2 != T.dst_len || nt > 8) return false;
        b0 = T.base[0];
        for (int k = 0; k < nt; k++) { wE[k] = T.w[k]; wO[k] = T.w[nt + k]; }
        for (int i = 0; i < T.dst_len; i++) {
            if (T.base[i] != (i >> 1) + b0 + (i & 1)) return false;
            if (memcmp(&T.w[(size_t)i * nt], (i & 1) ? wO : wE, nt * sizeof(float)) != 0) return false;
        }
        return true;
    };
    float wE2[8], wO2[8]; int b0y = 0;
    fp.periodic2x = (periodic(m_axisX.host, fp.wE, fp.wO, fp.b0) && periodic(m_axisY.host, wE2, wO2, b0y) && b0y == fp.b0 &&
                     m_axisX.host.ntaps == m_axisY.host.ntaps &&
                     memcmp(fp.wE, wE2, m_axisX.host.ntaps * sizeof(float)) == 0 && memcmp(fp.wO, wO2, m_axisX.host.ntaps * sizeof(float)) == 0) ? 1 : 0;
    return FusedSupported(fp, why);
}

// ------------------------------------------------------------------------------------------------
// images
// ------------------------------------------------------------------------------------------------
int32_t CB200VideoProcessor::GetCurentImage(void* host, size_t pitch, int32_t* pw, int32_t* ph)
{
    if (m_device < 0) return Fail(MPCVR_E_UNEXPECTED, "no CUDA device");
    if (!m_bInit || !m_haveFrame) return Fail(MPCVR_E_UNEXPECTED, "no frame");
    unsigned w = m_srcRectWidth, h = m_srcRectHeight;                          // :3495-3503
    if (m_srcAnamorphic) w = (unsigned)(((int64_t)h * m_srcAspectRatioX + m_srcAspectRatioY / 2) / m_srcAspectRatioY);   // MulDiv rounds to nearest
    if (m_iRotation == 90 || m_iRotation == 270) std::swap(w, h);               // :3500-3502
    if (pw) *pw = (int32_t)w; if (ph) *ph = (int32_t)h;
    if (!host) return MPCVR_S_OK;                                              // size query
    if (pitch < (size_t)w * 4) return MPCVR_E_INVALIDARG;
    cudaSetDevice(m_device);
    const CRect imageRect(0, 0, (int)w, (int)h);
    // temporarily m_videoRect = m_windowRect = imageRect, swap chain format B8G8R8X8 (:3519, :3547-3557)
    const CRect backupVid = m_videoRect, backupWnd = m_windowRect; const int backupFmt = m_SwapChainFmt; const Tex2D backupUser = m_UserBackBuffer;
    const bool backupFinal = m_bFinalPass;
    m_videoRect = imageRect; m_windowRect = imageRect; m_SwapChainFmt = MPCVR_SURF_BGRA8; m_UserBackBuffer = Tex2D();
    m_bFinalPass = m_bUseDither && (m_InternalTexFmt != MPCVR_SURF_BGRA8);
    m_planDirty = true;
    int32_t hr = MPCVR_S_OK;
    if (!Cuda(m_TexImage.CheckCreate(MPCVR_SURF_BGRA8, (int)w, (int)h), "image target")) hr = MPCVR_E_FAIL;
    if (hr == MPCVR_S_OK) hr = Process(m_TexImage, m_srcRect, imageRect, false);
    if (hr == MPCVR_S_OK && !Cuda(cudaMemcpy2DAsync(host, pitch, m_TexImage.ptr, m_TexImage.pitch, (size_t)w * 4, h, cudaMemcpyDeviceToHost, m_stream), "D2H image")) hr = MPCVR_E_FAIL;
    if (hr == MPCVR_S_OK && !Cuda(cudaStreamSynchronize(m_stream), "sync")) hr = MPCVR_E_FAIL;
    m_videoRect = backupVid; m_windowRect = backupWnd; m_SwapChainFmt = backupFmt; m_UserBackBuffer = backupUser; m_bFinalPass = backupFinal;
    m_planDirty = true;
    return hr;
}

int32_t CB200VideoProcessor::GetDisplayedImage(void* host, size_t pitch, int32_t* pw, int32_t* ph)
{
    if (m_device < 0) return Fail(MPCVR_E_UNEXPECTED, "no CUDA device");
    Tex2D& t = CurrentTarget();
    if (!t.ptr) return Fail(MPCVR_E_UNEXPECTED, "nothing displayed");
    if (pw) *pw = t.w; if (ph) *ph = t.h;
    if (!host) return MPCVR_S_OK;
    if (pitch < (size_t)t.w * 4) return MPCVR_E_INVALIDARG;
    cudaSetDevice(m_device);
    const Tex2D* src = &t;
    if (t.fmt != MPCVR_SURF_BGRA8) {
        if (!Cuda(m_TexImage.CheckCreate(MPCVR_SURF_BGRA8, t.w, t.h), "image target")) return MPCVR_E_FAIL;
        if (!Cuda(LaunchToBGRX(t.desc(), m_TexImage.desc(), m_stream), "to BGRX")) return MPCVR_E_FAIL;
        src = &m_TexImage;
    }
    if (!Cuda(cudaMemcpy2DAsync(host, pitch, src->ptr, src->pitch, (size_t)t.w * 4, t.h, cudaMemcpyDeviceToHost, m_stream), "D2H")) return MPCVR_E_FAIL;
    return Cuda(cudaStreamSynchronize(m_stream), "sync") ? MPCVR_S_OK : MPCVR_E_FAIL;
}

int32_t CB200VideoProcessor::GetVPInfo(char* dst, size_t n)
{
    if (!dst || !n) return MPCVR_E_POINTER;
    static const char* up[] = {"Nearest-neighbor", "Mitchell-Netravali", "Catmull-Rom", "Lanczos2", "Lanczos3", "Jinc2m"};
    static const char* down[] = {"Box", "Bilinear", "Hamming", "Bicubic", "Bicubic sharp", "Lanczos"};
    static const char* surf[] = {"B8G8R8A8_UNORM", "R10G10B10A2_UNORM", "R16G16B16A16_FLOAT"};
    cudaDeviceProp prop; memset(&prop, 0, sizeof(prop));
    if (m_device >= 0) cudaGetDeviceProperties(&prop, m_device);
    snprintf(dst, n,
             "MPC VR B200 frame path (ABI %d)\nGraphics adapter: %s\nVideoProcessor: CUDA shaders (%s)\n"
             "Input format : %s %ux%u\nInternalFormat: %s\nUpscaling : %s\nDownscaling: %s\nPostProcessing:%s%s\n",
             MPCVR_ABI_VERSION, m_device >= 0 ? prop.name : "none (planning only)",
             m_path == MPCVR_PATH_FUSED ? "fused" : "multipass", m_srcParams.str, m_srcRectWidth, m_srcRectHeight,
             surf[m_InternalTexFmt], up[std::min(std::max(m_iUpscaling, 0), 5)], down[std::min(std::max(m_iDownscaling, 0), 5)],
             m_color.transfer == TRANSFER_PQ_TO_SDR ? " PQ to SDR," : (m_color.transfer == TRANSFER_HLG_TO_SDR ? " HLG to SDR," : ""),
             m_bFinalPass ? " dither" : "");
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::SetSubCallback(mpcvr_sub_render_fn r, mpcvr_sub_set_device_fn d, void* user)
{
    if (m_subSetDevice && m_subUser) m_subSetDevice(m_subUser, -1);           // SetDevice11(NULL) on the outgoing consumer
    m_subRender = r; m_subSetDevice = d; m_subUser = user;
    if (m_subSetDevice) m_subSetDevice(m_subUser, m_device);
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::ActivateDeviceDecoding(int device, void* decodeStream, uint32_t)
{
    // ActivateD3D11Decoding -> SetDevice: the renderer adopts the decoder's device (Source/VideoRendererInputPin.cpp:186-197)
    (void)decodeStream;
    if (device != m_device) return Fail(MPCVR_E_FAIL, "decoder is on another CUDA device: create the renderer instance on the decoder's device");
    return MPCVR_S_OK;
}

// ------------------------------------------------------------------------------------------------
// debug hooks
// ------------------------------------------------------------------------------------------------
int32_t CB200VideoProcessor::DebugSetPath(int path) { if (path < 0 || path > 2) return MPCVR_E_INVALIDARG; m_forcePath = path; return MPCVR_S_OK; }

int32_t CB200VideoProcessor::DebugGetPlan(mpcvr_plan_info* info)
{
    if (!info) return MPCVR_E_POINTER;
    if (!m_bInit) return Fail(MPCVR_E_UNEXPECTED, "no media type");
    int32_t hr = UpdatePlan();
    if (hr != MPCVR_S_OK) return hr;
    memset(info, 0, sizeof(*info));
    info->internal_fmt = m_InternalTexFmt; info->final_pass = m_bFinalPass;
    info->scaler_x = m_kx; info->scaler_y = m_ky; info->correction = m_corr.mode;
    info->filter_x = m_axisX.host.filter; info->filter_y = m_axisY.host.filter;
    info->taps_x = m_axisX.host.ntaps; info->taps_y = m_axisY.host.ntaps;
    info->path = m_path; info->transfer = m_color.transfer; info->src_exfmt = m_srcExFmt.value;
    info->src_pitch = m_srcPitch; info->src_lines = (int32_t)m_srcLines; info->variant = m_bUseD3D11 ? 0 : 1;
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::DebugGetColorParams(float m9[9], float c3[3], float g9[9], float* lum)
{
    if (!m_bInit) return Fail(MPCVR_E_UNEXPECTED, "no media type");
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) { if (m9) m9[i*3+j] = m_color.cm[i][j]; if (g9) g9[i*3+j] = m_color.gamut[i][j]; } if (c3) c3[i] = m_color.cc[i]; }
    if (lum) *lum = m_color.lum_scale;
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::DebugGetDeviceParams(int which, void* dst, size_t cap)
{
    if (!m_bInit) return Fail(MPCVR_E_UNEXPECTED, "no media type");
    const void* src = nullptr; size_t n = 0;
    switch (which) {
    case 0: src = &m_color; n = sizeof(m_color); break;
    case 1: if (!m_Dovi.bValid) return Fail(MPCVR_E_UNEXPECTED, "no Dolby Vision data"); src = &m_doviHost; n = sizeof(m_doviHost); break;
    case 2: src = &m_tm; n = sizeof(m_tm); break;
    default: return Fail(MPCVR_E_INVALIDARG, "which: 0 ColorParams, 1 DoviParams, 2 TonemapParams");
    }
    if (dst && cap) memcpy(dst, src, std::min(cap, n));
    return (int32_t)n;
}

int32_t CB200VideoProcessor::DebugSetDovi(const mpcvr_dovi_metadata* md)
{
    if (!m_bInit) return Fail(MPCVR_E_UNEXPECTED, "no media type");
    if (!md) return Fail(MPCVR_E_POINTER, "null metadata");
    if (m_srcParams.CSType != CS_YUV || !CheckDoviMetadata(md)) return Fail(MPCVR_E_INVALIDARG, "metadata not accepted (CopySample would ignore it)");
    m_Dovi.msd = *md; m_Dovi.bValid = true;
    const int32_t hr = SetShaderDoviCurves();
    if (hr != MPCVR_S_OK) return hr;
    SetShaderConvertColorParams();
    return MPCVR_S_OK;
}

int32_t CB200VideoProcessor::DebugBuildSrcDesc(const void* const planes[3], const int32_t pitches[3], void* dst, size_t cap)
{
    if (!m_bInit) return Fail(MPCVR_E_UNEXPECTED, "no media type");
    if (!planes || !pitches) return Fail(MPCVR_E_POINTER, "null planes");
    // the assignments CopySample makes for a device-memory sample (see there), on a local block
    const FmtConvParams& f = m_srcParams;
    SrcDesc S;
    memset(&S, 0, sizeof(S));
    S.W = m_srcWidth; S.H = m_srcHeight;
    S.cw = m_srcWidth / f.div_chroma_w; S.ch = m_srcHeight / f.div_chroma_h;
    S.layout = f.layout; S.bytes = f.bytes; S.lsb_shift = f.lsb_shift; S.packed_kind = f.packed_kind; S.subsampling = f.Subsampling;
    S.chroma_mode = m_iChromaScaling; S.siting = m_srcExFmt.ChromaSubsampling();
    S.deint_blend = (m_bDeintBlend && m_bInterlaced && f.Subsampling == 420) ? 1 : 0;
    S.rect_l = m_srcRect.left; S.rect_t = m_srcRect.top; S.rect_w = m_srcRectWidth; S.rect_h = m_srcRectHeight;
    const int np = f.layout == LAYOUT_BIPLANAR ? 2 : (f.layout == LAYOUT_PLANAR3 ? 3 : 1);
    for (int i = 0; i < np; i++) {
        if (!planes[i] || pitches[i] == 0) return Fail(MPCVR_E_POINTER, "missing plane");
        S.pl[i] = (const uint8_t*)planes[i]; S.pitch[i] = pitches[i];
    }
    if (f.layout == LAYOUT_PLANAR3 && f.v_before_u) { std::swap(S.pl[1], S.pl[2]); std::swap(S.pitch[1], S.pitch[2]); }
    if (dst && cap) memcpy(dst, &S, std::min(cap, sizeof(S)));
    return (int32_t)sizeof(S);
}

int32_t CB200VideoProcessor::DebugAxisTaps(int axis, int i, int32_t* base, float* w, int maxw, float* wsum)
{
    int32_t hr = UpdatePlan();
    if (hr != MPCVR_S_OK) return hr;
    const HostAxisTable& T = axis == 0 ? m_axisX.host : m_axisY.host;
    if (i < 0 || i >= T.dst_len) return MPCVR_E_INVALIDARG;
    if (base) *base = T.base[i];
    for (int k = 0; k < T.ntaps && k < maxw; k++) w[k] = T.w[(size_t)i * T.ntaps + k];
    if (wsum) *wsum = T.wsum.empty() ? 1.0f : T.wsum[i];
    return T.ntaps;
}

int32_t CB200VideoProcessor::DebugCaptureF32(bool enable) { m_captureF32 = enable; return MPCVR_S_OK; }

int32_t CB200VideoProcessor::DebugReadSurface(int which, void* host, size_t pitch, int32_t* w, int32_t* h, int32_t* fmt)
{
    if (m_device < 0) return Fail(MPCVR_E_UNEXPECTED, "no CUDA device");
    cudaSetDevice(m_device);
    if (which == MPCVR_DBG_SURF_CONVERT_F32) {
        if (!m_dbgConvertF32.ptr) return Fail(MPCVR_E_UNEXPECTED, "f32 capture is off");
        if (w) *w = m_TexConvertOutput.w; if (h) *h = m_TexConvertOutput.h; if (fmt) *fmt = -1;
        if (!host) return MPCVR_S_OK;
        if (!Cuda(cudaMemcpy2DAsync(host, pitch, m_dbgConvertF32.ptr, (size_t)m_TexConvertOutput.w * 16, (size_t)m_TexConvertOutput.w * 16, m_TexConvertOutput.h, cudaMemcpyDeviceToHost, m_stream), "D2H")) return MPCVR_E_FAIL;
        return Cuda(cudaStreamSynchronize(m_stream), "sync") ? MPCVR_S_OK : MPCVR_E_FAIL;
    }
    const Tex2D* t = which == MPCVR_DBG_SURF_CONVERT ? &m_TexConvertOutput : (which == MPCVR_DBG_SURF_RESIZE_X ? &m_TexResize : &m_TexsPostScale);
    if (!t->ptr) return Fail(MPCVR_E_UNEXPECTED, "surface does not exist for this plan");
    if (w) *w = t->w; if (h) *h = t->h; if (fmt) *fmt = t->fmt;
    if (!host) return MPCVR_S_OK;
    if (!Cuda(cudaMemcpy2DAsync(host, pitch, t->ptr, t->pitch, (size_t)t->w * SurfBpp(t->fmt), t->h, cudaMemcpyDeviceToHost, m_stream), "D2H")) return MPCVR_E_FAIL;
    return Cuda(cudaStreamSynchronize(m_stream), "sync") ? MPCVR_S_OK : MPCVR_E_FAIL;
}

int32_t CB200VideoProcessor::DebugRenderFromConvert(const void* host, size_t pitch)
{
    if (m_device < 0) return Fail(MPCVR_E_UNEXPECTED, "no CUDA device");
    if (!m_bInit || !m_TexConvertOutput.ptr) return Fail(MPCVR_E_UNEXPECTED, "no media type");
    cudaSetDevice(m_device);
    if (!Cuda(cudaMemcpy2DAsync(m_TexConvertOutput.ptr, m_TexConvertOutput.pitch, host, pitch, (size_t)m_TexConvertOutput.w * SurfBpp(m_TexConvertOutput.fmt),
                                m_TexConvertOutput.h, cudaMemcpyHostToDevice, m_stream), "H2D convert surface")) return MPCVR_E_FAIL;
    m_lastLaunches = 0;
    if (!CurrentTarget().ptr) { int32_t hr = UpdatePostScaleTexures(); if (hr != MPCVR_S_OK) return hr; }
    return Process(CurrentTarget(), m_srcRect, m_videoRect, true);
}

int32_t CB200VideoProcessor::DebugProbePQ(const float* yuv3, float* rgb3, int n, int mode)
{
    if (m_device < 0) return Fail(MPCVR_E_UNEXPECTED, "no CUDA device");
    cudaSetDevice(m_device);
    float* d = nullptr;
    if (!Cuda(cudaMalloc(&d, (size_t)n * 6 * sizeof(float)), "probe alloc")) return MPCVR_E_FAIL;
    std::vector<float> planar((size_t)3 * n);
    for (int i = 0; i < n; i++) { planar[i] = yuv3[3 * i]; planar[n + i] = yuv3[3 * i + 1]; planar[2 * n + i] = yuv3[3 * i + 2]; }
    int32_t hr = MPCVR_E_FAIL;
    if (Cuda(cudaMemcpyAsync(d, planar.data(), (size_t)3 * n * sizeof(float), cudaMemcpyHostToDevice, m_stream), "H2D") &&
        Cuda(LaunchProbePQ(d, d + n, d + 2 * n, d + 3 * n, m_color, n, mode, m_stream), "probe") &&
        Cuda(cudaMemcpyAsync(rgb3, d + 3 * n, (size_t)3 * n * sizeof(float), cudaMemcpyDeviceToHost, m_stream), "D2H") &&
        Cuda(cudaStreamSynchronize(m_stream), "sync")) hr = MPCVR_S_OK;
    cudaFree(d);
    return hr;
}

}  // namespace b200vr
