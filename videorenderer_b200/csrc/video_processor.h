// video_processor.h -- CB200VideoProcessor: the host-side object behind the C ABI.
//
// Mirrors the slice of CVideoProcessor / CDX11VideoProcessor that drives the shader chain
// (Source/VideoProcessor.h:40-236, Source/DX11VideoProcessor.h:351-357): same member names where
// a counterpart exists, same decision rules, CUDA buffers and kernels in place of D3D11 textures
// and draw calls.  Everything about windows, swap chains, OSD, DXVA and DirectShow is absent
// (out of scope).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/mpcvr_b200.h"
#include "../../include/mpcvr_b200_debug.h"
#include "fmt_table.h"
#include "cspmath.h"
#include "frame_plan.h"
#include "resize_tables.h"
#include "kernels.h"

namespace b200vr {

struct CRect {
    int left = 0, top = 0, right = 0, bottom = 0;
    CRect() = default;
    CRect(int l, int t, int r, int b) : left(l), top(t), right(r), bottom(b) {}
    int Width() const { return right - left; }
    int Height() const { return bottom - top; }
    bool IsRectNull() const { return left == 0 && top == 0 && right == 0 && bottom == 0; }
    bool IsRectEmpty() const { return right <= left || bottom <= top; }
    bool operator==(const CRect& o) const { return left == o.left && top == o.top && right == o.right && bottom == o.bottom; }
    bool operator!=(const CRect& o) const { return !(*this == o); }
    bool IntersectRect(const CRect& a, const CRect& b) {
        left = a.left > b.left ? a.left : b.left; top = a.top > b.top ? a.top : b.top;
        right = a.right < b.right ? a.right : b.right; bottom = a.bottom < b.bottom ? a.bottom : b.bottom;
        if (IsRectEmpty()) { *this = CRect(); return false; }
        return true;
    }
};

// device allocation with a pitch (the role ID3D11Texture2D plays in the reference)
struct Tex2D {
    void*  ptr = nullptr;
    size_t pitch = 0;
    int    w = 0, h = 0, fmt = 0;
    size_t bytes = 0;
    bool   owned = true;
    SurfDesc desc() const { SurfDesc d; d.ptr = ptr; d.pitch = (int32_t)pitch; d.w = w; d.h = h; d.fmt = fmt; return d; }
    void Release();
    cudaError_t CheckCreate(int fmt, int w, int h);     // Tex2D_t::CheckCreate, Source/DX11Helper.h
};

template <typename T>
struct DevArray {
    T* ptr = nullptr; size_t n = 0;
    void Release() { if (ptr) cudaFree(ptr); ptr = nullptr; n = 0; }
    cudaError_t Upload(const std::vector<T>& v, cudaStream_t st) {
        if (v.size() > n) { Release(); cudaError_t e = cudaMalloc(&ptr, v.size() * sizeof(T)); if (e != cudaSuccess) return e; n = v.size(); }
        if (v.empty()) return cudaSuccess;
        return cudaMemcpyAsync(ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    }
};

struct DeviceAxis {
    HostAxisTable host;
    DevArray<int32_t> base; DevArray<float> w, wsum;
    AxisTable table() const;
    cudaError_t Upload(cudaStream_t st);
    void Release() { base.Release(); w.Release(); wsum.Release(); }
};

class CB200VideoProcessor {
public:
    CB200VideoProcessor();
    ~CB200VideoProcessor();

    int32_t Init(const mpcvr_settings* s, int device, void* stream);          // CVideoProcessor::Init

    // Source/VideoProcessor.h:178-189
    int32_t VerifyMediaType(const mpcvr_media_type* mt);
    int32_t InitMediaType(const mpcvr_media_type* mt);
    int32_t GetAlignmentSize(const mpcvr_media_type* mt, int32_t* cx, int32_t* cy);
    int32_t ProcessSample(const mpcvr_sample* s);
    int32_t Render(int field, int64_t frameStartTime);
    int32_t FillBlack();
    void    Flush();
    int32_t Reset(bool displayModeChange);
    int32_t WaitIdle();

    void    GetSourceRect(CRect& r) const { r = m_srcRect; }
    void    GetVideoRect(CRect& r) const { r = m_videoRect; }
    void    SetVideoRect(const CRect& r);
    int32_t SetWindowRect(const CRect& r);
    int32_t SetRotation(int v);
    int32_t SetFlip(bool v);
    int32_t SetDisplayHdr(bool supported);
    int32_t SetUpstreamConverter(bool present, bool pqSupported);
    void    GetVideoSize(int32_t& w, int32_t& h) const;
    void    GetAspectRatio(int32_t& ax, int32_t& ay) const;

    void    GetSettings(mpcvr_settings& s) const { s = m_settings; }
    void    Configure(const mpcvr_settings& s);                                // Source/DX11VideoProcessor.cpp:3800-4050
    bool    GetActive() const { return m_bInit && m_device >= 0; }
    int32_t SetProcAmpValues(uint32_t flags, const mpcvr_procamp* v);
    int32_t GetProcAmpValues(uint32_t flags, mpcvr_procamp* v) const;

    int32_t SetBackbufferFormat(int fmt);
    int32_t BindBackbuffer(void* dev, size_t pitch);
    int32_t GetBackbuffer(void** dev, size_t* pitch, int32_t* w, int32_t* h, int32_t* fmt);
    int32_t ReadBackbuffer(void* host, size_t pitch, bool sync = true);
    int32_t GetCurentImage(void* host, size_t pitch, int32_t* w, int32_t* h);  // sic, Source/VideoProcessor.h:220
    int32_t GetDisplayedImage(void* host, size_t pitch, int32_t* w, int32_t* h);
    int32_t GetVPInfo(char* dst, size_t n);

    int32_t SetSubCallback(mpcvr_sub_render_fn r, mpcvr_sub_set_device_fn d, void* user);
    int32_t ActivateDeviceDecoding(int device, void* decodeStream, uint32_t flags);
    int     GetDeviceIndex() const { return m_device; }
    cudaStream_t Stream() const { return m_stream; }
    int     LastFrameKernels() const { return m_lastLaunches; }
    const char* LastError() const { return m_lastError.c_str(); }

    // debug hooks (include/mpcvr_b200_debug.h)
    int32_t DebugSetPath(int path);
    int64_t DebugVerifyGammaTable() { if (m_device < 0) return -2; cudaSetDevice(m_device); return VerifyGammaTable(m_stream); }
    int32_t DebugGetPlan(mpcvr_plan_info* info);
    int32_t DebugGetColorParams(float m9[9], float c3[3], float g9[9], float* lum);
    int32_t DebugGetDeviceParams(int which, void* dst, size_t cap);
    int32_t DebugSetDovi(const mpcvr_dovi_metadata* md);
    int32_t DebugBuildSrcDesc(const void* const planes[3], const int32_t pitches[3], void* dst, size_t cap);
    int32_t DebugAxisTaps(int axis, int i, int32_t* base, float* w, int maxw, float* wsum);
    int32_t DebugCaptureF32(bool enable);
    int32_t DebugReadSurface(int which, void* host, size_t pitch, int32_t* w, int32_t* h, int32_t* fmt);
    int32_t DebugRenderFromConvert(const void* host, size_t pitch);
    int32_t DebugProbePQ(const float* yuv3, float* rgb3, int n, int mode);

private:
    // ---- settings (names as in Source/VideoProcessor.h:48-73)
    mpcvr_settings m_settings;
    int  m_iTexFormat = MPCVR_TEXFMT_AUTOINT;
    int  m_iChromaScaling = MPCVR_CHROMA_Bilinear;
    int  m_iUpscaling = MPCVR_UPSCALE_CatmullRom;
    int  m_iDownscaling = MPCVR_DOWNSCALE_Hamming;
    bool m_bInterpolateAt50pct = true;
    bool m_bUseDither = true;
    bool m_bDeintBlend = false;
    bool m_bConvertToSdr = true;
    int  m_iSDRDisplayNits = MPCVR_SDR_NITS_DEF;
    bool m_bUseD3D11 = true;                 // which shader set is restated (D3D11 / D3D9 twins)
    // HDR output (Source/VideoProcessor.h:66-71): the display capability comes from the caller (mpcvr_set_display_hdr)
    bool m_bHdrPassthroughSupport = false, m_bHdrPassthrough = true, m_bHdrLocalToneMapping = false;
    int  m_iHdrLocalToneMappingType = 1, m_iHdrDisplayMaxNits = 1000;
    void ApplySettings(const mpcvr_settings& s);

    // ---- input parameters (Source/VideoProcessor.h:80-95)
    FmtConvParams m_srcParams = GetFmtConvParams(MPCVR_CF_NONE);
    unsigned m_srcWidth = 0, m_srcHeight = 0, m_srcRectWidth = 0, m_srcRectHeight = 0;
    int      m_srcPitch = 0;
    unsigned m_srcLines = 0;
    uint32_t m_srcAspectRatioX = 0, m_srcAspectRatioY = 0;
    bool     m_srcAnamorphic = false;
    CRect    m_srcRect;
    ExtendedFormat m_decExFmt, m_srcExFmt;
    bool     m_bInterlaced = false;
    int64_t  m_rtAvgTimePerFrame = 0;
    CRect    m_videoRect, m_windowRect, m_renderRect;
    int      m_iRotation = 0;
    bool     m_bFlip = false;
    bool     m_bFinalPass = false, m_bDitherUsed = false;
    ProcAmp  m_ProcAmp;
    int64_t  m_rtStart = 0, m_rtEnd = 0;

    // ---- device state
    int          m_device = -1;
    cudaStream_t m_stream = nullptr;
    bool         m_ownStream = false;
    bool         m_bInit = false;
    bool         m_haveFrame = false;
    std::string  m_lastError;
    int          m_lastLaunches = 0;

    int m_InternalTexFmt = MPCVR_SURF_BGRA8;     // DXGI_FORMAT of intermediates, UpdateTexParams
    int m_SwapChainFmt = MPCVR_SURF_BGRA8;

    Tex2D m_TexSrcVideo[3];                      // plane buffers of the source "texture"
    SrcDesc m_src = {};                          // what the kernels read (own planes or the sample's device planes)
    Tex2D m_TexSrcUnpacked;      // RGB source as the texture the reference's CPU repack produces (only when no convert pass runs)
    Tex2D m_TexConvertOutput, m_TexResize, m_TexsPostScale, m_BackBuffer, m_TexImage;
    Tex2D m_UserBackBuffer;                      // caller-bound target (not owned)
    DevArray<float4> m_dbgConvertF32; bool m_captureF32 = false;

    // DXGI_HDR_METADATA_HDR10 subset kept per sample / last applied (HDRMetadata, Source/DX11VideoProcessor.h)
    struct Hdr10 { bool bValid = false; uint32_t MaxMasteringLuminance = 0, MinMasteringLuminance = 0; uint16_t MaxCLL = 0, MaxFALL = 0; };
    Hdr10 m_hdr10, m_lastHdr10;
    bool  m_hdrColorSpaceActive = false;         // m_currentSwapChainColorSpace == RGB_FULL_G2084_NONE_P2020
    bool  m_bTonemap = false;                    // m_pPSHDR10ToneMapping exists
    TonemapParams m_tm = {};
    Tex2D m_TexsPostScale2;                      // second texture of the m_TexsPostScale ring (two post-scale steps)
    bool  HdrOut() const { return m_bHdrPassthroughSupport && (m_bHdrPassthrough || m_bHdrLocalToneMapping); }
    bool  SourceIsHDR10orHLG() const {           // Source/VideoProcessor.h:239-242
        return (m_decExFmt.TransferFunction() == Trc_2084 && m_decExFmt.TransferMatrix() != Matrix_Unknown) || m_decExFmt.TransferFunction() == Trc_HLG; }
    void  UpdateHdrState();                      // the HDR part of Render(), :2626-2727
    void  SetHDR10ShaderParams(float minNits, float maxNits, float maxCLL, float maxFALL, float displayMaxNits, int type);   // :907-951
    int32_t TonemapPass(const Tex2D& tex, const Tex2D& rt, const CRect& rect);

    // Dolby Vision (m_Dovi, Source/VideoProcessor.h:110-114): reshaping data of the current frame and its device copy
    struct { mpcvr_dovi_metadata msd = {}; bool bValid = false; bool bHasMMR = false; } m_Dovi;
    bool m_bHdrPreferDoVi = false;
    DoviParams m_doviHost = {};
    // m_DoviExtensionMetadata (Source/DX11VideoProcessor.h:211-237): L1 (min, max, avg in nits) and the L2 trims stay set once a frame
    // carried them -- later frames without those blocks keep the last values -- until Flush() or a bHdrPreferDoVi change clears them
    bool  m_doviL1Present = false; float m_doviL1Nits[3] = {0, 0, 0};
    bool  m_doviL2Present = false; float m_doviL2[5] = {0, 0, 0, 0, 0};     // slope, offset, power, chroma weight, saturation gain (as DolbyConstants take them)
    // static HDR10 metadata derived from the RPU (:2472-2490): source_max_pq / source_min_pq, overridden by a Level 6 block
    uint32_t m_DoviMaxMasteringLuminance = 0, m_DoviMinMasteringLuminance = 0, m_DoviMaxContentLightLevel = 0, m_DoviMaxFrameAverageLightLevel = 0;
    DevArray<DoviParams> m_doviDev;
    bool    CheckDoviMetadata(const mpcvr_dovi_metadata* p) const;          // Source/VideoProcessor.cpp:273-297 (curve part)
    int32_t SetShaderDoviCurves();                                          // :990-1141

    // path-B correction step (m_pPSCorrection): the caller's RGB samples come from a fixed-function YCbCr->RGB converter
    bool m_upstreamVP = false, m_upstreamPq = false;
    ExtendedFormat m_upstreamExFmt;              // colour tags of the ORIGINAL stream (the RGB media type itself is untagged)
    CorrectionParams m_corr = {};                // .mode = the shader picked by UpdateCorrection (:1893-1931)
    void    UpdateCorrection();
    int32_t CorrectionPass(const Tex2D& tex, const Tex2D& rt, const CRect& rect);

    ColorTransform m_PSConvColorData = {};
    ColorParams    m_color = {};
    DeviceAxis m_axisX, m_axisY;                 // filtered axes
    DeviceAxis m_pointX, m_pointY;               // point-sampled index maps (other axis / copies)
    DeviceAxis m_pointRX, m_pointRY;             // second pass maps of the two-pass resize
    // rotated / flipped first draw (TextureResizeShader / TextureCopyRect with iRotation, bFlip; FillVertices :130-179)
    bool m_oriented = false;                     // m_iRotation != 0 || m_bFlip
    DeviceAxis m_oFilt, m_oU, m_oV;              // filtered table of the first draw; point maps output index -> texture x / y
    int  m_oInAxis = -1, m_oOutAxis = 0;         // texture axis the first draw filters (-1: ps_simple) and the output axis it feeds
    bool m_oTwoPass = false;
    // UPSCALE_Jinc2: m_pShaderUpscaleX/Y are the same one-pass 2-D shader (Source/DX11VideoProcessor.cpp:94-101)
    bool m_jincX = false, m_jincY = false;
    int  m_kx = 0, m_ky = 0;                     // scaler kind per direction as ResizeShaderPass picks it (:3106-3127)
    bool m_planDirty = true;
    int  m_forcePath = MPCVR_PATH_AUTO, m_path = MPCVR_PATH_MULTIPASS;

    mpcvr_sub_render_fn m_subRender = nullptr; mpcvr_sub_set_device_fn m_subSetDevice = nullptr; void* m_subUser = nullptr;

    // ---- helpers named after their reference counterparts
    void    UpdateTexParams(int cdepth);                         // :1143-1155
    void    SetShaderConvertColorParams();                       // :813-887
    void    SetShaderLuminanceParams();                          // :889-905
    void    UpdateRenderRect();
    int32_t UpdateTexures();                                     // :2869-2892 (sic)
    int32_t UpdatePostScaleTexures();                            // :2894-2912 (sic)
    int32_t UpdatePlan();
    int32_t CopySample(const mpcvr_sample* s);                   // :2202-2597
    int32_t MemCopyToTexSrcVideo(const uint8_t* data, int pitch);// :1213-1252
    int32_t Process(const Tex2D& target, const CRect& srcRect, const CRect& dstRect, bool skipConvert);   // :3285-3424
    int32_t ConvertColorPass(const Tex2D& rt);                   // :3048-3101
    int32_t ResizeShaderPass(const Tex2D& tex, const Tex2D& rt, const CRect& srcRect, const CRect& dstRect);  // :3103-3187
    int32_t FinalPass(const Tex2D& tex, const Tex2D& rt, const CRect& rect);                                  // :3189-3233
    int32_t ProcessFused(const Tex2D& target, const CRect& dstRect);
    bool    PlanFused(const Tex2D& target, const CRect& dstRect, FusedPlan& fp, const char** why);
    int32_t DrawSubtitles(const Tex2D& rt);                      // :3235-3282
    Tex2D&  CurrentTarget();
    int32_t Fail(int32_t hr, const char* what, cudaError_t e = cudaSuccess);
    bool    Cuda(cudaError_t e, const char* what);
    int     Transfer() const;
};

}  // namespace b200vr
