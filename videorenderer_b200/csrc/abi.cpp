// abi.cpp -- the extern "C" surface declared in include/mpcvr_b200.h and include/mpcvr_b200_debug.h.
// Thin: argument checks, then CB200VideoProcessor.  No exception crosses this boundary.
#include <cstring>
#include <new>
#include "video_processor.h"
#include "kernel_strip.h"

using b200vr::CB200VideoProcessor;
using b200vr::CRect;

struct mpcvr { CB200VideoProcessor vp; };

#define VP_OR(ret) if (!h) return (ret); CB200VideoProcessor& vp = h->vp
#define GUARD(expr) try { return (expr); } catch (...) { return MPCVR_E_FAIL; }

extern "C" {

void mpcvr_settings_default(mpcvr_settings* s)
{
    if (!s) return;
    memset(s, 0, sizeof(*s));
    s->bUseD3D11 = 1;                       // IsWindows8OrGreater()
    s->bShowStats = 0; s->iResizeStats = 0;
    s->iTexFormat = MPCVR_TEXFMT_AUTOINT;
    s->VPFmts_bNV12 = s->VPFmts_bP01x = s->VPFmts_bYUY2 = s->VPFmts_bOther = 1;
    s->iVPDeinterlacing = 1; s->bDeintDouble = 1; s->bVPScaling = 1; s->iVPSuperRes = 0; s->bVPRTXVideoHDR = 0;
    s->iChromaScaling = MPCVR_CHROMA_Bilinear;
    s->iUpscaling = MPCVR_UPSCALE_CatmullRom;
    s->iDownscaling = MPCVR_DOWNSCALE_Hamming;
    s->bInterpolateAt50pct = 1; s->bUseDither = 1; s->bDeintBlend = 0;
    s->iSwapEffect = 1; s->bExclusiveFS = 0; s->bVBlankBeforePresent = 0; s->bAdjustPresentTime = 1;
    s->bReinitByDisplay = 0; s->bHdrPreferDoVi = 0;
    s->bHdrPassthrough = 1; s->bHdrLocalToneMapping = 0; s->iHdrLocalToneMappingType = 1; s->iHdrDisplayMaxNits = 1000;
    s->iHdrToggleDisplay = 0; s->bConvertToSdr = 1; s->iHdrOsdBrightness = 0; s->iSDRDisplayNits = MPCVR_SDR_NITS_DEF;
}

int32_t mpcvr_abi_version(void) { return MPCVR_ABI_VERSION; }

int32_t mpcvr_create(const mpcvr_settings* settings, int32_t cuda_device, void* stream, mpcvr** out)
{
    if (!out) return MPCVR_E_POINTER;
    *out = nullptr;
    mpcvr* h = new (std::nothrow) mpcvr;
    if (!h) return MPCVR_E_OUTOFMEMORY;
    int32_t hr;
    try { hr = h->vp.Init(settings, cuda_device, stream); } catch (...) { hr = MPCVR_E_FAIL; }
    if (hr != MPCVR_S_OK) { delete h; return hr; }
    *out = h;
    return MPCVR_S_OK;
}

void mpcvr_destroy(mpcvr* h) { delete h; }

const char* mpcvr_last_error(const mpcvr* h) { return h ? h->vp.LastError() : "null instance"; }

int32_t mpcvr_verify_media_type(mpcvr* h, const mpcvr_media_type* mt) { VP_OR(0); GUARD(vp.VerifyMediaType(mt)); }
int32_t mpcvr_init_media_type(mpcvr* h, const mpcvr_media_type* mt)   { VP_OR(0); GUARD(vp.InitMediaType(mt)); }
int32_t mpcvr_get_alignment_size(mpcvr* h, const mpcvr_media_type* mt, int32_t* cx, int32_t* cy) { VP_OR(0); GUARD(vp.GetAlignmentSize(mt, cx, cy)); }

int32_t mpcvr_process_sample(mpcvr* h, const mpcvr_sample* s) { VP_OR(MPCVR_E_POINTER); GUARD(vp.ProcessSample(s)); }
int32_t mpcvr_render(mpcvr* h, int32_t field, int64_t t)      { VP_OR(MPCVR_E_POINTER); GUARD(vp.Render(field, t)); }
int32_t mpcvr_fill_black(mpcvr* h)                            { VP_OR(MPCVR_E_POINTER); GUARD(vp.FillBlack()); }
void    mpcvr_flush(mpcvr* h)                                 { if (h) h->vp.Flush(); }
int32_t mpcvr_reset(mpcvr* h, int32_t dm)                     { VP_OR(MPCVR_E_POINTER); GUARD(vp.Reset(dm != 0)); }
int32_t mpcvr_wait_idle(mpcvr* h)                             { VP_OR(MPCVR_E_POINTER); GUARD(vp.WaitIdle()); }

static void to_rect(const CRect& c, mpcvr_rect* r) { if (r) { r->left = c.left; r->top = c.top; r->right = c.right; r->bottom = c.bottom; } }
void mpcvr_get_source_rect(mpcvr* h, mpcvr_rect* r) { if (!h) return; CRect c; h->vp.GetSourceRect(c); to_rect(c, r); }
void mpcvr_get_video_rect(mpcvr* h, mpcvr_rect* r)  { if (!h) return; CRect c; h->vp.GetVideoRect(c); to_rect(c, r); }
void mpcvr_set_video_rect(mpcvr* h, const mpcvr_rect* r) { if (h && r) h->vp.SetVideoRect(CRect(r->left, r->top, r->right, r->bottom)); }
int32_t mpcvr_set_window_rect(mpcvr* h, const mpcvr_rect* r) { VP_OR(MPCVR_E_POINTER); if (!r) return MPCVR_E_POINTER; GUARD(vp.SetWindowRect(CRect(r->left, r->top, r->right, r->bottom))); }
int32_t mpcvr_set_rotation(mpcvr* h, int32_t d) { VP_OR(MPCVR_E_POINTER); return vp.SetRotation(d); }
int32_t mpcvr_set_flip(mpcvr* h, int32_t f)     { VP_OR(MPCVR_E_POINTER); return vp.SetFlip(f != 0); }
int32_t mpcvr_set_upstream_converter(mpcvr* h, int32_t present, int32_t pq) { VP_OR(MPCVR_E_POINTER); return vp.SetUpstreamConverter(present != 0, pq != 0); }
int32_t mpcvr_set_display_hdr(mpcvr* h, int32_t s) { VP_OR(MPCVR_E_POINTER); return vp.SetDisplayHdr(s != 0); }
void mpcvr_get_video_size(mpcvr* h, int32_t* w, int32_t* hh) { if (!h || !w || !hh) return; h->vp.GetVideoSize(*w, *hh); }
void mpcvr_get_aspect_ratio(mpcvr* h, int32_t* ax, int32_t* ay) { if (!h || !ax || !ay) return; h->vp.GetAspectRatio(*ax, *ay); }

void mpcvr_get_settings(mpcvr* h, mpcvr_settings* s) { if (h && s) h->vp.GetSettings(*s); }
void mpcvr_configure(mpcvr* h, const mpcvr_settings* s) { if (h && s) { try { h->vp.Configure(*s); } catch (...) {} } }
int32_t mpcvr_get_active(mpcvr* h) { return h && h->vp.GetActive() ? 1 : 0; }
int32_t mpcvr_set_procamp(mpcvr* h, uint32_t flags, const mpcvr_procamp* v) { VP_OR(MPCVR_E_POINTER); GUARD(vp.SetProcAmpValues(flags, v)); }
int32_t mpcvr_get_procamp(mpcvr* h, uint32_t flags, mpcvr_procamp* v) { VP_OR(MPCVR_E_POINTER); return vp.GetProcAmpValues(flags, v); }

int32_t mpcvr_set_backbuffer_format(mpcvr* h, int32_t fmt) { VP_OR(MPCVR_E_POINTER); GUARD(vp.SetBackbufferFormat(fmt)); }
int32_t mpcvr_bind_backbuffer(mpcvr* h, void* p, size_t pitch) { VP_OR(MPCVR_E_POINTER); GUARD(vp.BindBackbuffer(p, pitch)); }
int32_t mpcvr_get_backbuffer(mpcvr* h, void** p, size_t* pitch, int32_t* w, int32_t* hh, int32_t* fmt) { VP_OR(MPCVR_E_POINTER); return vp.GetBackbuffer(p, pitch, w, hh, fmt); }
int32_t mpcvr_read_backbuffer(mpcvr* h, void* dst, size_t pitch) { VP_OR(MPCVR_E_POINTER); GUARD(vp.ReadBackbuffer(dst, pitch, true)); }
int32_t mpcvr_read_backbuffer_async(mpcvr* h, void* dst, size_t pitch) { VP_OR(MPCVR_E_POINTER); GUARD(vp.ReadBackbuffer(dst, pitch, false)); }
int32_t mpcvr_get_current_image(mpcvr* h, void* dst, size_t pitch, int32_t* w, int32_t* hh) { VP_OR(MPCVR_E_POINTER); GUARD(vp.GetCurentImage(dst, pitch, w, hh)); }
int32_t mpcvr_get_displayed_image(mpcvr* h, void* dst, size_t pitch, int32_t* w, int32_t* hh) { VP_OR(MPCVR_E_POINTER); GUARD(vp.GetDisplayedImage(dst, pitch, w, hh)); }
int32_t mpcvr_get_vp_info(mpcvr* h, char* dst, size_t n) { VP_OR(MPCVR_E_POINTER); GUARD(vp.GetVPInfo(dst, n)); }

int32_t mpcvr_set_sub_callback(mpcvr* h, mpcvr_sub_render_fn r, mpcvr_sub_set_device_fn d, void* user) { VP_OR(MPCVR_E_POINTER); return vp.SetSubCallback(r, d, user); }
int32_t mpcvr_activate_device_decoding(mpcvr* h, int32_t dev, void* st, uint32_t flags) { VP_OR(MPCVR_E_POINTER); return vp.ActivateDeviceDecoding(dev, st, flags); }
int32_t mpcvr_get_device_index(mpcvr* h) { return h ? h->vp.GetDeviceIndex() : -1; }
int32_t mpcvr_get_stream(mpcvr* h, void** st) { VP_OR(MPCVR_E_POINTER); if (!st) return MPCVR_E_POINTER; *st = (void*)vp.Stream(); return MPCVR_S_OK; }
int32_t mpcvr_get_last_frame_kernels(mpcvr* h, int32_t* n) { VP_OR(MPCVR_E_POINTER); if (!n) return MPCVR_E_POINTER; *n = vp.LastFrameKernels(); return MPCVR_S_OK; }

// ---- debug surface
int64_t mpcvr_debug_verify_gamma_table(mpcvr* h) { VP_OR(-2); return vp.DebugVerifyGammaTable(); }
int32_t mpcvr_debug_set_resize_four(int32_t enable) { b200vr::SetResizeFour(enable); return MPCVR_S_OK; }
int32_t mpcvr_debug_set_convert_two_lane(int32_t enable) { b200vr::SetConvertTwoLane(enable); return MPCVR_S_OK; }
int32_t mpcvr_debug_set_gamma_table(int32_t enable) { b200vr::SetGammaTableEnabled(enable != 0); return MPCVR_S_OK; }
int32_t mpcvr_debug_set_fused1x_vector(int32_t enable) { b200vr::SetFused1xVector(enable); return MPCVR_S_OK; }    // 0 per-pixel kernel, 1 span kernel, 2 span kernel with 16 pixels per thread (8-bit sources)
int32_t mpcvr_debug_set_strip_kernel(int32_t enable, int32_t cfg) { b200vr::SetStripKernel(enable != 0, cfg); return MPCVR_S_OK; }
int32_t mpcvr_debug_strip_status(uint32_t out8[8], int32_t clear) { return b200vr::StripStatus(out8, clear) == 0 ? MPCVR_S_OK : MPCVR_E_FAIL; }
int32_t mpcvr_debug_fmt_params(int32_t cformat, int32_t out[12])
{
    if (!out || cformat <= 0 || cformat >= MPCVR_CF_COUNT) return MPCVR_E_INVALIDARG;
    const b200vr::FmtConvParams& p = b200vr::GetFmtConvParams(cformat);
    out[0] = p.cformat; out[1] = p.Packsize; out[2] = p.PitchCoeff; out[3] = p.CSType; out[4] = p.Subsampling; out[5] = p.CDepth;
    out[6] = p.layout; out[7] = p.bytes; out[8] = p.div_chroma_w; out[9] = p.div_chroma_h; out[10] = p.lsb_shift; out[11] = p.v_before_u ? 1 : 0;
    return MPCVR_S_OK;
}
uint32_t mpcvr_debug_specify_extended_format(uint32_t exfmt, int32_t cformat, uint32_t width, uint32_t height)
{
    if (cformat <= 0 || cformat >= MPCVR_CF_COUNT) return 0;
    b200vr::ExtendedFormat e; e.value = exfmt;
    return b200vr::SpecifyExtendedFormat(e, b200vr::GetFmtConvParams(cformat), width, height).value;
}
int32_t mpcvr_debug_render_n(mpcvr* h, int32_t n)
{
    VP_OR(MPCVR_E_POINTER);
    for (int32_t i = 0; i < n; i++) { const int32_t hr = vp.Render(1, 0); if (hr != MPCVR_S_OK) return hr; }
    return MPCVR_S_OK;
}
int32_t mpcvr_debug_set_path(mpcvr* h, int32_t path) { VP_OR(MPCVR_E_POINTER); return vp.DebugSetPath(path); }
int32_t mpcvr_debug_get_plan(mpcvr* h, mpcvr_plan_info* info) { VP_OR(MPCVR_E_POINTER); GUARD(vp.DebugGetPlan(info)); }
int32_t mpcvr_debug_get_color_params(mpcvr* h, float m9[9], float c3[3], float g9[9], float* lum) { VP_OR(MPCVR_E_POINTER); return vp.DebugGetColorParams(m9, c3, g9, lum); }
int32_t mpcvr_debug_get_device_params(mpcvr* h, int32_t which, void* dst, size_t cap) { VP_OR(MPCVR_E_POINTER); GUARD(vp.DebugGetDeviceParams(which, dst, cap)); }
int32_t mpcvr_debug_set_dovi(mpcvr* h, const mpcvr_dovi_metadata* md) { VP_OR(MPCVR_E_POINTER); GUARD(vp.DebugSetDovi(md)); }
int32_t mpcvr_debug_build_src_desc(mpcvr* h, const void* const planes[3], const int32_t pitches[3], void* dst, size_t cap) { VP_OR(MPCVR_E_POINTER); GUARD(vp.DebugBuildSrcDesc(planes, pitches, dst, cap)); }
int32_t mpcvr_debug_axis_taps(mpcvr* h, int32_t axis, int32_t i, int32_t* base, float* w, int32_t maxw, float* wsum) { VP_OR(MPCVR_E_POINTER); GUARD(vp.DebugAxisTaps(axis, i, base, w, maxw, wsum)); }
int32_t mpcvr_debug_capture_f32(mpcvr* h, int32_t en) { VP_OR(MPCVR_E_POINTER); return vp.DebugCaptureF32(en != 0); }
int32_t mpcvr_debug_read_surface(mpcvr* h, int32_t which, void* dst, size_t pitch, int32_t* w, int32_t* hh, int32_t* fmt) { VP_OR(MPCVR_E_POINTER); GUARD(vp.DebugReadSurface(which, dst, pitch, w, hh, fmt)); }
int32_t mpcvr_debug_render_from_convert(mpcvr* h, const void* surf, size_t pitch) { VP_OR(MPCVR_E_POINTER); GUARD(vp.DebugRenderFromConvert(surf, pitch)); }

int32_t mpcvr_debug_probe_pq(mpcvr* h, const float* yuv3, float* rgb3, int32_t n, int32_t mode) { VP_OR(MPCVR_E_POINTER); GUARD(vp.DebugProbePQ(yuv3, rgb3, n, mode)); }

int32_t mpcvr_debug_math(int32_t dev, int32_t op, const float* x, const float* y, float* out, int64_t n)
{
    if (!x || !out || n <= 0) return MPCVR_E_POINTER;
    if (cudaSetDevice(dev) != cudaSuccess) return MPCVR_E_FAIL;
    float *dx = nullptr, *dy = nullptr, *dout = nullptr;
    int32_t hr = MPCVR_E_FAIL;
    if (cudaMalloc(&dx, n * 4) == cudaSuccess && cudaMalloc(&dout, n * 4) == cudaSuccess && (!y || cudaMalloc(&dy, n * 4) == cudaSuccess)) {
        cudaMemcpy(dx, x, n * 4, cudaMemcpyHostToDevice);
        if (y) cudaMemcpy(dy, y, n * 4, cudaMemcpyHostToDevice);
        if (b200vr::LaunchMathProbe(op, dx, dy ? dy : dx, dout, n, 0) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess &&
            cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost) == cudaSuccess) hr = MPCVR_S_OK;
    }
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return hr;
}

}  // extern "C"
