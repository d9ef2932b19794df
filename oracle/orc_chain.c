/*
 * oracle/orc_chain.c -- TEST INFRASTRUCTURE ONLY (see orc_types.h).
 * Instantiates orc_core.inc for the three arithmetic modes and exports the C API.
 * Small decision rules restated here:
 *   UpdateTexParams           Source/DX11VideoProcessor.cpp:1143-1155
 *   UpdatePostScaleTexures    Source/DX11VideoProcessor.cpp:2894-2907
 *   ResizeShaderPass scaler   Source/DX11VideoProcessor.cpp:3106-3127
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <omp.h>
#include "orc_types.h"
#include "orc_api.h"
#include "orc_math.h"

int orc_internal_fmt(int iTexFormat, int cdepth)
{
    switch (iTexFormat) {
    case ORC_TEXFMT_8INT:    return ORC_SURF_BGRA8;
    case ORC_TEXFMT_10INT:   return ORC_SURF_RGB10A2;
    case ORC_TEXFMT_16FLOAT: return ORC_SURF_RGBA16F;
    default:                 return cdepth > 8 ? ORC_SURF_RGB10A2 : ORC_SURF_BGRA8;
    }
}

int orc_need_final_pass(int backbuffer_fmt, int internal_fmt, int bUseDither)
{
    const int need = (backbuffer_fmt == ORC_SURF_BGRA8 && internal_fmt != ORC_SURF_BGRA8)
                  || (backbuffer_fmt == ORC_SURF_RGB10A2 && internal_fmt == ORC_SURF_RGBA16F);
    return bUseDither && need;
}

int orc_scaler_kind(int src, int dst, int bInterpolateAt50pct)
{
    const int k = bInterpolateAt50pct ? 2 : 1;
    if (src == dst) return 0;
    return (src > k * dst) ? 2 : 1;
}

#define ORC_MODE 0
#define ORC_SUFFIX _spec
#include "orc_core.inc"
#undef ORC_MODE
#undef ORC_SUFFIX

#define ORC_MODE 1
#define ORC_SUFFIX _lit
#include "orc_core.inc"
#undef ORC_MODE
#undef ORC_SUFFIX

#define ORC_MODE 2
#define ORC_SUFFIX _f64
#include "orc_core.inc"
#undef ORC_MODE
#undef ORC_SUFFIX

int orc_num_threads(void) { return omp_get_max_threads(); }

/* The transfer pieces of the convert shader as the oracle restates them (orc_core.inc: st2084_to_linear, linear_to_st2084, inverse_hlg,
 * hable / ToneMappingHable, the HLG OOTF lines of convert_color), in any arithmetic mode -- tests/test_reference_hlsl_pin.py runs the
 * reference's own .hlsl include files as C++ (oracle/_ref/libref_hlsl.so) and compares LITERAL mode bit for bit.
 * fn: 0 ST2084ToLinear(v, param)  1 LinearToST2084(v, param)  2 inverse_HLG(v)  3 ToneMappingHable(v);  mode: 0 SPEC, 1 LITERAL, 2 F64 */
int orc_test_transfer(int fn, int mode, const float* in, float* out, int n, float param)
{
    for (int i = 0; i < n; i++) {
        const float v = in[i];
        if (mode == 0)      out[i] = fn == 0 ? st2084_to_linear_spec(v, param) : fn == 1 ? linear_to_st2084_spec(v, param) : fn == 2 ? inverse_hlg_spec(v)
                                   : hable_spec(v) * (1.0f / hable_spec(4.8f));
        else if (mode == 1) out[i] = fn == 0 ? st2084_to_linear_lit(v, param) : fn == 1 ? linear_to_st2084_lit(v, param) : fn == 2 ? inverse_hlg_lit(v)
                                   : hable_lit(v) / hable_lit(4.8f);
        else                out[i] = (float)(fn == 0 ? st2084_to_linear_f64(v, param) : fn == 1 ? linear_to_st2084_f64(v, param) : fn == 2 ? inverse_hlg_f64(v)
                                   : hable_f64(v) / hable_f64(4.8));
    }
    return (fn < 0 || fn > 3) ? -1 : 0;
}
/* HLGtoLinear (hlg.hlsl:13-20) as convert_color states it (orc_core.inc, "hlg.hlsl:13-20"): inverse_HLG per channel, the OOTF gain from the
 * BT.2020 luminance, n RGB triplets, LITERAL mode */
int orc_test_hlg_to_linear_literal(const float* in, float* out, int n)
{
    for (int i = 0; i < n; i++) {
        float l[3];
        for (int k = 0; k < 3; k++) l[k] = inverse_hlg_lit(in[3 * i + k]);
        float dt = 0.2627f * l[0]; dt = 0.6780f * l[1] + dt; dt = 0.0593f * l[2] + dt;
        const float ys = 2000.0f * dt;
        const float g = orc_pow_literal(ys, 0.2f);
        for (int k = 0; k < 3; k++) out[3 * i + k] = l[k] * g;
    }
    return 0;
}

/* One resize draw (resize_pass: TextureResizeShader with one of ps_interpolation_*.hlsl / ps_convolution.hlsl along `axis`) and the final
 * pass (final_pass: ps_final_pass.hlsl) on RAW float surfaces, i.e. the shader's float4 result before any render-target conversion --
 * tests/test_reference_hlsl_pin.py runs the reference's own pixel shaders as C++ (oracle/_ref/libref_hlsl_shaders.so) on the same texels and
 * compares LITERAL mode bit for bit.  in: w*h RGB texels; out: ow*oh RGB (the out rect is the whole output surface).  mode 0 SPEC, 1 LITERAL. */
int orc_test_resize_draw(int mode, int axis, int kind, int filter, int variant, const float* in, int w, int h, const int in_rect[4],
                         int ow, int oh, float* out)
{
    const int out_rect[4] = {0, 0, ow, oh};
    if (mode == 0) {
        surf_spec a = {w, h, ORC_SURF_RAW, (float*)in}, b = {ow, oh, ORC_SURF_RAW, out};
        return resize_pass_spec(&a, &b, axis, in_rect, out_rect, kind, filter, variant, 1);
    }
    if (mode == 1) {
        surf_lit a = {w, h, ORC_SURF_RAW, (float*)in}, b = {ow, oh, ORC_SURF_RAW, out};
        return resize_pass_lit(&a, &b, axis, in_rect, out_rect, kind, filter, variant, 1);
    }
    return -1;
}
/* one Jinc2 draw (jinc_pass: ps_resize_onepass_jinc2.hlsl) of in_rect of a w x h RAW surface onto ow x oh pixels, mode 0 SPEC / 1 LITERAL */
int orc_test_jinc_draw(int mode, const float* in, int w, int h, const int in_rect[4], int ow, int oh, float* out)
{
    const int out_rect[4] = {0, 0, ow, oh};
    if (mode == 0) { surf_spec a = {w, h, ORC_SURF_RAW, (float*)in}, b = {ow, oh, ORC_SURF_RAW, out}; return jinc_pass_spec(&a, &b, in_rect, out_rect, 1); }
    if (mode == 1) { surf_lit a = {w, h, ORC_SURF_RAW, (float*)in}, b = {ow, oh, ORC_SURF_RAW, out}; return jinc_pass_lit(&a, &b, in_rect, out_rect, 1); }
    return -1;
}
/* the first draw of a turned frame (oriented_pass: SetRotation / SetFlip, FillVertices DX11VideoProcessor.cpp:130-179) with the X shader
 * (in_axis 0), the Y shader (1) or ps_simple (-1), and the turned Jinc2 draw (in_axis 2) -- RAW surfaces, mode 0 SPEC / 1 LITERAL */
int orc_test_oriented_draw(int mode, int in_axis, int kind, int filter, int variant, int rot, int flip, const float* in, int w, int h,
                           const int in_rect[4], int ow, int oh, float* out)
{
    const int out_rect[4] = {0, 0, ow, oh};
    if (mode == 0) {
        surf_spec a = {w, h, ORC_SURF_RAW, (float*)in}, b = {ow, oh, ORC_SURF_RAW, out};
        return in_axis == 2 ? jinc_pass_o_spec(&a, &b, in_rect, out_rect, rot, flip, 1) : oriented_pass_spec(&a, &b, in_rect, out_rect, rot, flip, in_axis, kind, filter, variant, 1);
    }
    if (mode == 1) {
        surf_lit a = {w, h, ORC_SURF_RAW, (float*)in}, b = {ow, oh, ORC_SURF_RAW, out};
        return in_axis == 2 ? jinc_pass_o_lit(&a, &b, in_rect, out_rect, rot, flip, 1) : oriented_pass_lit(&a, &b, in_rect, out_rect, rot, flip, in_axis, kind, filter, variant, 1);
    }
    return -1;
}
int orc_test_final_pass(int mode, const float* in, int w, int h, float Q, float* out)
{
    const int rect[4] = {0, 0, w, h};
    if (mode == 0) { surf_spec a = {w, h, ORC_SURF_RAW, (float*)in}, b = {w, h, ORC_SURF_RAW, out}; final_pass_spec(&a, &b, rect, Q, 1); return 0; }
    if (mode == 1) { surf_lit a = {w, h, ORC_SURF_RAW, (float*)in}, b = {w, h, ORC_SURF_RAW, out}; final_pass_lit(&a, &b, rect, Q, 1); return 0; }
    return -1;
}

/* The D3D11 render-target write followed by the read a later pass performs (rt_roundtrip, SPEC), and the final pass's quantiser
 * floor(p * Q + d) / Q (final_pass) -- tests/test_device_code_on_host.py holds the product's surf_store / surf_load / final_quant, compiled for
 * the host, to them.  fmt = ORC_SURF_* */
int orc_test_rt_roundtrip(int fmt, const float* in, float* out, int n)
{
    for (int i = 0; i < n; i++) out[i] = rt_roundtrip_spec(fmt, in[i]);
    return 0;
}
int orc_test_final_quant(const float* p, const float* d, float Q, float* out, int n)
{
    for (int i = 0; i < n; i++) out[i] = floorf(fmaf(p[i], Q, d[i])) / Q;
    return 0;
}

/* tm_pixel (ps_hdr10_tonemap.hlsl's main) on n RGB PQ pixels with the constant-buffer values given directly: hp = MasteringMinLuminanceNits,
 * MasteringMaxLuminanceNits, maxCLL, maxFALL, displayMaxNits, selection; l2 = {ChromaWeight, SaturationGain, TrimSlope, TrimOffset, TrimPower}
 * or NULL -- tests/test_reference_convert_exec_pin.py runs the reference's shader file as C++ on the same pixels.  mode 0 SPEC, 1 LITERAL. */
int orc_test_tonemap(int mode, const float hp[6], const float* l2, const float* in, float* out, int n)
{
    if (mode == 0) {
        tmparams_spec T = { hp[0], hp[1], hp[2], hp[3], hp[4], (int)hp[5], l2 != NULL, 0, 0, 0, 0, 0 };
        if (l2) { T.l2_cw = l2[0]; T.l2_sg = l2[1]; T.l2_slope = l2[2]; T.l2_offset = l2[3]; T.l2_power = l2[4]; }
        for (int i = 0; i < n; i++) tm_pixel_spec(&T, in + 3 * i, out + 3 * i);
        return 0;
    }
    if (mode == 1) {
        tmparams_lit T = { hp[0], hp[1], hp[2], hp[3], hp[4], (int)hp[5], l2 != NULL, 0, 0, 0, 0, 0 };
        if (l2) { T.l2_cw = l2[0]; T.l2_sg = l2[1]; T.l2_slope = l2[2]; T.l2_offset = l2[3]; T.l2_power = l2[4]; }
        for (int i = 0; i < n; i++) tm_pixel_lit(&T, in + 3 * i, out + 3 * i);
        return 0;
    }
    return -1;
}

/* corr_pixel (the six post-scale correction shaders) on n RGB pixels; consts (18 floats, optional) = the fix-BT.2020 / fix-YCgCo matrix of the
 * mode and the 2020 -> 709 gamut matrix as corr_constants evaluates them -- tests/test_reference_convert_exec_pin.py runs the shader files. */
int orc_test_correction(int mode, int corr, float lum_scale, const float* in, float* out, int n, float* consts)
{
    if (mode == 0) {
        float f20[3][3], fyc[3][3], gm[3][3];
        corr_constants_spec(f20, fyc, gm);
        if (consts) for (int i = 0; i < 9; i++) { consts[i] = corr == ORC_CORR_FIX_YCGCO ? fyc[i / 3][i % 3] : f20[i / 3][i % 3]; consts[9 + i] = gm[i / 3][i % 3]; }
        for (int i = 0; i < n; i++) corr_pixel_spec(corr, lum_scale, f20, fyc, gm, in + 3 * i, out + 3 * i);
        return 0;
    }
    if (mode == 1) {
        float f20[3][3], fyc[3][3], gm[3][3];
        corr_constants_lit(f20, fyc, gm);
        if (consts) for (int i = 0; i < 9; i++) { consts[i] = corr == ORC_CORR_FIX_YCGCO ? fyc[i / 3][i % 3] : f20[i / 3][i % 3]; consts[9 + i] = gm[i / 3][i % 3]; }
        for (int i = 0; i < n; i++) corr_pixel_lit(corr, lum_scale, f20, fyc, gm, in + 3 * i, out + 3 * i);
        return 0;
    }
    return -1;
}

/* The DolbyConstants the oracle derives for a display (SPEC mode's build_dovi: Level 2 trim selection / interpolation,
 * DX11VideoProcessor.cpp:2383-2469, and the +-0.5 of SetDolbyVisionDynamicParams :956-960): {chroma_weight, saturation_gain, slope, offset, power}.
 * Returns l2_enabled. */
int orc_test_dovi_trims(const orc_dovi* dv, int display_nits, float out5[5])
{
    static cparams_spec P;
    memset(&P, 0, sizeof(P));
    build_dovi_spec(dv, &P, display_nits);
    out5[0] = (float)P.l2_chroma_weight; out5[1] = (float)P.l2_saturation_gain; out5[2] = (float)P.l2_slope; out5[3] = (float)P.l2_offset; out5[4] = (float)P.l2_power;
    return P.l2_enabled;
}

/* Trace of the resize stage's draws (orc_types.h: orc_trace_entry): begin with a buffer, run ONE orc_process_from_convert / orc_process
 * call in SPEC mode from this thread, end returns the number of records written. */
extern orc_trace_entry* orc_g_trace; extern int orc_g_trace_cap, orc_g_trace_n;
void orc_test_trace_begin(orc_trace_entry* buf, int cap) { orc_g_trace_n = 0; orc_g_trace_cap = cap; orc_g_trace = buf; }
int  orc_test_trace_end(void) { orc_g_trace = NULL; return orc_g_trace_n; }

/* PS_COLOR_TRANSFORM and m_PSConvColorData.bEnable as the oracle derives them (SPEC mode's build_cparams: the restatement of
 * SetShaderConvertColorParams, DX11VideoProcessor.cpp:813-873), handed out so that tests/test_reference_shader_pin.py can compare
 * them with the reference's own function compiled in place (oracle/_ref/libref_shaders.so: ref_convert_color_params).  exfmt_out = the
 * extended format after SpecifyExtendedFormat, which is what the reference's m_srcExFmt holds at that point. */
int orc_test_convert_params(const orc_config* cfg, const orc_dovi* dv, float cm[3][3], float cc[3], int32_t* enable, uint32_t* exfmt_out)
{
    const orc_fmt* f = orc_get_fmt(cfg->cformat);
    if (!f) return -10;
    const int sw = cfg->src_rect[2] - cfg->src_rect[0], sh = cfg->src_rect[3] - cfg->src_rect[1];
    const uint32_t exfmt = orc_specify_extended_format(f->cstype == 0 ? cfg->exfmt : 0u, f->cstype, f->subsampling, sw, sh);
    static cparams_spec P;
    memset(&P, 0, sizeof(P));
    if (build_cparams_spec(cfg, f, exfmt, &P, dv)) return -11;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) cm[i][j] = (float)P.cm[i][j]; cc[i] = (float)P.cc[i]; }
    *enable = P.enable; *exfmt_out = exfmt;
    return 0;
}

/* convert_color (the body of the generated convert shader) on n yuv triplets, with the constants build_cparams derives for cfg / dv exactly as
 * orc_process does -- tests/test_device_code_on_host.py runs the product's device function convert_color() compiled for the host on the same
 * triplets.  mode 0 SPEC, 1 LITERAL, 2 F64. */
int orc_test_convert_pixels(const orc_config* cfg, const orc_dovi* dv, int mode, const float* yuv, float* out, int n)
{
    const orc_fmt* f = orc_get_fmt(cfg->cformat);
    if (!f) return -10;
    const int sw = cfg->src_rect[2] - cfg->src_rect[0], sh = cfg->src_rect[3] - cfg->src_rect[1];
    const uint32_t exfmt = orc_specify_extended_format(f->cstype == 0 ? cfg->exfmt : 0u, f->cstype, f->subsampling, sw, sh);
    if (mode == 0) {
        static cparams_spec P; memset(&P, 0, sizeof(P));
        if (build_cparams_spec(cfg, f, exfmt, &P, dv)) return -11;
        for (int i = 0; i < n; i++) convert_color_spec(&P, yuv + 3 * i, out + 3 * i);
        return 0;
    }
    if (mode == 1) {
        static cparams_lit P; memset(&P, 0, sizeof(P));
        if (build_cparams_lit(cfg, f, exfmt, &P, dv)) return -11;
        for (int i = 0; i < n; i++) convert_color_lit(&P, yuv + 3 * i, out + 3 * i);
        return 0;
    }
    if (mode == 2) {                                        /* float64 twin: the same text evaluated in double, rounded to float at the end */
        static cparams_f64 P; memset(&P, 0, sizeof(P));
        if (build_cparams_f64(cfg, f, exfmt, &P, dv)) return -11;
        for (int i = 0; i < n; i++) {
            const double in[3] = { yuv[3 * i], yuv[3 * i + 1], yuv[3 * i + 2] }; double o[3];
            convert_color_f64(&P, in, o);
            out[3 * i] = (float)o[0]; out[3 * i + 1] = (float)o[1]; out[3 * i + 2] = (float)o[2];
        }
        return 0;
    }
    return -1;
}

/* The PS_DOVI_CURVE constants the oracle derives from Dolby Vision reshaping metadata (SPEC mode's build_dovi: the restatement of
 * SetShaderDoviCurves / SetShaderDoviCurvesPoly, DX11VideoProcessor.cpp:990-1122), handed out so that tests/test_reference_dovi_pin.py
 * can compare them with the reference's own functions compiled in place (oracle/_ref/libref_dovi.so).  par[c] = {methods, mmr_single,
 * min_order, max_order}. */
int orc_test_dovi_curves(const orc_dovi* dv, float piv[3][7], float coef[3][8][4], float mmr[3][48][4], int32_t par[3][4])
{
    static cparams_spec P;                  /* large (mmr tables): static; the tests call this from one thread */
    memset(&P, 0, sizeof(P));
    build_dovi_spec(dv, &P, 100);
    for (int c = 0; c < 3; c++) {
        for (int i = 0; i < 7; i++) piv[c][i] = (float)P.dv_pivots[c][i];
        for (int i = 0; i < 8; i++) for (int k = 0; k < 4; k++) coef[c][i][k] = (float)P.dv_coeffs[c][i][k];
        for (int i = 0; i < 48; i++) for (int k = 0; k < 4; k++) mmr[c][i][k] = (float)P.dv_mmr[c][i][k];
        par[c][0] = P.dv_methods[c]; par[c][1] = P.dv_mmr_single[c]; par[c][2] = P.dv_min_order[c]; par[c][3] = P.dv_max_order[c];
    }
    return P.dovi_mmr;
}

/* Level 1 (+ Level 3 offsets) of the Dolby Vision extension blocks -> nits, as CopySample does before feeding
 * SetHDR10ShaderParams (DX11VideoProcessor.cpp:2347-2381, :2717-2720).  Returns 1 when a Level 1 block exists. */
int orc_dovi_l1_nits(const orc_dovi* dv, float out_min_max_avg[3])
{
    const float PQ_M1 = 2610.f / (4096.f * 4.f), PQ_M2 = 2523.f / 4096.f * 128.f, PQ_C1 = 3424.f / 4096.f,
                PQ_C2 = 2413.f / 4096.f * 32.f, PQ_C3 = 2392.f / 4096.f * 32.f;
    for (int i = 0; i < 32; i++) {
        if (dv->ext[i].level != 1) continue;
        uint16_t v[3] = { dv->ext[i].v[0], dv->ext[i].v[1], dv->ext[i].v[2] };
        for (int k = 0; k < 32; k++)
            if (dv->ext[k].level == 3) { for (int j = 0; j < 3; j++) v[j] = (uint16_t)(v[j] + dv->ext[k].v[j] - 2048); break; }
        for (int j = 0; j < 3; j++) {
            float x = v[j] / 4095.f;
            x = powf(x, 1.0f / PQ_M2);
            x = fmaxf(x - PQ_C1, 0.0f) / (PQ_C2 - PQ_C3 * x);
            x = powf(x, 1.0f / PQ_M1);
            out_min_max_avg[j] = (float)(uint16_t)(unsigned int)(x * 10000.0f);
        }
        return 1;
    }
    return 0;
}

/* PqToLinearNits of CopySample (DX11VideoProcessor.cpp:2329-2334): host arithmetic as written there (fp32, libm powf) */
float orc_pq_to_linear_nits(float x)
{
    const float PQ_M1 = 2610.f / (4096.f * 4.f), PQ_M2 = 2523.f / 4096.f * 128.f, PQ_C1 = 3424.f / 4096.f,
                PQ_C2 = 2413.f / 4096.f * 32.f, PQ_C3 = 2392.f / 4096.f * 32.f;
    x = powf(x, 1.0f / PQ_M2);
    x = fmaxf(x - PQ_C1, 0.0f) / (PQ_C2 - PQ_C3 * x);
    x = powf(x, 1.0f / PQ_M1);
    return x * 10000.0f;
}

int orc_process_dovi(const orc_config* cfg, const void* const planes[3], const int32_t pitches[3],
                     void* out, int32_t out_pitch, orc_debug* dbg, int mode, int threads, const orc_dovi* dovi)
{
    if (threads <= 0) threads = omp_get_max_threads();
    switch (mode) {
    case ORC_MODE_SPEC:    return process_spec(cfg, planes, pitches, out, out_pitch, dbg, threads, dovi);
    case ORC_MODE_LITERAL: return process_lit(cfg, planes, pitches, out, out_pitch, dbg, threads, dovi);
    case ORC_MODE_F64:     return process_f64(cfg, planes, pitches, out, out_pitch, dbg, threads, dovi);
    }
    return -100;
}

int orc_describe_convert(const orc_config* cfg, const orc_dovi* dovi, char* out, int cap) { return describe_spec(cfg, dovi, out, cap); }

int orc_process(const orc_config* cfg, const void* const planes[3], const int32_t pitches[3],
                void* out, int32_t out_pitch, orc_debug* dbg, int mode, int threads)
{
    return orc_process_dovi(cfg, planes, pitches, out, out_pitch, dbg, mode, threads, NULL);
}

int orc_process_from_convert(const orc_config* cfg, const void* convert_surf, int32_t pitch,
                             void* out, int32_t out_pitch, orc_debug* dbg, int mode, int threads)
{
    if (threads <= 0) threads = omp_get_max_threads();
    switch (mode) {
    case ORC_MODE_SPEC:    return process_from_convert_spec(cfg, convert_surf, pitch, out, out_pitch, dbg, threads);
    case ORC_MODE_LITERAL: return process_from_convert_lit(cfg, convert_surf, pitch, out, out_pitch, dbg, threads);
    case ORC_MODE_F64:     return process_from_convert_f64(cfg, convert_surf, pitch, out, out_pitch, dbg, threads);
    }
    return -100;
}

int orc_axis_taps(int kind, int filter, int variant, int src_origin, int src_len, int dst_len,
                  int i, int* base, float* weights, float* wsum)
{
    return axis_taps_export_spec(kind, filter, variant, src_origin, src_len, dst_len, i, base, weights, wsum);
}

/* vector forms of the scalar contract for exhaustive sweeps: op 0 lg2, 1 ex2, 2 pow(x, y), 3 sin, 4 ex2p(x, y), 5 x / y (IEEE) */
int orc_test_math_vec(int op, const float* x, const float* y, float* out, int64_t n)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        switch (op) {
        case 0: out[i] = orc_lg2_spec(x[i]); break;
        case 1: out[i] = orc_ex2_spec(x[i]); break;
        case 2: out[i] = orc_pow_spec(x[i], y[i]); break;
        case 3: out[i] = orc_sin_spec(x[i]); break;
        case 4: out[i] = orc_ex2p_spec(x[i], y[i]); break;
        default: out[i] = x[i] / y[i]; break;
        }
    }
    return 0;
}
float orc_test_lg2(float x) { return orc_lg2_spec(x); }
float orc_test_sin(float x) { return orc_sin_spec(x); }
float orc_test_ex2(float x) { return orc_ex2_spec(x); }
float orc_test_pow(float x, float y, int mode)
{
    switch (mode) {
    case ORC_MODE_SPEC:    return orc_pow_spec(x, y);
    case ORC_MODE_LITERAL: return orc_pow_literal(x, y);
    default:               return (float)orc_pow_f64(x, y);
    }
}
