// tonemap_ops.cuh -- the per-pixel functions of the HDR10 local tone-mapping pass (k_tonemap, kernels_multipass.cu):
// Shaders/d3d11/ps_hdr10_tonemap.hlsl:34-336, drawn by TextureCopyRect over rect (Source/DX11VideoProcessor.cpp:3359-3367).
// Same contract as the other passes: written-order fp32, a*b+c fused, runtime quotients IEEE, pow = 2^(y log2 x).
// A header of its own so that tests/test_device_code_on_host.py can compile exactly this code for the host and hold it to the oracle.
#pragma once
#include "pixel_ops.cuh"

namespace b200vr {
namespace tm {
__device__ __forceinline__ float pq2l(const ColorParams& P, float v)          // ST2084ToLinear(v, 10000), st2084.hlsl:9-16
{
    const float a = powr(v, P.inv_m2);
    const float b = __fdiv_rn(fmaxf(__fsub_rn(a, ST2084_C1), 0.0f), fma_rn(-ST2084_C3, a, ST2084_C2));
    return __fmul_rn(powr(b, P.inv_m1), 10000.0f);
}
__device__ __forceinline__ float l2pq(float v)                                // LinearToST2084(v, 10000), st2084.hlsl:18-25
{
    const float p = powr(__fmul_rn(v, 1.0f / 10000.0f), ST2084_M1);
    return powr(__fdiv_rn(fma_rn(ST2084_C2, p, ST2084_C1), fma_rn(ST2084_C3, p, 1.0f)), ST2084_M2);
}
__device__ __forceinline__ float lerpm(float a, float b, float t) { return fma_rn(t, __fsub_rn(b, a), a); }
__device__ __forceinline__ float dot2020(f3 c) { return fma_rn(0.0593f, c.z, fma_rn(0.6780f, c.y, __fmul_rn(0.2627f, c.x))); }
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
__device__ __forceinline__ float smoothstep(float e0, float e1, float x)      // pl_smoothstep, :118-123
{
    const float t = clamp01(__fdiv_rn(__fsub_rn(x, e0), __fsub_rn(e1, e0)));
    return __fmul_rn(__fmul_rn(t, t), fma_rn(-2.0f, t, 3.0f));
}
__device__ __forceinline__ f3 scale3(f3 c, float k) { c.x = __fmul_rn(c.x, k); c.y = __fmul_rn(c.y, k); c.z = __fmul_rn(c.z, k); return c; }

__device__ __forceinline__ f3 bt2390(const ColorParams& P, const TonemapParams& T, f3 c)        // :66-116
{
    float safe = T.max_cll;
    if (safe <= 10.0f) safe = T.max_nits;
    if (safe <= 10.0f) safe = 1000.0f;
    if (T.display >= safe) return c;
    const float avg = dot2020(c);
    if (avg <= 0.000001f) return c;
    const float m_pq = l2pq(safe), t_pq = l2pq(T.display), E1 = l2pq(avg);
    const float KS = fmaxf(0.0f, fma_rn(1.5f, t_pq, -__fmul_rn(0.5f, m_pq)));
    float E2 = E1;
    if (E1 > KS) {
        const float mk = __fsub_rn(m_pq, KS);
        const float t = __fdiv_rn(__fsub_rn(E1, KS), fmaxf(0.000001f, mk));
        const float t2 = __fmul_rn(t, t), t3 = __fmul_rn(t2, t);
        const float a = __fadd_rn(fma_rn(2.0f, t3, -__fmul_rn(3.0f, t2)), 1.0f);
        const float b = __fadd_rn(fma_rn(-2.0f, t2, t3), t);
        const float d = fma_rn(-2.0f, t3, __fmul_rn(3.0f, t2));
        E2 = fma_rn(d, t_pq, fma_rn(b, mk, __fmul_rn(a, KS)));
    }
    return scale3(c, __fdiv_rn(pq2l(P, E2), avg));
}

__device__ __forceinline__ f3 st2094(const ColorParams& P, const TonemapParams& T, f3 c)        // :126-205
{
    if (T.display >= T.max_cll) return c;
    const float src_min = l2pq(T.min_nits), src_max = l2pq(T.max_cll), src_avg = l2pq(T.max_fall);
    const float dst_min = l2pq(0.0f), dst_max = l2pq(T.display);
    const float min_knee = 0.1f, max_knee = 0.8f, def_knee = 0.4f, knee_adaptation = 0.4f;
    const float src_knee_min = lerpm(src_min, src_max, min_knee), src_knee_max = lerpm(src_min, src_max, max_knee);
    const float dst_knee_min = lerpm(dst_min, dst_max, min_knee), dst_knee_max = lerpm(dst_min, dst_max, max_knee);
    float src_knee = (T.max_fall > 0.0f) ? src_avg : lerpm(src_min, src_max, def_knee);
    src_knee = fminf(fmaxf(src_knee, src_knee_min), src_knee_max);
    const float target = __fdiv_rn(__fsub_rn(src_knee, src_min), __fsub_rn(src_max, src_min));
    const float adapted = lerpm(dst_min, dst_max, target);
    const float tuning = __fsub_rn(1.0f, __fmul_rn(smoothstep(max_knee, def_knee, target), smoothstep(min_knee, def_knee, target)));
    const float adaptation = lerpm(knee_adaptation, 1.0f, tuning);
    float dst_knee = lerpm(src_knee, adapted, adaptation);
    dst_knee = fminf(fmaxf(dst_knee, dst_knee_min), dst_knee_max);
    const float x1 = T.min_nits, x3 = T.max_cll, x2 = pq2l(P, src_knee);
    const float y1 = 0.0f, y3 = T.display, y2 = pq2l(P, dst_knee);
    const float m00 = __fmul_rn(__fmul_rn(x2, x3), __fsub_rn(y2, y3)), m01 = __fmul_rn(__fmul_rn(x1, x3), __fsub_rn(y3, y1));
    const float m02 = __fmul_rn(__fmul_rn(x1, x2), __fsub_rn(y1, y2));
    const float m10 = fma_rn(x3, y3, -__fmul_rn(x2, y2)), m11 = fma_rn(x1, y1, -__fmul_rn(x3, y3)), m12 = fma_rn(x2, y2, -__fmul_rn(x1, y1));
    const float m20 = __fsub_rn(x3, x2), m21 = __fsub_rn(x1, x3), m22 = __fsub_rn(x2, x1);
    const float coef0 = fma_rn(m02, y3, fma_rn(m01, y2, __fmul_rn(m00, y1)));
    const float coef1 = fma_rn(m12, y3, fma_rn(m11, y2, __fmul_rn(m10, y1)));
    const float coef2 = fma_rn(m22, y3, fma_rn(m21, y2, __fmul_rn(m20, y1)));
    const float k = __fdiv_rn(1.0f, fma_rn(__fmul_rn(x1, y1), __fsub_rn(x2, x3),
                                    fma_rn(__fmul_rn(x2, y2), __fsub_rn(x3, x1), __fmul_rn(__fmul_rn(x3, y3), __fsub_rn(x1, x2)))));
    const float c1 = __fmul_rn(k, coef0), c2 = __fmul_rn(k, coef1), c3 = __fmul_rn(k, coef2);
    const float xn = dot2020(c);
    const float yn = __fdiv_rn(fma_rn(c2, xn, c1), fma_rn(c3, xn, 1.0f));
    return scale3(c, (xn > 0.0f) ? __fdiv_rn(yn, xn) : 1.0f);
}

__device__ __forceinline__ float curve(const ColorParams& P, const TonemapParams& T, float v)
{
    switch (T.sel) {
    case 2:  return __fdiv_rn(v, __fadd_rn(1.0f, v));                                                          // Reinhard :48-51
    case 3:  return __fsub_rn(__fdiv_rn(fma_rn(v, fma_rn(0.15f, v, P.hable_cb), P.hable_de), fma_rn(v, fma_rn(0.15f, v, 0.50f), P.hable_df)), P.hable_ef);   // Habel :53-57
    case 4:  return __fdiv_rn(v, __fadd_rn(1.0f, __fdiv_rn(v, __fadd_rn(T.display, 0.000001f))));              // Moebius :59-64
    default: return __fdiv_rn(__fmul_rn(v, fma_rn(2.51f, v, 0.03f)), fma_rn(v, fma_rn(2.43f, v, 0.59f), 0.14f));   // ACES :34-46
    }
}

// main(), :265-336: p = the PQ texel of the post-scale texture; returns the PQ value the pass writes
__device__ __forceinline__ f3 pixel(const ColorParams& P, const TonemapParams& T, f3 p)
{
    f3 c; c.x = pq2l(P, sat(p.x)); c.y = pq2l(P, sat(p.y)); c.z = pq2l(P, sat(p.z));     // main, :265-336
    // DolbyVisionTrims, :245-263.  Like its twin in the convert shader the text has no guard: a negative pow base (trimmed PQ value * slope +
    // offset < 0) or a zero trimmed luminance make EVERY channel NaN (through Y); a zero channel under a non-positive saturation gain
    // (0 * inf) or a value past ST2084ToLinear's pole (negative quotient) make THAT channel NaN.  nx / ny / nz: the channel is NaN in the
    // reference.  powr returns 0 for a non-positive base, so the cases are flagged explicitly -- the same tests as the oracle's tm_pixel.
    bool nx = false, ny = false, nz = false;
    if (T.l2_enabled) {
        const float b0 = fma_rn(l2pq(c.x), T.l2_slope, T.l2_offset), b1 = fma_rn(l2pq(c.y), T.l2_slope, T.l2_offset), b2 = fma_rn(l2pq(c.z), T.l2_slope, T.l2_offset);
        bool all = (b0 < 0.0f) || (b1 < 0.0f) || (b2 < 0.0f);
        f3 t; t.x = powr(b0, T.l2_power); t.y = powr(b1, T.l2_power); t.z = powr(b2, T.l2_power);
        const float Y = dot2020(t);
        const float g = __fadd_rn(1.0f, T.l2_chroma_weight);
        all = all || (Y == 0.0f);
        const bool sg_pos = T.l2_saturation_gain > 0.0f;
        nx = (t.x == 0.0f) && !sg_pos; ny = (t.y == 0.0f) && !sg_pos; nz = (t.z == 0.0f) && !sg_pos;
        const float s0 = __fmul_rn(t.x, powr(__fdiv_rn(__fmul_rn(g, t.x), Y), T.l2_saturation_gain));
        const float s1 = __fmul_rn(t.y, powr(__fdiv_rn(__fmul_rn(g, t.y), Y), T.l2_saturation_gain));
        const float s2 = __fmul_rn(t.z, powr(__fdiv_rn(__fmul_rn(g, t.z), Y), T.l2_saturation_gain));
        nx = nx || (fma_rn(-ST2084_C3, powr(s0, P.inv_m2), ST2084_C2) < 0.0f);
        ny = ny || (fma_rn(-ST2084_C3, powr(s1, P.inv_m2), ST2084_C2) < 0.0f);
        nz = nz || (fma_rn(-ST2084_C3, powr(s2, P.inv_m2), ST2084_C2) < 0.0f);
        c.x = pq2l(P, s0); c.y = pq2l(P, s1); c.z = pq2l(P, s2);
        if (all) { nx = true; ny = true; nz = true; }
    }
    f3 o;
    if (T.sel == 5 || T.sel == 6) {
        if (nx || ny || nz) {
            // BT.2390: with the curve active the NaN reaches the luminance and from there every channel; on a display that covers the content
            // the colour is returned as it is, NaN channels included.  ST 2094-10: the NaN luminance fails `x_nits > 0`, the factor is 1 and
            // the colour passes through unchanged, NaN channels included.  A NaN result is 0 in a UNORM target: 0 stands for it.
            bool whole = false;
            if (T.sel == 5) {
                float safe = T.max_cll;
                if (safe <= 10.0f) safe = T.max_nits;
                if (safe <= 10.0f) safe = 1000.0f;
                whole = !(T.display >= safe);
            }
            o.x = (whole || nx) ? 0.0f : l2pq(c.x); o.y = (whole || ny) ? 0.0f : l2pq(c.y); o.z = (whole || nz) ? 0.0f : l2pq(c.z);
            return o;
        }
        c = (T.sel == 5) ? bt2390(P, T, c) : st2094(P, T, c);
        o.x = l2pq(c.x); o.y = l2pq(c.y); o.z = l2pq(c.z);
    } else {
        if (nx) c.x = 0.0f;                                   // curves 1-4: color.rgb = saturate(color.rgb * ...) turns NaN into 0
        if (ny) c.y = 0.0f;
        if (nz) c.z = 0.0f;
        const float base = fmaxf(T.display, T.max_nits);
        const float eff = fminf(base, T.max_cll);
        const float fall_adj = fminf(__fdiv_rn(base, T.max_fall), 1.0f);
        const float reff = __fdiv_rn(1.0f, eff);
        o.x = l2pq(__fmul_rn(curve(P, T, __fmul_rn(sat(__fmul_rn(c.x, reff)), fall_adj)), T.display));
        o.y = l2pq(__fmul_rn(curve(P, T, __fmul_rn(sat(__fmul_rn(c.y, reff)), fall_adj)), T.display));
        o.z = l2pq(__fmul_rn(curve(P, T, __fmul_rn(sat(__fmul_rn(c.z, reff)), fall_adj)), T.display));
    }
    return o;
}
}  // namespace tm
}  // namespace b200vr
