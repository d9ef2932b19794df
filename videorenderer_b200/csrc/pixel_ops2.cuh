// pixel_ops2.cuh -- two-lane (FFMA2 / FADD2 / FMUL2) forms of the convert shader body: the two lanes are two horizontally
// adjacent pixels, every lane performs exactly the operations of convert_color<TR> in pixel_ops.cuh (same order, same
// roundings), so the results are bit-identical to the scalar form.  Used by stage A of the fused 2x kernel (PQ) and by
// k_convert2, the two-pixels-per-thread ConvertColorPass of the multipass family.
//
// ptxas 12.9 note ("Toolchain findings"): a packed multiply immediately followed by a packed add is contracted
// into FFMA2 even under -fmad=false.  Every product below therefore feeds a multiplicand, the addend of an explicit fma2, or
// a non-arithmetic consumer -- never an add2.
#pragma once
#include "pixel_ops.cuh"
#include "mpcvr_math2.cuh"

namespace b200vr {

struct f3x2 { float2 x, y, z; };

__device__ __forceinline__ float2 dp3_2(const float m[3], const f3x2& v) { return fma2(m[2], v.z, fma2(m[1], v.y, mul2(v.x, m[0]))); }
__device__ __forceinline__ f3x2 mat3_apply2(const float m[3][3], const f3x2& v) { f3x2 r; r.x = dp3_2(m[0], v); r.y = dp3_2(m[1], v); r.z = dp3_2(m[2], v); return r; }

__device__ __forceinline__ float2 pq_channel_to_hable2(const ColorParams& P, float2 c)
{
    const float2 a   = powr01_ylt1_2(sat2(c), P.inv_m2);                         // ST2084ToLinear, st2084.hlsl:9-16
    const float2 num = max2(add2(a, -ST2084_C1), 0.0f);
    const float2 den = fma2(-ST2084_C3, a, ST2084_C2);
    const float2 lin = mul2(powr01_ygt1_2(div_rn_safe_2(num, den), P.inv_m1), P.lum_scale);
    const float2 hn  = fma2(lin, fma2(0.15f, lin, P.hable_cb), P.hable_de);       // hable(), hdr_tone_mapping.hlsl:1-6
    const float2 hd  = fma2(lin, fma2(0.15f, lin, 0.50f), P.hable_df);
    return mul2(add2(div_rn_safe_2(hn, hd), -P.hable_ef), P.hable_rdiv);
}

__device__ __forceinline__ float2 sel2(bool cx, bool cy, float2 a, float2 b) { return make_float2(cx ? a.x : b.x, cy ? a.y : b.y); }

__device__ __forceinline__ float2 inverse_hlg2(const ColorParams& P, float2 x)                            // hlg.hlsl:1-11
{
    const float b = 0.28466892f, c = 0.55991073f;
    const float2 lo = mul2(mul2(x, x), 4.0f);
    const float2 hi = add2(ex2p_2(mul2(add2(x, -c), P.hlg_ra), 1.44269504f), b);
    return sel2(x.x <= 0.5f, x.y <= 0.5f, lo, hi);
}

__device__ __forceinline__ float2 linear_to_st2084_2(float2 v, float rdivider)                           // st2084.hlsl:18-25
{
    v = mul2(v, rdivider);
    const float2 p = powr_2(v, ST2084_M1);
    const float2 q = div_rn_safe_2(fma2(ST2084_C2, p, ST2084_C1), fma2(ST2084_C3, p, 1.0f));
    return powr_2(q, ST2084_M2);
}

// convert_color<-1> for two pixels (no Dolby Vision: those frames take the scalar kernel)
__device__ __forceinline__ f3x2 convert_color2(const ColorParams& P, const f3x2& yuv)
{
    const int tr = P.transfer;
    f3x2 c;
    c.x = add2(dp3_2(P.cm[0], yuv), P.cc[0]); c.y = add2(dp3_2(P.cm[1], yuv), P.cc[1]); c.z = add2(dp3_2(P.cm[2], yuv), P.cc[2]);
    if (tr == TRANSFER_NONE) return c;
    if (tr == TRANSFER_SDR_2020) {
        f3x2 l; l.x = powr_2(sat2(c.x), P.sdr_gamma); l.y = powr_2(sat2(c.y), P.sdr_gamma); l.z = powr_2(sat2(c.z), P.sdr_gamma);
        c = mat3_apply2(P.gamut, l);
    } else {
        if (tr == TRANSFER_HLG_TO_SDR || tr == TRANSFER_HLG_TO_PQ) {
            f3x2 l; l.x = inverse_hlg2(P, sat2(c.x)); l.y = inverse_hlg2(P, sat2(c.y)); l.z = inverse_hlg2(P, sat2(c.z));
            const float2 dt = fma2(0.0593f, l.z, fma2(0.6780f, l.y, mul2(l.x, 0.2627f)));
            const float2 g  = powr_2(mul2(dt, 2000.0f), 0.2f);
            c.x = linear_to_st2084_2(mul2(l.x, g), P.r1000); c.y = linear_to_st2084_2(mul2(l.y, g), P.r1000); c.z = linear_to_st2084_2(mul2(l.z, g), P.r1000);
            if (tr == TRANSFER_HLG_TO_PQ) return c;
        }
        f3x2 h; h.x = pq_channel_to_hable2(P, c.x); h.y = pq_channel_to_hable2(P, c.y); h.z = pq_channel_to_hable2(P, c.z);
        c = mat3_apply2(P.gamut, h);
    }
    c.x = powr_2(sat2(c.x), P.inv_gamma22); c.y = powr_2(sat2(c.y), P.inv_gamma22); c.z = powr_2(sat2(c.z), P.inv_gamma22);
    return c;
}

constexpr int kGammaN = 1023;      // codes of the UNORM10 threshold table (kernel_fused.cu: EnsureGammaTable)

// ---- two-lane (FFMA2) PQ->SDR convert: f3x2, dp3_2 and pq_channel_to_hable2 live in pixel_ops2.cuh (shared with k_convert2)
// everything before the final pow: saturate(mul(matrix_conv_prim, ToneMappingHable(ST2084ToLinear(saturate(rgb)))))
__device__ __forceinline__ f3x2 convert_color_pq2_pre(const ColorParams& P, const f3x2& yuv)
{
    f3x2 c;
    c.x = add2(dp3_2(P.cm[0], yuv), P.cc[0]); c.y = add2(dp3_2(P.cm[1], yuv), P.cc[1]); c.z = add2(dp3_2(P.cm[2], yuv), P.cc[2]);
    f3x2 h;
    h.x = pq_channel_to_hable2(P, c.x); h.y = pq_channel_to_hable2(P, c.y); h.z = pq_channel_to_hable2(P, c.z);
    f3x2 g;
    g.x = sat2(dp3_2(P.gamut[0], h)); g.y = sat2(dp3_2(P.gamut[1], h)); g.z = sat2(dp3_2(P.gamut[2], h));
    return g;
}

__device__ __forceinline__ f3x2 convert_color_pq2(const ColorParams& P, const f3x2& yuv)
{
    f3x2 c;
    c.x = add2(dp3_2(P.cm[0], yuv), P.cc[0]); c.y = add2(dp3_2(P.cm[1], yuv), P.cc[1]); c.z = add2(dp3_2(P.cm[2], yuv), P.cc[2]);
    f3x2 h;
    h.x = pq_channel_to_hable2(P, c.x); h.y = pq_channel_to_hable2(P, c.y); h.z = pq_channel_to_hable2(P, c.z);
    f3x2 o;
    o.x = powr01_ylt1_2(sat2(dp3_2(P.gamut[0], h)), P.inv_gamma22);
    o.y = powr01_ylt1_2(sat2(dp3_2(P.gamut[1], h)), P.inv_gamma22);
    o.z = powr01_ylt1_2(sat2(dp3_2(P.gamut[2], h)), P.inv_gamma22);
    return o;
}

// ---- the same PQ -> SDR chain with the bookkeeping trimmed (persistent strip kernel); bit-identical results:
//  * log2 through lg2s_2 with exponents pre-scaled by 2^-23 (ys = y * 2^-23, exact);
//  * the first pow drops pow(0,y) = 0's guard: for x < FLT_MIN the unguarded 2^(y lg2 x) is <= 2^-1.59 < ST2084_C1, so
//    max(a - c1, 0) -- the only consumer besides a positive denominator -- is the same +0, and 0 / den = +0 either way;
//  * saturates ride on the producing add / fma (add.sat, fma.sat: NaN -> 0 like sat()).
__device__ __forceinline__ float add_sat(float a, float b) { float d; asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fma_sat1(float a, float b, float c) { float d; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

__device__ __forceinline__ float2 pq_channel_to_hable2s(const ColorParams& P, float ym2s, float ym1s, float2 csat)
{
    float2 z;
    const float2 a   = ex2p_core2(lg2s_2(csat), ym2s, z);
    const float2 num = max2(add2(a, -ST2084_C1), 0.0f);
    const float2 den = fma2(-ST2084_C3, a, ST2084_C2);
    const float2 q   = div_rn_safe_2(num, den);
    const float2 v   = ex2p_core2(lg2s_2(q), ym1s, z);
    const float lo = 12582912.0f - 125.0f;
    const float2 lin = mul2(make_float2(z.x < lo ? 0.0f : v.x, z.y < lo ? 0.0f : v.y), P.lum_scale);
    const float2 hn  = fma2(lin, fma2(0.15f, lin, P.hable_cb), P.hable_de);
    const float2 hd  = fma2(lin, fma2(0.15f, lin, 0.50f), P.hable_df);
    return mul2(add2(div_rn_safe_2(hn, hd), -P.hable_ef), P.hable_rdiv);
}

// saturate(mul(matrix_conv_prim, ToneMappingHable(ST2084ToLinear(saturate(rgb))))) == convert_color_pq2_pre
__device__ __forceinline__ f3x2 convert_color_pq2_pre_s(const ColorParams& P, float ym2s, float ym1s, const f3x2& yuv)
{
    f3x2 h;
    #pragma unroll
    for (int i = 0; i < 3; i++) {
        const float2 t = dp3_2(P.cm[i], yuv);
        const float2 c = make_float2(add_sat(t.x, P.cc[i]), add_sat(t.y, P.cc[i]));
        const float2 r = pq_channel_to_hable2s(P, ym2s, ym1s, c);
        if (i == 0) h.x = r; else if (i == 1) h.y = r; else h.z = r;
    }
    f3x2 g;
    #pragma unroll
    for (int i = 0; i < 3; i++) {
        const float2 t = fma2(P.gamut[i][1], h.y, mul2(h.x, P.gamut[i][0]));
        const float2 r = make_float2(fma_sat1(P.gamut[i][2], h.z.x, t.x), fma_sat1(P.gamut[i][2], h.z.y, t.y));
        if (i == 0) g.x = r; else if (i == 1) g.y = r; else g.z = r;
    }
    return g;
}

// ---- measurement only (tools/strip_check.py --relaxed): the same stage with the float contract loosened -----------------------
// RELAX 1: both divisions as x * rcp.approx(y) (about 1.5 ULP), polynomials kept.  RELAX 2: also both pows on the MUFU unit
// (lg2.approx / ex2.approx, tens of ULP after the 78.8 exponent).  Never dispatched by the product; see DESIGN.md "relaxed float stage".
template <int RELAX>
__device__ __forceinline__ float2 pq_channel_to_hable2_relaxed(const ColorParams& P, float ym2s, float ym1s, float2 csat)
{
    auto fdiv = [](float2 x, float2 y) { float2 r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(y.x)); asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(y.y)); return mul2(x, r); };
    auto fpow = [](float2 x, float y) { float2 l, e; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l.x) : "f"(x.x)); asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l.y) : "f"(x.y));
                                        l = mul2(l, y); asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(l.x)); asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(l.y)); return e; };
    float2 z, v;
    const float2 a   = RELAX >= 2 ? fpow(csat, P.inv_m2) : ex2p_core2(lg2s_2(csat), ym2s, z);
    const float2 num = max2(add2(a, -ST2084_C1), 0.0f);
    const float2 den = fma2(-ST2084_C3, a, ST2084_C2);
    const float2 q   = fdiv(num, den);
    if (RELAX >= 2) v = fpow(q, P.inv_m1);
    else { const float2 t = ex2p_core2(lg2s_2(q), ym1s, z); const float lo = 12582912.0f - 125.0f; v = make_float2(z.x < lo ? 0.0f : t.x, z.y < lo ? 0.0f : t.y); }
    const float2 lin = mul2(v, P.lum_scale);
    const float2 hn  = fma2(lin, fma2(0.15f, lin, P.hable_cb), P.hable_de);
    const float2 hd  = fma2(lin, fma2(0.15f, lin, 0.50f), P.hable_df);
    return mul2(add2(fdiv(hn, hd), -P.hable_ef), P.hable_rdiv);
}
template <int RELAX>
__device__ __forceinline__ f3x2 convert_color_pq2_pre_relaxed(const ColorParams& P, float ym2s, float ym1s, const f3x2& yuv)
{
    f3x2 h;
    #pragma unroll
    for (int i = 0; i < 3; i++) {
        const float2 t = dp3_2(P.cm[i], yuv);
        const float2 c = make_float2(add_sat(t.x, P.cc[i]), add_sat(t.y, P.cc[i]));
        const float2 r = pq_channel_to_hable2_relaxed<RELAX>(P, ym2s, ym1s, c);
        if (i == 0) h.x = r; else if (i == 1) h.y = r; else h.z = r;
    }
    f3x2 g;
    #pragma unroll
    for (int i = 0; i < 3; i++) {
        const float2 t = fma2(P.gamut[i][1], h.y, mul2(h.x, P.gamut[i][0]));
        const float2 r = make_float2(fma_sat1(P.gamut[i][2], h.z.x, t.x), fma_sat1(P.gamut[i][2], h.z.y, t.y));
        if (i == 0) g.x = r; else if (i == 1) g.y = r; else g.z = r;
    }
    return g;
}

// gamma_q10_thr with a single exception float
__device__ __forceinline__ float gamma_q10_thr1(float x, const float* __restrict__ sT, const float* __restrict__ sD, uint32_t exc0)
{
    float l, e;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(l * (1.0f / 2.2f)));
    const int c0 = (int)fmaf(e, (float)kGammaN, 0.25f);                    // true code is c0 or c0 + 1
    int c = c0 + (x >= sT[c0 + 1] ? 1 : 0);
    c -= (__float_as_uint(x) == exc0) ? 1 : 0;
    return sD[c];
}

// gamma + UNORM10 write + read-back with ONE table access (strip kernel, OPT_GAMMA_LUT).  Entry c0 of the table is
// {T[c0 + 1], D[c0], D[c0 + 1], -}: c0 = floor(1023 * x^(1/2.2)) from the MUFU estimate (error << 1/2 code, so the true code
// round(1023 * pow) is c0 or c0 + 1), taken as the low mantissa bits of fma.rm(e, 1023, 2^23) -- no F2I, no second dependent
// load.  The exception float (code one below its neighbours') has c0 = its own code, so "x >= T and x is not the exception"
// selects for it as well.  k_gamma_verify checks every float of [0,1] against the direct evaluation.
struct GammaEntry { float t, d0, d1, pad; };
__device__ __forceinline__ uint32_t gamma_lut_index(float e)            // table index c0 + 0x4B000000 (the bits of 2^23 + c0)
{
    float z; asm("fma.rm.f32 %0, %1, 0f447FC000, 0f4B000000;" : "=f"(z) : "f"(e));      // e * 1023 + 2^23, rounded down
    return __float_as_uint(z);
}
__device__ __forceinline__ float gamma_q10_lut_g(float x, const GammaEntry* __restrict__ G, uint32_t exc0)    // generic-pointer form (self-check kernel)
{
    float l, e;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(l, 1.0f / 2.2f)));
    const GammaEntry g = G[gamma_lut_index(e) - 0x4B000000u];
    return (x >= g.t && __float_as_uint(x) != exc0) ? g.d1 : g.d0;
}
// two values; K = shared-memory address of the table - (0x4B000000 << 4) (mod 2^32)
__device__ __forceinline__ float2 gamma_q10_lut2(float2 x, uint32_t K, uint32_t exc0)
{
    float l0, l1, e0, e1;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(x.x));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(x.y));
    const float2 l = mul2(make_float2(l0, l1), 1.0f / 2.2f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(l.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(l.y));
    const float2 z = __ffma2_rd(make_float2(e0, e1), bc2(1023.0f), bc2(8388608.0f));
    const uint32_t a0 = (__float_as_uint(z.x) << 4) + K, a1 = (__float_as_uint(z.y) << 4) + K;
    float t0, p0, q0, t1, p1, q1, pad;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t0), "=f"(p0), "=f"(q0), "=f"(pad) : "r"(a0));
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t1), "=f"(p1), "=f"(q1), "=f"(pad) : "r"(a1));
    return make_float2((x.x >= t0 && __float_as_uint(x.x) != exc0) ? q0 : p0, (x.y >= t1 && __float_as_uint(x.y) != exc0) ? q1 : p1);
}

__device__ __forceinline__ float2 rt_roundtrip2(int fmt, float2 v)
{
    if (fmt == MPCVR_SURF_RGB10A2) return unorm_roundtrip2<1023>(v);
    if (fmt == MPCVR_SURF_BGRA8)   return unorm_roundtrip2<255>(v);
    return make_float2(f16_roundtrip(v.x), f16_roundtrip(v.y));
}

// gamma + UNORM10 write + read-back through the threshold table (see g_gammaT): x is already saturated
__device__ __forceinline__ float gamma_q10_thr(float x, const float* __restrict__ sT, const float* __restrict__ sD, uint32_t exc0, uint32_t exc1)
{
    float l, e;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(l * (1.0f / 2.2f)));
    const int c0 = (int)fmaf(e, (float)kGammaN, 0.25f);                    // true code is c0 or c0 + 1
    const uint32_t xb = __float_as_uint(x);
    int c = c0 + (x >= sT[c0 + 1] ? 1 : 0);
    c -= (xb == exc0 || xb == exc1) ? 1 : 0;
    return sD[c];
}

}  // namespace b200vr
