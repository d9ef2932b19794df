// mpcvr_math2.cuh -- two-lane (fma.rn.f32x2 / FFMA2) forms of the scalar arithmetic in mpcvr_math.cuh.
// Same operations in the same order on each lane, so results are bit-identical to the scalar functions; the
// packed forms only halve the issue slots of the FMA-class instructions (sm_100a: FFMA2 / FADD2 / FMUL2).
#pragma once
#include "mpcvr_math.cuh"

namespace b200vr {

__device__ __forceinline__ float2 bc2(float c) { return make_float2(c, c); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fma2(float2 a, float b, float2 c) { return __ffma2_rn(a, bc2(b), c); }
__device__ __forceinline__ float2 fma2(float a, float2 b, float2 c) { return __ffma2_rn(bc2(a), b, c); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float c) { return __ffma2_rn(a, b, bc2(c)); }
__device__ __forceinline__ float2 fma2(float2 a, float b, float c) { return __ffma2_rn(a, bc2(b), bc2(c)); }
__device__ __forceinline__ float2 fma2(float a, float2 b, float c) { return __ffma2_rn(bc2(a), b, bc2(c)); }
// NOTE: ptxas (12.9) contracts mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 even under -fmad=false (it does
// not for the scalar .rn forms).  No product below is therefore ever followed by an add of that product: products feed
// either a multiplicand or the ADDEND of an explicit fma (not contractible), and exp2's range reduction takes its
// product inside fused multiply-adds by definition (ex2p).  tests: test_two_lane_math_equals_scalar_math.
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float b) { return __fmul2_rn(a, bc2(b)); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float b) { return __fadd2_rn(a, bc2(b)); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 sat2(float2 a) { return make_float2(sat(a.x), sat(a.y)); }
__device__ __forceinline__ float2 max2(float2 a, float b) { return make_float2(fmaxf(a.x, b), fmaxf(a.y, b)); }

__device__ __forceinline__ float2 lg2_2(float2 x)
{
    // exponent as (v & 0xff800000) = e << 23, converted exactly and scaled by 2^-23 (both steps exact).  NOTE: the
    // textbook form ((int)v >> 23 then int->float) is miscompiled by ptxas 12.9 when two lanes are live in a longer
    // chain: it emits I2FP on the UNSHIFTED value for lane 0 (tools/f32x2_test.cu reproduces it).
    const uint32_t u0 = __float_as_uint(x.x), u1 = __float_as_uint(x.y);
    const uint32_t k0 = (u0 - 0x3f3504f3u) & 0xff800000u, k1 = (u1 - 0x3f3504f3u) & 0xff800000u;
    const float2   m  = make_float2(__uint_as_float(u0 - k0), __uint_as_float(u1 - k1));
    const float2   e  = mul2(make_float2((float)(int32_t)k0, (float)(int32_t)k1), 1.1920928955078125e-07f);
    const float2   f  = add2(m, -1.0f);
    float2 p = bc2(MPCVR_LG2_C8);
    p = fma2(p, f, MPCVR_LG2_C7);
    p = fma2(p, f, MPCVR_LG2_C6);
    p = fma2(p, f, MPCVR_LG2_C5);
    p = fma2(p, f, MPCVR_LG2_C4);
    p = fma2(p, f, MPCVR_LG2_C3);
    p = fma2(p, f, MPCVR_LG2_C2);
    p = fma2(p, f, MPCVR_LG2_C1);
    p = fma2(p, f, MPCVR_LG2_C0);
    return fma2(f, p, e);
}

// 2^23 * log2(x) per lane: lg2_2() with the coefficients and the exponent term scaled by 2^23 (every Horner step and the final
// fma scale exactly, so the result is exactly 2^23 * lg2_2(x)); saves the 2^-23 multiply of the exponent.  Callers pair it
// with an exponent y * 2^-23: ex2p's products a*b are then the same real numbers and every rounding is the same.
__device__ __forceinline__ float2 lg2s_2(float2 x)
{
    constexpr float S = 8388608.0f;
    const uint32_t u0 = __float_as_uint(x.x), u1 = __float_as_uint(x.y);
    const uint32_t k0 = (u0 - 0x3f3504f3u) & 0xff800000u, k1 = (u1 - 0x3f3504f3u) & 0xff800000u;
    const float2   m  = make_float2(__uint_as_float(u0 - k0), __uint_as_float(u1 - k1));
    const float2   e  = make_float2((float)(int32_t)k0, (float)(int32_t)k1);
    const float2   f  = add2(m, -1.0f);
    float2 p = bc2(MPCVR_LG2_C8 * S);
    p = fma2(p, f, MPCVR_LG2_C7 * S);
    p = fma2(p, f, MPCVR_LG2_C6 * S);
    p = fma2(p, f, MPCVR_LG2_C5 * S);
    p = fma2(p, f, MPCVR_LG2_C4 * S);
    p = fma2(p, f, MPCVR_LG2_C3 * S);
    p = fma2(p, f, MPCVR_LG2_C2 * S);
    p = fma2(p, f, MPCVR_LG2_C1 * S);
    p = fma2(p, f, MPCVR_LG2_C0 * S);
    return fma2(f, p, e);
}

// 2^(a*b) per lane, without the range selects (caller guarantees -125.5 < a*b < 127.5)
__device__ __forceinline__ float2 ex2p_core2(float2 a, float b, float2& z)
{
    z = fma2(a, b, 12582912.0f);
    const float2 nf = add2(z, -12582912.0f);
    const float2 r  = fma2(a, b, make_float2(-nf.x, -nf.y));
    const int32_t n0 = (int32_t)__float_as_uint(z.x) - 0x4b400000, n1 = (int32_t)__float_as_uint(z.y) - 0x4b400000;
    float2 p = bc2(MPCVR_EX2_C6);
    p = fma2(p, r, MPCVR_EX2_C5);
    p = fma2(p, r, MPCVR_EX2_C4);
    p = fma2(p, r, MPCVR_EX2_C3);
    p = fma2(p, r, MPCVR_EX2_C2);
    p = fma2(p, r, MPCVR_EX2_C1);
    p = fma2(p, r, MPCVR_EX2_C0);
    return make_float2(__uint_as_float(__float_as_uint(p.x) + ((uint32_t)n0 << 23)), __uint_as_float(__float_as_uint(p.y) + ((uint32_t)n1 << 23)));
}

__device__ __forceinline__ float2 ex2p_2(float2 a, float b)       // == ex2p() per lane
{
    float2 z;
    const float2 v = ex2p_core2(a, b, z);
    const float lo = 12582912.0f - 125.0f, hi = 12582912.0f + 127.0f, inf = __int_as_float(0x7f800000);
    return make_float2(z.x < lo ? 0.0f : (z.x > hi ? inf : v.x), z.y < lo ? 0.0f : (z.y > hi ? inf : v.y));
}

__device__ __forceinline__ float2 ex2_2(float2 t) { return ex2p_2(t, 1.0f); }

__device__ __forceinline__ float2 powr_2(float2 x, float y)      // == powr() per lane
{
    const float2 r = ex2p_2(lg2_2(max2(x, 1.17549435e-38f)), y);
    return make_float2(x.x >= 1.17549435e-38f ? r.x : 0.0f, x.y >= 1.17549435e-38f ? r.y : 0.0f);
}

// powr() for x in [0,1] and 0 < y < 0.99: y*log2(x) lies in (-125, 0], so neither range select can trigger.
__device__ __forceinline__ float2 powr01_ylt1_2(float2 x, float y)
{
    float2 z;
    const float2 r = ex2p_core2(lg2_2(max2(x, 1.17549435e-38f)), y, z);
    return make_float2(x.x >= 1.17549435e-38f ? r.x : 0.0f, x.y >= 1.17549435e-38f ? r.y : 0.0f);
}

// powr() for x in [0,1] and y > 1.01: x below FLT_MIN (zero, denormal) has lg2() <= -126.9 (the bit-level reduction
// reads it as 2^-127 * [1,2)), so y*lg2 < -128 and the low range select returns the 0 that powr() returns for such x;
// y*log2(x) <= 0 makes the high select unnecessary.
__device__ __forceinline__ float2 powr01_ygt1_2(float2 x, float y)
{
    float2 z;
    const float2 v = ex2p_core2(lg2_2(x), y, z);
    const float lo = 12582912.0f - 125.0f;
    return make_float2(z.x < lo ? 0.0f : v.x, z.y < lo ? 0.0f : v.y);
}

__device__ __forceinline__ float2 div_rn_safe_2(float2 x, float2 y)   // == div_rn_safe() per lane
{
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(y.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(y.y));
    const float2 ny = make_float2(-y.x, -y.y);
    const float2 e = fma2(ny, r, 1.0f);
    r = fma2(r, e, r);
    const float2 q = fma2(x, r, 0.0f);
    const float2 m = fma2(ny, q, x);
    return fma2(r, m, q);
}

// floor(z) as an integer-valued float for 0 <= z < 2^22 (FADD.RM on both lanes)
__device__ __forceinline__ float2 floor_pos2(float2 z) { return add2(__fadd2_rd(z, bc2(8388608.0f)), -8388608.0f); }

template <int N>
__device__ __forceinline__ float2 div_by_const_exact2(float2 c) { return fma2(c, RcpSplit<N>::hi, mul2(c, RcpSplit<N>::lo)); }   // == div_by_const_exact<N> per lane

// float -> UNORM render-target write followed by the UNORM read (what the next pass samples)
template <int N>
__device__ __forceinline__ float2 unorm_roundtrip2(float2 v)
{
    return div_by_const_exact2<N>(floor_pos2(fma2(sat2(v), (float)N, 0.5f)));
}

}  // namespace b200vr
