/*
 * oracle/orc_math.h -- TEST INFRASTRUCTURE ONLY (see orc_types.h).
 *
 * The frozen scalar arithmetic contract ("SPEC mode") for the transcendental
 * pieces of the shader chain.
 *
 * The reference evaluates pow/exp/log inside HLSL pixel shaders
 * (Shaders/convert/st2084.hlsl:9-25, hlg.hlsl:1-20, Source/Shaders.cpp:896-921).
 * fxc lowers  pow(x,y) -> exp2(y*log2(x))  and  exp(x) -> exp2(x*log2(e));
 * the GPU's exp2/log2 are vendor approximations (D3D11 tolerance ~2^-21), so the
 * reference's float results are hardware dependent ("parity unpinned").
 * This file therefore DEFINES exp2/log2 as short, fully specified fp32
 * algorithms built only from IEEE-754 correctly rounded operations (add, mul,
 * fma, integer bit moves), so that the CUDA path can reproduce them bit for bit.
 * Measured against libm over all fp32 inputs of interest: LG2 <= 2.1 ulp,
 * EX2 <= 1 ulp (tests/test_oracle_math.py re-measures this).
 *
 * Independent restatement: the CUDA implementation lives in
 * videorenderer_b200/csrc/mpcvr_math.cuh and shares only the numbers below
 * (they are the specification, DESIGN.md section "Arithmetic contract").
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <stdint.h>
#include <string.h>
#include <math.h>

static inline uint32_t orc_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    orc_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* log2(1+f)/f on f in [sqrt(1/2)-1, sqrt(2)-1], degree 8, Horner with fma */
static const float ORC_LG2_C[9] = {
    0x1.715476p+0f, -0x1.71547p-1f, 0x1.ec73d4p-2f, -0x1.715c9cp-2f, 0x1.26d41p-2f,
    -0x1.e94f12p-3f, 0x1.b9b11ep-3f, -0x1.a8cc5cp-3f, 0x1.025a2p-3f };
/* 2^r on r in [-0.5, 0.5], degree 6 */
static const float ORC_EX2_C[7] = {
    0x1p+0f, 0x1.62e43p-1f, 0x1.ebfbep-3f, 0x1.c6aeccp-5f, 0x1.3b29d8p-7f,
    0x1.5f48c8p-10f, 0x1.446c7ep-13f };

/* x must be a positive normal float */
static inline float orc_lg2_spec(float x)
{
    uint32_t u = orc_f2u(x);
    int32_t  e = (int32_t)(u - 0x3f3504f3u) >> 23;      /* m in [sqrt(.5), sqrt(2)) */
    float    m = orc_u2f(u - ((uint32_t)e << 23));
    float    f = m - 1.0f;                              /* exact */
    float    p = ORC_LG2_C[8];
    for (int i = 7; i >= 0; i--) p = fmaf(p, f, ORC_LG2_C[i]);
    return fmaf(f, p, (float)e);
}

/* 2^(a*b) with the product folded into the range reduction (two fused multiply-adds):
 *   z = fma(a, b, 1.5*2^23)  -> n = round(a*b) sits in the low mantissa bits
 *   r = fma(a, b, -n)        -> a*b - n with a single rounding, |r| <= 1/2
 * results below 2^-125 flush to 0, above 2^127 saturate to +inf. */
static inline float orc_ex2p_spec(float a, float b)
{
    const float z  = fmaf(a, b, 12582912.0f);
    const float nf = z - 12582912.0f;                   /* exact */
    const float r  = fmaf(a, b, -nf);
    float p = ORC_EX2_C[6];
    for (int i = 5; i >= 0; i--) p = fmaf(p, r, ORC_EX2_C[i]);
    if (z < 12582912.0f - 125.0f) return 0.0f;
    if (z > 12582912.0f + 127.0f) return INFINITY;
    const int32_t n = (int32_t)orc_f2u(z) - 0x4b400000;
    return orc_u2f(orc_f2u(p) + ((uint32_t)n << 23));
}

static inline float orc_ex2_spec(float t) { return orc_ex2p_spec(t, 1.0f); }

/* SIN (contract for the per-pixel sines of the Jinc2 shader; arguments are in [0, 16)):
 * k = nearest integer to x/pi via the 1.5*2^23 trick, three-term Cody-Waite reduction r = x - k*pi (k*PI_A is exact),
 * odd Taylor polynomial of degree 11 on [-pi/2, pi/2] evaluated with fused steps, sign from the parity of k.
 * |error| < 1.5e-7 absolute on the range used. */
static inline float orc_sin_spec(float x)
{
    const float magic = 12582912.0f;
    const float z = fmaf(x, 0.318309886f, magic);
    const float kf = z - magic;
    float r = fmaf(-kf, 3.140625f, x);
    r = fmaf(-kf, 9.67502593994140625e-4f, r);
    r = fmaf(-kf, 1.509957990978376e-7f, r);
    const float r2 = r * r;
    float p = -2.50521084e-8f;                 /* -1/11! */
    p = fmaf(p, r2, 2.75573192e-6f);           /*  1/9!  */
    p = fmaf(p, r2, -1.98412698e-4f);          /* -1/7!  */
    p = fmaf(p, r2, 8.33333333e-3f);           /*  1/5!  */
    p = fmaf(p, r2, -1.66666667e-1f);          /* -1/3!  */
    const float s = fmaf(r * r2, p, r);
    uint32_t zb; memcpy(&zb, &z, 4);
    return (zb & 1u) ? -s : s;
}

/* HLSL pow(x,y) for x >= 0:  exp2(y*log2(x)); pow(0,y>0) = 0 (log2(0) = -inf). */
static inline float orc_pow_spec(float x, float y)
{
    if (!(x >= 1.17549435e-38f)) return 0.0f;
    return orc_ex2p_spec(orc_lg2_spec(x), y);
}

/* std::lerp(float, float, float) as libstdc++ evaluates it (exactness / monotonicity guards of C++20 [c.math.lerp]) */
static inline float orc_lerpf(float a, float b, float t)
{
    if ((a <= 0 && b >= 0) || (a >= 0 && b <= 0)) return t * b + (1 - t) * a;
    if (t == 1) return b;
    const float x = a + t * (b - a);
    return ((t > 1) == (b > a)) ? (b < x ? x : b) : (b > x ? x : b);
}

static inline float orc_pow_literal(float x, float y)
{
    if (!(x > 0.0f)) return 0.0f;
    return exp2f(y * log2f(x));
}

static inline double orc_pow_f64(double x, double y)
{
    if (!(x > 0.0)) return 0.0;
    return exp2(y * log2(x));
}

#endif
