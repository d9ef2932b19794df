// pixel_ops.cuh -- device building blocks of the frame path, shared by the multipass kernels
// (one kernel per reference pass) and the fused per-frame kernel.
//
// What each block replaces in the reference:
//   fetch_yuv()      the pixel-fetch half of the generated convert shader, ShaderGetPixels
//                    (Source/Shaders.cpp:82-329, DX11 branches) incl. chroma reconstruction
//   convert_color()  the body GetShaderConvertColor emits (Source/Shaders.cpp:819-923) with
//                    Shaders/convert/{st2084,hlg,hdr_tone_mapping}.hlsl
//   surf_load/store  the D3D11 render-target write / shader-resource read conversions between passes
// Arithmetic: fp32, fused multiply-add exactly where fxc emits mad/dp3 (DESIGN.md "Arithmetic contract").
#pragma once
#include <cuda_fp16.h>
#include "frame_plan.h"
#include "fmt_table.h"
#include "mpcvr_math.cuh"

namespace b200vr {

struct f3 { float x, y, z; };

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---------------------------------------------------------------------------------------------
// surfaces
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ f3 surf_load(const SurfDesc& s, int x, int y)
{
    const uint8_t* row = (const uint8_t*)s.ptr + (size_t)y * s.pitch;
    f3 r;
    if (s.fmt == MPCVR_SURF_RGB10A2) {
        const uint32_t v = __ldg((const uint32_t*)row + x);
        r.x = unorm10(v & 1023u); r.y = unorm10((v >> 10) & 1023u); r.z = unorm10((v >> 20) & 1023u);
    } else if (s.fmt == MPCVR_SURF_BGRA8) {
        const uint32_t v = __ldg((const uint32_t*)row + x);
        r.z = unorm8(v & 255u); r.y = unorm8((v >> 8) & 255u); r.x = unorm8((v >> 16) & 255u);
    } else if (s.fmt == MPCVR_SURF_RGBA16) {
        const uint2 v = __ldg((const uint2*)row + x);
        r.x = unorm16(v.x & 0xFFFFu); r.y = unorm16(v.x >> 16); r.z = unorm16(v.y & 0xFFFFu);
    } else if (s.fmt == MPCVR_SURF_RGB8) {
        r.x = unorm8(__ldg(row + 3 * x)); r.y = unorm8(__ldg(row + 3 * x + 1)); r.z = unorm8(__ldg(row + 3 * x + 2));
    } else {
        const uint2 v = __ldg((const uint2*)row + x);
        const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
        r.x = __low2float(a); r.y = __high2float(a); r.z = __low2float(b);
    }
    return r;
}

__device__ __forceinline__ void surf_store(const SurfDesc& s, int x, int y, f3 v)
{
    uint8_t* row = (uint8_t*)s.ptr + (size_t)y * s.pitch;
    if (s.fmt == MPCVR_SURF_RGB10A2) {
        ((uint32_t*)row)[x] = qunorm(v.x, 1023.0f) | (qunorm(v.y, 1023.0f) << 10) | (qunorm(v.z, 1023.0f) << 20) | (3u << 30);
    } else if (s.fmt == MPCVR_SURF_BGRA8) {
        ((uint32_t*)row)[x] = qunorm(v.z, 255.0f) | (qunorm(v.y, 255.0f) << 8) | (qunorm(v.x, 255.0f) << 16) | (255u << 24);
    } else if (s.fmt == MPCVR_SURF_RGBA16) {
        uint2 o; o.x = qunorm(v.x, 65535.0f) | (qunorm(v.y, 65535.0f) << 16); o.y = qunorm(v.z, 65535.0f) | (65535u << 16);
        ((uint2*)row)[x] = o;
    } else if (s.fmt == MPCVR_SURF_RGB8) {                  // extension: the B8G8R8A8 codes, packed R,G,B
        row[3 * x] = (uint8_t)qunorm(v.x, 255.0f); row[3 * x + 1] = (uint8_t)qunorm(v.y, 255.0f); row[3 * x + 2] = (uint8_t)qunorm(v.z, 255.0f);
    } else {
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, 1.0f);
        uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b);
        ((uint2*)row)[x] = o;
    }
}

// the value a later pass reads back after `v` went through a render target of format fmt
__device__ __forceinline__ float rt_roundtrip(int fmt, float v)
{
    if (fmt == MPCVR_SURF_RGB10A2) return unorm10(qunorm(v, 1023.0f));
    if (fmt == MPCVR_SURF_BGRA8)   return unorm8(qunorm(v, 255.0f));
    if (fmt == MPCVR_SURF_RGBA16)  return unorm16(qunorm(v, 65535.0f));
    return f16_roundtrip(v);
}

// ---------------------------------------------------------------------------------------------
// source fetch
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float src_comp(const SrcDesc& S, int plane, int x, int y, int comp, int ncomp)
{
    const uint8_t* p = S.pl[plane] + (ptrdiff_t)y * S.pitch[plane];      // pitch < 0: bottom-up DIB (RGB samples)
    if (S.bytes == 1) return unorm8(__ldg(p + x * ncomp + comp));
    const uint32_t v = __ldg((const uint16_t*)p + x * ncomp + comp);
    return unorm16((v << S.lsb_shift) & 0xFFFFu);
}

struct f2 { float u, v; };

// chroma texel with CLAMP addressing (sampler state Source/DX11VideoProcessor.cpp:1308-1319)
__device__ __forceinline__ f2 chroma_texel(const SrcDesc& S, int cx, int cy)
{
    cx = clampi(cx, 0, S.cw - 1); cy = clampi(cy, 0, S.ch - 1);
    f2 r;
    if (S.layout == LAYOUT_BIPLANAR) { r.u = src_comp(S, 1, cx, cy, 0, 2); r.v = src_comp(S, 1, cx, cy, 1, 2); }
    else                             { r.u = src_comp(S, 1, cx, cy, 0, 1); r.v = src_comp(S, 2, cx, cy, 0, 1); }
    return r;
}

__device__ __forceinline__ float lerp_mad(float a, float b, float f) { return fma_rn(f, __fsub_rn(b, a), a); }

// code_CatmullRom_weights, Source/Shaders.cpp:66-72; t is a multiple of 1/4 so every term is exact
__device__ __forceinline__ void catmull_weights(float t, float w[4])
{
    const float t2 = t * t, t3 = t * t2;
    w[0] = t2 - (t3 + t) * 0.5f;
    w[1] = t3 * 1.5f + 1.0f - t2 * 2.5f;
    w[2] = t2 * 2.0f + t * 0.5f - t3 * 1.5f;
    w[3] = (t3 - t2) * 0.5f;
}

// CATMULLROM_05, Source/Shaders.cpp:145
__device__ __forceinline__ float catmull05(float c0, float c1, float c2, float c3)
{
    return __fmul_rn(fma_rn(9.0f, __fadd_rn(c1, c2), -__fadd_rn(c0, c3)), 0.0625f);
}

__device__ __forceinline__ f3 fetch_planar(const SrcDesc& S, int sx, int sy)
{
    f3 r;
    r.x = src_comp(S, 0, sx, sy, 0, 1);
    if (S.deint_blend && S.subsampling == 420) {                       // Source/Shaders.cpp:232-237
        const float y1 = src_comp(S, 0, sx, max(sy - 1, 0), 0, 1);
        const float y2 = src_comp(S, 0, sx, min(sy + 1, S.H - 1), 0, 1);
        r.x = __fmul_rn(__fadd_rn(fma_rn(r.x, 2.0f, y1), y2), 0.25f);
    }
    f2 c;
    if (S.chroma_mode == MPCVR_CHROMA_Nearest || S.subsampling == 444) {                 // :239-241
        c = chroma_texel(S, S.subsampling == 444 ? sx : (sx >> 1), S.subsampling == 420 ? (sy >> 1) : sy);
    } else if (S.chroma_mode == MPCVR_CHROMA_CatmullRom && S.subsampling == 420) {       // :242-251
        float ox, oy;
        if (S.siting == ChromaSub_Cosited)    { ox = -0.25f; oy = -0.25f; }
        else if (S.siting == ChromaSub_MPEG1) { ox = -0.5f;  oy = -0.5f;  }
        else                                  { ox = -0.25f; oy = -0.5f;  }
        float wx[4], wy[4];
        catmull_weights(((sx & 1) ? 0.75f : 0.25f) + ox, wx);
        catmull_weights(((sy & 1) ? 0.75f : 0.25f) + oy, wy);
        const int cx = sx >> 1, cy = sy >> 1;
        float qu[4], qv[4];
        #pragma unroll
        for (int j = 0; j < 4; j++) {                                                    // code_Bicubic_UV :74-79
            const f2 c0 = chroma_texel(S, cx - 1, cy + j - 1), c1 = chroma_texel(S, cx, cy + j - 1);
            const f2 c2 = chroma_texel(S, cx + 1, cy + j - 1), c3 = chroma_texel(S, cx + 2, cy + j - 1);
            qu[j] = fma_rn(c3.u, wx[3], fma_rn(c2.u, wx[2], fma_rn(c1.u, wx[1], __fmul_rn(c0.u, wx[0]))));
            qv[j] = fma_rn(c3.v, wx[3], fma_rn(c2.v, wx[2], fma_rn(c1.v, wx[1], __fmul_rn(c0.v, wx[0]))));
        }
        c.u = fma_rn(qu[3], wy[3], fma_rn(qu[2], wy[2], fma_rn(qu[1], wy[1], __fmul_rn(qu[0], wy[0]))));
        c.v = fma_rn(qv[3], wy[3], fma_rn(qv[2], wy[2], fma_rn(qv[1], wy[1], __fmul_rn(qv[0], wy[0]))));
    } else if (S.chroma_mode == MPCVR_CHROMA_CatmullRom && S.subsampling == 422) {       // :252-264
        if ((sx & 1) == 0) c = chroma_texel(S, sx >> 1, sy);
        else {
            const int j = sx >> 1;
            const f2 c0 = chroma_texel(S, j - 1, sy), c1 = chroma_texel(S, j, sy), c2 = chroma_texel(S, j + 1, sy), c3 = chroma_texel(S, j + 2, sy);
            c.u = catmull05(c0.u, c1.u, c2.u, c3.u);
            c.v = catmull05(c0.v, c1.v, c2.v, c3.v);
        }
    } else {                                                                             // CHROMA_Bilinear :265-270
        // sample point in quarter chroma texels: strChromaPos (:118-142) through the MIN_POINT_MAG_LINEAR sampler
        int qx, qy;
        if (S.subsampling == 420) {
            if (S.siting == ChromaSub_Cosited)    { qx = 2 * sx;     qy = 2 * sy;     }
            else if (S.siting == ChromaSub_MPEG1) { qx = 2 * sx - 1; qy = 2 * sy - 1; }
            else                                  { qx = 2 * sx;     qy = 2 * sy - 1; }
        } else { qx = 2 * sx; qy = 4 * sy; }
        const int ix = qx >> 2, iy = qy >> 2;
        const float fx = (float)(qx & 3) * 0.25f, fy = (float)(qy & 3) * 0.25f;
        const f2 c00 = chroma_texel(S, ix, iy), c10 = chroma_texel(S, ix + 1, iy);
        const f2 c01 = chroma_texel(S, ix, iy + 1), c11 = chroma_texel(S, ix + 1, iy + 1);
        c.u = lerp_mad(lerp_mad(c00.u, c10.u, fx), lerp_mad(c01.u, c11.u, fx), fy);
        c.v = lerp_mad(lerp_mad(c00.v, c10.v, fx), lerp_mad(c01.v, c11.v, fx), fy);
    }
    r.y = c.u; r.z = c.v;
    return r;
}

// v210 sample read in place.  Four little-endian words carry six pixels (Cb0 Y0 Cr0 | Y1 Cb1 Y2 | Cr1 Y3 Cb2 | Y4 Cr2 Y5,
// three 10-bit fields per word); the reference rewrites them on the CPU as Y210 macro-pixels with the codes in the top ten
// bits (CopyFrameV210, Source/Helper.cpp:709-748), so a component reads back as unorm16(code << 6).
__device__ __forceinline__ float v210_field(const uint8_t* row, int word, int field)
{
    const uint32_t w = __ldg((const uint32_t*)row + word);
    return unorm16(((w >> (10 * field)) & 1023u) << 6);
}
__device__ __forceinline__ float v210_luma(const uint8_t* row, int x)
{
    const int g = x / 6, k = x - 6 * g;                    // Y0:w0.1 Y1:w1.0 Y2:w1.2 Y3:w2.1 Y4:w3.0 Y5:w3.2
    const int word = (k == 0) ? 0 : ((k <= 2) ? 1 : ((k == 3) ? 2 : 3));
    const int field = (k == 0 || k == 3) ? 1 : ((k == 1 || k == 4) ? 0 : 2);
    return v210_field(row, 4 * g + word, field);
}
__device__ __forceinline__ f2 v210_chroma(const uint8_t* row, int m)      // m = macro-pixel (chroma sample) index
{
    const int g = m / 3, j = m - 3 * g;                    // Cb0:w0.0 Cr0:w0.2 | Cb1:w1.1 Cr1:w2.0 | Cb2:w2.2 Cr2:w3.1
    f2 c;
    if (j == 0)      { c.u = v210_field(row, 4 * g,     0); c.v = v210_field(row, 4 * g,     2); }
    else if (j == 1) { c.u = v210_field(row, 4 * g + 1, 1); c.v = v210_field(row, 4 * g + 2, 0); }
    else             { c.u = v210_field(row, 4 * g + 2, 2); c.v = v210_field(row, 4 * g + 3, 1); }
    return c;
}
__device__ __forceinline__ f3 fetch_v210(const SrcDesc& S, int sx, int sy)   // shader side: Source/Shaders.cpp:209-228 (packed422)
{
    const uint8_t* row = S.pl[0] + (ptrdiff_t)sy * S.pitch[0];
    const int mw = S.W >> 1, mx = sx >> 1;
    f3 r;
    r.x = v210_luma(row, sx);
    if ((sx & 1) == 0) { const f2 c = v210_chroma(row, mx); r.y = c.u; r.z = c.v; }
    else {
        const int m0 = max(mx - 1, 0), m2 = min(mx + 1, mw - 1), m3 = min(mx + 2, mw - 1);
        const f2 c1 = v210_chroma(row, mx), c2 = v210_chroma(row, m2);
        if (S.chroma_mode == MPCVR_CHROMA_CatmullRom) {
            const f2 c0 = v210_chroma(row, m0), c3 = v210_chroma(row, m3);
            r.y = catmull05(c0.u, c1.u, c2.u, c3.u); r.z = catmull05(c0.v, c1.v, c2.v, c3.v);
        } else {
            r.y = __fmul_rn(__fadd_rn(c1.u, c2.u), 0.5f); r.z = __fmul_rn(__fadd_rn(c1.v, c2.v), 0.5f);
        }
    }
    return r;
}

// packed 4:2:2 (Source/Shaders.cpp:195-228): one RGBA texel holds two pixels
__device__ __forceinline__ f3 fetch_packed422(const SrcDesc& S, int sx, int sy)
{
    if (S.packed_kind == 2) return fetch_v210(S, sx, sy);
    const int mw = S.W >> 1, mx = sx >> 1;
    const bool uyvy = S.packed_kind == 1;
    const int iy0 = uyvy ? 1 : 0, iu = uyvy ? 0 : 1, iy1 = uyvy ? 3 : 2, iv = uyvy ? 2 : 3;
    f3 r;
    if ((sx & 1) == 0) {
        r.x = src_comp(S, 0, mx, sy, iy0, 4); r.y = src_comp(S, 0, mx, sy, iu, 4); r.z = src_comp(S, 0, mx, sy, iv, 4);
    } else {
        r.x = src_comp(S, 0, mx, sy, iy1, 4);
        const int m0 = max(mx - 1, 0), m2 = min(mx + 1, mw - 1), m3 = min(mx + 2, mw - 1);
        if (S.chroma_mode == MPCVR_CHROMA_CatmullRom) {
            r.y = catmull05(src_comp(S, 0, m0, sy, iu, 4), src_comp(S, 0, mx, sy, iu, 4), src_comp(S, 0, m2, sy, iu, 4), src_comp(S, 0, m3, sy, iu, 4));
            r.z = catmull05(src_comp(S, 0, m0, sy, iv, 4), src_comp(S, 0, mx, sy, iv, 4), src_comp(S, 0, m2, sy, iv, 4), src_comp(S, 0, m3, sy, iv, 4));
        } else {
            r.y = __fmul_rn(__fadd_rn(src_comp(S, 0, mx, sy, iu, 4), src_comp(S, 0, m2, sy, iu, 4)), 0.5f);
            r.z = __fmul_rn(__fadd_rn(src_comp(S, 0, mx, sy, iv, 4), src_comp(S, 0, m2, sy, iv, 4)), 0.5f);
        }
    }
    return r;
}

// packed 4:4:4 (Source/Shaders.cpp:183-194)
__device__ __forceinline__ f3 fetch_packed444(const SrcDesc& S, int sx, int sy)
{
    const uint8_t* row = S.pl[0] + (ptrdiff_t)sy * S.pitch[0];
    f3 r;
    if (S.packed_kind == 0) {          // AYUV: V U Y A
        const uint32_t v = __ldg((const uint32_t*)row + sx);
        r.z = unorm8(v & 255u); r.y = unorm8((v >> 8) & 255u); r.x = unorm8((v >> 16) & 255u);
    } else if (S.packed_kind == 1) {   // Y410: U:10 Y:10 V:10 A:2
        const uint32_t v = __ldg((const uint32_t*)row + sx);
        r.y = unorm10(v & 1023u); r.x = unorm10((v >> 10) & 1023u); r.z = unorm10((v >> 20) & 1023u);
    } else {                           // Y416: U Y V A, 16 bit
        const uint2 v = __ldg((const uint2*)row + sx);
        r.y = unorm16(v.x & 0xFFFFu); r.x = unorm16(v.x >> 16); r.z = unorm16(v.y & 0xFFFFu);
    }
    return r;
}

// RGB samples: "float4 color = tex.Sample(samp, input.Tex)" (Source/Shaders.cpp:183-185) of the texture the reference's
// CPU repack would have produced (fmt_table.h RgbKind); alpha is not carried (opaque swap chain).
__device__ __forceinline__ f3 fetch_rgb(const SrcDesc& S, int sx, int sy)
{
    const uint8_t* row = S.pl[0] + (ptrdiff_t)sy * S.pitch[0];
    f3 r;
    switch (S.packed_kind) {
    case RGBK_BGRX32: {
        const uint32_t v = __ldg((const uint32_t*)row + sx);
        r.z = unorm8(v & 255u); r.y = unorm8((v >> 8) & 255u); r.x = unorm8((v >> 16) & 255u); break; }
    case RGBK_BGR24: {
        const uint8_t* p = row + 3 * sx;
        r.z = unorm8(__ldg(p)); r.y = unorm8(__ldg(p + 1)); r.x = unorm8(__ldg(p + 2)); break; }
    case RGBK_R210: {
        const uint32_t be = __byte_perm(__ldg((const uint32_t*)row + sx), 0, 0x0123);
        r.x = unorm10((be >> 20) & 1023u); r.y = unorm10((be >> 10) & 1023u); r.z = unorm10(be & 1023u); break; }
    case RGBK_RGB48: case RGBK_BGR48: {
        const uint16_t* p = (const uint16_t*)row + 3 * sx;
        const float a = unorm16(__ldg(p)), b = unorm16(__ldg(p + 1)), c = unorm16(__ldg(p + 2));
        if (S.packed_kind == RGBK_RGB48) { r.x = a; r.y = b; r.z = c; } else { r.x = c; r.y = b; r.z = a; }
        break; }
    case RGBK_BGRA64: {
        const uint2 v = __ldg((const uint2*)row + sx);
        r.z = unorm16(v.x & 0xFFFFu); r.y = unorm16(v.x >> 16); r.x = unorm16(v.y & 0xFFFFu); break; }
    default: {  // RGBK_B64A: bytes Ahi Alo Rhi Rlo Ghi Glo Bhi Blo
        const uint2 v = __ldg((const uint2*)row + sx);
        r.x = unorm16(__byte_perm(v.x, 0, 0x4423)); r.y = unorm16(__byte_perm(v.y, 0, 0x4401)); r.z = unorm16(__byte_perm(v.y, 0, 0x4423)); break; }
    }
    return r;
}

__device__ __forceinline__ f3 fetch_yuv(const SrcDesc& S, int sx, int sy)
{
    if (S.layout == LAYOUT_PACKED422) return fetch_packed422(S, sx, sy);
    if (S.layout == LAYOUT_PACKED444) return fetch_packed444(S, sx, sy);
    if (S.layout == LAYOUT_RGB)       return fetch_rgb(S, sx, sy);
    if (S.layout == LAYOUT_GRAY)      { f3 r; r.x = src_comp(S, 0, sx, sy, 0, 1); r.y = 0.0f; r.z = 0.0f; return r; }   // R8/R16 view: .gb = 0
    return fetch_planar(S, sx, sy);
}

// ---------------------------------------------------------------------------------------------
// colour conversion
// ---------------------------------------------------------------------------------------------
// Shaders/convert/st2084.hlsl:1-5 (exactly representable)
#define ST2084_M1 0.1593017578125f
#define ST2084_M2 78.84375f
#define ST2084_C1 0.8359375f
#define ST2084_C2 18.8515625f
#define ST2084_C3 18.6875f

// quotients of the tone-map: operands are in safe exponent ranges (den in [0.16, 19], numerators in [0, 1e3])
__device__ __forceinline__ float st2084_to_linear(const ColorParams& P, float v, float factor)          // st2084.hlsl:9-16
{
    const float a   = powr(v, P.inv_m2);
    const float num = fmaxf(__fsub_rn(a, ST2084_C1), 0.0f);
    const float den = fma_rn(-ST2084_C3, a, ST2084_C2);
    const float b   = div_rn_safe(num, den);
    return __fmul_rn(powr(b, P.inv_m1), factor);
}

__device__ __forceinline__ float linear_to_st2084(float v, float rdivider)                              // st2084.hlsl:18-25
{
    v = __fmul_rn(v, rdivider);
    const float p = powr(v, ST2084_M1);
    const float q = div_rn_safe(fma_rn(ST2084_C2, p, ST2084_C1), fma_rn(ST2084_C3, p, 1.0f));
    return powr(q, ST2084_M2);
}

__device__ __forceinline__ float hable(const ColorParams& P, float x)                                   // hdr_tone_mapping.hlsl:1-6
{
    const float A = 0.15f, B = 0.50f;
    return __fsub_rn(div_rn_safe(fma_rn(x, fma_rn(A, x, P.hable_cb), P.hable_de), fma_rn(x, fma_rn(A, x, B), P.hable_df)), P.hable_ef);
}

__device__ __forceinline__ float inverse_hlg(const ColorParams& P, float x)                             // hlg.hlsl:1-11
{
    const float b = 0.28466892f, c = 0.55991073f;
    const float lo = __fmul_rn(__fmul_rn(x, x), 4.0f);
    const float hi = __fadd_rn(ex2p(__fmul_rn(__fsub_rn(x, c), P.hlg_ra), 1.44269504f), b);
    return (x <= 0.5f) ? lo : hi;
}

__device__ __forceinline__ f3 mat3_apply(const float m[3][3], f3 v)               // mul(float3x3, float3): three dp3
{
    f3 r;
    r.x = fma_rn(m[0][2], v.z, fma_rn(m[0][1], v.y, __fmul_rn(m[0][0], v.x)));
    r.y = fma_rn(m[1][2], v.z, fma_rn(m[1][1], v.y, __fmul_rn(m[1][0], v.x)));
    r.z = fma_rn(m[2][2], v.z, fma_rn(m[2][1], v.y, __fmul_rn(m[2][0], v.x)));
    return r;
}

// ---- Dolby Vision reshaping (ShaderDoviReshape / ShaderDoviReshapePoly, Source/Shaders.cpp:531-589; reshape_mmr :734-765)
__device__ __forceinline__ float dv_dot3(const float m[4], float x, float y, float z) { return fma_rn(m[2], z, fma_rn(m[1], y, __fmul_rn(m[0], x))); }
__device__ __forceinline__ float dv_dot4(const float m[4], float x, float y, float z, float w) { return fma_rn(m[3], w, fma_rn(m[2], z, fma_rn(m[1], y, __fmul_rn(m[0], x)))); }

__device__ __forceinline__ float dovi_mmr(const DoviParams& D, int c, const float co[4], f3 sig)
{
    const int idx = D.mmr_single[c] ? 0 : (int)co[1];
    const float (*M)[4] = D.mmr[c];
    float s = co[0];
    const float x0 = __fmul_rn(sig.x, sig.y), x1 = __fmul_rn(sig.x, sig.z), x2 = __fmul_rn(sig.y, sig.z), x3 = __fmul_rn(x0, sig.z);
    s = __fadd_rn(s, dv_dot3(M[idx + 0], sig.x, sig.y, sig.z));
    s = __fadd_rn(s, dv_dot4(M[idx + 1], x0, x1, x2, x3));
    if (D.max_order[c] >= 2) {
        const int order = (int)co[3];
        if (D.min_order[c] < 2 && order < 2) return s;
        const float sx = __fmul_rn(sig.x, sig.x), sy = __fmul_rn(sig.y, sig.y), sz = __fmul_rn(sig.z, sig.z);
        const float y0 = __fmul_rn(x0, x0), y1 = __fmul_rn(x1, x1), y2 = __fmul_rn(x2, x2), y3 = __fmul_rn(x3, x3);
        s = __fadd_rn(s, dv_dot3(M[idx + 2], sx, sy, sz));
        s = __fadd_rn(s, dv_dot4(M[idx + 3], y0, y1, y2, y3));
        if (D.max_order[c] == 3) {
            if (D.min_order[c] < 3 && order < 3) return s;
            s = __fadd_rn(s, dv_dot3(M[idx + 4], __fmul_rn(sx, sig.x), __fmul_rn(sy, sig.y), __fmul_rn(sz, sig.z)));
            s = __fadd_rn(s, dv_dot4(M[idx + 5], __fmul_rn(y0, x0), __fmul_rn(y1, x1), __fmul_rn(y2, x2), __fmul_rn(y3, x3)));
        }
    }
    return s;
}

__device__ __forceinline__ f3 dovi_reshape(const DoviParams& D, f3 yuv)
{
    f3 sig; sig.x = sat(yuv.x); sig.y = sat(yuv.y); sig.z = sat(yuv.z);
    float r[3];
    #pragma unroll
    for (int c = 0; c < 3; c++) {
        const float s = c == 0 ? sig.x : (c == 1 ? sig.y : sig.z);
        const float* pv = D.pivots[c];
        const int k = (s < pv[3]) ? ((s < pv[1]) ? ((s < pv[0]) ? 0 : 1) : ((s < pv[2]) ? 2 : 3))
                                  : ((s < pv[5]) ? ((s < pv[4]) ? 4 : 5) : ((s < pv[6]) ? 6 : 7));
        const float* co = D.coeffs[c][k];
        bool poly = true;
        if (D.has_mmr) poly = (D.methods[c] == 3) ? (co[3] == 0.0f) : (D.methods[c] == 1);
        r[c] = sat(poly ? fma_rn(fma_rn(co[2], s, co[1]), s, co[0]) : dovi_mmr(D, c, co, sig));
    }
    f3 o; o.x = r[0]; o.y = r[1]; o.z = r[2];
    return o;
}

// convert shader body; returns the value before the render-target write.  TR >= 0 fixes the transfer at compile time.
// D (optional): Dolby Vision -- reshape before the matrix, PQ -> linear -> LMS matrix -> PQ after it (Source/Shaders.cpp:842-858).
template <int TR = -1>
__device__ __forceinline__ f3 convert_color(const ColorParams& P, f3 yuv, const DoviParams* D = nullptr)
{
    const int tr = (TR >= 0) ? TR : P.transfer;
    if (D) yuv = dovi_reshape(*D, yuv);
    f3 c = mat3_apply(P.cm, yuv);                                                   // Source/Shaders.cpp:820
    c.x = __fadd_rn(c.x, P.cc[0]); c.y = __fadd_rn(c.y, P.cc[1]); c.z = __fadd_rn(c.z, P.cc[2]);
    if (D) {
        f3 l;                                                                       // ST2084ToLinear(max(color, 0), 1.0)
        l.x = st2084_to_linear(P, fmaxf(c.x, 0.0f), 1.0f); l.y = st2084_to_linear(P, fmaxf(c.y, 0.0f), 1.0f); l.z = st2084_to_linear(P, fmaxf(c.z, 0.0f), 1.0f);
        const f3 m = mat3_apply(D->mat, l);
        c.x = linear_to_st2084(fmaxf(m.x, 0.0f), 1.0f); c.y = linear_to_st2084(fmaxf(m.y, 0.0f), 1.0f); c.z = linear_to_st2084(fmaxf(m.z, 0.0f), 1.0f);
    }
    if (tr == TRANSFER_NONE) return c;
    if (tr == TRANSFER_SDR_2020) {                                                  // :892-915
        f3 l; l.x = powr(sat(c.x), P.sdr_gamma); l.y = powr(sat(c.y), P.sdr_gamma); l.z = powr(sat(c.z), P.sdr_gamma);
        c = mat3_apply(P.gamut, l);
    } else {
        if (tr == TRANSFER_HLG_TO_SDR || tr == TRANSFER_HLG_TO_PQ) {                // :862-868 and :886-891: the same three lines
            f3 l; l.x = inverse_hlg(P, sat(c.x)); l.y = inverse_hlg(P, sat(c.y)); l.z = inverse_hlg(P, sat(c.z));
            const float dt = fma_rn(0.0593f, l.z, fma_rn(0.6780f, l.y, __fmul_rn(0.2627f, l.x)));   // hlg.hlsl:16-17
            const float g  = powr(__fmul_rn(2000.0f, dt), 0.2f);
            c.x = linear_to_st2084(__fmul_rn(l.x, g), P.r1000);
            c.y = linear_to_st2084(__fmul_rn(l.y, g), P.r1000);
            c.z = linear_to_st2084(__fmul_rn(l.z, g), P.r1000);
            if (tr == TRANSFER_HLG_TO_PQ) return c;
        }
        c.x = sat(c.x); c.y = sat(c.y); c.z = sat(c.z);                              // :870-872
        if (D && D->l2_enabled) {                                                   // DolbyVisionTrims, :787-795, :873-877
            // The shader's text has no guard here: a negative pow base, a zero luminance (0 / 0), a zero channel under a non-positive
            // saturation gain (0 * inf) or a value past the pole of ST2084ToLinear (c2 - c3 * v^(1/m2) <= 0) make NaN, which
            // mul(matrix_conv_prim, ...) spreads over the pixel and the final saturate turns into 0: the reference's pixel is (0, 0, 0).
            // powr returns 0 for such bases, so the cases are flagged explicitly (same tests as the oracle's dovi_trims).
            const float b0 = fma_rn(c.x, D->l2_slope, D->l2_offset), b1 = fma_rn(c.y, D->l2_slope, D->l2_offset), b2 = fma_rn(c.z, D->l2_slope, D->l2_offset);
            bool poisoned = (b0 < 0.0f) || (b1 < 0.0f) || (b2 < 0.0f);
            c.x = powr(b0, D->l2_power);
            c.y = powr(b1, D->l2_power);
            c.z = powr(b2, D->l2_power);
            const float Y = fma_rn(0.0593f, c.z, fma_rn(0.6780f, c.y, __fmul_rn(0.2627f, c.x)));
            const float g = __fadd_rn(1.0f, D->l2_chroma_weight);
            poisoned = poisoned || (Y == 0.0f) || (!(D->l2_saturation_gain > 0.0f) && (c.x == 0.0f || c.y == 0.0f || c.z == 0.0f));
            c.x = __fmul_rn(c.x, powr(__fdiv_rn(__fmul_rn(g, c.x), Y), D->l2_saturation_gain));
            c.y = __fmul_rn(c.y, powr(__fdiv_rn(__fmul_rn(g, c.y), Y), D->l2_saturation_gain));
            c.z = __fmul_rn(c.z, powr(__fdiv_rn(__fmul_rn(g, c.z), Y), D->l2_saturation_gain));
            poisoned = poisoned || !(fma_rn(-ST2084_C3, powr(c.x, P.inv_m2), ST2084_C2) > 0.0f)
                                || !(fma_rn(-ST2084_C3, powr(c.y, P.inv_m2), ST2084_C2) > 0.0f)
                                || !(fma_rn(-ST2084_C3, powr(c.z, P.inv_m2), ST2084_C2) > 0.0f);
            if (poisoned) { f3 z; z.x = 0.0f; z.y = 0.0f; z.z = 0.0f; return z; }
        }
        f3 h;                                                                       // :878-880; x / HABLE_DIV -> x * (1/HABLE_DIV)
        h.x = __fmul_rn(hable(P, st2084_to_linear(P, c.x, P.lum_scale)), P.hable_rdiv);
        h.y = __fmul_rn(hable(P, st2084_to_linear(P, c.y, P.lum_scale)), P.hable_rdiv);
        h.z = __fmul_rn(hable(P, st2084_to_linear(P, c.z, P.lum_scale)), P.hable_rdiv);
        c = mat3_apply(P.gamut, h);                                                 // :881
    }
    c.x = powr(sat(c.x), P.inv_gamma22); c.y = powr(sat(c.y), P.inv_gamma22); c.z = powr(sat(c.z), P.inv_gamma22);   // :917-923
    return c;
}

// ---------------------------------------------------------------------------------------------
// final pass, Shaders/d3d11/ps_final_pass.hlsl:23-31: floor(pixel*Q + dither)/Q written to a UNORM target
// ---------------------------------------------------------------------------------------------
// the quotient is correctly rounded either way: n is an integer-valued float and Q is 255 or 1023, so the residual-corrected
// reciprocal sequence (div_rn_safe) is exact unless n is huge / infinite (fp16 infinities in the pass input), which takes div.rn
__device__ __forceinline__ float final_quant(float p, float Q, float d)
{
    const float n = floorf(fma_rn(p, Q, d));
    return fabsf(n) < 1e30f ? div_rn_safe(n, Q) : __fdiv_rn(n, Q);
}

}  // namespace b200vr
