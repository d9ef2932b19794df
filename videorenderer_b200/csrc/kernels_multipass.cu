// kernels_multipass.cu -- one CUDA kernel per pass of the reference's chain
// (CDX11VideoProcessor::Process, Source/DX11VideoProcessor.cpp:3285-3424), intermediates in HBM in the
// reference's own surface formats.  This family handles every plan; the fused kernel
// (kernel_fused.cu) covers the plans that matter for throughput.  sm_100a, -fmad=false.
#include <cuda_runtime.h>
#include <atomic>
#include "kernels.h"
#include "pixel_ops.cuh"
#include "tonemap_ops.cuh"
#include "pixel_ops2.cuh"
#include "mpcvr_math2.cuh"
#include "dither32.h"

namespace b200vr {

__device__ __align__(16) const uint16_t g_dither32[1024] = { MPCVR_DITHER32_VALUES };

// ---------------------------------------------------------------------------------------------
// ConvertColorPass: one thread per pixel of m_TexConvertOutput (= m_srcRect size)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_convert(const __grid_constant__ SrcDesc S, const __grid_constant__ ColorParams P, const __grid_constant__ SurfDesc out,
          float4* __restrict__ dbg, const DoviParams* __restrict__ dovi)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out.w || y >= out.h) return;
    const f3 yuv = fetch_yuv(S, S.rect_l + x, S.rect_t + y);
    const f3 rgb = convert_color<-1>(P, yuv, dovi);
    if (dbg) dbg[(size_t)y * out.w + x] = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
    surf_store(out, x, y, rgb);
}

// ConvertColorPass, two horizontally adjacent pixels per thread: the shader body runs in the two lanes of FFMA2 / FADD2 /
// FMUL2 (pixel_ops2.cuh), bit-identical to k_convert; fetches stay scalar.  Frames with Dolby Vision data use k_convert.
__global__ void __launch_bounds__(256)
k_convert2(const __grid_constant__ SrcDesc S, const __grid_constant__ ColorParams P, const __grid_constant__ SurfDesc out,
           float4* __restrict__ dbg)
{
    const int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out.w || y >= out.h) return;
    const bool two = x + 1 < out.w;
    const f3 a = fetch_yuv(S, S.rect_l + x, S.rect_t + y);
    const f3 b = two ? fetch_yuv(S, S.rect_l + x + 1, S.rect_t + y) : a;
    f3x2 yuv; yuv.x = make_float2(a.x, b.x); yuv.y = make_float2(a.y, b.y); yuv.z = make_float2(a.z, b.z);
    const f3x2 rgb = convert_color2(P, yuv);
    f3 r0; r0.x = rgb.x.x; r0.y = rgb.y.x; r0.z = rgb.z.x;
    if (dbg) dbg[(size_t)y * out.w + x] = make_float4(r0.x, r0.y, r0.z, 1.0f);
    surf_store(out, x, y, r0);
    if (two) {
        f3 r1; r1.x = rgb.x.y; r1.y = rgb.y.y; r1.z = rgb.z.y;
        if (dbg) dbg[(size_t)y * out.w + x + 1] = make_float4(r1.x, r1.y, r1.z, 1.0f);
        surf_store(out, x + 1, y, r1);
    }
}

// bilinear chroma sample of fetch_planar at the PX luma positions of a span whose texels [t0 - 1, t0 + PX/2] sit in T[row][.]:
// qx = 2 (sx0 + i) + OX with 2 sx0 = 4 t0, so texel (qx >> 2) is T[.][((2 i + OX) >> 2) + 1]; OX is a template value to keep T in registers
template <int PX, int OX>
__device__ __forceinline__ void span_chroma(const float2 (&T)[2][PX / 2 + 2], float fy, float (&dU)[PX], float (&dV)[PX])
{
    #pragma unroll
    for (int i = 0; i < PX; i++) {
        const int q = 2 * i + OX;
        const int j = (q >> 2) + 1;
        const float fx = (float)(q & 3) * 0.25f;
        const float2 c00 = T[0][j], c10 = T[0][j + 1], c01 = T[1][j], c11 = T[1][j + 1];
        dU[i] = lerp_mad(lerp_mad(c00.x, c10.x, fx), lerp_mad(c01.x, c11.x, fx), fy);
        dV[i] = lerp_mad(lerp_mad(c00.y, c10.y, fx), lerp_mad(c01.y, c11.y, fx), fy);
    }
}

// ConvertColorPass for bi-planar sources (NV12 / P010 / P016 / P210 / P216) with bilinear chroma and for YUY2 / UYVY, in two phases per CTA so that the
// source is read with 128-bit loads and the shader body still exists once in the code:
//   phase 1  each thread fetches a span of 16 / BYTES pixels of one row: one 128-bit luma load, one 128-bit load per chroma row plus the
//            two neighbour texels, UNORM reads, the bilinear sample of fetch_planar at every luma position -> Y, U, V floats in shared memory
//   phase 2  the CTA walks its pixels two at a time through convert_color2 (any transfer chain) and writes texel pairs with 64/128-bit stores
// Same device functions and the same operation order per pixel as k_convert2, so the surface is bit-identical.
template <int BYTES, bool PACKED>
__global__ void __launch_bounds__(128)
k_convert_bp(const __grid_constant__ SrcDesc S, const __grid_constant__ ColorParams P, const __grid_constant__ SurfDesc out)
{
    constexpr int PX = PACKED ? 8 : 16 / BYTES;                  // pixels per span: 16 (8-bit planes), 8 (16-bit planes, YUY2 / UYVY words)
    constexpr int TW = 32 * PX, TH = 4;                          // CTA tile
    __shared__ __align__(16) float2 sY[TH][TW / 2], sU[TH][TW / 2], sV[TH][TW / 2];
    const int bx0 = blockIdx.x * TW, by0 = blockIdx.y * TH;
    {
        const int x0 = bx0 + threadIdx.x * PX, y = by0 + threadIdx.y;
        if (x0 < out.w && y < out.h) {
            const int sy = S.rect_t + y, sx0 = S.rect_l + x0;
            float* dY = reinterpret_cast<float*>(&sY[threadIdx.y][0]) + threadIdx.x * PX;
            float* dU = reinterpret_cast<float*>(&sU[threadIdx.y][0]) + threadIdx.x * PX;
            float* dV = reinterpret_cast<float*>(&sV[threadIdx.y][0]) + threadIdx.x * PX;
            if (x0 + PX > out.w) {                               // ragged last span: per-pixel fetch
                for (int x = x0; x < out.w; x++) { const f3 t = fetch_yuv(S, S.rect_l + x, sy); dY[x - x0] = t.x; dU[x - x0] = t.y; dV[x - x0] = t.z; }
            } else if (PACKED) {
                // YUY2 / UYVY (fetch_packed422): four macropixels in one 128-bit load plus the next one (CLAMP) for the last odd pixel;
                // even pixels take their macropixel's chroma, odd pixels the mean of this and the next macropixel
                const uint8_t* row = S.pl[0] + (ptrdiff_t)sy * S.pitch[0];
                const int mx0 = sx0 >> 1, mw = S.W >> 1;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + (size_t)mx0 * 4));
                const uint32_t w[5] = { v.x, v.y, v.z, v.w, __ldg(reinterpret_cast<const uint32_t*>(row) + min(mx0 + 4, mw - 1)) };
                const bool uyvy = S.packed_kind == 1;
                float cu[5], cv[5], vY[PX], vU[PX], vV[PX];
                #pragma unroll
                for (int m = 0; m < 5; m++) {
                    const uint32_t t = uyvy ? w[m] : (w[m] >> 8);            // bytes U ? V ?
                    cu[m] = unorm8(t & 255u); cv[m] = unorm8((t >> 16) & 255u);
                }
                #pragma unroll
                for (int m = 0; m < 4; m++) {
                    const uint32_t t = uyvy ? (w[m] >> 8) : w[m];            // bytes Y0 ? Y1 ?
                    vY[2 * m] = unorm8(t & 255u); vY[2 * m + 1] = unorm8((t >> 16) & 255u);
                    vU[2 * m] = cu[m]; vV[2 * m] = cv[m];
                    vU[2 * m + 1] = __fmul_rn(__fadd_rn(cu[m], cu[m + 1]), 0.5f); vV[2 * m + 1] = __fmul_rn(__fadd_rn(cv[m], cv[m + 1]), 0.5f);
                }
                #pragma unroll
                for (int i = 0; i < PX; i += 4) {
                    reinterpret_cast<float4*>(dY)[i >> 2] = make_float4(vY[i], vY[i + 1], vY[i + 2], vY[i + 3]);
                    reinterpret_cast<float4*>(dU)[i >> 2] = make_float4(vU[i], vU[i + 1], vU[i + 2], vU[i + 3]);
                    reinterpret_cast<float4*>(dV)[i >> 2] = make_float4(vV[i], vV[i + 1], vV[i + 2], vV[i + 3]);
                }
            } else {
                int ox, oy, my;
                if (S.subsampling == 420) {
                    my = 2;
                    if (S.siting == ChromaSub_Cosited)    { ox = 0;  oy = 0;  }
                    else if (S.siting == ChromaSub_MPEG1) { ox = -1; oy = -1; }
                    else                                  { ox = 0;  oy = -1; }
                } else { my = 4; ox = 0; oy = 0; }
                float vY[PX], vU[PX], vV[PX];
                {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(S.pl[0] + (ptrdiff_t)sy * S.pitch[0] + (size_t)sx0 * BYTES));
                    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
                    #pragma unroll
                    for (int i = 0; i < PX; i++)
                        vY[i] = BYTES == 1 ? unorm8((w[i >> 2] >> (8 * (i & 3))) & 255u) : unorm16((w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu);
                }
                // chroma texel columns [t0 - 1, t0 + PX/2] of the two rows the bilinear sample spans (sampler CLAMP on indices)
                constexpr int NTX = PX / 2 + 2;
                const int t0 = sx0 >> 1;
                const int qy = my * sy + oy;
                const float fy = (float)(qy & 3) * 0.25f;
                const int r0 = clampi(qy >> 2, 0, S.ch - 1), r1 = clampi((qy >> 2) + 1, 0, S.ch - 1);
                float2 T[2][NTX];
                #pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const uint8_t* row = S.pl[1] + (ptrdiff_t)(rr ? r1 : r0) * S.pitch[1];
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + (size_t)t0 * 2 * BYTES));
                    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
                    #pragma unroll
                    for (int i = 0; i < PX / 2; i++) {
                        if (BYTES == 1) { const uint32_t c = (w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu; T[rr][1 + i] = make_float2(unorm8(c & 255u), unorm8(c >> 8)); }
                        else            { const uint32_t c = w[i]; T[rr][1 + i] = make_float2(unorm16(c & 0xFFFFu), unorm16(c >> 16)); }
                    }
                    const int tl = clampi(t0 - 1, 0, S.cw - 1), tr = clampi(t0 + PX / 2, 0, S.cw - 1);
                    if (BYTES == 1) {
                        const uint32_t a = __ldg(reinterpret_cast<const uint16_t*>(row) + tl), b = __ldg(reinterpret_cast<const uint16_t*>(row) + tr);
                        T[rr][0] = make_float2(unorm8(a & 255u), unorm8(a >> 8)); T[rr][NTX - 1] = make_float2(unorm8(b & 255u), unorm8(b >> 8));
                    } else {
                        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(row) + tl), b = __ldg(reinterpret_cast<const uint32_t*>(row) + tr);
                        T[rr][0] = make_float2(unorm16(a & 0xFFFFu), unorm16(a >> 16)); T[rr][NTX - 1] = make_float2(unorm16(b & 0xFFFFu), unorm16(b >> 16));
                    }
                }
                if (ox == 0) span_chroma<PX, 0>(T, fy, vU, vV); else span_chroma<PX, -1>(T, fy, vU, vV);
                #pragma unroll
                for (int i = 0; i < PX; i += 4) {                // 128-bit shared stores
                    reinterpret_cast<float4*>(dY)[i >> 2] = make_float4(vY[i], vY[i + 1], vY[i + 2], vY[i + 3]);
                    reinterpret_cast<float4*>(dU)[i >> 2] = make_float4(vU[i], vU[i + 1], vU[i + 2], vU[i + 3]);
                    reinterpret_cast<float4*>(dV)[i >> 2] = make_float4(vV[i], vV[i + 1], vV[i + 2], vV[i + 3]);
                }
            }
        }
    }
    __syncthreads();
    const int tid = threadIdx.y * 32 + threadIdx.x;
    #pragma unroll 1
    for (int p = tid; p < TH * (TW / 2); p += 128) {
        const int r = p / (TW / 2), c = p - r * (TW / 2);
        const int x = bx0 + 2 * c, y = by0 + r;
        if (x >= out.w || y >= out.h) continue;
        f3x2 yuv; yuv.x = sY[r][c]; yuv.y = sU[r][c]; yuv.z = sV[r][c];
        const f3x2 rgb = convert_color2(P, yuv);
        uint8_t* row = (uint8_t*)out.ptr + (size_t)y * out.pitch;
        if (x + 1 < out.w) {
            if (out.fmt == MPCVR_SURF_RGBA16F) {
                const __half2 a0 = __floats2half2_rn(rgb.x.x, rgb.y.x), b0 = __floats2half2_rn(rgb.z.x, 1.0f), a1 = __floats2half2_rn(rgb.x.y, rgb.y.y), b1 = __floats2half2_rn(rgb.z.y, 1.0f);
                reinterpret_cast<uint4*>(row)[x >> 1] = make_uint4(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&b0),
                                                                   *reinterpret_cast<const uint32_t*>(&a1), *reinterpret_cast<const uint32_t*>(&b1));
            } else if (out.fmt == MPCVR_SURF_RGB10A2) {
                reinterpret_cast<uint2*>(row)[x >> 1] = make_uint2(qunorm(rgb.x.x, 1023.0f) | (qunorm(rgb.y.x, 1023.0f) << 10) | (qunorm(rgb.z.x, 1023.0f) << 20) | (3u << 30),
                                                                   qunorm(rgb.x.y, 1023.0f) | (qunorm(rgb.y.y, 1023.0f) << 10) | (qunorm(rgb.z.y, 1023.0f) << 20) | (3u << 30));
            } else {                                             // B8G8R8A8
                reinterpret_cast<uint2*>(row)[x >> 1] = make_uint2(qunorm(rgb.z.x, 255.0f) | (qunorm(rgb.y.x, 255.0f) << 8) | (qunorm(rgb.x.x, 255.0f) << 16) | (255u << 24),
                                                                   qunorm(rgb.z.y, 255.0f) | (qunorm(rgb.y.y, 255.0f) << 8) | (qunorm(rgb.x.y, 255.0f) << 16) | (255u << 24));
            }
        } else {
            f3 r0; r0.x = rgb.x.x; r0.y = rgb.y.x; r0.z = rgb.z.x;
            surf_store(out, x, y, r0);
        }
    }
}

// 128-bit path of the convert pass: bi-planar, bilinear chroma, planes / pitches / rect origin 16-byte aligned, an 8-, 10- or fp16 target
static bool ConvertSpanOk(const SrcDesc& S, const SurfDesc& out)
{
    if (out.fmt != MPCVR_SURF_BGRA8 && out.fmt != MPCVR_SURF_RGB10A2 && out.fmt != MPCVR_SURF_RGBA16F) return false;
    if (((uintptr_t)out.ptr & 15) || (out.pitch & 15)) return false;
    if (S.layout == LAYOUT_PACKED422)            // YUY2 / UYVY with the averaging chroma fetch
        return S.bytes == 1 && S.packed_kind <= 1 && S.chroma_mode != MPCVR_CHROMA_CatmullRom && !S.deint_blend && !((uintptr_t)S.pl[0] & 15) &&
               S.pitch[0] > 0 && !(S.pitch[0] & 15) && (S.rect_l % 8) == 0;
    if (S.layout != LAYOUT_BIPLANAR || S.chroma_mode != MPCVR_CHROMA_Bilinear || S.deint_blend || S.lsb_shift != 0) return false;
    if ((S.subsampling != 420 && S.subsampling != 422) || (S.bytes != 1 && S.bytes != 2)) return false;
    if (((uintptr_t)S.pl[0] & 15) || ((uintptr_t)S.pl[1] & 15) || S.pitch[0] <= 0 || S.pitch[1] <= 0 || (S.pitch[0] & 15) || (S.pitch[1] & 15)) return false;
    if ((S.rect_l % (16 / S.bytes)) != 0) return false;
    if (out.fmt != MPCVR_SURF_BGRA8 && out.fmt != MPCVR_SURF_RGB10A2 && out.fmt != MPCVR_SURF_RGBA16F) return false;
    if (((uintptr_t)out.ptr & 15) || (out.pitch & 15)) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------
// TextureResizeShader along one axis: one thread per output pixel, taps from the host table
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_resize(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out, const __grid_constant__ AxisTable T,
         const int axis, const int32_t* __restrict__ other_idx, const int other_dst0, const int other_dst_len, const int in_axis)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
    // (a, b) index the output rectangle: a along X, b along Y
    const int i  = axis == 0 ? a : b;      // index along the filtered axis
    const int j  = axis == 0 ? b : a;      // index along the other axis
    if (i >= T.dst_len || j >= other_dst_len) return;
    const int ox = axis == 0 ? T.dst0 + i : other_dst0 + j;
    const int oy = axis == 0 ? other_dst0 + j : T.dst0 + i;
    if (ox < 0 || oy < 0 || ox >= out.w || oy >= out.h) return;
    // in_axis: the texture axis the taps run along; differs from the output axis on a quad turned by 90/270 degrees
    const int lim   = in_axis == 0 ? in.w : in.h;
    const int fixed = clampi(__ldg(other_idx + j), 0, (in_axis == 0 ? in.h : in.w) - 1);
    const int base  = __ldg(T.base + i);
    const float* w  = T.w + (size_t)i * T.ntaps;
    f3 acc = {0.f, 0.f, 0.f};
    for (int k = 0; k < T.ntaps; k++) {
        int idx = base + k;
        if (T.dup_q1 && k == 1) idx = base;                 // ps_interpolation_lanczos3.hlsl:33-34
        idx = clampi(idx, 0, lim - 1);
        const f3 p = in_axis == 0 ? surf_load(in, idx, fixed) : surf_load(in, fixed, idx);
        const float wk = __ldg(w + k);
        if (k == 0 && T.kind != SCALER_CONV) { acc.x = __fmul_rn(wk, p.x); acc.y = __fmul_rn(wk, p.y); acc.z = __fmul_rn(wk, p.z); }
        else { acc.x = fma_rn(wk, p.x, acc.x); acc.y = fma_rn(wk, p.y, acc.y); acc.z = fma_rn(wk, p.z, acc.z); }
    }
    if (T.kind == SCALER_CONV) {                            // avg /= ww, ps_convolution.hlsl:47
        const float ww = __ldg(T.wsum + i);
        acc.x = __fdiv_rn(acc.x, ww); acc.y = __fdiv_rn(acc.y, ww); acc.z = __fdiv_rn(acc.z, ww);
    }
    surf_store(out, ox, oy, acc);
}

// Same pass, four consecutive outputs of the filtered axis per thread: neighbouring outputs share most of their taps, so the
// union window is loaded and decoded once and every texel feeds the (up to four) accumulators it belongs to.  Each output
// still receives its taps in ascending k with the same first-product / fma chain, so results are bit-identical to k_resize.
__global__ void __launch_bounds__(256)
k_resize4(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out, const __grid_constant__ AxisTable T,
          const int axis, const int32_t* __restrict__ other_idx, const int other_dst0, const int other_dst_len)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
    const int i0 = 4 * (axis == 0 ? a : b);        // first output index along the filtered axis
    const int j  = axis == 0 ? b : a;              // index along the other axis
    if (i0 >= T.dst_len || j >= other_dst_len) return;
    const int n = min(4, T.dst_len - i0);
    const int oj = other_dst0 + j;
    if (oj < 0 || oj >= (axis == 0 ? out.h : out.w)) return;
    const int lim   = axis == 0 ? in.w : in.h;
    const int fixed = clampi(__ldg(other_idx + j), 0, (axis == 0 ? in.h : in.w) - 1);
    const int nt = T.ntaps;
    int base[4]; const float* w[4];
    int lo = 0x7fffffff, hi = -0x7fffffff;
    #pragma unroll
    for (int q = 0; q < 4; q++) {
        const int i = i0 + min(q, n - 1);
        base[q] = __ldg(T.base + i); w[q] = T.w + (size_t)i * nt;
        lo = min(lo, base[q]); hi = max(hi, base[q] + nt);
    }
    f3 acc[4];
    #pragma unroll
    for (int q = 0; q < 4; q++) { acc[q].x = 0.f; acc[q].y = 0.f; acc[q].z = 0.f; }
    const bool conv = T.kind == SCALER_CONV, dup = T.dup_q1 != 0;
    f3 prev = {0.f, 0.f, 0.f};
    for (int t = lo; t < hi; t++) {
        const int idx = clampi(t, 0, lim - 1);
        const f3 p = axis == 0 ? surf_load(in, idx, fixed) : surf_load(in, fixed, idx);
        #pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = t - base[q];
            if (q < n && k >= 0 && k < nt) {
                const f3 v = (dup && k == 1) ? prev : p;                 // ps_interpolation_lanczos3.hlsl:33-34: tap 1 re-reads tap 0's texel
                const float wk = __ldg(w[q] + k);
                if (k == 0 && !conv) { acc[q].x = __fmul_rn(wk, v.x); acc[q].y = __fmul_rn(wk, v.y); acc[q].z = __fmul_rn(wk, v.z); }
                else { acc[q].x = fma_rn(wk, v.x, acc[q].x); acc[q].y = fma_rn(wk, v.y, acc[q].y); acc[q].z = fma_rn(wk, v.z, acc[q].z); }
            }
        }
        prev = p;
    }
    #pragma unroll
    for (int q = 0; q < 4; q++) {
        if (q >= n) break;
        const int i = i0 + q;
        f3 r = acc[q];
        if (conv) { const float ww = __ldg(T.wsum + i); r.x = __fdiv_rn(r.x, ww); r.y = __fdiv_rn(r.y, ww); r.z = __fdiv_rn(r.z, ww); }
        const int ox = axis == 0 ? T.dst0 + i : oj, oy = axis == 0 ? oj : T.dst0 + i;
        if (ox < 0 || oy < 0 || ox >= out.w || oy >= out.h) continue;
        surf_store(out, ox, oy, r);
    }
}

// Same pass for upright draws with the tap count, both surface formats, the axis and the shader kind fixed at compile time:
// one thread = one index of the filtered axis x TWO neighbouring lines of the other axis.  The two outputs share the tap
// window, the clamps and the weights and ride in the two lanes of FFMA2; each still sees its taps in ascending k with the same
// first-product / fma chain as k_resize, so the pass is bit-identical.  Y passes fetch the texel pair with one 64/128-bit load.
template <int FMT> __device__ __forceinline__ f3 surf_load_f(const SurfDesc& s, int x, int y) { SurfDesc t = s; t.fmt = FMT; return surf_load(t, x, y); }
template <int FMT> __device__ __forceinline__ void surf_store_f(const SurfDesc& s, int x, int y, f3 v) { SurfDesc t = s; t.fmt = FMT; surf_store(t, x, y, v); }
template <int FMT> __device__ __forceinline__ f3 texel_decode(uint32_t lo, uint32_t hi)
{
    f3 r;
    if (FMT == MPCVR_SURF_RGB10A2)    { r.x = unorm10(lo & 1023u); r.y = unorm10((lo >> 10) & 1023u); r.z = unorm10((lo >> 20) & 1023u); }
    else if (FMT == MPCVR_SURF_BGRA8) { r.z = unorm8(lo & 255u); r.y = unorm8((lo >> 8) & 255u); r.x = unorm8((lo >> 16) & 255u); }
    else { const __half2 a = *reinterpret_cast<const __half2*>(&lo), b = *reinterpret_cast<const __half2*>(&hi); r.x = __low2float(a); r.y = __high2float(a); r.z = __low2float(b); }
    return r;
}

template <int NT, int INF, int OUTF, int AXIS, bool CONV>
__global__ void __launch_bounds__(256)
k_resize_p(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out, const __grid_constant__ AxisTable T,
           const int32_t* __restrict__ other_idx, const int other_dst0, const int other_dst_len)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
    const int i  = AXIS == 0 ? a : b;                  // index along the filtered axis
    const int j0 = 2 * (AXIS == 0 ? b : a);            // first of the two lines of the other axis
    if (i >= T.dst_len || j0 >= other_dst_len) return;
    const bool two = j0 + 1 < other_dst_len;
    const int lim = AXIS == 0 ? in.w : in.h, olim = AXIS == 0 ? in.h : in.w;
    const int fA = clampi(__ldg(other_idx + j0), 0, olim - 1);
    const int fB = two ? clampi(__ldg(other_idx + j0 + 1), 0, olim - 1) : fA;
    constexpr int TEXEL = INF == MPCVR_SURF_RGBA16F ? 8 : 4;
    const bool vec = AXIS == 1 && fB == fA + 1 && (fA & 1) == 0;          // the pair is one aligned 2-texel word of the row
    const int base = __ldg(T.base + i);
    const int nt = NT > 0 ? NT : T.ntaps;
    const float* w = T.w + (size_t)i * nt;
    const bool dup = T.dup_q1 != 0;
    float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax;
    #pragma unroll
    for (int k = 0; k < (NT > 0 ? NT : nt); k++) {
        int idx = base + k;
        if (!CONV && dup && k == 1) idx = base;                          // ps_interpolation_lanczos3.hlsl:33-34
        idx = clampi(idx, 0, lim - 1);
        f3 pa, pb;
        if (AXIS == 1) {
            const uint8_t* row = (const uint8_t*)in.ptr + (size_t)idx * in.pitch;
            if (vec) {
                if (TEXEL == 8) { const uint4 v = __ldg((const uint4*)(row + (size_t)fA * 8)); pa = texel_decode<INF>(v.x, v.y); pb = texel_decode<INF>(v.z, v.w); }
                else            { const uint2 v = __ldg((const uint2*)(row + (size_t)fA * 4)); pa = texel_decode<INF>(v.x, 0u); pb = texel_decode<INF>(v.y, 0u); }
            } else { pa = surf_load_f<INF>(in, fA, idx); pb = surf_load_f<INF>(in, fB, idx); }
        } else { pa = surf_load_f<INF>(in, idx, fA); pb = surf_load_f<INF>(in, idx, fB); }
        const float wk = __ldg(w + k);
        const float2 vx = make_float2(pa.x, pb.x), vy = make_float2(pa.y, pb.y), vz = make_float2(pa.z, pb.z);
        if (k == 0 && !CONV) { ax = mul2(vx, wk); ay = mul2(vy, wk); az = mul2(vz, wk); }
        else { ax = fma2(wk, vx, ax); ay = fma2(wk, vy, ay); az = fma2(wk, vz, az); }
    }
    f3 rA = {ax.x, ay.x, az.x}, rB = {ax.y, ay.y, az.y};
    if (CONV) {                                                          // avg /= ww, ps_convolution.hlsl:47
        const float ww = __ldg(T.wsum + i);
        const bool wok = ww > 1e-10f && ww < 1e10f;
        // correctly rounded either way: the residual-corrected reciprocal sequence when every intermediate stays a normal
        // number (always, for pixel data), div.rn otherwise
        auto dv = [&](float x) { const float ax = fabsf(x); return (wok && ax < 1e30f && (ax > 1e-30f || ax == 0.0f)) ? div_rn_safe(x, ww) : __fdiv_rn(x, ww); };
        rA.x = dv(rA.x); rA.y = dv(rA.y); rA.z = dv(rA.z);
        rB.x = dv(rB.x); rB.y = dv(rB.y); rB.z = dv(rB.z);
    }
    const int oi = T.dst0 + i, oj = other_dst0 + j0;
    {
        const int ox = AXIS == 0 ? oi : oj, oy = AXIS == 0 ? oj : oi;
        if (ox >= 0 && oy >= 0 && ox < out.w && oy < out.h) surf_store_f<OUTF>(out, ox, oy, rA);
    }
    if (two) {
        const int ox = AXIS == 0 ? oi : oj + 1, oy = AXIS == 0 ? oj + 1 : oi;
        if (ox >= 0 && oy >= 0 && ox < out.w && oy < out.h) surf_store_f<OUTF>(out, ox, oy, rB);
    }
}

// ---------------------------------------------------------------------------------------------
// TextureCopyRect + ps_simple: point-sampled copy / nearest-neighbour stretch
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_copy(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out,
       const int32_t* __restrict__ idx_x, const int32_t* __restrict__ idx_y, int dx0, int dy0, int dw, int dh, int transpose)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int ox = dx0 + x, oy = dy0 + y;
    if (ox < 0 || oy < 0 || ox >= out.w || oy >= out.h) return;
    // transpose (quarter turns): the source column follows the output row and vice versa
    const int ix = clampi(__ldg(idx_x + (transpose ? y : x)), 0, in.w - 1), iy = clampi(__ldg(idx_y + (transpose ? x : y)), 0, in.h - 1);
    surf_store(out, ox, oy, surf_load(in, ix, iy));
}

// ---------------------------------------------------------------------------------------------
// FinalPass: dither texel = window coordinates mod 32 (ditherCoordScale, :3204)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_final(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out, int x0, int y0, int x1, int y1, float Q)
{
    const int x = x0 + blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= x1 || y >= y1) return;
    const float d = (float)g_dither32[(y & 31) * 32 + (x & 31)] * 0.0009765625f;
    const f3 p = surf_load(in, x, y);
    f3 o; o.x = final_quant(p.x, Q, d); o.y = final_quant(p.y, Q, d); o.z = final_quant(p.z, Q, d);
    surf_store(out, x, y, o);
}

// ---------------------------------------------------------------------------------------------
// Jinc2 one-pass resizer (UPSCALE_Jinc2), Shaders/examples/ps_resize_onepass_jinc2.hlsl:44-101: a 2-D shader, 4x4 texels
// around floor(pc - .5) weighted by sin(d wa) sin(d wb) / d^2 of their distance d, normalised, anti-ringing clamp.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_jinc2(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out,
        int in_l, int in_t, int in_w, int in_h, int out_l, int out_t, int out_w, int out_h, int orient)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out_w || y >= out_h) return;
    const int ox = out_l + x, oy = out_t + y;
    if (ox < 0 || oy < 0 || ox >= out.w || oy >= out.h) return;
    // orient (FillVertices, Source/DX11VideoProcessor.cpp:130-179): bit 0 = quarter turn (u follows y), bit 1 = u reversed, bit 2 = v reversed
    const bool turned = orient & 1, u_rev = orient & 2, v_rev = orient & 4;
    const int ui = turned ? y : x, vi = turned ? x : y;
    const float su = __fdiv_rn((float)in_w, (float)(turned ? out_h : out_w)), sv = __fdiv_rn((float)in_h, (float)(turned ? out_w : out_h));   // :351-354
    const float du = __fmul_rn(__fadd_rn((float)ui, 0.5f), su), dv = __fmul_rn(__fadd_rn((float)vi, 0.5f), sv);
    const float pcx = u_rev ? __fsub_rn((float)(in_l + in_w), du) : __fadd_rn(du, (float)in_l);
    const float pcy = v_rev ? __fsub_rn((float)(in_t + in_h), dv) : __fadd_rn(dv, (float)in_t);
    const float pi = 3.14159265358979323846f;
    const float wa = __fmul_rn(0.416f, pi), wb = __fmul_rn(0.985f, pi), wab = __fmul_rn(wa, wb);
    const float tcx = __fadd_rn(floorf(__fsub_rn(pcx, 0.5f)), 0.5f), tcy = __fadd_rn(floorf(__fsub_rn(pcy, 0.5f)), 0.5f);
    const int ix = (int)floorf(tcx), iy = (int)floorf(tcy);
    float tot = 0.0f;
    f3 col = {0.f, 0.f, 0.f};
    f3 mn = {0.f, 0.f, 0.f}, mx = {0.f, 0.f, 0.f};
    #pragma unroll
    for (int n = 0; n < 4; n++) {
        const float vy = __fsub_rn(__fadd_rn(tcy, (float)(n - 1)), pcy);
        const int ty = clampi(iy + n - 1, 0, in.h - 1);
        float w[4]; f3 c[4];
        #pragma unroll
        for (int m = 0; m < 4; m++) {
            const float vx = __fsub_rn(__fadd_rn(tcx, (float)(m - 1)), pcx);
            const float d = __fsqrt_rn(fma_rn(vy, vy, __fmul_rn(vx, vx)));
            w[m] = (d == 0.0f) ? wab : __fdiv_rn(__fmul_rn(sin_spec(__fmul_rn(d, wa)), sin_spec(__fmul_rn(d, wb))), __fmul_rn(d, d));
            c[m] = surf_load(in, clampi(ix + m - 1, 0, in.w - 1), ty);
        }
        const float rs = __fadd_rn(__fadd_rn(__fadd_rn(w[0], w[1]), w[2]), w[3]);
        f3 r;
        r.x = fma_rn(w[3], c[3].x, fma_rn(w[2], c[2].x, fma_rn(w[1], c[1].x, __fmul_rn(w[0], c[0].x))));
        r.y = fma_rn(w[3], c[3].y, fma_rn(w[2], c[2].y, fma_rn(w[1], c[1].y, __fmul_rn(w[0], c[0].y))));
        r.z = fma_rn(w[3], c[3].z, fma_rn(w[2], c[2].z, fma_rn(w[1], c[1].z, __fmul_rn(w[0], c[0].z))));
        if (n == 0) { tot = rs; col = r; }
        else { tot = __fadd_rn(tot, rs); col.x = __fadd_rn(col.x, r.x); col.y = __fadd_rn(col.y, r.y); col.z = __fadd_rn(col.z, r.z); }
        if (n == 1) {                                          // c11, c21
            mn.x = fminf(c[1].x, c[2].x); mn.y = fminf(c[1].y, c[2].y); mn.z = fminf(c[1].z, c[2].z);
            mx.x = fmaxf(c[1].x, c[2].x); mx.y = fmaxf(c[1].y, c[2].y); mx.z = fmaxf(c[1].z, c[2].z);
        } else if (n == 2) {                                   // c12, c22
            mn.x = fminf(fminf(mn.x, c[1].x), c[2].x); mn.y = fminf(fminf(mn.y, c[1].y), c[2].y); mn.z = fminf(fminf(mn.z, c[1].z), c[2].z);
            mx.x = fmaxf(fmaxf(mx.x, c[1].x), c[2].x); mx.y = fmaxf(fmaxf(mx.y, c[1].y), c[2].y); mx.z = fmaxf(fmaxf(mx.z, c[1].z), c[2].z);
        }
    }
    f3 o;
    { const float v = __fdiv_rn(col.x, tot); o.x = fma_rn(0.8f, __fsub_rn(fminf(fmaxf(v, mn.x), mx.x), v), v); }
    { const float v = __fdiv_rn(col.y, tot); o.y = fma_rn(0.8f, __fsub_rn(fminf(fmaxf(v, mn.y), mx.y), v), v); }
    { const float v = __fdiv_rn(col.z, tot); o.z = fma_rn(0.8f, __fsub_rn(fminf(fmaxf(v, mn.z), mx.z), v), v); }
    surf_store(out, ox, oy, o);
}

// ---------------------------------------------------------------------------------------------
// HDR10 local tone-map post-scale step: Shaders/d3d11/ps_hdr10_tonemap.hlsl:34-336 without the Dolby Vision trims
// (L2Enabled = 0), drawn by TextureCopyRect over rect (Source/DX11VideoProcessor.cpp:3359-3367).
// Same contract as the other passes: written-order fp32, a*b+c fused, runtime quotients IEEE, pow = 2^(y log2 x).
// ---------------------------------------------------------------------------------------------
// (the per-pixel functions of this pass live in tonemap_ops.cuh)

// ---------------------------------------------------------------------------------------------
// Path-B correction shaders as a post-scale step (m_pPSCorrection, Source/DX11VideoProcessor.cpp:3354-3357):
// ps_convert_pq_to_sdr / ps_fixconvert_pq_to_sdr / ps_fixconvert_hlg_to_sdr / ps_fix_bt2020 / ps_convert_hlg_to_pq / ps_fix_ycgco
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_correction(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out, const __grid_constant__ ColorParams P,
             const __grid_constant__ CorrectionParams K, int x0, int y0, int x1, int y1)
{
    const int x = x0 + blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= x1 || y >= y1) return;
    f3 c = surf_load(in, x, y);
    const int mode = K.mode;
    if (mode == CORR_FIX_YCGCO) { surf_store(out, x, y, mat3_apply(K.fixycgco, c)); return; }
    if (mode == CORR_FIXCONVERT_PQ_TO_SDR || mode == CORR_FIXCONVERT_HLG_TO_SDR || mode == CORR_FIX_BT2020) c = mat3_apply(K.fix2020, c);
    if (mode == CORR_FIXCONVERT_HLG_TO_SDR || mode == CORR_CONVERT_HLG_TO_PQ) {                  // HLG to PQ
        f3 l; l.x = inverse_hlg(P, sat(c.x)); l.y = inverse_hlg(P, sat(c.y)); l.z = inverse_hlg(P, sat(c.z));
        const float dt = fma_rn(0.0593f, l.z, fma_rn(0.6780f, l.y, __fmul_rn(0.2627f, l.x)));
        const float g  = powr(__fmul_rn(2000.0f, dt), 0.2f);
        c.x = linear_to_st2084(__fmul_rn(l.x, g), P.r1000); c.y = linear_to_st2084(__fmul_rn(l.y, g), P.r1000); c.z = linear_to_st2084(__fmul_rn(l.z, g), P.r1000);
        if (mode == CORR_CONVERT_HLG_TO_PQ) { surf_store(out, x, y, c); return; }
    }
    if (mode == CORR_FIX_BT2020) {
        c.x = powr(sat(c.x), 2.2f); c.y = powr(sat(c.y), 2.2f); c.z = powr(sat(c.z), 2.2f);
    } else {                                                                                      // PQ to linear, Hable
        c.x = __fmul_rn(hable(P, st2084_to_linear(P, sat(c.x), P.lum_scale)), P.hable_rdiv);
        c.y = __fmul_rn(hable(P, st2084_to_linear(P, sat(c.y), P.lum_scale)), P.hable_rdiv);
        c.z = __fmul_rn(hable(P, st2084_to_linear(P, sat(c.z), P.lum_scale)), P.hable_rdiv);
    }
    c = mat3_apply(K.gamut, c);
    c.x = powr(sat(c.x), P.inv_gamma22); c.y = powr(sat(c.y), P.inv_gamma22); c.z = powr(sat(c.z), P.inv_gamma22);
    surf_store(out, x, y, c);
}

__global__ void __launch_bounds__(256)
k_tonemap(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out, const __grid_constant__ ColorParams P,
          const __grid_constant__ TonemapParams T, int x0, int y0, int x1, int y1)
{
    const int x = x0 + blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= x1 || y >= y1) return;
    surf_store(out, x, y, tm::pixel(P, T, surf_load(in, x, y)));
}

__global__ void __launch_bounds__(256)
k_to_bgrx(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out.w || y >= out.h) return;
    surf_store(out, x, y, surf_load(in, x, y));
}

// op: 0 lg2, 1 ex2, 2 pow, 3 unorm16, 4 unorm10, 5 unorm8
__global__ void k_math_probe(int op, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = x[i];
    float r;
    switch (op) {
    case 0: r = lg2(a); break;
    case 1: r = ex2(a); break;
    case 2: r = powr(a, y[i]); break;
    case 3: r = unorm16((uint32_t)a); break;
    case 4: r = unorm10((uint32_t)a); break;
    case 5: r = unorm8((uint32_t)a); break;
    case 6: r = sin_spec(a); break;
    // two-lane forms (mpcvr_math2.cuh): lane x = this element, lane y = its successor; lane x is returned
    case 10: r = lg2_2(make_float2(a, x[(i + 1) % n])).x; break;
    case 11: r = ex2_2(make_float2(a, x[(i + 1) % n])).x; break;
    case 12: r = powr_2(make_float2(a, x[(i + 1) % n]), y[i]).x; break;
    case 13: r = powr01_ylt1_2(make_float2(a, x[(i + 1) % n]), y[i]).x; break;
    case 14: r = powr01_ygt1_2(make_float2(a, x[(i + 1) % n]), y[i]).x; break;
    case 15: r = div_rn_safe_2(make_float2(a, x[(i + 1) % n]), make_float2(y[i], y[(i + 1) % n])).x; break;
    case 16: r = div_rn_safe(a, y[i]); break;
    case 17: r = __fdiv_rn(a, y[i]); break;
    case 18: r = unorm_roundtrip2<1023>(make_float2(a, x[(i + 1) % n])).x; break;
    case 19: r = rt_roundtrip(MPCVR_SURF_RGB10A2, a); break;
    case 20: r = lg2_2(make_float2(x[(i + 1) % n], a)).y; break;
    case 21: r = powr01_ygt1_2(make_float2(x[(i + 1) % n], a), y[i]).y; break;
    default: r = 0.f; break;
    }
    out[i] = r;
}

// ---------------------------------------------------------------------------------------------
static inline dim3 Grid2D(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

static std::atomic<int> g_resizeFour{1};             // debug switch: 0 = k_resize (one output per thread), 2 = k_resize4, 1 = k_resize_p where it applies
void SetResizeFour(int on) { g_resizeFour = on; }
static std::atomic<int> g_convertTwoLane{1};         // debug switch: 0 = one pixel per thread (scalar shader body), 2 = k_convert2 everywhere, 1 = k_convert_bp where it applies
void SetConvertTwoLane(int on) { g_convertTwoLane = on; }

cudaError_t LaunchConvertPass(const SrcDesc& src, const ColorParams& color, const SurfDesc& out, float4* dbg_f32, cudaStream_t st,
                              const DoviParams* dovi_dev)
{
    const dim3 b(32, 8);
    if (!dovi_dev && !dbg_f32 && g_convertTwoLane == 1 && ConvertSpanOk(src, out)) {      // 128-bit spans + two-lane body
        const dim3 bs(32, 4);
        if (src.layout == LAYOUT_PACKED422) k_convert_bp<1, true><<<dim3((out.w + 255) / 256, (out.h + 3) / 4), bs, 0, st>>>(src, color, out);
        else if (src.bytes == 1)            k_convert_bp<1, false><<<dim3((out.w + 511) / 512, (out.h + 3) / 4), bs, 0, st>>>(src, color, out);
        else                                k_convert_bp<2, false><<<dim3((out.w + 255) / 256, (out.h + 3) / 4), bs, 0, st>>>(src, color, out);
        return cudaGetLastError();
    }
    if (!dovi_dev && g_convertTwoLane) {
        k_convert2<<<Grid2D((out.w + 1) / 2, out.h, b), b, 0, st>>>(src, color, out, dbg_f32);
        return cudaGetLastError();
    }
    k_convert<<<Grid2D(out.w, out.h, b), b, 0, st>>>(src, color, out, dbg_f32, dovi_dev);
    return cudaGetLastError();
}

// picks the compile-time variant of k_resize_p for the common (taps, formats) combinations; false = not covered
template <int NT, bool CONV>
static bool LaunchResizeP(const SurfDesc& in, const SurfDesc& out, const AxisTable& T, int axis, const int32_t* other_idx, int other_dst0, int other_dst_len, cudaStream_t st)
{
    const dim3 b(32, 8);
    const int w = axis == 0 ? T.dst_len : (other_dst_len + 1) / 2, h = axis == 0 ? (other_dst_len + 1) / 2 : T.dst_len;
    const dim3 g = Grid2D(w, h, b);
#define MPCVR_RP(INF, OUTF) \
    if (in.fmt == INF && out.fmt == OUTF) { \
        if (axis == 0) k_resize_p<NT, INF, OUTF, 0, CONV><<<g, b, 0, st>>>(in, out, T, other_idx, other_dst0, other_dst_len); \
        else           k_resize_p<NT, INF, OUTF, 1, CONV><<<g, b, 0, st>>>(in, out, T, other_idx, other_dst0, other_dst_len); \
        return true; }
    // first pass: internal format -> m_TexResize (RGBA16F); second pass: m_TexResize -> internal format / back buffer
    MPCVR_RP(MPCVR_SURF_RGBA16F, MPCVR_SURF_RGBA16F)
    MPCVR_RP(MPCVR_SURF_RGB10A2, MPCVR_SURF_RGBA16F)
    MPCVR_RP(MPCVR_SURF_BGRA8,   MPCVR_SURF_RGBA16F)
    MPCVR_RP(MPCVR_SURF_RGBA16F, MPCVR_SURF_RGB10A2)
    MPCVR_RP(MPCVR_SURF_RGBA16F, MPCVR_SURF_BGRA8)
#undef MPCVR_RP
    return false;
}

cudaError_t LaunchResizePass(const SurfDesc& in, const SurfDesc& out, const AxisTable& table, int axis,
                             const int32_t* other_idx, int other_dst0, int other_dst_len, cudaStream_t st, int in_axis)
{
    const dim3 b(32, 8);
    const int w = axis == 0 ? table.dst_len : other_dst_len, h = axis == 0 ? other_dst_len : table.dst_len;
    const int mode = g_resizeFour;                                         // 1: default, 2: k_resize4 everywhere, 0: k_resize everywhere
    if ((in_axis < 0 || in_axis == axis) && mode == 1 && (table.kind == SCALER_INTERP || table.kind == SCALER_CONV)) {
        bool done = false;
        if (table.kind == SCALER_INTERP) {
            if (table.ntaps == 4)      done = LaunchResizeP<4, false>(in, out, table, axis, other_idx, other_dst0, other_dst_len, st);
            else if (table.ntaps == 6) done = LaunchResizeP<6, false>(in, out, table, axis, other_idx, other_dst0, other_dst_len, st);
        } else {
            if (table.ntaps == 9)      done = LaunchResizeP<9, true>(in, out, table, axis, other_idx, other_dst0, other_dst_len, st);   // 2:1 reductions with the 2-lobe filters
            else if (table.ntaps <= 32) done = LaunchResizeP<0, true>(in, out, table, axis, other_idx, other_dst0, other_dst_len, st);
        }
        if (done) return cudaGetLastError();
    }
    if ((in_axis < 0 || in_axis == axis) && mode) {                       // upright draws: four outputs per thread
        const int w4 = axis == 0 ? (w + 3) / 4 : w, h4 = axis == 0 ? h : (h + 3) / 4;
        k_resize4<<<Grid2D(w4, h4, b), b, 0, st>>>(in, out, table, axis, other_idx, other_dst0, other_dst_len);
        return cudaGetLastError();
    }
    k_resize<<<Grid2D(w, h, b), b, 0, st>>>(in, out, table, axis, other_idx, other_dst0, other_dst_len, in_axis < 0 ? axis : in_axis);
    return cudaGetLastError();
}

cudaError_t LaunchCopyPass(const SurfDesc& in, const SurfDesc& out, const int32_t* idx_x, const int32_t* idx_y,
                           int dst_x0, int dst_y0, int dst_w, int dst_h, cudaStream_t st, int transpose)
{
    const dim3 b(32, 8);
    k_copy<<<Grid2D(dst_w, dst_h, b), b, 0, st>>>(in, out, idx_x, idx_y, dst_x0, dst_y0, dst_w, dst_h, transpose);
    return cudaGetLastError();
}

// FinalPass with both surface formats fixed at compile time, two neighbouring pixels per thread (one 64/128-bit load, one 64-bit store,
// one 32-bit read of the two dither thresholds).  Same per-pixel operations as k_final.
template <int INF, int OUTF>
__global__ void __launch_bounds__(256)
k_final_p(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out, int x0, int y0, int x1, int y1, float Q)
{
    const int x = x0 + 2 * (blockIdx.x * blockDim.x + threadIdx.x), y = y0 + blockIdx.y * blockDim.y + threadIdx.y;      // x0 is even
    if (x >= x1 || y >= y1) return;
    const uint32_t dd = *reinterpret_cast<const uint32_t*>(&g_dither32[(y & 31) * 32 + (x & 31)]);                        // thresholds of x and x + 1
    const float d0 = (float)(dd & 0xFFFFu) * 0.0009765625f, d1 = (float)(dd >> 16) * 0.0009765625f;
    const uint8_t* irow = (const uint8_t*)in.ptr + (size_t)y * in.pitch;
    f3 p0, p1;
    if (INF == MPCVR_SURF_RGBA16F) { const uint4 v = __ldg((const uint4*)(irow + (size_t)x * 8)); p0 = texel_decode<INF>(v.x, v.y); p1 = texel_decode<INF>(v.z, v.w); }
    else                           { const uint2 v = __ldg((const uint2*)(irow + (size_t)x * 4)); p0 = texel_decode<INF>(v.x, 0u); p1 = texel_decode<INF>(v.y, 0u); }
    f3 o0, o1;
    o0.x = final_quant(p0.x, Q, d0); o0.y = final_quant(p0.y, Q, d0); o0.z = final_quant(p0.z, Q, d0);
    o1.x = final_quant(p1.x, Q, d1); o1.y = final_quant(p1.y, Q, d1); o1.z = final_quant(p1.z, Q, d1);
    uint8_t* orow = (uint8_t*)out.ptr + (size_t)y * out.pitch;
    auto pack = [](f3 v) -> uint32_t {
        return OUTF == MPCVR_SURF_RGB10A2 ? (qunorm(v.x, 1023.0f) | (qunorm(v.y, 1023.0f) << 10) | (qunorm(v.z, 1023.0f) << 20) | (3u << 30))
                                          : (qunorm(v.z, 255.0f) | (qunorm(v.y, 255.0f) << 8) | (qunorm(v.x, 255.0f) << 16) | (255u << 24)); };
    if (x + 1 < x1) *reinterpret_cast<uint2*>(orow + (size_t)x * 4) = make_uint2(pack(o0), pack(o1));
    else            *reinterpret_cast<uint32_t*>(orow + (size_t)x * 4) = pack(o0);
}

cudaError_t LaunchFinalPass(const SurfDesc& in, const SurfDesc& out, int x0, int y0, int x1, int y1, float Q, cudaStream_t st)
{
    if (x1 <= x0 || y1 <= y0) return cudaSuccess;
    const dim3 b(32, 8);
    // paired variant: even start column, both row bases 16-byte aligned, and the input row readable one texel past an odd right edge
    if (g_resizeFour == 1 && (x0 & 1) == 0 && !((uintptr_t)in.ptr & 15) && !(in.pitch & 15) && !((uintptr_t)out.ptr & 7) && !(out.pitch & 7) && ((x1 & 1) == 0 || x1 < in.w)) {
        const dim3 g = Grid2D((x1 - x0 + 1) / 2, y1 - y0, b);
#define MPCVR_FP(INF, OUTF) if (in.fmt == INF && out.fmt == OUTF) { k_final_p<INF, OUTF><<<g, b, 0, st>>>(in, out, x0, y0, x1, y1, Q); return cudaGetLastError(); }
        MPCVR_FP(MPCVR_SURF_RGBA16F, MPCVR_SURF_BGRA8) MPCVR_FP(MPCVR_SURF_RGBA16F, MPCVR_SURF_RGB10A2)
        MPCVR_FP(MPCVR_SURF_RGB10A2, MPCVR_SURF_BGRA8) MPCVR_FP(MPCVR_SURF_RGB10A2, MPCVR_SURF_RGB10A2)
        MPCVR_FP(MPCVR_SURF_BGRA8, MPCVR_SURF_BGRA8)
#undef MPCVR_FP
    }
    k_final<<<Grid2D(x1 - x0, y1 - y0, b), b, 0, st>>>(in, out, x0, y0, x1, y1, Q);
    return cudaGetLastError();
}

cudaError_t LaunchCorrectionPass(const SurfDesc& in, const SurfDesc& out, const ColorParams& P, const CorrectionParams& K,
                                 int x0, int y0, int x1, int y1, cudaStream_t st)
{
    if (x1 <= x0 || y1 <= y0) return cudaSuccess;
    const dim3 b(32, 8);
    k_correction<<<Grid2D(x1 - x0, y1 - y0, b), b, 0, st>>>(in, out, P, K, x0, y0, x1, y1);
    return cudaGetLastError();
}

cudaError_t LaunchJinc2Pass(const SurfDesc& in, const SurfDesc& out, int in_l, int in_t, int in_w, int in_h,
                            int out_l, int out_t, int out_w, int out_h, cudaStream_t st, int orient)
{
    if (out_w <= 0 || out_h <= 0) return cudaSuccess;
    const dim3 b(32, 8);
    k_jinc2<<<Grid2D(out_w, out_h, b), b, 0, st>>>(in, out, in_l, in_t, in_w, in_h, out_l, out_t, out_w, out_h, orient);
    return cudaGetLastError();
}

cudaError_t LaunchTonemapPass(const SurfDesc& in, const SurfDesc& out, const ColorParams& P, const TonemapParams& T,
                              int x0, int y0, int x1, int y1, cudaStream_t st)
{
    if (x1 <= x0 || y1 <= y0) return cudaSuccess;
    const dim3 b(32, 8);
    k_tonemap<<<Grid2D(x1 - x0, y1 - y0, b), b, 0, st>>>(in, out, P, T, x0, y0, x1, y1);
    return cudaGetLastError();
}

cudaError_t LaunchClear(const SurfDesc& s, cudaStream_t st)
{
    return cudaMemset2DAsync(s.ptr, s.pitch, 0, (size_t)s.w * ((s.fmt == MPCVR_SURF_RGBA16F || s.fmt == MPCVR_SURF_RGBA16) ? 8 : (s.fmt == MPCVR_SURF_RGB8 ? 3 : 4)), s.h, st);
}

cudaError_t LaunchToBGRX(const SurfDesc& in, const SurfDesc& out, cudaStream_t st)
{
    const dim3 b(32, 8);
    k_to_bgrx<<<Grid2D(out.w, out.h, b), b, 0, st>>>(in, out);
    return cudaGetLastError();
}

cudaError_t LaunchMathProbe(int op, const float* x, const float* y, float* out, long long n, cudaStream_t st)
{
    const int b = 256;
    k_math_probe<<<(unsigned)((n + b - 1) / b), b, 0, st>>>(op, x, y, out, n);
    return cudaGetLastError();
}

cudaError_t UploadDitherTable() { return cudaSuccess; }   // statically initialised device table

}  // namespace b200vr
