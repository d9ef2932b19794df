// kernels_multipass.cu -- one CUDA kernel per pass of the reference's chain
// (CDX11VideoProcessor::Process, Source/DX11VideoProcessor.cpp:3285-3424), intermediates in HBM in the
// reference's own surface formats.  This family handles every plan; the fused kernel
// (kernel_fused.cu) covers the plans that matter for throughput.  sm_100a, -fmad=false.
#include <cuda_runtime.h>
#include "kernels.h"
#include "pixel_ops.cuh"
#include "mpcvr_math2.cuh"
#include "dither32.h"

namespace b200vr {

__device__ const uint16_t g_dither32[1024] = { MPCVR_DITHER32_VALUES };

// ---------------------------------------------------------------------------------------------
// ConvertColorPass: one thread per pixel of m_TexConvertOutput (= m_srcRect size)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_convert(const __grid_constant__ SrcDesc S, const __grid_constant__ ColorParams P, const __grid_constant__ SurfDesc out,
          float4* __restrict__ dbg)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out.w || y >= out.h) return;
    const f3 yuv = fetch_yuv(S, S.rect_l + x, S.rect_t + y);
    const f3 rgb = convert_color<-1>(P, yuv);
    if (dbg) dbg[(size_t)y * out.w + x] = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
    surf_store(out, x, y, rgb);
}

// ---------------------------------------------------------------------------------------------
// TextureResizeShader along one axis: one thread per output pixel, taps from the host table
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_resize(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out, const __grid_constant__ AxisTable T,
         const int axis, const int32_t* __restrict__ other_idx, const int other_dst0, const int other_dst_len)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
    // (a, b) index the output rectangle: a along X, b along Y
    const int i  = axis == 0 ? a : b;      // index along the filtered axis
    const int j  = axis == 0 ? b : a;      // index along the other axis
    if (i >= T.dst_len || j >= other_dst_len) return;
    const int ox = axis == 0 ? T.dst0 + i : other_dst0 + j;
    const int oy = axis == 0 ? other_dst0 + j : T.dst0 + i;
    if (ox < 0 || oy < 0 || ox >= out.w || oy >= out.h) return;
    const int lim   = axis == 0 ? in.w : in.h;
    const int fixed = clampi(__ldg(other_idx + j), 0, (axis == 0 ? in.h : in.w) - 1);
    const int base  = __ldg(T.base + i);
    const float* w  = T.w + (size_t)i * T.ntaps;
    f3 acc = {0.f, 0.f, 0.f};
    for (int k = 0; k < T.ntaps; k++) {
        int idx = base + k;
        if (T.dup_q1 && k == 1) idx = base;                 // ps_interpolation_lanczos3.hlsl:33-34
        idx = clampi(idx, 0, lim - 1);
        const f3 p = axis == 0 ? surf_load(in, idx, fixed) : surf_load(in, fixed, idx);
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

// ---------------------------------------------------------------------------------------------
// TextureCopyRect + ps_simple: point-sampled copy / nearest-neighbour stretch
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_copy(const __grid_constant__ SurfDesc in, const __grid_constant__ SurfDesc out,
       const int32_t* __restrict__ idx_x, const int32_t* __restrict__ idx_y, int dx0, int dy0, int dw, int dh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int ox = dx0 + x, oy = dy0 + y;
    if (ox < 0 || oy < 0 || ox >= out.w || oy >= out.h) return;
    const int ix = clampi(__ldg(idx_x + x), 0, in.w - 1), iy = clampi(__ldg(idx_y + y), 0, in.h - 1);
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
    // two-lane forms (mpcvr_math2.cuh): lane x = this element, lane y = its successor; lane x is returned
    case 10: r = lg2_2(make_float2(a, x[(i + 1) % n])).x; break;
    case 11: r = ex2_2(make_float2(a, x[(i + 1) % n])).x; break;
    case 12: r = powr_2(make_float2(a, x[(i + 1) % n]), y[i]).x; break;
    case 13: r = powr01_ylt1_2(make_float2(a, x[(i + 1) % n]), y[i]).x; break;
    case 14: r = powr01_ygt1_2(make_float2(a, x[(i + 1) % n]), y[i]).x; break;
    case 15: r = div_rn_safe_2(make_float2(a, x[(i + 1) % n]), make_float2(y[i], y[(i + 1) % n])).x; break;
    case 16: r = div_rn_safe(a, y[i]); break;
    case 17: r = __fdiv_rn(a, y[i]); break;
    case 18: r = unorm_roundtrip2(make_float2(a, x[(i + 1) % n]), 1023.0f, 1.0f / 1023.0f).x; break;
    case 19: r = rt_roundtrip(MPCVR_SURF_RGB10A2, a); break;
    case 20: r = lg2_2(make_float2(x[(i + 1) % n], a)).y; break;
    case 21: r = powr01_ygt1_2(make_float2(x[(i + 1) % n], a), y[i]).y; break;
    default: r = 0.f; break;
    }
    out[i] = r;
}

// ---------------------------------------------------------------------------------------------
static inline dim3 Grid2D(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

cudaError_t LaunchConvertPass(const SrcDesc& src, const ColorParams& color, const SurfDesc& out, float4* dbg_f32, cudaStream_t st)
{
    const dim3 b(32, 8);
    k_convert<<<Grid2D(out.w, out.h, b), b, 0, st>>>(src, color, out, dbg_f32);
    return cudaGetLastError();
}

cudaError_t LaunchResizePass(const SurfDesc& in, const SurfDesc& out, const AxisTable& table, int axis,
                             const int32_t* other_idx, int other_dst0, int other_dst_len, cudaStream_t st)
{
    const dim3 b(32, 8);
    const int w = axis == 0 ? table.dst_len : other_dst_len, h = axis == 0 ? other_dst_len : table.dst_len;
    k_resize<<<Grid2D(w, h, b), b, 0, st>>>(in, out, table, axis, other_idx, other_dst0, other_dst_len);
    return cudaGetLastError();
}

cudaError_t LaunchCopyPass(const SurfDesc& in, const SurfDesc& out, const int32_t* idx_x, const int32_t* idx_y,
                           int dst_x0, int dst_y0, int dst_w, int dst_h, cudaStream_t st)
{
    const dim3 b(32, 8);
    k_copy<<<Grid2D(dst_w, dst_h, b), b, 0, st>>>(in, out, idx_x, idx_y, dst_x0, dst_y0, dst_w, dst_h);
    return cudaGetLastError();
}

cudaError_t LaunchFinalPass(const SurfDesc& in, const SurfDesc& out, int x0, int y0, int x1, int y1, float Q, cudaStream_t st)
{
    if (x1 <= x0 || y1 <= y0) return cudaSuccess;
    const dim3 b(32, 8);
    k_final<<<Grid2D(x1 - x0, y1 - y0, b), b, 0, st>>>(in, out, x0, y0, x1, y1, Q);
    return cudaGetLastError();
}

cudaError_t LaunchClear(const SurfDesc& s, cudaStream_t st)
{
    return cudaMemset2DAsync(s.ptr, s.pitch, 0, (size_t)s.w * (s.fmt == MPCVR_SURF_RGBA16F ? 8 : 4), s.h, st);
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
