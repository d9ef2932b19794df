// tests/host_shim/pixel_host.cpp -- TEST INFRASTRUCTURE ONLY: the product's device pixel functions, compiled for the host (see cuda_host_shim.h)
#include "cuda_host_shim.h"
#include "pixel_ops.cuh"
using namespace b200vr;
extern "C" {
int host_sizeof_color_params(void) { return (int)sizeof(ColorParams); }
int host_sizeof_dovi_params(void) { return (int)sizeof(DoviParams); }
// convert_color<-1> (the body of k_convert, kernels_multipass.cu) on n yuv triplets; D may be null
int host_convert_color(const ColorParams* P, const DoviParams* D, const float* yuv, float* out, int n)
{
    for (int i = 0; i < n; i++) {
        f3 v; v.x = yuv[3 * i]; v.y = yuv[3 * i + 1]; v.z = yuv[3 * i + 2];
        const f3 r = convert_color<-1>(*P, v, D);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
    return 0;
}
}

#include "tonemap_ops.cuh"
extern "C" {
int host_sizeof_tonemap_params(void) { return (int)sizeof(TonemapParams); }
// tm::pixel (the body of k_tonemap) on n PQ texels
int host_tonemap_pixel(const ColorParams* P, const TonemapParams* T, const float* in, float* out, int n)
{
    for (int i = 0; i < n; i++) {
        f3 v; v.x = in[3 * i]; v.y = in[3 * i + 1]; v.z = in[3 * i + 2];
        const f3 r = tm::pixel(*P, *T, v);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
    return 0;
}
}

extern "C" {
int host_sizeof_src_desc(void) { return (int)sizeof(SrcDesc); }
// the body of k_convert (kernels_multipass.cu) for every pixel of the source rect: fetch_yuv + convert_color<-1>; out = rect_h * rect_w * 3 floats
int host_convert_frame(const SrcDesc* S, const ColorParams* P, const DoviParams* D, float* out)
{
    for (int y = 0; y < S->rect_h; y++)
        for (int x = 0; x < S->rect_w; x++) {
            const f3 yuv = fetch_yuv(*S, S->rect_l + x, S->rect_t + y);
            const f3 r = convert_color<-1>(*P, yuv, D);
            float* o = out + ((size_t)y * S->rect_w + x) * 3;
            o[0] = r.x; o[1] = r.y; o[2] = r.z;
        }
    return 0;
}
}

extern "C" {
// surf_store then surf_load of n RGB triplets through a one-row surface of format fmt (MPCVR_SURF_*): the render-target write and the read
// the next pass performs
int host_surface_roundtrip(int fmt, const float* in, float* out, int n)
{
    const int bpp = (fmt == MPCVR_SURF_RGBA16F || fmt == MPCVR_SURF_RGBA16) ? 8 : (fmt == MPCVR_SURF_RGB8 ? 3 : 4);
    uint8_t* row = new uint8_t[(size_t)n * bpp + 16];
    SurfDesc s; memset(&s, 0, sizeof(s));
    s.ptr = row; s.pitch = n * bpp; s.w = n; s.h = 1; s.fmt = fmt;
    for (int i = 0; i < n; i++) { f3 v; v.x = in[3 * i]; v.y = in[3 * i + 1]; v.z = in[3 * i + 2]; surf_store(s, i, 0, v); }
    for (int i = 0; i < n; i++) { const f3 v = surf_load(s, i, 0); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
    delete[] row;
    return 0;
}
int host_final_quant(const float* p, const float* d, float Q, float* out, int n)
{
    for (int i = 0; i < n; i++) out[i] = final_quant(p[i], Q, d[i]);
    return 0;
}
}

extern "C" {
// the scalar math of mpcvr_math.cuh: op 0 lg2, 1 ex2, 2 powr(x, y), 3 sin_spec, 4 ex2p(x, y), 5 div_rn_safe(x, y), 6 unorm8 / 7 unorm10 / 8 unorm16 of (uint)x
int host_math(int op, const float* x, const float* y, float* out, long long n)
{
    for (long long i = 0; i < n; i++) {
        switch (op) {
        case 0: out[i] = lg2(x[i]); break;
        case 1: out[i] = ex2(x[i]); break;
        case 2: out[i] = powr(x[i], y[i]); break;
        case 3: out[i] = sin_spec(x[i]); break;
        case 4: out[i] = ex2p(x[i], y[i]); break;
        case 5: out[i] = div_rn_safe(x[i], y[i]); break;
        case 6: out[i] = unorm8((uint32_t)x[i]); break;
        case 7: out[i] = unorm10((uint32_t)x[i]); break;
        default: out[i] = unorm16((uint32_t)x[i]); break;
        }
    }
    return 0;
}
}
