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
