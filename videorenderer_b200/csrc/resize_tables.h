// resize_tables.h -- host-side evaluation of the resize shaders' per-output-pixel constants.
//
// In the reference every output pixel recomputes its tap window and weights from the
// interpolated texture coordinate (Shaders/d3d11/ps_interpolation_{spline4,lanczos2,lanczos3}.hlsl,
// Shaders/d3d11/ps_convolution.hlsl + Shaders/resize/convolution_filters.hlsl; D3D9 twins under
// Shaders/d3d9/).  They depend only on the output index along the axis, so here they are computed
// once per (axis, rects, filter) on the host and the kernels consume tables.
#pragma once
#include <vector>
#include <cstdint>

namespace b200vr {

struct HostAxisTable {
    int kind = 0;              // ScalerKind
    int filter = 0;            // MPCVR_UPSCALE_* or MPCVR_DOWNSCALE_*
    int ntaps = 0;             // max taps over all outputs
    bool dup_q1 = false;
    int src0 = 0, src_len = 0, dst0 = 0, dst_len = 0;
    std::vector<int32_t> base;     // [dst_len]
    std::vector<float>   w;        // [dst_len * ntaps]
    std::vector<float>   wsum;     // [dst_len] (convolution)
};

// scaler choice per axis, ResizeShaderPass Source/DX11VideoProcessor.cpp:3106-3127
int ChooseScaler(int src_len, int dst_len, bool interpolateAt50pct);

// d3d9_variant: restate the Shaders/d3d9 twins instead
// reverse: the interpolated coordinate runs from the far edge down (rotated / flipped quads, FillVertices :130-179):
//          c = (src0 + src_len) - (i + .5) * scale.
// qscale : the `scale` constant the convolution shader is handed (TextureResizeShader :351-354 divides same-named axes of
//          srcRect and dstRect, which is not the geometric ratio once the quad is turned by 90 degrees); <= 0: geometric.
bool BuildAxisTable(int kind, int filter, bool d3d9_variant, int src0, int src_len, int dst0, int dst_len, HostAxisTable& out,
                    bool reverse = false, float qscale = 0.0f);

}  // namespace b200vr
