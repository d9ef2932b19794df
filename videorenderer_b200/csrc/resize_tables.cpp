// resize_tables.cpp -- see resize_tables.h.
// fp32 throughout, operations in the order the HLSL text writes them (host code is compiled
// with -ffp-contract=off), sin/cos from libm.
#include <cmath>
#include "resize_tables.h"
#include "frame_plan.h"
#include "../../include/mpcvr_b200.h"

namespace b200vr {

int ChooseScaler(int src_len, int dst_len, bool interpolateAt50pct)
{
    const int k = interpolateAt50pct ? 2 : 1;
    if (src_len == dst_len) return SCALER_NONE;
    return (src_len > k * dst_len) ? SCALER_CONV : SCALER_INTERP;
}

namespace {

const float kPi = 3.14159265358979323846f;      // HLSL: #define PI acos(-1.)

// Shaders/d3d11/ps_interpolation_spline4.hlsl:48-54
void Spline4Weights(int method, float t, float w[4])
{
    const float t2 = t * t, t3 = t * t2;
    if (method == MPCVR_UPSCALE_Mitchell) {
        const float a[4] = {1.f, 16.f, 1.f, 0.f}, b[4] = {-.5f, 0.f, .5f, 0.f}, c[4] = {5.f, -12.f, 9.f, -2.f}, d[4] = {-7.f, 21.f, -21.f, 7.f};
        for (int j = 0; j < 4; j++) w[j] = a[j] / 18.f + b[j] * t + c[j] / 6.f * t2 + d[j] / 18.f * t3;
    } else {
        const float b[4] = {-.5f, 0.f, .5f, 0.f}, c[4] = {1.f, -2.5f, 2.f, -.5f}, d[4] = {-.5f, 1.5f, -1.5f, .5f};
        for (int j = 0; j < 4; j++) w[j] = b[j] * t + c[j] * t2 + d[j] * t3;
        w[1] += 1.f;
    }
}

// Shaders/d3d11/ps_interpolation_lanczos2.hlsl:46-51
void Lanczos2Weights(float t, float w[4])
{
    if (t == 0.f) { w[0] = 0; w[1] = 1; w[2] = 0; w[3] = 0; return; }      // `if (t)` / return Q1, :31, :56
    const float ws[4] = {1.f + t, 0.f + t, 1.f - t, 2.f - t};
    for (int j = 0; j < 4; j++)
        w[j] = sinf(ws[j] * kPi) * sinf(ws[j] * kPi * .5f) / (ws[j] * ws[j] * kPi * kPi * .5f);
    const float wc = 1.f - (((w[0] + w[1]) + w[2]) + w[3]);
    w[1] += wc * (1.f - t);
    w[2] += wc * t;
}

// Shaders/d3d11/ps_interpolation_lanczos3.hlsl:50-59
void Lanczos3Weights(float t, float w[6])
{
    if (t == 0.f) { for (int j = 0; j < 6; j++) w[j] = (j == 2) ? 1.f : 0.f; return; }   // :31, :64
    float w0[3], w1[3];
    for (int j = 0; j < 3; j++) {
        const float a0 = (float)(2 - j) * kPi + t * kPi, a1 = (float)(1 + j) * kPi - t * kPi;
        const float a0s = a0 * .5f, a1s = a1 * .5f;
        w0[j] = sinf(a0) * sinf(a0s) / (a0 * a0s);
        w1[j] = sinf(a1) * sinf(a1s) / (a1 * a1s);
    }
    const float wc = 1.f - (((w0[0] + w1[0]) + (w0[1] + w1[1])) + (w0[2] + w1[2]));
    w0[2] += wc * (1.f - t);
    w1[0] += wc * t;
    w[0] = w0[0]; w[1] = w0[1]; w[2] = w0[2]; w[3] = w1[0]; w[4] = w1[1]; w[5] = w1[2];
}

// Shaders/resize/convolution_filters.hlsl:7-86
float ConvFilter(int filter, float x)
{
    switch (filter) {
    case MPCVR_DOWNSCALE_Box:
        return (x >= -0.5f && x < 0.5f) ? 1.f : 0.f;
    case MPCVR_DOWNSCALE_Bilinear:
        if (x < 0.f) x = -x;
        return x < 1.f ? 1.f - x : 0.f;
    case MPCVR_DOWNSCALE_Hamming:
        if (x < 0.f) x = -x;
        if (x == 0.f) return 1.f;
        if (x >= 1.f) return 0.f;
        x *= kPi;
        return sinf(x) / x * (0.54f + 0.46f * cosf(x));
    case MPCVR_DOWNSCALE_Bicubic:
    case MPCVR_DOWNSCALE_BicubicSharp: {
        const float A = (filter == MPCVR_DOWNSCALE_Bicubic) ? -0.5f : -1.5f;     // compile_shaders.cmd:97-100
        if (x < 0.f) x = -x;
        if (x < 1.f) return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
        if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * A;
        return 0.f; }
    default: {
        if (!(-3.f <= x && x < 3.f)) return 0.f;
        float a = x, b = x / 3.f, sa, sb;
        if (a == 0.f) sa = 1.f; else { a *= kPi; sa = sinf(a) / a; }
        if (b == 0.f) sb = 1.f; else { b *= kPi; sb = sinf(b) / b; }
        return sa * sb; }
    }
}

float ConvSupport(int filter)
{
    switch (filter) {
    case MPCVR_DOWNSCALE_Box: return 0.5f;
    case MPCVR_DOWNSCALE_Bilinear: case MPCVR_DOWNSCALE_Hamming: return 1.f;
    case MPCVR_DOWNSCALE_Bicubic: case MPCVR_DOWNSCALE_BicubicSharp: return 2.f;
    default: return 3.f;
    }
}

}  // namespace

bool BuildAxisTable(int kind, int filter, bool d3d9_variant, int src0, int src_len, int dst0, int dst_len, HostAxisTable& T,
                    bool reverse, float qscale)
{
    T = HostAxisTable();
    T.kind = kind; T.filter = filter; T.src0 = src0; T.src_len = src_len; T.dst0 = dst0; T.dst_len = dst_len;
    if (dst_len <= 0 || src_len <= 0) return false;
    // constants of TextureResizeShader, Source/DX11VideoProcessor.cpp:351-354: scale = (float)src/dst
    const float scale = (float)src_len / (float)dst_len;
    T.base.resize(dst_len);
    const float far_edge = (float)(src0 + src_len);
    auto coord = [&](int i) { return reverse ? far_edge - ((float)i + 0.5f) * scale : ((float)i + 0.5f) * scale + (float)src0; };

    if (kind == SCALER_NONE || kind == SCALER_NEAREST) {
        // point sample of the interpolated coordinate: texel floor(src0 + (i+.5)*scale)
        T.ntaps = 1; T.w.assign(dst_len, 1.f);
        for (int i = 0; i < dst_len; i++) T.base[i] = (int)floorf(coord(i));
        return true;
    }
    if (kind == SCALER_INTERP) {
        int nt;
        switch (filter) {
        case MPCVR_UPSCALE_Mitchell: case MPCVR_UPSCALE_CatmullRom: case MPCVR_UPSCALE_Lanczos2: nt = 4; break;
        case MPCVR_UPSCALE_Lanczos3: nt = 6; break;
        default: return false;                                   // Jinc2: SURVEY 8(f)-4
        }
        T.ntaps = nt; T.w.resize((size_t)dst_len * nt);
        T.dup_q1 = (filter == MPCVR_UPSCALE_Lanczos3) && !d3d9_variant;
        for (int i = 0; i < dst_len; i++) {
            const float c = coord(i);                                  // interpolated Tex*wh
            const float pos = c - 0.5f;
            const float fl = floorf(pos);
            const float t = pos - fl;
            float* w = &T.w[(size_t)i * nt];
            if (nt == 6) { Lanczos3Weights(t, w); T.base[i] = (int)fl - 2; }
            else {
                if (filter == MPCVR_UPSCALE_Lanczos2) Lanczos2Weights(t, w); else Spline4Weights(filter, t, w);
                T.base[i] = (int)fl - 1;
            }
        }
        return true;
    }
    // convolution: Shaders/d3d11/ps_convolution.hlsl:23-49 (pos = Tex*wh + 0.5), D3D9 twin convolution.hlsl:17 (pos = c)
    const float cscale = qscale > 0.0f ? qscale : scale;            // the shader's `scale` constant
    const float support = ConvSupport(filter) * cscale;
    const float ss = 1.0f / cscale;
    std::vector<int> lows(dst_len), counts(dst_len);
    int nt = 1;
    for (int i = 0; i < dst_len; i++) {
        const float c = coord(i);
        const float pos = d3d9_variant ? c : c + 0.5f;
        lows[i] = (int)floorf(pos - support);
        counts[i] = (int)ceilf(pos + support) - lows[i];
        if (counts[i] > nt) nt = counts[i];
    }
    if (nt > 1024) return false;                                 // 340x reduction with the widest filter; beyond that E_NOTIMPL
    T.ntaps = nt; T.w.assign((size_t)dst_len * nt, 0.f); T.wsum.resize(dst_len);
    for (int i = 0; i < dst_len; i++) {
        const float c = coord(i);
        const float pos = d3d9_variant ? c : c + 0.5f;
        float ww = 0.f;
        for (int k = 0; k < counts[i]; k++) {
            const float w = ConvFilter(filter, ((float)(lows[i] + k) - pos + 0.5f) * ss);
            T.w[(size_t)i * nt + k] = w;
            ww += w;
        }
        T.base[i] = lows[i]; T.wsum[i] = ww;
    }
    return true;
}

}  // namespace b200vr
