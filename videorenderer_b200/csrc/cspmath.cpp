// cspmath.cpp -- see cspmath.h.
//
// The arithmetic (operation order, float vs double intermediates) is dictated by the reference
// because the 12 + 9 constants must come out bit-identical (KATs: SURVEY 8 a3/a5, checked
// against the reference's own csputils.cpp in tests/test_oracle_pins.py and tests/test_host_logic.py); the structure is ours.
#include <cmath>
#include "cspmath.h"

namespace b200vr {
namespace {

struct Mat3 {
    float v[3][3];
    float* operator[](int i) { return v[i]; }
    const float* operator[](int i) const { return v[i]; }
};

// adjugate / determinant inverse in float, cofactors first then one reciprocal (csputils.cpp:14-40)
Mat3 Inverse(const Mat3& a)
{
    const float a00 = a[0][0], a01 = a[0][1], a02 = a[0][2];
    const float a10 = a[1][0], a11 = a[1][1], a12 = a[1][2];
    const float a20 = a[2][0], a21 = a[2][1], a22 = a[2][2];
    Mat3 r;
    r[0][0] =  (a11 * a22 - a21 * a12);
    r[0][1] = -(a01 * a22 - a21 * a02);
    r[0][2] =  (a01 * a12 - a11 * a02);
    r[1][0] = -(a10 * a22 - a20 * a12);
    r[1][1] =  (a00 * a22 - a20 * a02);
    r[1][2] = -(a00 * a12 - a10 * a02);
    r[2][0] =  (a10 * a21 - a20 * a11);
    r[2][1] = -(a00 * a21 - a20 * a01);
    r[2][2] =  (a00 * a11 - a10 * a01);
    float det = a00 * r[0][0] + a10 * r[0][1] + a20 * r[0][2];
    det = 1.0f / det;
    for (auto& row : r.v) for (float& x : row) x *= det;
    return r;
}

Mat3 Mul(const Mat3& a, const Mat3& b)      // csputils.cpp:43-54
{
    Mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            r[j][i] = a[j][0] * b[0][i] + a[j][1] * b[1][i] + a[j][2] * b[2][i];
    return r;
}

struct XY { float x, y; };
struct Primaries { XY r, g, b, w; };

Primaries GetPrimaries(Chromaticities c)    // csputils.cpp:57-205 (the sets set_colorspace can name)
{
    const XY d65 = {0.31271f, 0.32902f}, illC = {0.31006f, 0.31616f}, dci = {0.31400f, 0.35100f};
    switch (c) {
    case CHR_BT470M:    return {{0.670f, 0.330f}, {0.210f, 0.710f}, {0.140f, 0.080f}, illC};
    case CHR_BT601_525: return {{0.630f, 0.340f}, {0.310f, 0.595f}, {0.155f, 0.070f}, d65};
    case CHR_BT601_625: return {{0.640f, 0.330f}, {0.290f, 0.600f}, {0.150f, 0.060f}, d65};
    case CHR_BT2020:    return {{0.708f, 0.292f}, {0.170f, 0.797f}, {0.131f, 0.046f}, d65};
    case CHR_DCI_P3:    return {{0.680f, 0.320f}, {0.265f, 0.690f}, {0.150f, 0.060f}, dci};
    case CHR_BT709:
    default:            return {{0.640f, 0.330f}, {0.300f, 0.600f}, {0.150f, 0.060f}, d65};
    }
}

// RGB -> XYZ from chromaticities (Lindbloom), csputils.cpp:230-263
Mat3 RgbToXyz(const Primaries& p)
{
    const XY prim[3] = {p.r, p.g, p.b};
    float X[4], Z[4];
    for (int i = 0; i < 3; i++) { X[i] = prim[i].x / prim[i].y; Z[i] = (1 - prim[i].x - prim[i].y) / prim[i].y; }
    X[3] = p.w.x / p.w.y; Z[3] = (1 - p.w.x - p.w.y) / p.w.y;
    Mat3 m;
    for (int i = 0; i < 3; i++) { m[0][i] = X[i]; m[1][i] = 1; m[2][i] = Z[i]; }
    m = Inverse(m);
    float S[3];
    for (int i = 0; i < 3; i++) S[i] = m[i][0] * X[3] + m[i][1] * 1 + m[i][2] * Z[3];
    for (int i = 0; i < 3; i++) { m[0][i] = S[i] * X[i]; m[1][i] = S[i] * 1; m[2][i] = S[i] * Z[i]; }
    return m;
}

enum MatrixKind { MK_BT601, MK_BT709, MK_SMPTE240M, MK_BT2020NC, MK_YCGCO, MK_RGB };
enum LevelsKind { LV_TV, LV_PC, LV_ANYFULL };

void LumaWeightsToMatrix(float m[3][3], float lr, float lg, float lb)     // csputils.cpp:380-389
{
    m[0][0] = 1; m[0][1] = 0;                     m[0][2] = 2 * (1 - lr);
    m[1][0] = 1; m[1][1] = -2 * (1 - lb) * lb / lg; m[1][2] = -2 * (1 - lr) * lr / lg;
    m[2][0] = 1; m[2][1] = 2 * (1 - lb);          m[2][2] = 0;
}

}  // namespace

void BuildGamutMatrix(Chromaticities from, Chromaticities to, float out[3][3])
{
    const Mat3 in = RgbToXyz(GetPrimaries(from));
    const Mat3 m = Mul(Inverse(RgbToXyz(GetPrimaries(to))), in);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[i][j] = m[i][j];
}

bool BuildColorTransform(const ExtendedFormat& ex, const FmtConvParams& fmt, const ProcAmp& amp, ColorTransform& out)
{
    // set_colorspace, Source/Helper.cpp:949-1004
    MatrixKind mk; LevelsKind lv;
    if (ex.value == 0) { mk = MK_RGB; lv = LV_PC; }
    else {
        switch (ex.NominalRange()) { case Range_0_255: lv = LV_PC; break; default: lv = LV_TV; }   // AUTO -> TV (csputils.cpp:398-399)
        switch (ex.TransferMatrix()) {
        case Matrix_BT709:     mk = MK_BT709; break;
        case Matrix_SMPTE240M: mk = MK_SMPTE240M; break;
        case Matrix_BT2020_10: mk = MK_BT2020NC; break;
        case Matrix_YCgCo:     mk = MK_YCGCO; break;
        case Matrix_BT601:
        default:               mk = MK_BT601; break;                                               // AUTO -> 601 (csputils.cpp:395-396)
        }
    }
    const bool gray = fmt.CSType == CS_GRAY;
    // DX11VideoProcessor.cpp:839-845
    const float brightness = amp.brightness / 255;
    const float contrast   = amp.contrast;
    const float hue        = (float)(amp.hue / 180 * std::acos(-1.0));
    const float saturation = amp.saturation;
    const int   bits       = fmt.CDepth;

    float (&m)[3][3] = out.m;
    bool rotates = true;
    switch (mk) {
    case MK_BT601:     LumaWeightsToMatrix(m, 0.299f,  0.587f,  0.114f);  break;
    case MK_BT709:     LumaWeightsToMatrix(m, 0.2126f, 0.7152f, 0.0722f); break;
    case MK_SMPTE240M: LumaWeightsToMatrix(m, 0.2122f, 0.7013f, 0.0865f); break;
    case MK_BT2020NC:  LumaWeightsToMatrix(m, 0.2627f, 0.6780f, 0.0593f); break;
    case MK_YCGCO: { const float t[3][3] = {{1, -1, 1}, {1, 1, 0}, {1, -1, -1}};
                     for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = t[i][j]; rotates = false; break; }
    case MK_RGB:   { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = (i == j); lv = LV_ANYFULL; rotates = false; break; }
    }
    if (rotates) {                                               // hue/saturation, csputils.cpp:444-456
        const float hc = gray ? 0 : saturation * std::cos(hue);
        const float hs = gray ? 0 : saturation * std::sin(hue);
        for (int i = 0; i < 3; i++) {
            const float u = m[i][1], v = m[i][2];
            m[i][1] = hc * u - hs * v;
            m[i][2] = hs * u + hc * v;
        }
    }
    // range expansion in double, csputils.cpp:341-358, :459-497
    double mul;
    if (!bits) mul = 1;
    else if (mk == MK_RGB) mul = ((1LL << bits) - 1.) / ((1LL << bits) - 1.);
    else mul = (1LL << bits) / ((1LL << bits) - 1.) * 255 / 256;
    const double s = mul / 255;
    double ymin, ymax, cmax, cmid;
    switch (lv) {
    case LV_TV: ymin = 16 * s; ymax = 235 * s; cmax = 240 * s;     cmid = 128 * s; break;
    case LV_PC: ymin = 0 * s;  ymax = 255 * s; cmax = 255 * s;     cmid = 128 * s; break;
    default:    ymin = 0 * s;  ymax = 255 * s; cmax = 255 * s / 2; cmid = 0;       break;
    }
    double ymul = (1.0 - 0.0) / (ymax - ymin);
    double cmul = (1.0 - 0.0) / (cmax - cmid) / 2;
    ymul *= contrast;
    cmul *= contrast;
    for (int i = 0; i < 3; i++) {
        m[i][0] *= ymul;
        m[i][1] *= cmul;
        m[i][2] *= cmul;
        out.c[i] = 0.0 - m[i][0] * ymin - (m[i][1] + m[i][2]) * cmid + brightness;
    }
    out.enable = fmt.CSType == CS_YUV || gray || std::fabs(brightness) > 1e-4f || std::fabs(contrast - 1.0f) > 1e-4f;
    return true;
}

}  // namespace b200vr
