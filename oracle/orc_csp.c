/*
 * oracle/orc_csp.c -- TEST INFRASTRUCTURE ONLY (see orc_types.h).
 *
 * Host colour math restated from the reference:
 *   mp_invert_matrix3x3              Source/csputils.cpp:14-40
 *   mp_mul_matrix3x3                 Source/csputils.cpp:43-54
 *   mp_get_csp_primaries             Source/csputils.cpp:57-205
 *   mp_get_rgb2xyz_matrix            Source/csputils.cpp:230-263
 *   mp_get_csp_mul                   Source/csputils.cpp:341-358
 *   luma_coeffs                      Source/csputils.cpp:380-389
 *   mp_get_csp_matrix                Source/csputils.cpp:392-509
 *   GetColorspaceGamutConversionMatrix  Source/csputils.cpp:549-557
 *   set_colorspace                   Source/Helper.cpp:949-1004
 *   SpecifyExtendedFormat            Source/Helper.cpp:1169-1211
 * Pinned against the reference's own csputils.cpp compiled into
 * oracle/_ref/libref_csputils.so (tests/test_oracle_csp.py) and the KATs of
 * SURVEY 8 a3/a5 (tests/golden/kat_matrices.json).
 *
 * Float/double mix follows the reference statement by statement: the
 * reference carries `double s, ymul, cmul` and float matrix entries.
 * GCC note: levels_in is carried as int here so the
 * reference's `case -1` (RGB/XYZ/is_float) is reachable.
 */
#include <math.h>
#include <stdlib.h>
#include "orc_types.h"
#include "orc_api.h"

/* mp_csp / mp_csp_levels / mp_csp_prim values, Source/csputils.h:11-51 */
enum { CSP_AUTO, CSP_BT_601, CSP_BT_709, CSP_SMPTE_240M, CSP_BT_2020_NC, CSP_BT_2020_C, CSP_RGB, CSP_XYZ, CSP_YCGCO, CSP_COUNT };
enum { LEVELS_AUTO, LEVELS_TV, LEVELS_PC, LEVELS_COUNT };
enum { PRIM_AUTO, PRIM_BT_601_525, PRIM_BT_601_625, PRIM_BT_709, PRIM_BT_2020, PRIM_BT_470M, PRIM_APPLE, PRIM_ADOBE,
       PRIM_PRO_PHOTO, PRIM_CIE_1931, PRIM_DCI_P3, PRIM_DISPLAY_P3, PRIM_V_GAMUT, PRIM_S_GAMUT, PRIM_EBU_3213,
       PRIM_FILM_C, PRIM_ACES_AP0, PRIM_ACES_AP1, PRIM_COUNT };

typedef struct { float x, y; } xy_t;
typedef struct { xy_t red, green, blue, white; } prim_t;

static void invert3x3(float m[3][3])     /* csputils.cpp:14-40 */
{
    float m00 = m[0][0], m01 = m[0][1], m02 = m[0][2],
          m10 = m[1][0], m11 = m[1][1], m12 = m[1][2],
          m20 = m[2][0], m21 = m[2][1], m22 = m[2][2];
    m[0][0] =  (m11 * m22 - m21 * m12);
    m[0][1] = -(m01 * m22 - m21 * m02);
    m[0][2] =  (m01 * m12 - m11 * m02);
    m[1][0] = -(m10 * m22 - m20 * m12);
    m[1][1] =  (m00 * m22 - m20 * m02);
    m[1][2] = -(m00 * m12 - m10 * m02);
    m[2][0] =  (m10 * m21 - m20 * m11);
    m[2][1] = -(m00 * m21 - m20 * m01);
    m[2][2] =  (m00 * m11 - m10 * m01);
    float det = m00 * m[0][0] + m10 * m[0][1] + m20 * m[0][2];
    det = 1.0f / det;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            m[i][j] *= det;
}

static void mul3x3(float a[3][3], float b[3][3])   /* A := A*B, csputils.cpp:43-54 */
{
    float a00 = a[0][0], a01 = a[0][1], a02 = a[0][2],
          a10 = a[1][0], a11 = a[1][1], a12 = a[1][2],
          a20 = a[2][0], a21 = a[2][1], a22 = a[2][2];
    for (int i = 0; i < 3; i++) {
        a[0][i] = a00 * b[0][i] + a01 * b[1][i] + a02 * b[2][i];
        a[1][i] = a10 * b[0][i] + a11 * b[1][i] + a12 * b[2][i];
        a[2][i] = a20 * b[0][i] + a21 * b[1][i] + a22 * b[2][i];
    }
}

static prim_t get_primaries(int spc)     /* csputils.cpp:57-205 (subset reachable from set_colorspace) */
{
    static const xy_t d65 = {0.31271, 0.32902}, c = {0.31006, 0.31616}, dci = {0.31400, 0.35100};
    prim_t p = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    switch (spc) {
    case PRIM_BT_470M:    p = (prim_t){{0.670, 0.330}, {0.210, 0.710}, {0.140, 0.080}, c};   break;
    case PRIM_BT_601_525: p = (prim_t){{0.630, 0.340}, {0.310, 0.595}, {0.155, 0.070}, d65}; break;
    case PRIM_BT_601_625: p = (prim_t){{0.640, 0.330}, {0.290, 0.600}, {0.150, 0.060}, d65}; break;
    case PRIM_AUTO:
    case PRIM_BT_709:     p = (prim_t){{0.640, 0.330}, {0.300, 0.600}, {0.150, 0.060}, d65}; break;
    case PRIM_BT_2020:    p = (prim_t){{0.708, 0.292}, {0.170, 0.797}, {0.131, 0.046}, d65}; break;
    case PRIM_DCI_P3:     p = (prim_t){{0.680, 0.320}, {0.265, 0.690}, {0.150, 0.060}, dci}; break;
    case PRIM_DISPLAY_P3: p = (prim_t){{0.680, 0.320}, {0.265, 0.690}, {0.150, 0.060}, d65}; break;
    default: break;
    }
    return p;
}

static void rgb2xyz(prim_t space, float m[3][3])   /* csputils.cpp:230-263 */
{
    float S[3], X[4], Z[4];
    X[0] = space.red.x   / space.red.y;
    X[1] = space.green.x / space.green.y;
    X[2] = space.blue.x  / space.blue.y;
    X[3] = space.white.x / space.white.y;
    Z[0] = (1 - space.red.x   - space.red.y)   / space.red.y;
    Z[1] = (1 - space.green.x - space.green.y) / space.green.y;
    Z[2] = (1 - space.blue.x  - space.blue.y)  / space.blue.y;
    Z[3] = (1 - space.white.x - space.white.y) / space.white.y;
    for (int i = 0; i < 3; i++) { m[0][i] = X[i]; m[1][i] = 1; m[2][i] = Z[i]; }
    invert3x3(m);
    for (int i = 0; i < 3; i++)
        S[i] = m[i][0] * X[3] + m[i][1] * 1 + m[i][2] * Z[3];
    for (int i = 0; i < 3; i++) { m[0][i] = S[i] * X[i]; m[1][i] = S[i] * 1; m[2][i] = S[i] * Z[i]; }
}

/* GetColorspaceGamutConversionMatrix, csputils.cpp:549-557; prim values = mp_csp_prim */
void orc_gamut_matrix(int prim_in, int prim_out, float matrix[3][3])
{
    float in[3][3];
    rgb2xyz(get_primaries(prim_in), in);
    rgb2xyz(get_primaries(prim_out), matrix);
    invert3x3(matrix);
    mul3x3(matrix, in);
}

void orc_gamut_matrix_2020_to_709(float matrix[3][3])   /* call site Source/Shaders.cpp:634-645 */
{
    orc_gamut_matrix(PRIM_BT_2020, PRIM_BT_709, matrix);
}

static double csp_mul(int csp, int input_bits, int texture_bits)   /* csputils.cpp:341-358 */
{
    if (!input_bits) return 1;
    if (csp == CSP_RGB) return ((1LL << input_bits) - 1.) / ((1LL << texture_bits) - 1.);
    if (csp == CSP_XYZ) return 1;
    return (1LL << input_bits) / ((1LL << texture_bits) - 1.) * 255 / 256;
}

static void luma_coeffs(float m[3][3], float lr, float lg, float lb)   /* csputils.cpp:380-389 */
{
    m[0][0] = 1; m[0][1] = 0;                    m[0][2] = 2 * (1-lr);
    m[1][0] = 1; m[1][1] = -2 * (1-lb) * lb/lg;  m[1][2] = -2 * (1-lr) * lr/lg;
    m[2][0] = 1; m[2][1] = 2 * (1-lb);           m[2][2] = 0;
}

/* set_colorspace, Source/Helper.cpp:949-1004 (space + levels are what the matrix needs) */
void orc_set_colorspace(uint32_t exfmt, int* space, int* levels)
{
    if (exfmt == 0) { *space = CSP_RGB; *levels = LEVELS_PC; return; }
    switch (ORC_EXFMT_RANGE(exfmt)) {
    case ORC_RANGE_0_255:  *levels = LEVELS_PC; break;
    case ORC_RANGE_16_235: *levels = LEVELS_TV; break;
    default:               *levels = LEVELS_AUTO;
    }
    switch (ORC_EXFMT_MATRIX(exfmt)) {
    case ORC_MATRIX_BT709:     *space = CSP_BT_709;     break;
    case ORC_MATRIX_BT601:     *space = CSP_BT_601;     break;
    case ORC_MATRIX_SMPTE240M: *space = CSP_SMPTE_240M; break;
    case ORC_MATRIX_BT2020_10: *space = CSP_BT_2020_NC; break;
    case ORC_MATRIX_YCgCo:     *space = CSP_YCGCO;      break;
    default:                   *space = CSP_AUTO;
    }
}

/* mp_get_csp_matrix, Source/csputils.cpp:392-509, called as in
 * CDX11VideoProcessor::SetShaderConvertColorParams, Source/DX11VideoProcessor.cpp:837-847 */
int orc_csp_matrix(uint32_t exfmt, int cdepth, int gray,
                   float brightness_dxva, float contrast, float hue_deg, float saturation,
                   float m[3][3], float c[3])
{
    int colorspace, levels_in;
    orc_set_colorspace(exfmt, &colorspace, &levels_in);
    /* DX11VideoProcessor.cpp:839-842 */
    const float p_brightness = brightness_dxva / 255;
    const float p_contrast   = contrast;
    const float p_hue        = hue_deg / 180 * acos(-1);   /* float/int -> float, * double -> rounded to float */
    const float p_saturation = saturation;
    const int input_bits = cdepth, texture_bits = cdepth;

    if (colorspace <= CSP_AUTO || colorspace >= CSP_COUNT) colorspace = CSP_BT_601;
    if (levels_in <= LEVELS_AUTO || levels_in >= LEVELS_COUNT) levels_in = LEVELS_TV;

    switch (colorspace) {
    case CSP_BT_601:     luma_coeffs(m, 0.299,  0.587,  0.114 ); break;
    case CSP_BT_709:     luma_coeffs(m, 0.2126, 0.7152, 0.0722); break;
    case CSP_SMPTE_240M: luma_coeffs(m, 0.2122, 0.7013, 0.0865); break;
    case CSP_BT_2020_NC: luma_coeffs(m, 0.2627, 0.6780, 0.0593); break;
    case CSP_BT_2020_C: { const float t[3][3] = {{0, 0, 1}, {1, 0, 0}, {0, 1, 0}};
        for (int i = 0; i < 9; i++) m[i/3][i%3] = t[i/3][i%3]; break; }
    case CSP_RGB: { const float t[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int i = 0; i < 9; i++) m[i/3][i%3] = t[i/3][i%3]; levels_in = -1; break; }
    case CSP_YCGCO: { const float t[3][3] = {{1, -1, 1}, {1, 1, 0}, {1, -1, -1}};
        for (int i = 0; i < 9; i++) m[i/3][i%3] = t[i/3][i%3]; break; }
    default: return -1;   /* XYZ is never produced by set_colorspace */
    }

    if (colorspace == CSP_BT_601 || colorspace == CSP_BT_709 ||
        colorspace == CSP_SMPTE_240M || colorspace == CSP_BT_2020_NC) {
        float huecos = gray ? 0 : p_saturation * cosf(p_hue);
        float huesin = gray ? 0 : p_saturation * sinf(p_hue);
        for (int i = 0; i < 3; i++) {
            float u = m[i][1], v = m[i][2];
            m[i][1] = huecos * u - huesin * v;
            m[i][2] = huesin * u + huecos * v;
        }
    }

    double s = csp_mul(colorspace, input_bits, texture_bits) / 255;
    double ymin, ymax, cmax, cmid;
    switch (levels_in) {
    case LEVELS_TV: ymin = 16*s; ymax = 235*s; cmax = 240*s;   cmid = 128*s; break;
    case LEVELS_PC: ymin = 0*s;  ymax = 255*s; cmax = 255*s;   cmid = 128*s; break;
    case -1:        ymin = 0*s;  ymax = 255*s; cmax = 255*s/2; cmid = 0;     break;
    default: return -1;
    }
    const double rgbmin = 0, rgbmax = 1;        /* levels_out = PC (csputils.h:111) */
    double ymul = (rgbmax - rgbmin) / (ymax - ymin);
    double cmul = (rgbmax - rgbmin) / (cmax - cmid) / 2;
    ymul *= p_contrast;
    cmul *= p_contrast;
    for (int i = 0; i < 3; i++) {
        m[i][0] *= ymul;
        m[i][1] *= cmul;
        m[i][2] *= cmul;
        c[i] = rgbmin - m[i][0] * ymin - (m[i][1] + m[i][2]) * cmid + p_brightness;
    }
    return 0;
}

/* SpecifyExtendedFormat, Source/Helper.cpp:1169-1211; cstype 0 YUV 1 RGB 2 GRAY */
uint32_t orc_specify_extended_format(uint32_t ex, int cstype, int subsampling, int width, int height)
{
    if (cstype == 1) return 0u;
    if (cstype != 0) return ex;
    uint32_t sample = ex & 0xFF, chroma = ORC_EXFMT_CHROMA(ex), range = ORC_EXFMT_RANGE(ex),
             matrix = ORC_EXFMT_MATRIX(ex), light = ORC_EXFMT_LIGHT(ex), prim = ORC_EXFMT_PRIM(ex),
             trc = ORC_EXFMT_TRC(ex);
    if (subsampling != 420) chroma = ORC_CHROMA_SUB_UNKNOWN;
    else if (chroma == ORC_CHROMA_SUB_UNKNOWN) chroma = ORC_CHROMA_SUB_MPEG2;
    if (range == ORC_RANGE_UNKNOWN) range = ORC_RANGE_16_235;
    if (matrix == ORC_MATRIX_UNKNOWN) matrix = (width <= 1024 && height <= 576) ? ORC_MATRIX_BT601 : ORC_MATRIX_BT709;
    if (light == ORC_LIGHT_UNKNOWN) light = ORC_LIGHT_dim;
    if (prim == ORC_PRIM_UNKNOWN) prim = ORC_PRIM_BT709;
    if (trc == ORC_TRC_UNKNOWN) trc = ORC_TRC_709;
    return ORC_EXFMT(sample, chroma, range, matrix, light, prim, trc);
}
