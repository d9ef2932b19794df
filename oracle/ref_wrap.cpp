// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the REFERENCE's own colour math (Source/csputils.{h,cpp}, compiled in place by
// oracle/Makefile) so tests can pin the restatements in oracle/orc_csp.c and
// videorenderer_b200/csrc/cspmath.cpp against it.  Only YUV inputs are exercised: GCC drops the
// reference's `case -1`, which only affects RGB/XYZ/is_float.
#include "csputils.h"
extern "C" {
// mirrors the call in CDX11VideoProcessor::SetShaderConvertColorParams (Source/DX11VideoProcessor.cpp:837-847)
int ref_csp_matrix(int space, int levels, float brightness, float contrast, float hue, float saturation,
                   int gray, int bits, float* m9, float* c3)
{
    mp_csp_params p;
    p.color.space = (mp_csp)space;
    p.color.levels = (mp_csp_levels)levels;
    p.brightness = brightness; p.contrast = contrast; p.hue = hue; p.saturation = saturation;
    p.gray = gray != 0;
    p.input_bits = p.texture_bits = bits;
    mp_cmat cm;
    mp_get_csp_matrix(&p, &cm);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) m9[i*3+j] = cm.m[i][j]; c3[i] = cm.c[i]; }
    return 0;
}
void ref_gamut_matrix(int prim_in, int prim_out, float* m9)
{
    float m[3][3];
    GetColorspaceGamutConversionMatrix(m, (mp_csp_prim)prim_in, (mp_csp_prim)prim_out);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m9[i*3+j] = m[i][j];
}
}
