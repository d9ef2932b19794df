/*
 * oracle/orc_api.h -- TEST INFRASTRUCTURE ONLY (see orc_types.h).
 * Exported C entry points of liborc.so (loaded by tests/ and bench.py via ctypes).
 */
#ifndef ORC_API_H
#define ORC_API_H
#include "orc_types.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- host colour math (orc_csp.c) ---- */
void     orc_gamut_matrix(int mp_prim_in, int mp_prim_out, float matrix[3][3]);
void     orc_gamut_matrix_2020_to_709(float matrix[3][3]);
void     orc_set_colorspace(uint32_t exfmt, int* mp_csp, int* mp_levels);
int      orc_csp_matrix(uint32_t exfmt, int cdepth, int gray,
                        float brightness_dxva, float contrast, float hue_deg, float saturation,
                        float m[3][3], float c[3]);
uint32_t orc_specify_extended_format(uint32_t ex, int cstype, int subsampling, int width, int height);

/* ---- format table (orc_fmt.c), Source/Helper.cpp:309-359 ---- */
typedef struct orc_fmt {
    int32_t cformat;
    int32_t layout;        /* 0 none, 1 biplanar, 2 planar3, 3 packed422, 4 packed444, 5 rgb texel, 6 gray plane */
    int32_t bytes;         /* bytes per stored component */
    int32_t cdepth;        /* CDepth column */
    int32_t cstype;        /* 0 YUV 1 RGB 2 GRAY */
    int32_t subsampling;   /* 420 / 422 / 444 / 400 */
    int32_t div_w, div_h;  /* chroma plane divisors */
    int32_t lsb_shift;     /* CopyPlane10to16: stored code << shift before the 16-bit UNORM read */
    int32_t packsize, pitchcoeff;
    int32_t vu_swapped;    /* YV12/YV16/YV24: plane 2 is V */
    int32_t packed_kind;   /* packed422: 0 YUY2-order(Y0 U Y1 V) 1 UYVY ; packed444: 0 AYUV 1 Y410 2 Y416 */
    const char* name;
} orc_fmt;
const orc_fmt* orc_get_fmt(int cformat);
/* sample buffer -> source texture memory (the CopyFrame* family, Source/Helper.cpp:377-803); -1 = no repack for this format */
int orc_repack(int cformat, int width, int lines, const uint8_t* src, int src_pitch, uint8_t* dst, int dst_pitch);

/* ---- dither matrix (orc_dither.c): thresholds k with D[y][x] = k/1024 ---- */
const uint16_t* orc_dither_thresholds(void);   /* 1024 entries, row-major 32x32 */

/* ---- scalar math for tests ---- */
float  orc_test_lg2(float x);
float  orc_test_sin(float x);
float  orc_test_ex2(float x);
float  orc_test_pow(float x, float y, int mode);

/* ---- decisions (mirrors of small reference rules) ---- */
int orc_internal_fmt(int iTexFormat, int cdepth);                 /* UpdateTexParams, DX11VideoProcessor.cpp:1143-1155 */
int orc_need_final_pass(int backbuffer_fmt, int internal_fmt, int bUseDither); /* :2894-2907 */
/* scaler per axis: 0 none, 1 interpolation (iUpscaling), 2 convolution (iDownscaling); :3106-3127 */
int orc_scaler_kind(int src, int dst, int bInterpolateAt50pct);

/* per-axis resize weights for one output index, returns tap count; base = first source texel index.
 * kind/filter as above; origin = srcRect.left/top; used by tests to pin KATs. */
int orc_axis_taps(int kind, int filter, int variant, int src_origin, int src_len, int dst_len,
                  int i, int* base, float* weights /* >= 1024 */, float* wsum);

/* ---- the pass chain ---- */
/* planes[]/pitches[] as the DX11 texture views see them:
 *  biplanar: Y, interleaved UV; planar3: Y, U(or V for YV*), V; packed: single plane.
 * out: back buffer in cfg->backbuffer_fmt memory layout (BGRA8 / RGB10A2 / RGBA16F), win_h rows of out_pitch bytes,
 *      cleared to zero outside dst_rect.  mode = ORC_MODE_*.  threads <=0: all cores.  Returns 0 on success. */
int orc_process(const orc_config* cfg, const void* const planes[3], const int32_t pitches[3],
                void* out, int32_t out_pitch, orc_debug* dbg, int mode, int threads);

/* same with the Dolby Vision reshaping data of the frame (NULL = none) */
int orc_process_dovi(const orc_config* cfg, const void* const planes[3], const int32_t pitches[3],
                     void* out, int32_t out_pitch, orc_debug* dbg, int mode, int threads, const orc_dovi* dovi);

/* the fetch branch and the convert-shader steps the oracle takes for cfg, as text (see orc_core.inc FN(describe)) */
int orc_describe_convert(const orc_config* cfg, const orc_dovi* dovi, char* out, int cap);
float orc_pq_to_linear_nits(float x);   /* PqToLinearNits, DX11VideoProcessor.cpp:2329-2334 */
int orc_dovi_l1_nits(const orc_dovi* dv, float out_min_max_avg[3]);
/* test hooks: the transfer pieces of the convert shader in a given arithmetic mode (compared with the reference's .hlsl files run as C++) */
int orc_test_transfer(int fn, int mode, const float* in, float* out, int n, float param);
int orc_test_hlg_to_linear_literal(const float* in, float* out, int n);
/* test hook: the Level-2 trim constants build_dovi derives for a display (compared with the reference's CopySample block compiled in place) */
int orc_test_dovi_trims(const orc_dovi* dv, int display_nits, float out5[5]);
/* test hook: records the draws of the resize stage of the next SPEC-mode call made from this thread (orc_types.h: orc_trace_entry) */
void orc_test_trace_begin(orc_trace_entry* buf, int cap);
int  orc_test_trace_end(void);
/* test hook: the colour-transform constants build_cparams derives (compared with the reference's SetShaderConvertColorParams) */
int orc_test_convert_params(const orc_config* cfg, const orc_dovi* dv, float cm[3][3], float cc[3], int32_t* enable, uint32_t* exfmt_out);
/* test hook: the reshaping constants build_dovi derives (compared with the reference's SetShaderDoviCurves compiled in place) */
int orc_test_dovi_curves(const orc_dovi* dv, float piv[3][7], float coef[3][8][4], float mmr[3][48][4], int32_t par[3][4]);

/* run only the post-convert passes on a given m_TexConvertOutput surface (packed, internal fmt) */
int orc_process_from_convert(const orc_config* cfg, const void* convert_surf, int32_t pitch,
                             void* out, int32_t out_pitch, orc_debug* dbg, int mode, int threads);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
