/*
 * mpcvr_b200_debug.h -- test/inspection hooks of libmpcvr_b200.so.  NOT part of the drop-in
 * boundary (include/mpcvr_b200.h is); used by tests/ and bench.py to look inside the pass chain
 * the way the reference's DumpTexture2D debugging aid does (Source/DX11Helper.cpp:114-170).
 */
#ifndef MPCVR_B200_DEBUG_H
#define MPCVR_B200_DEBUG_H
#include "mpcvr_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* which kernel family Process() uses */
enum { MPCVR_PATH_AUTO = 0,       /* planner decides                                                   */
       MPCVR_PATH_MULTIPASS = 1,  /* one kernel per reference pass, intermediates in HBM               */
       MPCVR_PATH_FUSED = 2 };    /* single fused kernel per frame (fails if the plan has no fused form) */

typedef struct mpcvr_plan_info {
    int32_t internal_fmt;        /* m_InternalTexFmt as mpcvr_surface_format                             */
    int32_t final_pass;          /* m_bFinalPass                                                         */
    int32_t scaler_x, scaler_y;  /* 0 none, 1 interpolation, 2 convolution, 3 nearest stretch            */
    int32_t filter_x, filter_y;  /* UPSCALE_* / DOWNSCALE_* value used                                   */
    int32_t taps_x, taps_y;      /* max taps per output on each axis                                     */
    int32_t path;                /* MPCVR_PATH_* actually selected                                       */
    int32_t transfer;            /* 0 matrix only, 1 PQ->SDR, 2 HLG->SDR, 3 BT.2020 SDR gamut           */
    uint32_t src_exfmt;          /* m_srcExFmt.value after SpecifyExtendedFormat                         */
    int32_t src_pitch, src_lines;/* m_srcPitch, m_srcLines                                               */
    int32_t variant;             /* 0 D3D11 shader set, 1 D3D9 shader set (Settings_t::bUseD3D11)        */
    int32_t correction;          /* m_pPSCorrection: 0 none, 1 convert_pq_to_sdr, 2 fixconvert_pq_to_sdr, 3 fixconvert_hlg_to_sdr,
                                    4 fix_bt2020, 5 convert_hlg_to_pq                                        */
} mpcvr_plan_info;

int32_t mpcvr_debug_set_path(mpcvr* vp, int32_t path);
/* Render() the current frame n times back to back from C (measurement aid: the per-frame cost without the caller's language overhead) */
int32_t mpcvr_debug_render_n(mpcvr* vp, int32_t n);
/* the row of the input format table (Source/Helper.cpp:311-359 columns the chain uses): cformat, Packsize, PitchCoeff, CSType, Subsampling,
 * CDepth, layout (SrcLayout), bytes per component, div_chroma_w, div_chroma_h, lsb_shift, v_before_u.  No device needed. */
int32_t mpcvr_debug_fmt_params(int32_t cformat, int32_t out[12]);
/* SpecifyExtendedFormat (Source/Helper.cpp:1169-1211) as the product applies it.  No device needed. */
uint32_t mpcvr_debug_specify_extended_format(uint32_t exfmt, int32_t cformat, uint32_t width, uint32_t height);
int32_t mpcvr_debug_get_plan(mpcvr* vp, mpcvr_plan_info* info);
/* PS_COLOR_TRANSFORM cm_r,cm_g,cm_b (row-major 3x3) + cm_c, and matrix_conv_prim */
/* process-wide: 1 (default) lets the fused 2x PQ kernel use the gamma+UNORM10 threshold table, 0 forces the direct pow */
int32_t mpcvr_debug_set_gamma_table(int32_t enable);
/* process-wide: 1 (default) lets 1:1 plans of bi-planar sources run the span kernel (8 pixels per thread), 0 keeps one pixel per
 * thread, 2 = the span kernel with 16 pixels per thread for 8-bit sources (the round-2a form) */
int32_t mpcvr_debug_set_fused1x_vector(int32_t enable);
/* process-wide: enable = 1 (default) lets exact-2x plans of bi-planar sources run the persistent warp-specialised kernel (TMA-staged
 * source rows), 0 keeps them on the tile kernel.  cfg, low byte: configuration of the headline variant (kernel_strip.cu) -- 0 the
 * product's (32 warps, role P as two teams of four, setmaxnreg register budgets); 1-3 other warp-role splits; 6 gamma through the two-array table and the read-back quotient formed (round-2a arithmetic);
 * 7 32 warps at 64 registers; 8 single-lookup gamma only; 9, 10 32 warps with setmaxnreg budgets; 11 table read-back in role C;
 * 12 column-per-thread role P; 13 rolled channel loop in role C; 14 the 28-warp kernel (72 registers, role P as two teams of two warps)
 * -- all bit-identical; 4, 5 = relaxed float stage, measurement only, NOT
 * bit-exact.  Bits 8-11: role-isolation mask (P, A, B, C keep their handshakes but skip their arithmetic); bits 12-15: poll with
 * nanosleep(16 << n) instead of the suspend hint. */
int32_t mpcvr_debug_set_strip_kernel(int32_t enable, int32_t cfg);
/* watchdog record of that kernel on the current CUDA device: out8[0] = number of bounded mbarrier waits that timed out (0 expected),
 * [1] barrier slot, [2] parity, [3] row-group iteration, [4] block, [5] warp of the first one; clear != 0 resets the record */
int32_t mpcvr_debug_strip_status(uint32_t out8[8], int32_t clear);
/* process-wide: 1 (default) runs the multipass ConvertColorPass two pixels per thread in the packed-fp32 lanes, 0 one pixel per thread */
int32_t mpcvr_debug_set_convert_two_lane(int32_t enable);
/* process-wide: 1 (default) runs upright resize passes four outputs per thread, 0 one output per thread */
int32_t mpcvr_debug_set_resize_four(int32_t enable);
/* exhaustive check of that table against the direct evaluation over every float in [0,1] on the instance's device:
 * returns the number of mismatching floats (0 expected), -1 if the table was found unusable, -2 on a CUDA error */
int64_t mpcvr_debug_verify_gamma_table(mpcvr* vp);
int32_t mpcvr_debug_get_color_params(mpcvr* vp, float m9[9], float c3[3], float gamut9[9], float* luminance_scale);
/* The constant blocks the kernels receive, byte for byte as the host built them (videorenderer_b200/csrc/frame_plan.h): which 0 = ColorParams,
 * 1 = DoviParams (MPCVR_E_UNEXPECTED without Dolby Vision data), 2 = TonemapParams.  Copies min(size, cap) bytes, returns the struct's size.
 * Works on a planning-only instance (cuda_device -1): tests/test_device_code_on_host.py feeds them to the device pixel functions compiled
 * for the host. */
int32_t mpcvr_debug_get_device_params(mpcvr* vp, int32_t which, void* dst, size_t cap);
/* Takes Dolby Vision metadata the way CopySample takes it from a sample (acceptance check, SetShaderDoviCurves, SetShaderConvertColorParams)
 * without a frame -- so that the blocks above can be inspected without a GPU. */
int32_t mpcvr_debug_set_dovi(mpcvr* vp, const mpcvr_dovi_metadata* md);
/* The SrcDesc block (frame_plan.h) the convert kernels would receive for a sample whose planes lie at planes[] / pitches[] (the
 * MPCVR_SAMPLE_DEVICE_MEMORY form of mpcvr_process_sample), without touching a GPU: which 3 of mpcvr_debug_get_device_params, built from
 * the current media type and settings.  The pointers are stored as given (host addresses are fine for the host-compiled pixel functions).
 * Copies min(size, cap) bytes into dst, returns the struct's size. */
int32_t mpcvr_debug_build_src_desc(mpcvr* vp, const void* const planes[3], const int32_t pitches[3], void* dst, size_t cap);
/* per-axis resize table entry of output index i (host table the kernels consume) */
int32_t mpcvr_debug_axis_taps(mpcvr* vp, int32_t axis, int32_t i, int32_t* base, float* weights, int32_t max_w, float* wsum);

/* surfaces of the multipass chain (valid after a frame ran with MPCVR_PATH_MULTIPASS) */
enum { MPCVR_DBG_SURF_CONVERT = 0,   /* m_TexConvertOutput, internal fmt                   */
       MPCVR_DBG_SURF_RESIZE_X = 1,  /* m_TexResize, RGBA16F                               */
       MPCVR_DBG_SURF_POSTSCALE = 2, /* m_TexsPostScale[0], internal fmt                   */
       MPCVR_DBG_SURF_CONVERT_F32 = 3 /* convert shader output before the RT write: float4 per pixel (capture must be on) */ };
int32_t mpcvr_debug_capture_f32(mpcvr* vp, int32_t enable);
int32_t mpcvr_debug_read_surface(mpcvr* vp, int32_t which, void* host_dst, size_t dst_pitch, int32_t* w, int32_t* h, int32_t* fmt);
/* run Process() downstream of ConvertColorPass on a caller-supplied m_TexConvertOutput (host memory, internal fmt) */
int32_t mpcvr_debug_render_from_convert(mpcvr* vp, const void* host_surf, size_t pitch);

/* the PQ->SDR convert body on n normalised (y,u,v) triples with the instance's colour parameters: mode 0 scalar form, mode 1 two-lane (FFMA2) form */
int32_t mpcvr_debug_probe_pq(mpcvr* vp, const float* yuv3, float* rgb3, int32_t n, int32_t mode);

/* device scalar math of the arithmetic contract, elementwise on host arrays (GPU required):
 * op 0 LG2(x), 1 EX2(x), 2 POW(x,y), 3 UNORM16 read code/65535 (x = code), 4 UNORM10 read, 5 UNORM8 read */
int32_t mpcvr_debug_math(int32_t cuda_device, int32_t op, const float* x, const float* y, float* out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif
