// kernels.h -- host-callable launchers of the CUDA kernels (kernels_multipass.cu, kernel_fused.cu).
#pragma once
#include <cuda_runtime.h>
#include "frame_plan.h"

namespace b200vr {

// ---- multipass family: one kernel per reference pass (Process, Source/DX11VideoProcessor.cpp:3285-3424)
// ConvertColorPass (:3048-3101). dbg_f32 (optional): float4 per pixel, shader output before the RT write.
// dovi_dev (optional, device memory): Dolby Vision reshaping data of the frame
cudaError_t LaunchConvertPass(const SrcDesc& src, const ColorParams& color, const SurfDesc& out, float4* dbg_f32, cudaStream_t st,
                              const DoviParams* dovi_dev = nullptr);
// TextureResizeShader (:332-377): output axis `axis` (0 = X, 1 = Y) is filtered along texture axis `in_axis` (default: the same;
// differs on quads turned by 90/270 degrees); `other` maps the other output axis to the other texture axis (point sample).
cudaError_t LaunchResizePass(const SurfDesc& in, const SurfDesc& out, const AxisTable& table, int axis,
                             const int32_t* other_idx, int other_dst0, int other_dst_len, cudaStream_t st, int in_axis = -1);
// TextureCopyRect with ps_simple (:300-330, :3178-3181): point-sampled copy / stretch.
cudaError_t LaunchCopyPass(const SurfDesc& in, const SurfDesc& out, const int32_t* idx_x, const int32_t* idx_y,
                           int dst_x0, int dst_y0, int dst_w, int dst_h, cudaStream_t st, int transpose = 0);
// FinalPass (:3189-3233): ordered dither + quantise inside rect.
cudaError_t LaunchFinalPass(const SurfDesc& in, const SurfDesc& out, int x0, int y0, int x1, int y1, float Q, cudaStream_t st);
long long VerifyGammaTable(cudaStream_t st);   // mismatches of table vs direct over every float of [0,1]; -1 = table unusable
void SetResizeFour(int mode);        // debug: resize passes 0 = k_resize (one output per thread), 2 = k_resize4, 1 = default (k_resize_p where it applies)
void SetConvertTwoLane(int mode);    // debug: ConvertColorPass 0 = scalar shader body (k_convert), 2 = k_convert2, 1 = default (k_convert_bp where it applies)
void SetGammaTableEnabled(int on);
void SetFused1xVector(int on);      // debug: 0 = 1:1 plans one pixel per thread (k_fused1x) instead of the 128-bit span kernel (k_fused1x_bp)   // debug: 0 = the 2x PQ kernel evaluates the final pow directly instead of the threshold table
// path-B correction shader (m_pPSCorrection) over rect, Source/DX11VideoProcessor.cpp:3354-3357
cudaError_t LaunchCorrectionPass(const SurfDesc& in, const SurfDesc& out, const ColorParams& P, const CorrectionParams& K,
                                 int x0, int y0, int x1, int y1, cudaStream_t st);
// UPSCALE_Jinc2: the one-pass 2-D resizer (Shaders/examples/ps_resize_onepass_jinc2.hlsl) from in rect to out rect.
// orient: the quad's orientation (FillVertices): bit 0 quarter turn (u follows y), bit 1 u reversed, bit 2 v reversed
cudaError_t LaunchJinc2Pass(const SurfDesc& in, const SurfDesc& out, int in_l, int in_t, int in_w, int in_h,
                            int out_l, int out_t, int out_w, int out_h, cudaStream_t st, int orient = 0);
cudaError_t LaunchTonemapPass(const SurfDesc& in, const SurfDesc& out, const ColorParams& P, const TonemapParams& T,
                              int x0, int y0, int x1, int y1, cudaStream_t st);      // ps_hdr10_tonemap post-scale step
cudaError_t LaunchClear(const SurfDesc& s, cudaStream_t st);
// any surface -> B8G8R8X8 (GetDisplayedImage)
cudaError_t LaunchToBGRX(const SurfDesc& in, const SurfDesc& out, cudaStream_t st);

// ---- fused family: the whole chain in one kernel per frame
struct FusedPlan {
    SrcDesc     src;
    ColorParams color;
    AxisTable   ax, ay;          // device tables
    SurfDesc    out;             // back buffer
    int32_t     internal_fmt;    // m_InternalTexFmt: rounding of the convert output and of the post-scale surface
    int32_t     final_pass;      // dither + quantise at the end
    int32_t     two_pass;        // both axes resample -> fp16 intermediate (m_TexResize)
    float       Q;               // 255 / 1023
    // exact-2x plans: the two weight vectors (even / odd outputs) and the window origin relative to j = i/2
    float       wE[8], wO[8];
    int32_t     b0;
    int32_t     periodic2x;      // host verified the tables are 2-periodic
    // host copies of the per-output window origins (tile sizing of the generic fused kernel)
    const int32_t* hbase_x; const int32_t* hbase_y;
};
bool        FusedSupported(const FusedPlan& p, const char** why);
// which fused kernel would run: 1 = 1:1 kernel, 2 = register-blocked 2x kernel, 3 = generic table-driven kernel, 0 = none
int         FusedKind(const FusedPlan& p);
cudaError_t LaunchFused(const FusedPlan& p, cudaStream_t st, int* launches);

// PQ->SDR convert body, scalar (mode 0) vs two-lane (mode 1) forms, on arrays of normalised y,u,v (test hook)
cudaError_t LaunchProbePQ(const float* y, const float* u, const float* v, float* out, const ColorParams& P, int n, int mode, cudaStream_t st);

// ---- scalar math probe (mpcvr_debug_math)
cudaError_t LaunchMathProbe(int op, const float* x, const float* y, float* out, long long n, cudaStream_t st);

// upload of the dither thresholds into constant memory (once per device)
cudaError_t UploadDitherTable();

}  // namespace b200vr
