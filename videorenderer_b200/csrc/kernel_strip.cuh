// kernel_strip.cuh -- the per-frame chain of exact 2x interpolation upscales as ONE persistent, warp-specialised kernel
// (sm_100a): the Blackwell-native form of CDX11VideoProcessor::Process (Source/DX11VideoProcessor.cpp:3285-3424) for bi-planar
// sources (NV12 / P010 / P016 / P210 / P216) with bilinear chroma -- BASELINE configs 2 and 5.
//
//   one CTA per SM walks column strips of SW convert-surface pixels from top to bottom (a contiguous range of the frame's
//   (strip, 4-row quad) list, so every CTA gets the same amount of work and the vertical halo is converted once, not per tile);
//   inside the CTA the frame flows through four warp roles coupled by mbarriers, with no CTA-wide barrier in steady state:
//
//   TMA      one elected thread: cp.async.bulk.tensor.2d of the luma and chroma rows of the next row group into shared memory
//            (raw codes; out-of-frame texels arrive as zeros and are never used: CLAMP addressing is applied downstream)
//   role P   raw codes -> floats: UNORM read of luma, chroma texels (sampler CLAMP), horizontal then vertical bilinear
//            reconstruction at the luma positions   (ShaderGetPixels, Source/Shaders.cpp:230-272)
//   role A   the convert shader body on two pixels per thread in the packed-fp32 lanes, render-target rounding
//            (Source/Shaders.cpp:819-923; m_TexConvertOutput) -> fp32 rows in shared memory.  Pure arithmetic: three 64-bit
//            shared loads, three 64-bit shared stores, addresses linear in the item index.
//   role B   X pass: 8 outputs per item from 10 converted values, two weight phases, fp16 rounding (= m_TexResize) into a
//            32-row ring in shared memory   (ResizeShaderPass :3103-3187, ps_interpolation_*.hlsl)
//   role C   Y pass from the ring + post-scale rounding + final pass (ordered dither) + packing, coalesced 64-bit stores to
//            the back buffer   (FinalPass :3189-3233, ps_final_pass.hlsl)
//
// Arithmetic is the same device code as the tile kernel / multipass kernels (pixel_ops2.cuh), so the output is bit-identical.
// CLAMP addressing of the resize taps: columns are clamped when the X pass reads the converted rows (edge strips only), rows
// when the Y pass picks ring rows (first / last quads only); rows and columns outside the convert surface are never computed.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "kernels.h"
#include "pixel_ops.cuh"
#include "mpcvr_math2.cuh"
#include "pixel_ops2.cuh"
#include "dither32f.cuh"

namespace b200vr {

static __device__ const DitherTableF g_dither32_strip = MakeDitherTableF(0.0009765625f);    // d = k/1024, compile-time constant

enum { SCONV_PQ_THR = 0,      // PQ -> SDR, 10-bit internal format: gamma + UNORM10 write/read through the threshold table
       SCONV_PQ = 1,          // PQ -> SDR, any internal format
       SCONV_ANY = 2,         // any transfer (runtime switch), any internal format
       SCONV_PQ_THR_RELAX1 = 3, SCONV_PQ_THR_RELAX2 = 4 };   // measurement only: SCONV_PQ_THR with a loosened float stage (pixel_ops2.cuh)
enum { SSPEC_RUNTIME = 0, SSPEC_10_BGRA8_DITHER = 1, SSPEC_8_BGRA8 = 2,
       SSPEC_10_RGB8_DITHER = 3 };   // as 1, written as packed R,G,B bytes (MPCVR_SURF_RGB8); lane pairs exchange two bytes by warp shuffle

struct StripParams {
    alignas(64) CUtensorMap tmY;          // luma plane: W x H elements of 1 or 2 bytes, box LUMA_BOX x R
    alignas(64) CUtensorMap tmC;          // interleaved chroma plane: cw x ch texels of 2 or 4 bytes, box CP x rc_rows
    ColorParams color;
    SurfDesc    out;
    float       wE[6], wO[6];             // weights of even / odd outputs (the 2x ratio has exactly two phases)
    const float* gT; const float* gD;     // threshold table (SCONV_PQ_THR)
    const GammaEntry* gG;                 // the same table as 16-byte entries (OPT_GAMMA_LUT)
    uint32_t    exc0, exc1;
    int32_t     b0;                       // window start relative to j (even output); odd = b0 + 1
    int32_t     w1, h1;                   // convert surface size (= m_srcRect size)
    int32_t     rect_l, rect_t;
    int32_t     cw, ch;                   // chroma plane size in texels
    int32_t     bytes;                    // per component: 1 or 2
    int32_t     ox, oy, my;               // chroma sample position of luma (sx, sy) in quarter texels: (2 sx + ox, my sy + oy)
    int32_t     rc_rows;                  // chroma rows per row group (tmC box height)
    int32_t     luma_pitch, chroma_pitch; // shared-memory row pitches of the raw boxes, bytes
    int32_t     luma_bytes, chroma_bytes; // bytes per TMA box
    int32_t     dst_x0, dst_y0;
    int32_t     internal_fmt, final_pass;
    float       Q;
    int32_t     strips, nq;               // column strips, quads (4 convert rows) per strip
    int32_t     wait_ns;                  // 0: mbarrier.try_wait with a suspend hint; else nanosleep this long between polls
    int32_t     dbg_skip;                 // debug (role-balance experiments): bit 0/1/2/3 = role P/A/B/C keeps its handshakes but skips its work
};

// ---- PTX helpers -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug makes the role give up (the frame is then wrong and g_stripStatus says where it stopped; the
// launcher's debug query reports it) instead of hanging the device.  try_wait suspends the thread in hardware for up to the
// hinted time, so a waiting warp costs a few instructions per microsecond, not per cycle.
#ifndef MPCVR_STRIP_WAIT_HINT_NS
#define MPCVR_STRIP_WAIT_HINT_NS 2000     /* suspend-time hint of mbarrier.try_wait; 20 .. 20000 ns measure the same frame time */
#endif
static __device__ unsigned int g_stripStatus[8];       // [0] = failures, [1] = barrier slot, [2] = parity, [3] = iteration, [4] = block, [5] = warp
__device__ __forceinline__ bool mbar_try_hint(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"((unsigned)MPCVR_STRIP_WAIT_HINT_NS) : "memory");
    return ok != 0;
}
// the watchdog record is cold code: kept out of line so that the roles' loops stay small (the SM's instruction cache holds four of them)
__device__ __noinline__ void strip_wait_failed(uint32_t bar, uint32_t parity, int it)
{
    if (atomicAdd(&g_stripStatus[0], 1u) == 0) {
        g_stripStatus[1] = (bar & 255u) >> 3; g_stripStatus[2] = parity; g_stripStatus[3] = (unsigned)it;
        g_stripStatus[4] = blockIdx.x; g_stripStatus[5] = threadIdx.x >> 5;
    }
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int it, uint32_t wns)
{
    if (mbar_try(bar, parity)) return true;
    if (wns) {
        // poll with a real sleep in between: a waiting warp then issues a handful of instructions per wns nanoseconds
        #pragma unroll 1
        for (int n = 0; n < (1 << 22); n++) {
            asm volatile("nanosleep.u32 %0;" :: "r"(wns));
            if (mbar_try(bar, parity)) return true;
        }
    } else {
        #pragma unroll 1
        for (int n = 0; n < (1 << 17); n++)
            if (mbar_try_hint(bar, parity)) return true;
    }
    strip_wait_failed(bar, parity, it);
    return false;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ float fma_sat(float a, float b, float c) { float d; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
// prmt.b32 in its default mode: selector nibbles 0-7 pick a byte of {a, b}; bit 3 of a nibble replicates that byte's sign bit
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; }

// ---- geometry ---------------------------------------------------------------------------------------------------------
// optional code paths of the kernel (all bit-identical; kept selectable so that each one's effect can be measured on its own)
enum { OPT_GAMMA_LUT = 1,         // role A: gamma + UNORM10 through ONE 16-byte table access (gamma_q10_lut2) instead of estimate -> F2I -> T[] -> D[]
       OPT_SETMAXNREG = 2,
       OPT_C_LUT = 4,
       OPT_C_ROLL = 8,            // role C: channel loop not unrolled (instruction-cache footprint)
       OPT_C_FUSEDQ = 16,
       OPT_P_COLS = 32 };         // role P: one column pair per thread, rows in a loop (no index divisions; the UNORM reads in the packed lanes)       // role C: UNORM10 read-back and final pass as ONE fma on the integer code (see kFinalK1 below)

// m_TexsPostScale (RGB10A2) read-back followed by ps_final_pass on an 8-bit target:  out = floor(fma(RN(c / 1023), 255, d)),  c the
// 10-bit code, d = k / 1024 the dither threshold.  For every one of the 1024 x 1024 (c, k) pairs this equals
// floor(fma(c, K1, d + 2^-20)) with K1 = RN(255 / 1023): the quotient never has to be formed.  (The offset has to lie in
// [2^-21, 1.4e-6]; checked exhaustively in tests/test_host_logic.py::test_fused_readback_final_pass_is_exact.)
constexpr float kFinalK1 = (float)(255.0 / 1023.0);
constexpr float kFinalDelta = 9.5367431640625e-07f;           // role C: the UNORM10 read-back c/1023 of m_TexsPostScale from the 1024-entry table (LSU) instead of three packed FMA-pipe operations      // 32 warps at 64 registers; roles P and B give registers back (setmaxnreg.dec 40), role C takes them (setmaxnreg.inc)

template <int SW_, int R_, int NP_, int NA_, int NB_, int NC_, int OPT_ = 0>
struct StripCfg {
    static constexpr int OPT = OPT_;
    static constexpr int SW = SW_;                 // convert-surface pixels per strip
    static constexpr int R = R_;                   // convert rows per row group
    static constexpr int NP = NP_, NA = NA_, NB = NB_, NC = NC_;      // warps per role
    static constexpr int CB = SW / 32;             // column blocks of 32 core pixels (= 64 outputs, one warp wide)
    static constexpr int THREADS = 32 * (NP + NA + NB + NC);
    // register budget per role under OPT_SETMAXNREG (warp groups of four: every role's warp count is a multiple of 4)
    static constexpr int REG_PB = 40, REG_A = 64;
    static constexpr int REG_C = ((2048 - REG_PB * (NP + NB) - REG_A * NA) / NC) / 8 * 8;
    static_assert(!(OPT_ & OPT_SETMAXNREG) || (NP % 4 == 0 && NA % 4 == 0 && NB % 4 == 0 && NC % 4 == 0 && NP + NA + NB + NC == 32 && REG_C >= 64 && REG_C <= 232),
                  "setmaxnreg works on aligned groups of four warps and the launch holds 32 warps x 64 registers");
    static constexpr int WP = SW + 8;              // window pitch in pixels (window = SW + NT <= WP)
    static constexpr int PAIRS = WP / 2;
    static constexpr int CP = SW / 2 + 8;          // chroma texels per box row
    static constexpr int RCMAX = R + 1;            // chroma rows per group: R/2 + 2 (4:2:0) or R + 1 (4:2:2)
    static constexpr int RING = 32;                // rows of the X-pass ring (>= 2R + NT + 6)
    static constexpr int NRAW = 4;                 // TMA stages (two per P team)
    static constexpr int PT = NP / 2;              // role P works as two teams of PT warps on alternate row groups
    static constexpr int NS = 3;                   // stages of the P -> A (prep) and A -> C (conv) rings
    static_assert(RING >= 2 * R + 6 + 6, "ring too small");
    static_assert(R == 8 && RING == 4 * R && (NB % 2) == 0, "the X-pass ring is four row groups; X-pass warps work in pairs");
    static_assert((NP % 2) == 0, "two P teams");
    static constexpr int align128(int x) { return (x + 127) / 128 * 128; }
    // shared memory layout (byte offsets)
    static constexpr int oBars = 0;                                   // 28 mbarriers
    static constexpr int oT = 256;
    static constexpr int oD = oT + align128((kGammaN + 2) * ((OPT_ & OPT_GAMMA_LUT) ? 16 : 4));   // 16-byte entries with OPT_GAMMA_LUT (sD then unused)
    static constexpr int oDith = oD + align128((kGammaN + 1) * 4);
    static constexpr int oRawY = oDith + 8192;                         // dither table stored twice (64 rows): 8 consecutive rows never wrap
    // TMA boxes start on 16-byte boundaries of the plane rows (the hardware rejects other inner coordinates), so the boxes are
    // wider than the window by the worst-case misalignment
    static constexpr int LBOX16 = WP + 8, LBOX8 = WP + 24;            // luma box width in texels (16-bit / 8-bit components)
    static constexpr int CBOX32 = CP + 4, CBOX16 = CP + 8;            // chroma box width in texels (4-byte / 2-byte texels)
    static constexpr int szRawY = align128(R * LBOX16 * 2);
    static constexpr int oRawC = oRawY + NRAW * szRawY;
    static constexpr int szRawC = align128(RCMAX * CBOX32 * 4);
    static constexpr int oCu = oRawC + NRAW * szRawC;
    static constexpr int szC = align128(RCMAX * CP * 4);
    static constexpr int oCv = oCu + szC;
    static constexpr int oHU = oCv + szC;
    static constexpr int szH = align128(RCMAX * WP * 4);
    static constexpr int oHV = oHU + szH;
    static constexpr int szScratch = 2 * szC + 2 * szH;              // per P team: chroma texels (u, v) + horizontally reconstructed rows (u, v)
    static constexpr int oPrep = oCu + 2 * szScratch;
    static constexpr int szPlane = R * WP * 4;                        // one fp32 plane of a row group
    static constexpr int oConv = oPrep + NS * 3 * szPlane;
    static constexpr int oX = oConv + NS * 3 * szPlane;
    static constexpr int szX = 3 * RING * 2 * SW * 2;                 // [3][RING][2 SW] halfs
    static constexpr int smem_bytes = oX + szX;
};

// barrier slots
enum { B_RAWFULL = 0, B_RAWEMPTY = 4, B_PREPFULL = 8, B_PREPEMPTY = 11, B_CONVFULL = 14, B_CONVEMPTY = 17, B_XFULL = 20, B_XDONE = 24, B_COUNT = 28 };

// the (strip, quad range) segments a CTA owns and their row groups; every role walks the same sequence
template <class C, int NT>
struct GroupIter {
    int Qc, Q1, nq, b0, h1;
    int s, qa, qb, wyA, wyB, ngroups, g;
    bool valid;
    __device__ __forceinline__ bool next_seg()
    {
        if (Qc >= Q1) return false;
        s = Qc / nq; qa = Qc - s * nq;
        const int n = min(nq - qa, Q1 - Qc);
        qb = qa + n; Qc += n;
        wyA = max(4 * qa + b0, 0);                              // first / last convert row the segment's outputs read (CLAMP applied)
        wyB = min(4 * qb - 1 + b0 + NT, h1 - 1);
        ngroups = (wyB - wyA + C::R) / C::R;
        g = 0;
        return true;
    }
    __device__ __forceinline__ void init(int q0, int q1, int nq_, int b0_, int h1_) { Qc = q0; Q1 = q1; nq = nq_; b0 = b0_; h1 = h1_; valid = next_seg(); }
    __device__ __forceinline__ void advance() { if (++g >= ngroups) valid = next_seg(); }
    __device__ __forceinline__ int cy0() const { return wyA + g * C::R; }
    __device__ __forceinline__ int rows() const { return min(C::R, wyB - cy0() + 1); }
};

// X pass (TextureResizeShader AXIS 0, 2x interpolation) of one row group for one column block: item = (channel, row, 8 outputs).
// XEDGE: the block's window crosses the left / right border of the convert surface -> CLAMP addressing of the taps.
template <class C, int NT, bool DUP, bool XEDGE>
__device__ __forceinline__ void xpass_rows(const float* __restrict__ conv, __half* __restrict__ sX, const float (&wE)[NT], const float (&wO)[NT],
                                           int rows, int ring0, int cb, int tCB, int cx0, int w1)
{
    constexpr int TCB = 64;                                          // a pair of warps per column block
    #pragma unroll
    for (int c = 0; c < 3; c++) {
        const float* plane = conv + c * (C::szPlane / 4);
        for (int idx = tCB; idx < rows * 8; idx += TCB) {
            const int row = idx >> 3, b = idx & 7;
            const float* src = plane + row * C::WP + 32 * cb + 4 * b;
            float v[NT + 4];
            if (!XEDGE) {
                const float4 a = *reinterpret_cast<const float4*>(src), d = *reinterpret_cast<const float4*>(src + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
                if (NT == 6) { const float2 e = *reinterpret_cast<const float2*>(src + 8); v[8] = e.x; v[9] = e.y; }
            } else {
                #pragma unroll
                for (int i = 0; i < NT + 4; i++) {
                    const int wx = min(max(cx0 + 32 * cb + 4 * b + i, 0), w1 - 1) - cx0;
                    v[i] = plane[row * C::WP + wx];
                }
            }
            __half2 o[4];
            #pragma unroll
            for (int q = 0; q < 4; q++) {
                float e = __fmul_rn(wE[0], v[q]), d = __fmul_rn(wO[0], v[q + 1]);
                #pragma unroll
                for (int k = 1; k < NT; k++) {
                    // D3D11 lanczos3 re-reads tap 0's texel for tap 1 (ps_interpolation_lanczos3.hlsl:33-34)
                    const float te = (NT == 6 && k == 1 && DUP) ? v[q] : v[q + k];
                    const float to = (NT == 6 && k == 1 && DUP) ? v[q + 1] : v[q + 1 + k];
                    e = fma_rn(wE[k], te, e);
                    d = fma_rn(wO[k], to, d);
                }
                o[q] = __floats2half2_rn(e, d);                       // m_TexResize is RGBA16F
            }
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&o[0]); pk.y = *reinterpret_cast<uint32_t*>(&o[1]);
            pk.z = *reinterpret_cast<uint32_t*>(&o[2]); pk.w = *reinterpret_cast<uint32_t*>(&o[3]);
            const int ring = (ring0 + row) & (C::RING - 1);
            *reinterpret_cast<uint4*>(sX + ((size_t)(c * C::RING + ring) * 2 * C::SW) + 64 * cb + 8 * b) = pk;
        }
    }
}

template <class C, int NT, bool DUP, int SPEC, int CONV>
__global__ void __launch_bounds__(C::THREADS, 1)
k_strip2x(const __grid_constant__ StripParams prm)
{
    constexpr bool HOT = SPEC == SSPEC_10_BGRA8_DITHER || SPEC == SSPEC_10_RGB8_DITHER;   // 10-bit internal format, final pass to 8 bits
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wns = (uint32_t)prm.wait_ns;
    float* sT = reinterpret_cast<float*>(smem + C::oT);
    float* sD = reinterpret_cast<float*>(smem + C::oD);
    float* sDith = reinterpret_cast<float*>(smem + C::oDith);      // NEGATED dither values when SPEC == SSPEC_10_BGRA8_DITHER

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NRAW; i++) { mbar_init(sbase + 8 * (B_RAWFULL + i), 1); mbar_init(sbase + 8 * (B_RAWEMPTY + i), C::PT); }
        for (int i = 0; i < C::NS; i++) {
            mbar_init(sbase + 8 * (B_PREPFULL + i), C::PT);  mbar_init(sbase + 8 * (B_PREPEMPTY + i), C::NA);
            mbar_init(sbase + 8 * (B_CONVFULL + i), C::NA);  mbar_init(sbase + 8 * (B_CONVEMPTY + i), C::NB);
        }
        for (int i = 0; i < 4; i++) { mbar_init(sbase + 8 * (B_XFULL + i), C::NB); mbar_init(sbase + 8 * (B_XDONE + i), C::NC); }
        reinterpret_cast<volatile uint32_t*>(smem)[60] = sbase + (uint32_t)C::oT - (0x4B000000u << 4);       // bytes 240..247: table bases as the lookups use them
        reinterpret_cast<volatile uint32_t*>(smem)[61] = sbase + (uint32_t)C::oD - (0x4B000000u << 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr bool LUT = (C::OPT & OPT_GAMMA_LUT) && CONV == SCONV_PQ_THR;
    constexpr bool CLUT = (C::OPT & OPT_C_LUT) && CONV == SCONV_PQ_THR && HOT;
    if (LUT) {
        for (int i = threadIdx.x; i < kGammaN + 2; i += C::THREADS)
            reinterpret_cast<float4*>(sT)[i] = reinterpret_cast<const float4*>(prm.gG)[i];
        if (CLUT) for (int i = threadIdx.x; i < kGammaN + 1; i += C::THREADS) sD[i] = prm.gD[i];
    } else if (CONV == SCONV_PQ_THR || CONV >= SCONV_PQ_THR_RELAX1) {
        for (int i = threadIdx.x; i < kGammaN + 2; i += C::THREADS) sT[i] = prm.gT[i];
        for (int i = threadIdx.x; i < kGammaN + 1; i += C::THREADS) sD[i] = prm.gD[i];
    }
    constexpr bool FQ = (C::OPT & OPT_C_FUSEDQ) && HOT;
    for (int i = threadIdx.x; i < 2048; i += C::THREADS)
        sDith[i] = FQ ? -__fadd_rn(g_dither32_strip.v[i & 1023], kFinalDelta) : (HOT ? -g_dither32_strip.v[i & 1023] : g_dither32_strip.v[i & 1023]);
    __syncthreads();

    const long long total = (long long)prm.strips * prm.nq;
    const int Q0 = (int)(total * blockIdx.x / gridDim.x), Q1 = (int)(total * (blockIdx.x + 1) / gridDim.x);
    const int IFMT = HOT ? (int)MPCVR_SURF_RGB10A2 : (SPEC == SSPEC_8_BGRA8 ? (int)MPCVR_SURF_BGRA8 : prm.internal_fmt);

    if (warp < C::NP) {
        // ================================================================================================ TMA + role P
        asm volatile("pmevent 1;");                                      // marks the role in SASS listings / ncu source pages (PMTRIG)
        if (C::OPT & OPT_SETMAXNREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(C::REG_PB));
        constexpr int TP = C::PT * 32;                                    // threads of one team
        const int team = warp / C::PT, tP = threadIdx.x - team * TP;
        GroupIter<C, NT> gi; gi.init(Q0, Q1, prm.nq, prm.b0, prm.h1);
        GroupIter<C, NT> iss = gi; int it_iss = 0;                       // issue cursor (thread 0 of each team, for the team's own groups)
        float* sCu = reinterpret_cast<float*>(smem + C::oCu + team * C::szScratch); float* sCv = sCu + C::szC / 4;
        float* sHU = sCv + C::szC / 4; float* sHV = sHU + C::szH / 4;
        // the PQ instantiations are only dispatched for 16-bit components (LaunchStrip2x): a compile-time constant halves role P's code
        const int bytes = CONV == SCONV_ANY ? prm.bytes : 2, ox = prm.ox, oy = prm.oy, my = prm.my, RC = prm.rc_rows;
        const int lmask = 16 / bytes - 1, cmask = 8 / bytes - 1;          // box origins are multiples of 16 bytes
        const int cbox = bytes == 2 ? C::CBOX32 : C::CBOX16;
        for (int it = 0; gi.valid; it++, gi.advance()) {
            if ((it & 1) != team) continue;
            if (tP == 0) {
                while (iss.valid && it_iss < it + C::NRAW) {
                    if ((it_iss & 1) == team) {
                        const int slot = it_iss % C::NRAW;
                        if (!mbar_wait(sbase + 8 * (B_RAWEMPTY + slot), ((it_iss / C::NRAW) & 1) ^ 1, it_iss, wns)) break;
                        const uint32_t full = sbase + 8 * (B_RAWFULL + slot);
                        mbar_expect_tx(full, (uint32_t)(prm.luma_bytes + prm.chroma_bytes));
                        const int X0 = prm.rect_l + iss.s * C::SW + prm.b0, sy0 = prm.rect_t + iss.cy0();
                        tma_load_2d(sbase + C::oRawY + slot * C::szRawY, &prm.tmY, X0 & ~lmask, sy0, full);
                        tma_load_2d(sbase + C::oRawC + slot * C::szRawC, &prm.tmC, ((2 * X0 + ox) >> 2) & ~cmask, (my * sy0 + oy) >> 2, full);
                    }
                    it_iss++; iss.advance();
                }
            }
            __syncwarp();
            const int slot = it % C::NRAW, pslot = it % C::NS;
            const int rows = gi.rows();
            const int X0 = prm.rect_l + gi.s * C::SW + prm.b0, sy0 = prm.rect_t + gi.cy0();
            const int ux0 = (2 * X0 + ox) >> 2, uy0 = (my * sy0 + oy) >> 2;
            const int xoff = X0 & lmask, uoff = ux0 & cmask;              // window origin inside the aligned boxes
            const unsigned char* rawY = smem + C::oRawY + slot * C::szRawY;
            const unsigned char* rawC = smem + C::oRawC + slot * C::szRawC;
            float2* pY = reinterpret_cast<float2*>(smem + C::oPrep + (pslot * 3 + 0) * C::szPlane);
            float2* pU = reinterpret_cast<float2*>(smem + C::oPrep + (pslot * 3 + 1) * C::szPlane);
            float2* pV = reinterpret_cast<float2*>(smem + C::oPrep + (pslot * 3 + 2) * C::szPlane);
            if (!mbar_wait(sbase + 8 * (B_RAWFULL + slot), (it / C::NRAW) & 1, it, wns)) return;
            if (!mbar_wait(sbase + 8 * (B_PREPEMPTY + pslot), ((it / C::NS) & 1) ^ 1, it, wns)) return;
            if (!(prm.dbg_skip & 1)) {
            if constexpr ((C::OPT & OPT_P_COLS) != 0) {
            // ---- column-per-thread form: thread tP owns column pair (texel column) tP of every row, so nothing is divided and the
            // column terms stay in registers; the 4 (8) columns beyond the 64th are one extra pass of a few lanes.  Same operations
            // per value as the flat loops below.
            static_assert(C::PAIRS - TP == 4 && C::CP - TP == 8 && TP == 64, "column-per-thread role P: 64 threads, 68 pairs, 72 texels");
            const int lp = prm.luma_pitch, cpch = prm.chroma_pitch;
            // P0: luma codes -> floats (UNORM read)
            auto luma = [&](int r, int p) {
                float2 cfl;
                if (bytes == 2) { const uint16_t* q = reinterpret_cast<const uint16_t*>(rawY + r * lp) + xoff + 2 * p; cfl = make_float2((float)q[0], (float)q[1]); pY[r * C::PAIRS + p] = div_by_const_exact2<65535>(cfl); }
                else            { const uint8_t* q = rawY + r * lp + xoff + 2 * p; cfl = make_float2((float)q[0], (float)q[1]); pY[r * C::PAIRS + p] = div_by_const_exact2<255>(cfl); }
            };
            #pragma unroll 4
            for (int r = 0; r < rows; r++) luma(r, tP);
            if (tP >= 32 && ((tP - 32) >> 2) < rows) luma((tP - 32) >> 2, 64 + (tP & 3));
            // P1: chroma texels -> floats, sampler CLAMP at the plane borders (interior boxes skip the clamps)
            const bool cedge = ux0 < 0 || uy0 < 0 || ux0 + C::CP > prm.cw || uy0 + RC > prm.ch;
            auto texel = [&](int rr, int c) {
                int cl = c + uoff, rl = rr;
                if (cedge) {
                    cl = min(max(min(max(ux0 + c, 0), prm.cw - 1) - ux0 + uoff, 0), cbox - 1);
                    rl = min(max(min(max(uy0 + rr, 0), prm.ch - 1) - uy0, 0), RC - 1);
                }
                float2 uv;
                if (bytes == 2) { const uint32_t w = *reinterpret_cast<const uint32_t*>(rawC + rl * cpch + 4 * cl); uv = div_by_const_exact2<65535>(make_float2((float)(w & 0xFFFFu), (float)(w >> 16))); }
                else            { const uint32_t w = *reinterpret_cast<const uint16_t*>(rawC + rl * cpch + 2 * cl); uv = div_by_const_exact2<255>(make_float2((float)(w & 255u), (float)(w >> 8))); }
                sCu[rr * C::CP + c] = uv.x; sCv[rr * C::CP + c] = uv.y;
            };
            for (int rr = 0; rr < RC; rr++) texel(rr, tP);
            for (int j = tP; j < 8 * RC; j += TP) texel(j >> 3, 64 + (j & 7));
            named_bar(1 + team, TP);
            // P2: horizontal bilinear reconstruction at the luma columns (top / bot of the bilinear sample)
            auto hrec = [&](int rr, int p) {
                const int qx0 = 2 * (X0 + 2 * p) + ox, qx1 = qx0 + 2;
                const int t0 = rr * C::CP + (qx0 >> 2) - ux0, t1 = rr * C::CP + (qx1 >> 2) - ux0;
                const float2 fx = make_float2((float)(qx0 & 3) * 0.25f, (float)(qx1 & 3) * 0.25f);
                const float2 ua = make_float2(sCu[t0], sCu[t1]), ub = make_float2(sCu[t0 + 1], sCu[t1 + 1]);
                const float2 va = make_float2(sCv[t0], sCv[t1]), vb = make_float2(sCv[t0 + 1], sCv[t1 + 1]);
                reinterpret_cast<float2*>(sHU)[rr * C::PAIRS + p] = fma2(fx, sub2(ub, ua), ua);
                reinterpret_cast<float2*>(sHV)[rr * C::PAIRS + p] = fma2(fx, sub2(vb, va), va);
            };
            for (int rr = 0; rr < RC; rr++) hrec(rr, tP);
            if (tP < 4 * RC) hrec(tP >> 2, 64 + (tP & 3));
            named_bar(1 + team, TP);
            // P3: vertical interpolation at the luma rows
            auto vrec = [&](int r, int p) {
                const int qy = my * (sy0 + r) + oy;
                const int h0 = ((qy >> 2) - uy0) * C::PAIRS + p;
                const float fy = (float)(qy & 3) * 0.25f;
                const float2 ut = reinterpret_cast<const float2*>(sHU)[h0], ub = reinterpret_cast<const float2*>(sHU)[h0 + C::PAIRS];
                const float2 vt = reinterpret_cast<const float2*>(sHV)[h0], vb = reinterpret_cast<const float2*>(sHV)[h0 + C::PAIRS];
                pU[r * C::PAIRS + p] = fma2(fy, sub2(ub, ut), ut);
                pV[r * C::PAIRS + p] = fma2(fy, sub2(vb, vt), vt);
            };
            #pragma unroll 4
            for (int r = 0; r < rows; r++) vrec(r, tP);
            if (tP >= 32 && ((tP - 32) >> 2) < rows) vrec((tP - 32) >> 2, 64 + (tP & 3));
            } else {
            // P0: luma codes -> floats (UNORM read)
            for (int k = tP; k < rows * C::PAIRS; k += TP) {
                const int r = k / C::PAIRS, p = k - r * C::PAIRS;
                float2 y;
                if (bytes == 2) { const uint16_t* q = reinterpret_cast<const uint16_t*>(rawY + r * prm.luma_pitch) + xoff + 2 * p; y = make_float2(unorm16(q[0]), unorm16(q[1])); }
                else            { const uint8_t* q = rawY + r * prm.luma_pitch + xoff + 2 * p; y = make_float2(unorm8(q[0]), unorm8(q[1])); }
                pY[k] = y;
            }
            // P1: chroma texels -> floats, sampler CLAMP at the plane borders (interior boxes skip the clamps)
            const bool cedge = ux0 < 0 || uy0 < 0 || ux0 + C::CP > prm.cw || uy0 + RC > prm.ch;
            if (!cedge) {
                for (int k = tP; k < RC * C::CP; k += TP) {
                    const int rr = k / C::CP, c = k - rr * C::CP;
                    float u, v;
                    if (bytes == 2) { const uint32_t w = *reinterpret_cast<const uint32_t*>(rawC + rr * prm.chroma_pitch + 4 * (c + uoff)); u = unorm16(w & 0xFFFFu); v = unorm16(w >> 16); }
                    else            { const uint32_t w = *reinterpret_cast<const uint16_t*>(rawC + rr * prm.chroma_pitch + 2 * (c + uoff)); u = unorm8(w & 255u); v = unorm8(w >> 8); }
                    sCu[k] = u; sCv[k] = v;
                }
            } else {
                for (int k = tP; k < RC * C::CP; k += TP) {
                    const int rr = k / C::CP, c = k - rr * C::CP;
                    const int cl = min(max(min(max(ux0 + c, 0), prm.cw - 1) - ux0 + uoff, 0), cbox - 1);
                    const int rl = min(max(min(max(uy0 + rr, 0), prm.ch - 1) - uy0, 0), RC - 1);
                    float u, v;
                    if (bytes == 2) { const uint32_t w = *reinterpret_cast<const uint32_t*>(rawC + rl * prm.chroma_pitch + 4 * cl); u = unorm16(w & 0xFFFFu); v = unorm16(w >> 16); }
                    else            { const uint32_t w = *reinterpret_cast<const uint16_t*>(rawC + rl * prm.chroma_pitch + 2 * cl); u = unorm8(w & 255u); v = unorm8(w >> 8); }
                    sCu[k] = u; sCv[k] = v;
                }
            }
            named_bar(1 + team, TP);
            // P2: horizontal bilinear reconstruction at the luma columns (top / bot of the bilinear sample)
            for (int k = tP; k < RC * C::PAIRS; k += TP) {
                const int rr = k / C::PAIRS, p = k - rr * C::PAIRS;
                const int qx0 = 2 * (X0 + 2 * p) + ox, qx1 = qx0 + 2;
                const int t0 = rr * C::CP + (qx0 >> 2) - ux0, t1 = rr * C::CP + (qx1 >> 2) - ux0;
                const float2 fx = make_float2((float)(qx0 & 3) * 0.25f, (float)(qx1 & 3) * 0.25f);
                const float2 ua = make_float2(sCu[t0], sCu[t1]), ub = make_float2(sCu[t0 + 1], sCu[t1 + 1]);
                const float2 va = make_float2(sCv[t0], sCv[t1]), vb = make_float2(sCv[t0 + 1], sCv[t1 + 1]);
                reinterpret_cast<float2*>(sHU)[k] = fma2(fx, sub2(ub, ua), ua);
                reinterpret_cast<float2*>(sHV)[k] = fma2(fx, sub2(vb, va), va);
            }
            named_bar(1 + team, TP);
            // P3: vertical interpolation at the luma rows
            for (int k = tP; k < rows * C::PAIRS; k += TP) {
                const int r = k / C::PAIRS, p = k - r * C::PAIRS;
                const int qy = my * (sy0 + r) + oy;
                const int h0 = ((qy >> 2) - uy0) * C::PAIRS + p;
                const float fy = (float)(qy & 3) * 0.25f;
                const float2 ut = reinterpret_cast<const float2*>(sHU)[h0], ub = reinterpret_cast<const float2*>(sHU)[h0 + C::PAIRS];
                const float2 vt = reinterpret_cast<const float2*>(sHV)[h0], vb = reinterpret_cast<const float2*>(sHV)[h0 + C::PAIRS];
                pU[k] = fma2(fy, sub2(ub, ut), ut);
                pV[k] = fma2(fy, sub2(vb, vt), vt);
            }
            }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(sbase + 8 * (B_RAWEMPTY + slot)); mbar_arrive(sbase + 8 * (B_PREPFULL + pslot)); }
        }
    } else if (warp < C::NP + C::NA) {
        // ================================================================================================ role A
        asm volatile("pmevent 2;");
        constexpr int TA = C::NA * 32;
        const int tA = threadIdx.x - C::NP * 32;
        GroupIter<C, NT> gi; gi.init(Q0, Q1, prm.nq, prm.b0, prm.h1);
        int stream = 0;                                                 // items handed out so far, modulo TA
        const uint32_t exc0 = prm.exc0;
        const float ym2s = prm.color.inv_m2 * 1.1920928955078125e-07f, ym1s = prm.color.inv_m1 * 1.1920928955078125e-07f;   // exponents * 2^-23 (exact)
        // table address - (0x4B000000 << 4) as ONE register: read back from shared memory so that ptxas cannot split it into
        // base + constant again (an entry's address is then a single LEA on the float's bits)
        uint32_t K; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(K) : "r"(sbase + 240u));

        int slot = 0; uint32_t ph = 0;                                 // ring slot it % NS and its phase (it / NS) & 1, kept as counters
        for (int it = 0; gi.valid; it++, gi.advance(), ph ^= (slot + 1 == C::NS), slot = (slot + 1 == C::NS) ? 0 : slot + 1) {
            const int items = gi.rows() * C::PAIRS;
            const float2* pY = reinterpret_cast<const float2*>(smem + C::oPrep + (slot * 3 + 0) * C::szPlane);
            const float2* pU = reinterpret_cast<const float2*>(smem + C::oPrep + (slot * 3 + 1) * C::szPlane);
            const float2* pV = reinterpret_cast<const float2*>(smem + C::oPrep + (slot * 3 + 2) * C::szPlane);
            float2* c0 = reinterpret_cast<float2*>(smem + C::oConv + (slot * 3 + 0) * C::szPlane);
            float2* c1 = reinterpret_cast<float2*>(smem + C::oConv + (slot * 3 + 1) * C::szPlane);
            float2* c2 = reinterpret_cast<float2*>(smem + C::oConv + (slot * 3 + 2) * C::szPlane);
            if (!mbar_wait(sbase + 8 * (B_PREPFULL + slot), ph, it, wns)) return;
            if (!mbar_wait(sbase + 8 * (B_CONVEMPTY + slot), ph ^ 1, it, wns)) return;
            int k = tA - stream; if (k < 0) k += TA;
            if (prm.dbg_skip & 2) k = items;
            for (; k < items; k += TA) {
                f3x2 yuv; yuv.x = pY[k]; yuv.y = pU[k]; yuv.z = pV[k];
                if (CONV == SCONV_PQ_THR || CONV >= SCONV_PQ_THR_RELAX1) {
                    const f3x2 g = CONV == SCONV_PQ_THR_RELAX1 ? convert_color_pq2_pre_relaxed<1>(prm.color, ym2s, ym1s, yuv)
                                 : CONV == SCONV_PQ_THR_RELAX2 ? convert_color_pq2_pre_relaxed<2>(prm.color, ym2s, ym1s, yuv)
                                 : convert_color_pq2_pre_s(prm.color, ym2s, ym1s, yuv);       // saturated gamut-mapped linear light
                    if (LUT) {
                        c0[k] = gamma_q10_lut2(g.x, K, exc0); c1[k] = gamma_q10_lut2(g.y, K, exc0); c2[k] = gamma_q10_lut2(g.z, K, exc0);
                    } else {
                    c0[k] = make_float2(gamma_q10_thr1(g.x.x, sT, sD, exc0), gamma_q10_thr1(g.x.y, sT, sD, exc0));
                    c1[k] = make_float2(gamma_q10_thr1(g.y.x, sT, sD, exc0), gamma_q10_thr1(g.y.y, sT, sD, exc0));
                    c2[k] = make_float2(gamma_q10_thr1(g.z.x, sT, sD, exc0), gamma_q10_thr1(g.z.y, sT, sD, exc0));
                    }
                } else {
                    const f3x2 rgb = (CONV == SCONV_PQ) ? convert_color_pq2(prm.color, yuv) : convert_color2(prm.color, yuv);
                    c0[k] = rt_roundtrip2(IFMT, rgb.x); c1[k] = rt_roundtrip2(IFMT, rgb.y); c2[k] = rt_roundtrip2(IFMT, rgb.z);
                }
            }
            stream += items; while (stream >= TA) stream -= TA;
            __syncwarp();
            if (lane == 0) { mbar_arrive(sbase + 8 * (B_CONVFULL + slot)); mbar_arrive(sbase + 8 * (B_PREPEMPTY + slot)); }
        }
    } else if (warp < C::NP + C::NA + C::NB) {
        // ================================================================================================ role B: X pass
        asm volatile("pmevent 3;");
        if (C::OPT & OPT_SETMAXNREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(C::REG_PB));
        const int wb = warp - C::NP - C::NA, pair = wb >> 1, tB = (wb & 1) * 32 + lane;
        __half* sX = reinterpret_cast<__half*>(smem + C::oX);
        float wE[NT], wO[NT];
        #pragma unroll
        for (int k = 0; k < NT; k++) { wE[k] = prm.wE[k]; wO[k] = prm.wO[k]; }
        GroupIter<C, NT> gi; gi.init(Q0, Q1, prm.nq, prm.b0, prm.h1);
        int it0 = 0;
        for (int it = 0; gi.valid; it++, gi.advance()) {
            if (gi.g == 0) it0 = it;
            const int slot = it % C::NS; const uint32_t ph = (it / C::NS) & 1;
            const int cx0 = gi.s * C::SW + prm.b0;                       // convert-surface column of window column 0
            if (!mbar_wait(sbase + 8 * (B_CONVFULL + slot), ph, it, wns)) return;
            // ring slot it % 4 held the rows of group it - 4; their last readers are the quads completed by group it - 3
            if (it >= 3 && !mbar_wait(sbase + 8 * (B_XDONE + ((it - 3) & 3)), ((it - 3) >> 2) & 1, it, wns)) return;
            const int rows = gi.rows(), ring0 = gi.cy0() - gi.wyA + C::R * it0;
            const float* conv = reinterpret_cast<const float*>(smem + C::oConv + slot * 3 * C::szPlane);
            for (int cb = pair; cb < C::CB; cb += C::NB / 2) {
                if (gi.s * C::SW + 32 * cb >= prm.w1 || (prm.dbg_skip & 4)) continue;
                const bool xedge = (cx0 + 32 * cb < 0) || (cx0 + 32 * cb + 32 + NT > prm.w1);
                if (!xedge) xpass_rows<C, NT, DUP, false>(conv, sX, wE, wO, rows, ring0, cb, tB, 0, 0);
                else        xpass_rows<C, NT, DUP, true>(conv, sX, wE, wO, rows, ring0, cb, tB, cx0, prm.w1);
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(sbase + 8 * (B_CONVEMPTY + slot)); mbar_arrive(sbase + 8 * (B_XFULL + (it & 3))); }
        }
    } else {
        // ================================================================================================ role C: Y pass + final pass
        asm volatile("pmevent 4;");
        if (C::OPT & OPT_SETMAXNREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(C::REG_C));
        const int wc = warp - C::NP - C::NA - C::NB;
        const int  OFMT  = SPEC != SSPEC_RUNTIME ? (int)MPCVR_SURF_BGRA8 : prm.out.fmt;
        const bool FINAL = HOT ? true : (SPEC == SSPEC_8_BGRA8 ? false : prm.final_pass != 0);
        __half* sX = reinterpret_cast<__half*>(smem + C::oX);
        float wE[NT], wO[NT];
        #pragma unroll
        for (int k = 0; k < NT; k++) { wE[k] = prm.wE[k]; wO[k] = prm.wO[k]; }
        GroupIter<C, NT> gi; gi.init(Q0, Q1, prm.nq, prm.b0, prm.h1);
        // items (quad, column block) go round the NC warps: with NC a multiple of CB a warp always takes column block wc % CB, of every
        // (NC / CB)-th quad -- qphase counts quads modulo NC / CB
        static_assert(C::NC % C::CB == 0, "role C hands out the column blocks of a quad to consecutive warps");
        constexpr int QCYC = C::NC / C::CB;
        const int myphase = wc / C::CB, cb = wc % C::CB;
        int qnext = 0, it0 = 0, qphase = 0;
        uint32_t KD;       // address of D[] - (0x4B000000 << 2) as one register (OPT_C_LUT), see K in role A
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(KD) : "r"(sbase + 244u));

        for (int it = 0; gi.valid; it++, gi.advance()) {
            if (gi.g == 0) { it0 = it; qnext = gi.qa; }
            if (!mbar_wait(sbase + 8 * (B_XFULL + (it & 3)), (it >> 2) & 1, it, wns)) return;
            // ---- Y pass + rounding + final pass for the quads whose rows are now complete; item = (quad, column block)
            const int cyLast = gi.cy0() + gi.rows() - 1;
            for (; qnext < gi.qb && min(4 * qnext + 3 + prm.b0 + NT, prm.h1 - 1) <= cyLast; qnext++) {
              {
                const bool mine = qphase == myphase;
                qphase = (qphase + 1 == QCYC) ? 0 : qphase + 1;
                if (!mine || (prm.dbg_skip & 8)) continue;
                const int q = qnext;
                const int col = gi.s * C::SW + 32 * cb + lane;                     // convert-surface column of this lane's output pair
                const bool valid = col < prm.w1;                                   // lanes past the right edge compute (on in-bounds garbage) but never store
                if (SPEC != SSPEC_10_RGB8_DITHER && !valid) continue;
                const int jy = 4 * q;                                              // first convert row of the quad
                const int gx = prm.dst_x0 + 2 * col, gy = prm.dst_y0 + 2 * jy;
                const bool yedge = (jy + prm.b0 < 0) || (jy + 3 + prm.b0 + NT > prm.h1 - 1);
                // byte offsets of the ring rows the quad reads (row pitch 4*SW bytes; RING = 32 rows wrap through the top 5 bits)
                constexpr int ROWB = 2 * C::SW * 2;
                static_assert(C::RING == 32, "ring arithmetic below assumes 32 rows");
                const int rbase = C::R * it0 - gi.wyA;                             // ring row of convert row cy: (cy + rbase) mod 32
                uint32_t roff[NT + 4];
                if (!yedge) {
                    const uint32_t t = (uint32_t)(jy + prm.b0 + rbase) << 27;
                    #pragma unroll
                    for (int r = 0; r < NT + 4; r++) roff[r] = ((t + ((uint32_t)r << 27)) >> 27) * ROWB;
                } else {
                    #pragma unroll
                    for (int r = 0; r < NT + 4; r++) {
                        const int cy = min(max(jy + prm.b0 + r, 0), prm.h1 - 1);         // CLAMP addressing of the Y taps
                        roff[r] = (uint32_t)((cy + rbase) & (C::RING - 1)) * ROWB;
                    }
                }
                const int nrows = min(8, 2 * (prm.h1 - jy));                       // output rows of this quad that exist
                const char* colp = reinterpret_cast<const char*>(sX + 64 * cb + 2 * lane);
                if (HOT) {
                    // channel-sequential: B, G, R (R, G, B for the packed RGB8 target); codes are packed as they appear (two PRMTs per pixel)
                    float2 nd[8];
                    #pragma unroll
                    for (int r = 0; r < 8; r++) nd[r] = *reinterpret_cast<const float2*>(sDith + ((gy & 31) << 5) + (gx & 31) + 32 * r);
                    uint32_t w0[8], w1[8];
                    // OPT_C_ROLL: the three channels as a real loop (a third of the code: the SM's instruction cache holds four roles'
                    // loops); the packing step then takes its byte selector from the channel number
                    constexpr bool CROLL = (C::OPT & OPT_C_ROLL) != 0;
                    auto channel = [&](const int cc, const uint32_t sel) {
                        const int c = SPEC == SSPEC_10_RGB8_DITHER ? cc : 2 - cc;
                        float2 v[NT + 4];
                        #pragma unroll
                        for (int r = 0; r < NT + 4; r++)
                            v[r] = __half22float2(*reinterpret_cast<const __half2*>(colp + roff[r] + c * C::RING * ROWB));
                        #pragma unroll
                        for (int qq = 0; qq < 4; qq++) {
                            float2 e = mul2(v[qq], wE[0]), d = mul2(v[qq + 1], wO[0]);
                            #pragma unroll
                            for (int k = 1; k < NT - 1; k++) {
                                const float2 te = (NT == 6 && k == 1 && DUP) ? v[qq] : v[qq + k];
                                const float2 to = (NT == 6 && k == 1 && DUP) ? v[qq + 1] : v[qq + 1 + k];
                                e = fma2(wE[k], te, e);
                                d = fma2(wO[k], to, d);
                            }
                            // last tap with the saturate of the UNORM write folded in (fma.rn.sat: same value as sat(fma))
                            const float2 le = v[qq + NT - 1], lo = v[qq + NT];
                            e = make_float2(fma_sat(wE[NT - 1], le.x, e.x), fma_sat(wE[NT - 1], le.y, e.y));
                            d = make_float2(fma_sat(wO[NT - 1], lo.x, d.x), fma_sat(wO[NT - 1], lo.y, d.y));
                            #pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const int r = 2 * qq + h;
                                const float2 val = h ? d : e;
                                // m_TexsPostScale (RGB10A2) write + read, then ps_final_pass: floor(p*255 + dither), negated so that the
                                // sign byte doubles as the alpha byte: fl = -(2^23 + floor(t))
                                if (FQ) {
                                    const float2 cf = floor_pos2(fma2(val, 1023.0f, 0.5f));                      // the UNORM10 code as a float
                                    const float2 fl = __fadd2_ru(fma2(cf, -kFinalK1, nd[r]), bc2(-8388608.0f));  // nd = -(d + 2^-20)
                                    const uint32_t f0 = __float_as_uint(fl.x), f1 = __float_as_uint(fl.y);
                                    if (CROLL)        { w0[r] = prmt(w0[r], f0, sel); w1[r] = prmt(w1[r], f1, sel); }
                                    else if (cc == 0) { w0[r] = f0; w1[r] = f1; }
                                    else if (cc == 1) { w0[r] = prmt(w0[r], f0, 0x0040); w1[r] = prmt(w1[r], f1, 0x0040); }
                                    else              { w0[r] = prmt(w0[r], f0, 0xF410); w1[r] = prmt(w1[r], f1, 0xF410); }
                                    continue;
                                }
                                float2 p;
                                if (CLUT) {
                                    // 2^23 + code in the low mantissa bits -> D[code] = code / 1023 (correctly rounded, the table of the gamma stage)
                                    const float2 cb = __fadd2_rd(fma2(val, 1023.0f, 0.5f), bc2(8388608.0f));
                                    asm("ld.shared.f32 %0, [%1];" : "=f"(p.x) : "r"((__float_as_uint(cb.x) << 2) + KD));
                                    asm("ld.shared.f32 %0, [%1];" : "=f"(p.y) : "r"((__float_as_uint(cb.y) << 2) + KD));
                                } else p = div_by_const_exact2<1023>(floor_pos2(fma2(val, 1023.0f, 0.5f)));
                                const float2 fl = __fadd2_ru(fma2(p, -255.0f, nd[r]), bc2(-8388608.0f));
                                const uint32_t f0 = __float_as_uint(fl.x), f1 = __float_as_uint(fl.y);
                                if (CROLL)        { w0[r] = prmt(w0[r], f0, sel); w1[r] = prmt(w1[r], f1, sel); }          // selector per channel (below)
                                else if (cc == 0) { w0[r] = f0; w1[r] = f1; }                                        // B       (R)
                                else if (cc == 1) { w0[r] = prmt(w0[r], f0, 0x0040); w1[r] = prmt(w1[r], f1, 0x0040); }   // B, G    (R, G)
                                else              { w0[r] = prmt(w0[r], f0, 0xF410); w1[r] = prmt(w1[r], f1, 0xF410); }   // B, G, R, 0xFF  (R, G, B, x)
                            }
                        }
                    };
                    if (CROLL) {
                        #pragma unroll
                        for (int r = 0; r < 8; r++) { w0[r] = 0u; w1[r] = 0u; }
                        #pragma unroll 1
                        for (int cc = 0; cc < 3; cc++) channel(cc, cc == 0 ? 0x4444u : (cc == 1 ? 0x0040u : 0xF410u));
                    } else {
                        #pragma unroll
                        for (int cc = 0; cc < 3; cc++) channel(cc, 0u);
                    }
                    const uint32_t opitch = (uint32_t)prm.out.pitch;
                    if (SPEC == SSPEC_10_RGB8_DITHER) {
                        // 6 bytes per lane, 12 contiguous bytes per lane pair = three aligned 32-bit words: the even lane writes words 0 and 1
                        // (its pixels + the partner's first two bytes, fetched by shuffle), the odd lane word 2.  dst_x0 % 4 == 0 (launcher).
                        uint8_t* orow = (uint8_t*)prm.out.ptr + (size_t)gy * prm.out.pitch + (size_t)gx * 3;
                        const bool odd = lane & 1;
                        const bool pvalid = __shfl_xor_sync(0xFFFFFFFFu, (int)valid, 1) != 0;
                        #pragma unroll
                        for (int r = 0; r < 8; r++) {
                            const uint32_t lo = prmt(w0[r], w1[r], 0x4210);              // R0 G0 B0 R1
                            const uint32_t hi = prmt(w1[r], w1[r], 0x0021);              // G1 B1 . .
                            const uint32_t plo = __shfl_xor_sync(0xFFFFFFFFu, lo, 1);
                            if (r < nrows && valid) {
                                uint8_t* q = orow + (size_t)((uint32_t)r * opitch);
                                if (!odd) {
                                    *reinterpret_cast<uint32_t*>(q) = lo;
                                    if (pvalid) *reinterpret_cast<uint32_t*>(q + 4) = prmt(hi, plo, 0x5410);      // G1 B1 | partner R0 G0
                                    else        *reinterpret_cast<uint16_t*>(q + 4) = (uint16_t)hi;
                                } else {
                                    *reinterpret_cast<uint32_t*>(q + 2) = prmt(lo, hi, 0x5432);                  // B0 R1 G1 B1
                                }
                            }
                        }
                    } else {
                        uint8_t* orow = (uint8_t*)prm.out.ptr + (size_t)gy * prm.out.pitch + (size_t)gx * 4;
                        #pragma unroll
                        for (int r = 0; r < 8; r++)
                            if (r < nrows) *reinterpret_cast<uint2*>(orow + (size_t)((uint32_t)r * opitch)) = make_uint2(w0[r], w1[r]);
                    }
                } else {
                    float2 res[3][8];
                    #pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float2 v[NT + 4];
                        #pragma unroll
                        for (int r = 0; r < NT + 4; r++)
                            v[r] = __half22float2(*reinterpret_cast<const __half2*>(colp + roff[r] + c * C::RING * ROWB));
                        #pragma unroll
                        for (int qq = 0; qq < 4; qq++) {
                            float2 e = mul2(v[qq], wE[0]), d = mul2(v[qq + 1], wO[0]);
                            #pragma unroll
                            for (int k = 1; k < NT; k++) {
                                const float2 te = (NT == 6 && k == 1 && DUP) ? v[qq] : v[qq + k];
                                const float2 to = (NT == 6 && k == 1 && DUP) ? v[qq + 1] : v[qq + 1 + k];
                                e = fma2(wE[k], te, e);
                                d = fma2(wO[k], to, d);
                            }
                            res[c][2 * qq] = e; res[c][2 * qq + 1] = d;
                        }
                    }
                    #pragma unroll
                    for (int r = 0; r < 8; r++) {
                        if (r >= nrows) break;
                        const int y = gy + r;
                        uint32_t code[2][3];
                        __half hv[2][3];
                        float2 dth = make_float2(0.f, 0.f);
                        if (FINAL) dth = *reinterpret_cast<const float2*>(sDith + ((y & 31) << 5) + (gx & 31));
                        #pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const float2 val = res[c][r];
                            if (FINAL) {
                                const float2 p = rt_roundtrip2(IFMT, val);
                                const float2 t = fma2(p, prm.Q, dth);
                                if (IFMT == MPCVR_SURF_RGBA16F) {
                                    code[0][c] = (uint32_t)min(max((int)floorf(t.x), 0), (int)prm.Q);
                                    code[1][c] = (uint32_t)min(max((int)floorf(t.y), 0), (int)prm.Q);
                                } else {
                                    const float2 fl = __fadd2_rd(t, bc2(8388608.0f));
                                    code[0][c] = __float_as_uint(fl.x) & 0x3FFFFFu; code[1][c] = __float_as_uint(fl.y) & 0x3FFFFFu;
                                }
                            } else if (OFMT == MPCVR_SURF_RGBA16F) {
                                hv[0][c] = __float2half_rn(val.x); hv[1][c] = __float2half_rn(val.y);
                            } else {
                                const float2 fl = __fadd2_rd(fma2(sat2(val), OFMT == MPCVR_SURF_RGB10A2 ? 1023.0f : 255.0f, 0.5f), bc2(8388608.0f));
                                code[0][c] = __float_as_uint(fl.x) & 0x3FFFFFu; code[1][c] = __float_as_uint(fl.y) & 0x3FFFFFu;
                            }
                        }
                        uint8_t* orow = (uint8_t*)prm.out.ptr + (size_t)y * prm.out.pitch;
                        if (OFMT == MPCVR_SURF_RGB8) {                              // packed R,G,B: three 16-bit stores (gx is even)
                            uint16_t* q = reinterpret_cast<uint16_t*>(orow + (size_t)gx * 3);
                            q[0] = (uint16_t)(code[0][0] | (code[0][1] << 8)); q[1] = (uint16_t)(code[0][2] | (code[1][0] << 8));
                            q[2] = (uint16_t)(code[1][1] | (code[1][2] << 8));
                        } else
                        if (OFMT == MPCVR_SURF_RGBA16F) {
                            uint4 o;
                            const __half2 a0 = __halves2half2(hv[0][0], hv[0][1]), a1 = __halves2half2(hv[0][2], __float2half_rn(1.0f));
                            const __half2 b0 = __halves2half2(hv[1][0], hv[1][1]), b1 = __halves2half2(hv[1][2], __float2half_rn(1.0f));
                            o.x = *reinterpret_cast<const uint32_t*>(&a0); o.y = *reinterpret_cast<const uint32_t*>(&a1);
                            o.z = *reinterpret_cast<const uint32_t*>(&b0); o.w = *reinterpret_cast<const uint32_t*>(&b1);
                            *reinterpret_cast<uint4*>(orow + (size_t)gx * 8) = o;
                        } else {
                            uint2 o;
                            if (OFMT == MPCVR_SURF_RGB10A2) {
                                o.x = code[0][0] | (code[0][1] << 10) | (code[0][2] << 20) | (3u << 30);
                                o.y = code[1][0] | (code[1][1] << 10) | (code[1][2] << 20) | (3u << 30);
                            } else {
                                o.x = code[0][2] | (code[0][1] << 8) | (code[0][0] << 16) | (255u << 24);
                                o.y = code[1][2] | (code[1][1] << 8) | (code[1][0] << 16) | (255u << 24);
                            }
                            *reinterpret_cast<uint2*>(orow + (size_t)gx * 4) = o;
                        }
                    }
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sbase + 8 * (B_XDONE + (it & 3)));
        }
    }
}

}  // namespace b200vr
