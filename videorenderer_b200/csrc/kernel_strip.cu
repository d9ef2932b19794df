// kernel_strip.cu -- host side of the persistent strip kernel (kernel_strip.cuh): eligibility, TMA tensor maps, launch.
#include <mutex>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include "kernel_strip.cuh"
#include "kernel_strip.h"

namespace b200vr {

// cuTensorMapEncodeTiled through the runtime's driver entry point query: libmpcvr_b200.so links the CUDA runtime statically
// and does not link libcuda.so directly.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn GetEncodeTiled()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    });
    return fn;
}

static bool MakeMap2D(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t w, uint64_t h, uint64_t pitch_bytes, uint32_t box_w, uint32_t box_h)
{
    EncodeTiledFn enc = GetEncodeTiled();
    if (!enc) return false;
    const cuuint64_t dims[2] = { w, h };
    const cuuint64_t strides[1] = { pitch_bytes };
    const cuuint32_t box[2] = { box_w, box_h };
    const cuuint32_t es[2] = { 1, 1 };
    return enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// watchdog record of the kernel's bounded mbarrier waits (all zero = no wait ever timed out); clears it when asked
int StripStatus(unsigned int out[8], int clear)
{
    if (cudaMemcpyFromSymbol(out, g_stripStatus, 8 * sizeof(unsigned int)) != cudaSuccess) return -1;
    if (clear) { const unsigned int z[8] = {}; if (cudaMemcpyToSymbol(g_stripStatus, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}

// process-wide debug switches (tests flip them between frames; atomics so a concurrent instance reads a whole value)
// MPCVR_DEBUG_STRIP_KERNEL=0 in the environment starts the process on the tile kernel (sanitizer runs: racecheck does not model
// mbarrier-ordered shared-memory handoffs and reports every one of them)
static int StripEnabledAtLoad() { const char* e = getenv("MPCVR_DEBUG_STRIP_KERNEL"); return e ? atoi(e) != 0 : 1; }
static std::atomic<int> g_stripEnabled{StripEnabledAtLoad()}, g_stripCfg{0}, g_stripSkip{0}, g_stripWaitNs{0};
void SetStripKernel(int enabled, int cfg) { g_stripEnabled = enabled; g_stripCfg = cfg & 0xFF; g_stripSkip = (cfg >> 8) & 0xF; g_stripWaitNs = ((cfg >> 12) & 0xF) ? (16 << ((cfg >> 12) & 0xF)) : 0; }   // cfg bits 8..11: role-skip mask (debug)

bool StripEligible(const FusedPlan& p)
{
    if (!g_stripEnabled) return false;
    const SrcDesc& S = p.src;
    if (S.layout != LAYOUT_BIPLANAR || S.chroma_mode != MPCVR_CHROMA_Bilinear || S.deint_blend || S.lsb_shift != 0) return false;
    if (S.subsampling != 420 && S.subsampling != 422) return false;
    if (S.bytes != 1 && S.bytes != 2) return false;
    // TMA: 16-byte aligned planes and pitches
    if (((uintptr_t)S.pl[0] & 15) || ((uintptr_t)S.pl[1] & 15) || S.pitch[0] <= 0 || S.pitch[1] <= 0 || (S.pitch[0] & 15) || (S.pitch[1] & 15)) return false;
    if (p.ax.ntaps != 4 && p.ax.ntaps != 6) return false;
    return GetEncodeTiled() != nullptr;
}

template <class C, int NT, bool DUP, int SPEC, int CONV>
static cudaError_t LaunchK(const StripParams& prm, int grid, cudaStream_t st)
{
    auto kern = k_strip2x<C, NT, DUP, SPEC, CONV>;
    static bool configured[64] = {};
    static std::mutex mtx;
    {
        std::lock_guard<std::mutex> lock(mtx);
        int dev = 0; cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    kern<<<grid, C::THREADS, C::smem_bytes, st>>>(prm);
    return cudaGetLastError();
}

// Configurations measured on B200 (C2 frame; DESIGN.md 4.1, profiles/r02_strip_experiments.txt).  Cfg0 is what the product runs; the
// others stay selectable through mpcvr_debug_set_strip_kernel so that the measurements can be repeated -- every one is bit-identical.
typedef StripCfg<128, 8, 4, 12, 4, 8, OPT_GAMMA_LUT | OPT_C_FUSEDQ> Cfg0;        // 28 warps at 72 registers: 4 P (two teams of 2) + 12 A + 4 B + 8 C; every other instantiation runs this one
typedef StripCfg<128, 8, 4, 10, 4, 8, OPT_GAMMA_LUT> Cfg1;        // 26 warps
typedef StripCfg<128, 8, 6, 10, 4, 8, OPT_GAMMA_LUT> Cfg2;        // 28 warps (P teams of 3)
typedef StripCfg<128, 8, 8, 10, 2, 8, OPT_GAMMA_LUT> Cfg3;        // 28 warps (P teams of 4)
typedef StripCfg<128, 8, 4, 12, 4, 8, 0> Cfg6;                    // gamma through T[] / D[] (two dependent loads, F2I), read-back quotient formed: the round-2a form
typedef StripCfg<128, 8, 4, 12, 4, 8, OPT_GAMMA_LUT> Cfg8;        // single-lookup gamma, read-back quotient formed
typedef StripCfg<128, 8, 4, 12, 4, 8, OPT_GAMMA_LUT | OPT_C_FUSEDQ | OPT_P_COLS> Cfg12;   // + column-per-thread role P (slower: role P's serial path per group grows)
typedef StripCfg<128, 8, 4, 16, 4, 8, OPT_GAMMA_LUT> Cfg7;        // 32 warps at 64 registers (role C spills)
typedef StripCfg<128, 8, 4, 16, 4, 8, OPT_GAMMA_LUT | OPT_SETMAXNREG> Cfg9;        // 32 warps, registers P/B 40, A 64, C 88
typedef StripCfg<128, 8, 4, 12, 4, 12, OPT_GAMMA_LUT | OPT_SETMAXNREG> Cfg10;      // 32 warps, registers P/B 40, A 64, C 80
typedef StripCfg<128, 8, 4, 12, 4, 8, OPT_GAMMA_LUT | OPT_C_LUT> Cfg11;            // + UNORM10 read-back of role C from the table (FMA pipe -> LSU)
// the product's configuration of the two PQ / 10-bit / dithered instantiations (BASELINE C2, C5): 32 warps -- role P as two teams of FOUR
// warps (its serial path per row group is what the other roles wait for when it slips), registers by setmaxnreg: P / B 40, A 64, C 96
typedef StripCfg<128, 8, 8, 12, 4, 8, OPT_GAMMA_LUT | OPT_C_FUSEDQ | OPT_SETMAXNREG> CfgP;
typedef StripCfg<128, 8, 4, 12, 4, 8, OPT_GAMMA_LUT | OPT_C_ROLL> Cfg13;           // + rolled channel loop in role C (instruction-cache footprint)

template <int NT, bool DUP>
static cudaError_t LaunchV(const StripParams& prm, int spec, int conv, int grid, cudaStream_t st)
{
    if (spec == SSPEC_10_BGRA8_DITHER && conv == SCONV_PQ_THR) {
        if constexpr (NT == 6 && DUP) {           // the headline variant carries the alternative warp-role splits (debug switch)
            switch (g_stripCfg.load()) {
            case 1:  return LaunchK<Cfg1, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 2:  return LaunchK<Cfg2, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 3:  return LaunchK<Cfg3, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 4:  return LaunchK<Cfg6, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR_RELAX1>(prm, grid, st);   // measurement only, NOT bit-exact
            case 5:  return LaunchK<Cfg6, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR_RELAX2>(prm, grid, st);   // measurement only, NOT bit-exact
            case 6:  return LaunchK<Cfg6, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 7:  return LaunchK<Cfg7, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 8:  return LaunchK<Cfg8, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 9:  return LaunchK<Cfg9, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 10: return LaunchK<Cfg10, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 11: return LaunchK<Cfg11, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 12: return LaunchK<Cfg12, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            case 14: return LaunchK<Cfg0, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);    // the 28-warp kernel (product until round 2b's last step)
            case 13: return LaunchK<Cfg13, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
            default: break;
            }
        }
        return LaunchK<CfgP, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_PQ_THR>(prm, grid, st);
    }
    if (spec == SSPEC_10_RGB8_DITHER && conv == SCONV_PQ_THR) return LaunchK<CfgP, NT, DUP, SSPEC_10_RGB8_DITHER, SCONV_PQ_THR>(prm, grid, st);
    if (spec == SSPEC_10_BGRA8_DITHER) return LaunchK<Cfg0, NT, DUP, SSPEC_10_BGRA8_DITHER, SCONV_ANY>(prm, grid, st);
    if (spec == SSPEC_8_BGRA8)         return LaunchK<Cfg0, NT, DUP, SSPEC_8_BGRA8, SCONV_ANY>(prm, grid, st);
    if (conv == SCONV_PQ)              return LaunchK<Cfg0, NT, DUP, SSPEC_RUNTIME, SCONV_PQ>(prm, grid, st);
    return LaunchK<Cfg0, NT, DUP, SSPEC_RUNTIME, SCONV_ANY>(prm, grid, st);
}

cudaError_t LaunchStrip2x(const FusedPlan& p, const float* gT, const float* gD, const GammaEntry* gG, const uint32_t exc[2], cudaStream_t st)
{
    const SrcDesc& S = p.src;
    StripParams prm;
    memset(&prm, 0, sizeof(prm));
    constexpr int SW = Cfg0::SW, R = Cfg0::R;
    const int nt = p.ax.ntaps;
    // raw boxes (origins on 16-byte boundaries of the plane rows): luma LBOX texels x R rows; chroma CBOX texels x rc rows
    const uint32_t lbox = S.bytes == 2 ? Cfg0::LBOX16 : Cfg0::LBOX8, cbox = S.bytes == 2 ? Cfg0::CBOX32 : Cfg0::CBOX16;
    prm.luma_pitch = (int32_t)(lbox * S.bytes); prm.luma_bytes = prm.luma_pitch * R;
    prm.rc_rows = S.subsampling == 420 ? R / 2 + 2 : R + 1;
    prm.chroma_pitch = (int32_t)(cbox * 2 * S.bytes); prm.chroma_bytes = prm.chroma_pitch * prm.rc_rows;
    if (!MakeMap2D(&prm.tmY, S.bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, S.pl[0], (uint64_t)S.W, (uint64_t)S.H,
                   (uint64_t)S.pitch[0], lbox, R)) return cudaErrorNotSupported;
    if (!MakeMap2D(&prm.tmC, S.bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, S.pl[1], (uint64_t)S.cw, (uint64_t)S.ch,
                   (uint64_t)S.pitch[1], cbox, (uint32_t)prm.rc_rows)) return cudaErrorNotSupported;
    prm.color = p.color; prm.out = p.out;
    for (int k = 0; k < nt; k++) { prm.wE[k] = p.wE[k]; prm.wO[k] = p.wO[k]; }
    prm.gT = gT; prm.gD = gD; prm.gG = gG; prm.exc0 = exc ? exc[0] : 0xFFFFFFFFu; prm.exc1 = exc ? exc[1] : 0xFFFFFFFFu;
    prm.b0 = p.b0;
    prm.w1 = p.ax.src_len; prm.h1 = p.ay.src_len;
    prm.rect_l = S.rect_l; prm.rect_t = S.rect_t; prm.cw = S.cw; prm.ch = S.ch; prm.bytes = S.bytes;
    if (S.subsampling == 420) {
        prm.my = 2;
        if (S.siting == ChromaSub_Cosited)    { prm.ox = 0;  prm.oy = 0;  }
        else if (S.siting == ChromaSub_MPEG1) { prm.ox = -1; prm.oy = -1; }
        else                                  { prm.ox = 0;  prm.oy = -1; }
    } else { prm.my = 4; prm.ox = 0; prm.oy = 0; }
    prm.dst_x0 = p.ax.dst0; prm.dst_y0 = p.ay.dst0;
    prm.internal_fmt = p.internal_fmt; prm.final_pass = p.final_pass; prm.Q = p.Q;
    prm.strips = (prm.w1 + SW - 1) / SW; prm.nq = (prm.h1 + 3) / 4;
    prm.dbg_skip = g_stripSkip; prm.wait_ns = g_stripWaitNs;

    int spec = SSPEC_RUNTIME;
    if (p.internal_fmt == MPCVR_SURF_RGB10A2 && p.out.fmt == MPCVR_SURF_BGRA8 && p.final_pass && p.Q == 255.0f) spec = SSPEC_10_BGRA8_DITHER;
    else if (p.internal_fmt == MPCVR_SURF_RGB10A2 && p.out.fmt == MPCVR_SURF_RGB8 && p.final_pass && p.Q == 255.0f && (p.ax.dst0 & 3) == 0 &&
             p.color.transfer == TRANSFER_PQ_TO_SDR && S.bytes == 2 && gT && gD && gG) spec = SSPEC_10_RGB8_DITHER;
    else if (p.internal_fmt == MPCVR_SURF_BGRA8 && p.out.fmt == MPCVR_SURF_BGRA8 && !p.final_pass) spec = SSPEC_8_BGRA8;
    int conv = SCONV_ANY;
    if (p.color.transfer == TRANSFER_PQ_TO_SDR && S.bytes == 2) conv = ((spec == SSPEC_10_BGRA8_DITHER || spec == SSPEC_10_RGB8_DITHER) && gT && gD && gG) ? SCONV_PQ_THR : SCONV_PQ;

    static int sms[64] = {};
    int dev = 0; cudaGetDevice(&dev);
    int nsm = (dev >= 0 && dev < 64) ? sms[dev] : 0;
    if (!nsm) { cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); if (dev >= 0 && dev < 64) sms[dev] = nsm; }
    const long long total = (long long)prm.strips * prm.nq;
    const int grid = (int)(total < nsm ? total : nsm);
    if (grid <= 0) return cudaErrorInvalidValue;
    const bool dup = p.ax.dup_q1 != 0;
    if (nt == 6) return dup ? LaunchV<6, true>(prm, spec, conv, grid, st) : LaunchV<6, false>(prm, spec, conv, grid, st);
    return LaunchV<4, false>(prm, spec, conv, grid, st);
}

}  // namespace b200vr
