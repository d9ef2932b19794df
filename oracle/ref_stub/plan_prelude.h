// oracle/ref_stub/plan_prelude.h -- TEST INFRASTRUCTURE ONLY.
// What CDX11VideoProcessor::UpdateTexParams (Source/DX11VideoProcessor.cpp:1143-1155: internal texture format), UpdatePostScaleTexures
// (:2893-2912: the final-pass decision m_bFinalPass and the size of the post-scale ring), GetPostScaleSteps (:778-795) and the
// convertType expression of UpdateConvertColorShader (:2946-2948) need to compile where they lie.  oracle/Makefile pipes this file, the
// reference's TEXFMT_* / SHADER_CONVERT_* enums and those lines, and ref_stub/plan_wrap.inc into g++ -> oracle/_ref/libref_plan.so.
#include <cstdint>
#include <vector>
typedef unsigned int UINT; typedef long HRESULT;
#define S_OK ((HRESULT)0)
#define SUCCEEDED(hr) (((HRESULT)(hr)) >= 0)
#define ASSERT(x) ((void)0)
#define FALSE 0
enum DXGI_FORMAT { DXGI_FORMAT_UNKNOWN = 0, DXGI_FORMAT_R16G16B16A16_FLOAT = 10, DXGI_FORMAT_R10G10B10A2_UNORM = 24, DXGI_FORMAT_B8G8R8A8_UNORM = 87 };
enum { IDF_PS_11_FINAL_PASS = 1, IDF_PS_11_FINAL_PASS_10 = 2, MFVideoTransFunc_HLG = 16 };
struct ID3D11PixelShader {}; struct ID3D11Device {};
template <class T> struct CComPtr { T* p = nullptr; operator T*() const { return p; } T** operator&() { return &p; } void Release() { p = nullptr; } };
struct RingRecord { DXGI_FORMAT fmt; UINT w, h, num; };
static RingRecord g_ring;
struct CRect { long left, top, right, bottom; int Width() const { return (int)(right - left); } int Height() const { return (int)(bottom - top); } };
struct DXVA2_ExtendedFormat { unsigned VideoTransferFunction; };
struct CDX11VideoProcessor {
    int m_iTexFormat; DXGI_FORMAT m_InternalTexFmt, m_SwapChainFmt; bool m_bVPUseRTXVideoHDR, m_bUseDither, m_bFinalPass;
    bool m_bConvertToSdr, m_bHdrPassthroughSupport, m_bHdrPassthrough, m_bHdrLocalToneMapping;
    DXVA2_ExtendedFormat m_srcExFmt;
    struct { void* pTexture; } m_TexDither;
    CComPtr<ID3D11PixelShader> m_pPSFinalPass, m_pPSCorrection, m_pPSHDR10ToneMapping, m_pPSHalfOUtoInterlace;
    std::vector<int> m_pPostScaleShaders;
    ID3D11Device* m_pDevice; CRect m_windowRect;
    struct { HRESULT CheckCreate(ID3D11Device*, DXGI_FORMAT f, UINT w, UINT h, UINT num) { g_ring = { f, w, h, num }; return S_OK; } } m_TexsPostScale;
    HRESULT CreatePShaderFromResource(ID3D11PixelShader** pp, UINT resid) { static ID3D11PixelShader ps[3]; *pp = &ps[resid]; g_final_resid = (int)resid; return S_OK; }
    static int g_final_resid;
    void UpdateTexParams(int cdepth);
    void UpdatePostScaleTexures();
    UINT GetPostScaleSteps();
    int  ConvertType();
};
int CDX11VideoProcessor::g_final_resid = 0;
