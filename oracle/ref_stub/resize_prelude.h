// oracle/ref_stub/resize_prelude.h -- TEST INFRASTRUCTURE ONLY.
// What CDX11VideoProcessor::ResizeShaderPass (Source/DX11VideoProcessor.cpp:3103-3187), TextureResizeShader (:332-377), TextureCopyRect
// (:300-330) and their helpers FillVertices / FillVertexBuffer / TextureBlt11 (:114-118, :130-179, :200-259) need to compile where they
// lie: a recording stand-in for the D3D11 device context (every Draw() is written down with the state it was issued under -- pixel
// shader, source texture, render target, viewport, constant buffer, the four vertices), CRect, CComPtr, Tex2D_t, and a class
// declaration with the members those functions touch.  oracle/Makefile pipes this file, the reference's own lines (sed -n) and
// ref_stub/resize_wrap.inc into g++ -> oracle/_ref/libref_resize.so.  Nothing of the reference is copied into the repository.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <utility>
typedef unsigned int UINT; typedef float FLOAT; typedef long HRESULT; typedef long LONG;
#define S_OK ((HRESULT)0)
#define E_FAIL ((HRESULT)0x80004005L)
#define FAILED(hr) (((HRESULT)(hr)) < 0)
#define ASSERT(x) ((void)0)
#define DLog(...) ((void)0)
#define DLogIf(...) ((void)0)
#define HR2Str(x) 0
struct POINT { LONG x, y; };
struct RECT { LONG left, top, right, bottom; };
struct CRect : RECT {
    CRect() { left = top = right = bottom = 0; }
    CRect(LONG l, LONG t, LONG r, LONG b) { left = l; top = t; right = r; bottom = b; }
    int Width() const { return (int)(right - left); }
    int Height() const { return (int)(bottom - top); }
};
namespace DirectX { struct XMFLOAT2 { float x, y; }; struct XMFLOAT3 { float x, y, z; }; }
enum DXGI_FORMAT { DXGI_FORMAT_UNKNOWN = 0, DXGI_FORMAT_R16G16B16A16_FLOAT = 10 };
enum Tex2DType { Tex2D_DefaultShaderRTarget = 1 };
struct D3D11_TEXTURE2D_DESC { UINT Width, Height; DXGI_FORMAT Format; };
struct D3D11_VIEWPORT { FLOAT TopLeftX, TopLeftY, Width, Height, MinDepth, MaxDepth; };
struct D3D11_MAPPED_SUBRESOURCE { void* pData; };
enum { D3D11_MAP_WRITE_DISCARD = 4, D3D11_PRIMITIVE_TOPOLOGY_TRIANGLESTRIP = 5 };
#define D3D11_DEFAULT_SAMPLE_MASK 0xffffffffu
struct ID3D11Buffer { unsigned char mem[256]; };
struct ID3D11Texture2D { int id; D3D11_TEXTURE2D_DESC d; void GetDesc(D3D11_TEXTURE2D_DESC* o) { *o = d; } };
struct ID3D11ShaderResourceView { int tex_id; };
struct ID3D11RenderTargetView { ID3D11Texture2D* tex; };
struct ID3D11PixelShader { int id; };
struct ID3D11VertexShader {}; struct ID3D11InputLayout {}; struct ID3D11SamplerState {}; struct ID3D11BlendState {}; struct ID3D11DepthStencilView {};
template <class T> struct CComPtr {
    T* p = nullptr;
    T** operator&() { return &p; }
    operator T*() const { return p; }
    T* operator->() const { return p; }
};
struct RefDraw { int ps, src_tex, src_w, src_h, target; float vp[4], cb[8], vtx[4][5]; int has_cb; };
static RefDraw g_draws[8]; static int g_ndraws;
struct ID3D11Device {
    ID3D11RenderTargetView views[8]; int nviews = 0;
    HRESULT CreateRenderTargetView(ID3D11Texture2D* t, const void*, ID3D11RenderTargetView** out) { views[nviews & 7].tex = t; *out = &views[nviews++ & 7]; return S_OK; }
};
struct ID3D11DeviceContext {
    ID3D11RenderTargetView* rt = nullptr; D3D11_VIEWPORT vp = {}; ID3D11PixelShader* ps = nullptr; ID3D11ShaderResourceView* srv = nullptr;
    ID3D11Buffer* cb = nullptr; ID3D11Buffer* vb = nullptr; int srv_w = 0, srv_h = 0;
    HRESULT Map(ID3D11Buffer* b, UINT, int, UINT, D3D11_MAPPED_SUBRESOURCE* mr) { mr->pData = b->mem; return S_OK; }
    void Unmap(ID3D11Buffer*, UINT) {}
    void IASetInputLayout(ID3D11InputLayout*) {}
    void OMSetRenderTargets(UINT, ID3D11RenderTargetView** v, ID3D11DepthStencilView*) { rt = *v; }
    void RSSetViewports(UINT, const D3D11_VIEWPORT* v) { vp = *v; }
    void OMSetBlendState(ID3D11BlendState*, const FLOAT*, UINT) {}
    void VSSetShader(ID3D11VertexShader*, void*, UINT) {}
    void PSSetShader(ID3D11PixelShader* p, void*, UINT) { ps = p; }
    void PSSetShaderResources(UINT, UINT, ID3D11ShaderResourceView** v) { if (*v) srv = *v; }
    void PSSetSamplers(UINT, UINT, ID3D11SamplerState**) {}
    void PSSetConstantBuffers(UINT, UINT, ID3D11Buffer** b) { cb = *b; }
    void IASetPrimitiveTopology(int) {}
    void IASetVertexBuffers(UINT, UINT, ID3D11Buffer** b, const UINT*, const UINT*) { vb = *b; }
    void Draw(UINT, UINT);
};
struct Tex2D_t {
    ID3D11Texture2D* pTexture = nullptr; D3D11_TEXTURE2D_DESC desc = {}; ID3D11ShaderResourceView* pShaderResource = nullptr;
    ID3D11Texture2D tex_store; ID3D11ShaderResourceView srv_store;
    void Release() { pTexture = nullptr; pShaderResource = nullptr; desc = D3D11_TEXTURE2D_DESC{}; }
    HRESULT Create(ID3D11Device*, DXGI_FORMAT f, UINT w, UINT h, Tex2DType) { tex_store.id = 1; tex_store.d = { w, h, f }; srv_store.tex_id = 1; pTexture = &tex_store; pShaderResource = &srv_store; desc = tex_store.d; return S_OK; }
};
// the class, as far as these member functions are concerned
struct CDX11VideoProcessor {
    ID3D11Device* m_pDevice; ID3D11DeviceContext* m_pDeviceContext;
    ID3D11Buffer* m_pVertexBuffer; ID3D11Buffer* m_pResizeShaderConstantBuffer;
    ID3D11InputLayout* m_pVSimpleInputLayout; ID3D11VertexShader* m_pVS_Simple; ID3D11SamplerState* m_pSamplerPoint; ID3D11PixelShader* m_pPS_Simple;
    CComPtr<ID3D11PixelShader> m_pShaderUpscaleX, m_pShaderUpscaleY, m_pShaderDownscaleX, m_pShaderDownscaleY;
    Tex2D_t m_TexResize;
    bool m_bInterpolateAt50pct, m_bFlip;
    HRESULT TextureCopyRect(const Tex2D_t& Tex, ID3D11Texture2D* pRenderTarget, const CRect& srcRect, const CRect& destRect,
                            ID3D11PixelShader* pPixelShader, ID3D11Buffer* pConstantBuffer, const int iRotation, const bool bFlip);
    HRESULT TextureResizeShader(const Tex2D_t& Tex, ID3D11Texture2D* pRenderTarget, const CRect& srcRect, const CRect& dstRect,
                                ID3D11PixelShader* pPixelShader, const int iRotation, const bool bFlip);
    HRESULT ResizeShaderPass(const Tex2D_t& Tex, ID3D11Texture2D* pRenderTarget, const CRect& srcRect, const CRect& dstRect, const int rotation);
};
static int g_srcW[2], g_srcH[2];        // sizes of texture 0 (the pass input) and 1 (m_TexResize), for the record
