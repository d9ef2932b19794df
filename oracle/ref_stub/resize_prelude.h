// oracle/ref_stub/resize_prelude.h -- TEST INFRASTRUCTURE ONLY.
// What CDX11VideoProcessor::Process (Source/DX11VideoProcessor.cpp:3285-3424), ResizeShaderPass (:3103-3187), FinalPass (:3189-3233),
// GetPostScaleSteps (:778-795), TextureResizeShader (:332-377), TextureCopyRect (:300-330), their helpers FillVertices / FillVertexBuffer /
// TextureBlt11 (:112-128, :130-179, :200-259) and the post-scale texture ring CTex2DRing (Source/DX11Helper.h:178-240) need to compile
// where they lie: a recording stand-in for the D3D11 device context (every Draw() is written down with the state it was issued under -- pixel
// shader, source texture, render target, viewport, constant buffer, the four vertices), CRect, CComPtr, Tex2D_t, and a class
// declaration with the members those functions touch (ref_stub/resize_class.inc, after the reference's CTex2DRing).  oracle/Makefile
// pipes this file, the reference's own lines (sed -n), resize_class.inc, more reference lines and ref_stub/resize_wrap.inc into g++
// -> oracle/_ref/libref_resize.so.  Nothing of the reference is copied into the repository.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <utility>
#include <vector>
#include <string>
typedef unsigned int UINT; typedef float FLOAT; typedef long HRESULT; typedef long LONG;
#define S_OK ((HRESULT)0)
#define S_FALSE ((HRESULT)1)
#define SUCCEEDED(hr) (((HRESULT)(hr)) >= 0)
typedef long long __int64;
static inline long GetTickCount() { return 0; }
#define E_FAIL ((HRESULT)0x80004005L)
#define FAILED(hr) (((HRESULT)(hr)) < 0)
#define ASSERT(x) ((void)0)
#define DLog(...) ((void)0)
#define DLogIf(...) ((void)0)
#define HR2Str(x) 0
struct POINT { LONG x, y; };
struct RECT { LONG left, top, right, bottom; };
struct CRect : RECT {      // atltypes.h semantics of the members used
    CRect() { left = top = right = bottom = 0; }
    CRect(LONG l, LONG t, LONG r, LONG b) { left = l; top = t; right = r; bottom = b; }
    int Width() const { return (int)(right - left); }
    int Height() const { return (int)(bottom - top); }
    void SetRect(LONG l, LONG t, LONG r, LONG b) { left = l; top = t; right = r; bottom = b; }
    bool operator!=(const RECT& o) const { return left != o.left || top != o.top || right != o.right || bottom != o.bottom; }
    bool IntersectRect(const RECT& a, const RECT& b) {      // ::IntersectRect: the empty rectangle when they do not overlap
        left = a.left > b.left ? a.left : b.left; top = a.top > b.top ? a.top : b.top;
        right = a.right < b.right ? a.right : b.right; bottom = a.bottom < b.bottom ? a.bottom : b.bottom;
        if (left >= right || top >= bottom) { left = top = right = bottom = 0; return false; }
        return true;
    }
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
    void PSSetShaderResources(UINT slot, UINT, ID3D11ShaderResourceView** v) { if (slot == 0 && *v) srv = *v; }
    void PSSetSamplers(UINT, UINT, ID3D11SamplerState**) {}
    void PSSetConstantBuffers(UINT slot, UINT, ID3D11Buffer** b) { if (slot == 0) cb = *b; }
    void UpdateSubresource(ID3D11Buffer*, UINT, const void*, const void*, UINT, UINT) {}
    void IASetPrimitiveTopology(int) {}
    void IASetVertexBuffers(UINT, UINT, ID3D11Buffer** b, const UINT*, const UINT*) { vb = *b; }
    void Draw(UINT, UINT);
};
static int g_texW[8], g_texH[8];        // sizes of the textures by id, for the record
struct Tex2D_t {                        // Source/DX11Helper.h:37-100, as far as the pass functions use it; id = the texture's name in the draw records
    CComPtr<ID3D11Texture2D> pTexture; D3D11_TEXTURE2D_DESC desc = {}; CComPtr<ID3D11ShaderResourceView> pShaderResource;
    ID3D11Texture2D tex_store; ID3D11ShaderResourceView srv_store; int id = 1;
    void Release() { pTexture.p = nullptr; pShaderResource.p = nullptr; desc = D3D11_TEXTURE2D_DESC{}; }
    HRESULT Create(ID3D11Device*, DXGI_FORMAT f, UINT w, UINT h, Tex2DType) {
        tex_store.id = id; tex_store.d = { w, h, f }; srv_store.tex_id = id; pTexture.p = &tex_store; pShaderResource.p = &srv_store; desc = tex_store.d;
        g_texW[id & 7] = (int)w; g_texH[id & 7] = (int)h; return S_OK;
    }
    HRESULT CheckCreate(ID3D11Device* d, DXGI_FORMAT f, UINT w, UINT h, Tex2DType t) {
        if (!w || !h) return E_FAIL;
        if (f == desc.Format && w == desc.Width && h == desc.Height) return S_OK;
        return Create(d, f, w, h, t);
    }
};
