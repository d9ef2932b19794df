// oracle/ref_stub/hlsl_compat_shaders.h -- TEST INFRASTRUCTURE ONLY (continues hlsl_compat.h).
// What the reference's resize and final-pass pixel shaders (Shaders/d3d11/ps_interpolation_*.hlsl, ps_convolution.hlsl, ps_final_pass.hlsl)
// use beyond float3 / float4: float2, component indexing, the two swizzles of ps_interpolation_lanczos2, frac / floor / ceil / sin / cos / acos
// on scalars and vectors, and Texture2D::Sample through a point sampler (the reference binds m_pSamplerPoint to these draws; the dither
// texture's sampler wraps).  The unit is compiled with -fsingle-precision-constant: as in HLSL, `1.`, `0.5`, `acos(-1.)` are floats.
#define acos(x)  hlsl_acosf(x)
#define sin(x)   hlsl_sin(x)
#define cos(x)   hlsl_cos(x)
#define floor(x) hlsl_floor(x)
#define ceil(x)  hlsl_ceil(x)
static inline float hlsl_acosf(float x) { return acosf(x); }
static inline float hlsl_sin(float x) { return sinf(x); }
static inline float hlsl_cos(float x) { return cosf(x); }
static inline float hlsl_floor(float x) { return floorf(x); }
static inline float hlsl_ceil(float x) { return ceilf(x); }
static inline float frac(float x) { return x - floorf(x); }
struct float2 {
    float x, y;
    float4 xxyy;                    // the swizzle ps_interpolation_lanczos2.hlsl takes of a freshly constructed float2(a, b)
    float2() : x(0), y(0) {}
    float2(double a, double b) : x((float)a), y((float)b), xxyy(a, a, b, b) {}
    float operator[](int i) const { return i ? y : x; }
};
static inline float2 operator*(const float2& a, const float2& b) { float2 r; r.x = a.x * b.x; r.y = a.y * b.y; return r; }
static inline float3 hlsl_sin(const float3& a) { float3 r; r.x = sinf(a.x); r.y = sinf(a.y); r.z = sinf(a.z); return r; }
static inline float4 hlsl_sin(const float4& a) { float4 r; r.x = sinf(a.x); r.y = sinf(a.y); r.z = sinf(a.z); r.w = sinf(a.w); return r; }
static inline float4 hlsl_floor(const float4& a) { float4 r; r.x = floorf(a.x); r.y = floorf(a.y); r.z = floorf(a.z); r.w = floorf(a.w); return r; }
static inline float  dot(const float4& a, const float4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }      // DP4: mul, mad, mad, mad -- unfused (LITERAL)
static inline float4 make4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
struct SamplerState { int wrap = 0; };
struct Texture2D {
    const float* data = nullptr; int w = 0, h = 0;                  // RGBA float texels, row-major
    void bind(const float* d, int w_, int h_) { data = d; w = w_; h = h_; }
    float4 Sample(const SamplerState& s, const float2& uv) const   // point sampling: texel floor(u * w), CLAMP or WRAP addressing
    {
        int ix = (int)floorf(uv.x * (float)w), iy = (int)floorf(uv.y * (float)h);
        if (s.wrap) { ix %= w; if (ix < 0) ix += w; iy %= h; if (iy < 0) iy += h; }
        else { ix = ix < 0 ? 0 : (ix >= w ? w - 1 : ix); iy = iy < 0 ? 0 : (iy >= h ? h - 1 : iy); }
        const float* p = data + ((size_t)iy * w + ix) * 4;
        return make4(p[0], p[1], p[2], p[3]);
    }
};
static inline float4 operator*(float a, const float4& b) { return make4(a * b.x, a * b.y, a * b.z, a * b.w); }
