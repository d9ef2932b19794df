// oracle/ref_stub/hlsl_vec.h -- TEST INFRASTRUCTURE ONLY.
// HLSL's vector vocabulary for C++, complete enough for the convert shader the reference GENERATES (Source/Shaders.cpp:593-930,
// GetShaderConvertColor -> ShaderGetPixels + the Shaders/convert/*.hlsl include files) to compile as the text it is and run on the CPU:
//   * float2 / float3 / float4 with x y z w / r g b a and the swizzles the text uses, readable and ASSIGNABLE (color.rgb = ...,
//     sigX.xyz = ...), component indexing, the constructors float4(float, float2, float), float4(float3, float), scalar broadcast;
//   * componentwise + - * / against vectors and scalars (an int operand converts to float, as in HLSL), compound assignment;
//   * intrinsics with HLSL's meaning: saturate, frac, max, exp, dot, mul(vector, vector) = dot, mul(float3x3, float3) = one dot per row
//     (each dot: mul, mad, mad in component order), pow(x, y) as fxc lowers it = exp2(y * log2(x)), fmod;
//   * Texture2D::Sample with a point or a bilinear SamplerState (CLAMP addressing, optional integer texel offset).  The bilinear filter
//     is a MODEL of the D3D11 sampler, not reference code: texel fractions from u * size - 0.5, result a*(1-f) + b*f horizontally, then
//     vertically.  The chroma positions the reference's shaders ask for are multiples of a quarter texel, which D3D11's 8-bit
//     fraction represents exactly;
//   * `cond ? a : b` on a bool VECTOR cannot be spelled in C++: the generator rewrites the one such statement (hlg.hlsl:8-10) to
//     hlsl_select(cond, a, b), both sides evaluated for every lane, which is what fxc emits (movc).
// The unit is compiled with -fsingle-precision-constant -ffp-contract=off: unsuffixed literals are floats and every operation is one
// IEEE single-precision operation in the order the text states -- the meaning of the oracle's LITERAL mode.
#include <cmath>
#include <cstdlib>
#include <cstddef>
#include <cstring>
typedef unsigned int uint;
static inline float hlsl_pow(float x, float y) { return exp2f(y * log2f(x)); }
// D3D11 functional spec, min / max / saturate with NaN: max(NaN, b) = b, max(a, NaN) = a, saturate(NaN) = 0.  (The Dolby Vision trims
// produce NaN on pixels whose trimmed luminance is 0 -- 0 / 0 in DolbyVisionTrims, Shaders.cpp:787-795 -- and the following saturate /
// max turn it into 0 on the GPU.)
static inline float hlsl_max(float a, float b) { return (a >= b || b != b) ? a : b; }
static inline float hlsl_min(float a, float b) { return (a <= b || b != b) ? a : b; }
static inline float hlsl_saturate(float a) { return !(a > 0.0f) ? 0.0f : (a > 1.0f ? 1.0f : a); }

template<int N> struct vec;
template<int N, int... I> struct swz {                     // a view of components I... of the N floats it overlays (member of vec<N>'s union)
    float v[N];
    static constexpr int M = (int)sizeof...(I);
    operator vec<M>() const { vec<M> r; const int idx[M] = { I... }; for (int k = 0; k < M; k++) r.v[k] = v[idx[k]]; return r; }
    swz& operator=(const vec<M>& o) { const int idx[M] = { I... }; for (int k = 0; k < M; k++) v[idx[k]] = o.v[k]; return *this; }
    swz& operator=(const swz& o) { const vec<M> t = o; return *this = t; }
    swz& operator*=(float s) { const int idx[M] = { I... }; for (int k = 0; k < M; k++) v[idx[k]] = v[idx[k]] * s; return *this; }
};

template<> struct vec<2> {
    union { float v[2]; struct { float x, y; }; struct { float r, g; }; swz<2, 0, 0, 1, 1> xxyy; };
    vec() : v{0, 0} {}
    vec(float s) : v{s, s} {}
    vec(float a, float b) : v{a, b} {}
    vec(const vec& o) : v{o.v[0], o.v[1]} {}
    vec& operator=(const vec& o) { v[0] = o.v[0]; v[1] = o.v[1]; return *this; }
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};
template<> struct vec<3> {
    union { float v[3]; struct { float x, y, z; }; struct { float r, g, b; };
            swz<3, 0, 0, 1> xxy; swz<3, 1, 2, 2> yzz; swz<3, 0, 1> xy, rg; swz<3, 1, 2> yz; swz<3, 0, 1, 2> xyz, rgb; swz<3, 1, 0, 1, 2> yxyz; };
    vec() : v{0, 0, 0} {}
    vec(float s) : v{s, s, s} {}
    vec(float a, float b, float c) : v{a, b, c} {}
    vec(const vec& o) : v{o.v[0], o.v[1], o.v[2]} {}
    vec& operator=(const vec& o) { v[0] = o.v[0]; v[1] = o.v[1]; v[2] = o.v[2]; return *this; }
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};
template<> struct vec<4> {
    union { float v[4]; struct { float x, y, z, w; }; struct { float r, g, b, a; };
            swz<4, 0, 1> xy, rg; swz<4, 0, 2> xz; swz<4, 1, 3> yw; swz<4, 2, 0> zx;
            swz<4, 0, 1, 2> xyz, rgb; swz<4, 1, 0, 2> yxz; swz<4, 0, 1, 3> xyw; swz<4, 1, 2, 0> yzx; swz<4, 2, 1, 3> zyw;
            swz<4, 2, 1, 0, 3> zyxw; swz<4, 1, 0, 2, 3> yxzw; swz<4, 0, 1, 2, 3> xyzw; };
    vec() : v{0, 0, 0, 0} {}
    vec(float s) : v{s, s, s, s} {}
    vec(float a, float b, float c, float d) : v{a, b, c, d} {}
    vec(float a, const vec<2>& bc, float d) : v{a, bc.v[0], bc.v[1], d} {}
    vec(const vec<3>& abc, float d) : v{abc.v[0], abc.v[1], abc.v[2], d} {}
    vec(const vec& o) : v{o.v[0], o.v[1], o.v[2], o.v[3]} {}
    vec& operator=(const vec& o) { for (int k = 0; k < 4; k++) v[k] = o.v[k]; return *this; }
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};
typedef vec<2> float2; typedef vec<3> float3; typedef vec<4> float4;
struct int2 { int x, y; int2(int a, int b) : x(a), y(b) {} };
struct float3x3 {                                          // three float3 rows; built from 9 scalars (row-major) or from 3 rows, as HLSL allows
    float3 r[3];
    float3x3() {}
    float3x3(float a, float b, float c, float d, float e, float f, float g, float h, float i) : r{float3(a, b, c), float3(d, e, f), float3(g, h, i)} {}
    float3x3(const float3& a, const float3& b, const float3& c) : r{a, b, c} {}
    float3& operator[](int i) { return r[i]; }
    const float3& operator[](int i) const { return r[i]; }
};

#define HLSL_BINOP(N, OP) \
    static inline vec<N> operator OP(const vec<N>& a, const vec<N>& b) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = a.v[k] OP b.v[k]; return r; } \
    static inline vec<N> operator OP(const vec<N>& a, float b) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = a.v[k] OP b; return r; } \
    static inline vec<N> operator OP(float a, const vec<N>& b) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = a OP b.v[k]; return r; } \
    static inline vec<N>& operator OP##=(vec<N>& a, const vec<N>& b) { for (int k = 0; k < N; k++) a.v[k] = a.v[k] OP b.v[k]; return a; } \
    static inline vec<N>& operator OP##=(vec<N>& a, float b) { for (int k = 0; k < N; k++) a.v[k] = a.v[k] OP b; return a; }
#define HLSL_VEC(N) \
    HLSL_BINOP(N, +) HLSL_BINOP(N, -) HLSL_BINOP(N, *) HLSL_BINOP(N, /) \
    struct bool##N { bool v[N]; }; \
    static inline bool##N operator==(const vec<N>& a, const vec<N>& b) { bool##N r; for (int k = 0; k < N; k++) r.v[k] = a.v[k] == b.v[k]; return r; } \
    static inline vec<N> hlsl_sin(const vec<N>& a) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = sinf(a.v[k]); return r; } \
    static inline vec<N> hlsl_floor(const vec<N>& a) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = floorf(a.v[k]); return r; } \
    static inline vec<N> min(const vec<N>& a, const vec<N>& b) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = hlsl_min(a.v[k], b.v[k]); return r; } \
    static inline vec<N> clamp(const vec<N>& x, const vec<N>& lo, const vec<N>& hi) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = hlsl_min(hlsl_max(x.v[k], lo.v[k]), hi.v[k]); return r; } \
    static inline vec<N> lerp(const vec<N>& a, const vec<N>& b, float t) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = t * (b.v[k] - a.v[k]) + a.v[k]; return r; } \
    static inline bool##N operator<=(const vec<N>& a, float b) { bool##N r; for (int k = 0; k < N; k++) r.v[k] = a.v[k] <= b; return r; } \
    static inline vec<N> hlsl_select(const bool##N& c, const vec<N>& a, const vec<N>& b) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = c.v[k] ? a.v[k] : b.v[k]; return r; } \
    static inline vec<N> operator-(const vec<N>& a) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = -a.v[k]; return r; } \
    static inline vec<N> pow(const vec<N>& a, float b) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = hlsl_pow(a.v[k], b); return r; } \
    static inline vec<N> pow(const vec<N>& a, const vec<N>& b) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = hlsl_pow(a.v[k], b.v[k]); return r; } \
    static inline vec<N> max(const vec<N>& a, float b) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = hlsl_max(a.v[k], b); return r; } \
    static inline vec<N> max(const vec<N>& a, const vec<N>& b) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = hlsl_max(a.v[k], b.v[k]); return r; } \
    static inline vec<N> exp(const vec<N>& a) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = expf(a.v[k]); return r; } \
    static inline vec<N> frac(const vec<N>& a) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = a.v[k] - floorf(a.v[k]); return r; } \
    static inline vec<N> saturate(const vec<N>& a) { vec<N> r; for (int k = 0; k < N; k++) r.v[k] = hlsl_saturate(a.v[k]); return r; } \
    static inline float  dot(const vec<N>& a, const vec<N>& b) { float d = a.v[0] * b.v[0]; for (int k = 1; k < N; k++) d = a.v[k] * b.v[k] + d; return d; } \
    static inline float  mul(const vec<N>& a, const vec<N>& b) { return dot(a, b); }
HLSL_VEC(2) HLSL_VEC(3) HLSL_VEC(4)
static inline float  pow(float a, float b) { return hlsl_pow(a, b); }
static inline float  saturate(float a) { return hlsl_saturate(a); }
static inline float  frac(float a) { return a - floorf(a); }
static inline float  max(float a, float b) { return hlsl_max(a, b); }
static inline float  min(float a, float b) { return hlsl_min(a, b); }                              // D3D11: min(NaN, b) = b
static inline float  clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
static inline float  lerp(float a, float b, float t) { return t * (b - a) + a; }                  // fxc: mad(t, b - a, a)
static inline float3 mul(const float3x3& M, const float3& a)
{
    float3 r;
    for (int i = 0; i < 3; i++) r.v[i] = dot(M.r[i], a);
    return r;
}
static inline float3x3 transpose(const float3x3& M)
{
    float3x3 t;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t.r[i].v[j] = M.r[j].v[i];
    return t;
}
static inline float3x3 mul(const float3x3& A, const float3x3& B)                   // (A B)[i][j] = dot(row i of A, column j of B)
{
    const float3x3 Bt = transpose(B);
    float3x3 o;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o.r[i].v[j] = dot(A.r[i], Bt.r[j]);
    return o;
}

// matrices of ps_resize_onepass_jinc2.hlsl: float4x4 = four float4 rows, float4x3 = four float3 rows; mul(row vector, matrix) and
// mul(matrix, column vector), each component a mul followed by mads in row order
struct float4x4 {
    float4 r[4];
    float4x4() {}
    float4x4(const float4& a, const float4& b, const float4& c, const float4& d) : r{a, b, c, d} {}
    float4x4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3, float c0, float c1, float c2, float c3, float d0, float d1, float d2, float d3)
        : r{float4(a0, a1, a2, a3), float4(b0, b1, b2, b3), float4(c0, c1, c2, c3), float4(d0, d1, d2, d3)} {}
    float4& operator[](int i) { return r[i]; }
    const float4& operator[](int i) const { return r[i]; }
};
struct float4x3 { float3 r[4]; float4x3(const float3& a, const float3& b, const float3& c, const float3& d) : r{a, b, c, d} {} };
static inline float3 mul(const float4& w, const float4x3& M)
{
    float3 o;
    for (int k = 0; k < 3; k++) { float d = w.v[0] * M.r[0].v[k]; for (int i = 1; i < 4; i++) d = w.v[i] * M.r[i].v[k] + d; o.v[k] = d; }
    return o;
}
static inline float4 mul(const float4x4& M, const float4& a) { float4 o; for (int i = 0; i < 4; i++) o.v[i] = dot(M.r[i], a); return o; }
static inline float4x4 mul(const float4x4& A, const float4x4& B)
{
    float4x4 o;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) o.r[i].v[j] = dot(A.r[i], float4(B.r[0].v[j], B.r[1].v[j], B.r[2].v[j], B.r[3].v[j]));
    return o;
}

struct SamplerState { int linear = 0; };
struct hlsl_sample : float4 {                               // what Sample returns: a float4 that also truncates to its first component
    hlsl_sample(float a, float b, float c, float d) : float4(a, b, c, d) {}      // ("float y1 = texY.Sample(...)", Shaders.cpp:234-235)
    operator float() const { return v[0]; }
};
struct Texture2D {
    const float* data = nullptr; int tw = 0, th = 0;                          // RGBA float texels as the UNORM view returns them, row-major
    void bind(const float* d, int w_, int h_) { data = d; tw = w_; th = h_; }
    const float* texel(int ix, int iy) const
    {
        ix = ix < 0 ? 0 : (ix >= tw ? tw - 1 : ix); iy = iy < 0 ? 0 : (iy >= th ? th - 1 : iy);
        return data + ((size_t)iy * tw + ix) * 4;
    }
    hlsl_sample Sample(const SamplerState& s, const float2& uv, const int2& off = int2(0, 0)) const
    {
        if (!s.linear) {
            const float* p = texel((int)floorf(uv.v[0] * (float)tw) + off.x, (int)floorf(uv.v[1] * (float)th) + off.y);
            return hlsl_sample(p[0], p[1], p[2], p[3]);
        }
        const float px = uv.v[0] * (float)tw - 0.5f, py = uv.v[1] * (float)th - 0.5f;
        const float x0 = floorf(px), y0 = floorf(py), fx = px - x0, fy = py - y0;
        const int ix = (int)x0 + off.x, iy = (int)y0 + off.y;
        const float *c00 = texel(ix, iy), *c10 = texel(ix + 1, iy), *c01 = texel(ix, iy + 1), *c11 = texel(ix + 1, iy + 1);
        float r[4];
        for (int k = 0; k < 4; k++) {
            const float top = c00[k] * (1.0f - fx) + c10[k] * fx, bot = c01[k] * (1.0f - fx) + c11[k] * fx;
            r[k] = top * (1.0f - fy) + bot * fy;
        }
        return hlsl_sample(r[0], r[1], r[2], r[3]);
    }
};

// D3D9 flavour (Shaders/d3d9/*.hlsl): a `sampler` is a texture with its sampling state, read by tex2D(s, uv [, ddx, ddy])
struct sampler {
    const float* data = nullptr; int tw = 0, th = 0, wrap = 0;
    void bind(const float* d, int w_, int h_, int wrap_ = 0) { data = d; tw = w_; th = h_; wrap = wrap_; }
};
static inline float4 tex2D(const sampler& s, const float2& uv)
{
    int ix = (int)floorf(uv.v[0] * (float)s.tw), iy = (int)floorf(uv.v[1] * (float)s.th);
    if (s.wrap) { ix %= s.tw; if (ix < 0) ix += s.tw; iy %= s.th; if (iy < 0) iy += s.th; }
    else { ix = ix < 0 ? 0 : (ix >= s.tw ? s.tw - 1 : ix); iy = iy < 0 ? 0 : (iy >= s.th ? s.th - 1 : iy); }
    const float* p = s.data + ((size_t)iy * s.tw + ix) * 4;
    return float4(p[0], p[1], p[2], p[3]);
}
static inline float4 tex2D(const sampler& s, const float2& uv, const float2&, const float2&) { return tex2D(s, uv); }   // explicit gradients: no mip levels here

// scalar intrinsics whose C names would pick the double-precision C functions: HLSL's are single precision (libm's sinf, cosf, ...)
static inline float hlsl_sin(float x) { return sinf(x); }
static inline float hlsl_cos(float x) { return cosf(x); }
static inline float hlsl_acos(float x) { return acosf(x); }
static inline float hlsl_sqrt(float x) { return sqrtf(x); }
static inline float hlsl_floor(float x) { return floorf(x); }
static inline float hlsl_ceil(float x) { return ceilf(x); }
#define sin hlsl_sin
#define cos hlsl_cos
#define acos hlsl_acos
#define sqrt hlsl_sqrt
#define floor hlsl_floor
#define ceil hlsl_ceil
