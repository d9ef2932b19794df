// oracle/ref_stub/hlsl_compat.h -- TEST INFRASTRUCTURE ONLY.
// Just enough of HLSL's vector types and intrinsics, with HLSL's meaning, for the reference's own shader include files
// (Shaders/convert/st2084.hlsl, hlg.hlsl, hdr_tone_mapping.hlsl) to compile AS THEY ARE as C++ and be executed on the CPU:
//   * float3 / float4 with componentwise arithmetic against vectors and scalars; unsuffixed literals are floats in HLSL, so a double
//     operand is narrowed to float before it takes part in anything;
//   * pow(x, y) is what fxc emits for it: exp2(y * log2(x)) (LOG, MUL, EXP); exp, exp2, log2 are the C library's single-precision
//     functions; max, dot, saturate as defined by HLSL;
//   * a comparison of vectors yields a bool vector; `cond ? a : b` on vectors cannot be overloaded in C++, so the bool vector converts
//     to bool only when all its lanes agree (checked: the tests feed such vectors) -- the only construct that needs this is the
//     branch of inverse_HLG.
// Compiled with -ffp-contract=off: every operation is one IEEE single-precision operation, in the order the shader text states.
#include <cmath>
#include <cstdlib>
static inline float hlsl_pow(float x, float y) { return exp2f(y * log2f(x)); }
struct float4 {
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(double v) : x((float)v), y((float)v), z((float)v), w((float)v) {}
    float4(double a, double b, double c, double d) : x((float)a), y((float)b), z((float)c), w((float)d) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    float4& operator+=(const float4& o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
    float4& operator*=(const float4& o) { x *= o.x; y *= o.y; z *= o.z; w *= o.w; return *this; }
    float4& operator/=(const float4& o) { x /= o.x; y /= o.y; z /= o.z; w /= o.w; return *this; }
};
static inline float4 operator+(const float4& a, const float4& b) { float4 r; r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z; r.w = a.w + b.w; return r; }
static inline float4 operator-(const float4& a, const float4& b) { float4 r; r.x = a.x - b.x; r.y = a.y - b.y; r.z = a.z - b.z; r.w = a.w - b.w; return r; }
static inline float4 operator*(const float4& a, const float4& b) { float4 r; r.x = a.x * b.x; r.y = a.y * b.y; r.z = a.z * b.z; r.w = a.w * b.w; return r; }
static inline float4 operator/(const float4& a, const float4& b) { float4 r; r.x = a.x / b.x; r.y = a.y / b.y; r.z = a.z / b.z; r.w = a.w / b.w; return r; }
static inline float4 pow(const float4& a, const float4& b) { float4 r; r.x = hlsl_pow(a.x, b.x); r.y = hlsl_pow(a.y, b.y); r.z = hlsl_pow(a.z, b.z); r.w = hlsl_pow(a.w, b.w); return r; }
static inline float4 max(const float4& a, const float4& b) { float4 r; r.x = a.x > b.x ? a.x : b.x; r.y = a.y > b.y ? a.y : b.y; r.z = a.z > b.z ? a.z : b.z; r.w = a.w > b.w ? a.w : b.w; return r; }
struct bool3 { bool x, y, z; operator bool() const { if (x != y || y != z) abort(); return x; } };
struct float3 {
    float x, y, z;
    float4 yxyz;                    // the swizzle ps_interpolation_lanczos2.hlsl takes of a freshly constructed float3(a, b, c)
    float3() : x(0), y(0), z(0) {}
    float3(double v) : x((float)v), y((float)v), z((float)v) {}
    float3(double a, double b, double c) : x((float)a), y((float)b), z((float)c), yxyz(b, a, b, c) {}
    float3& operator*=(const float3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    float3& operator/=(const float3& o) { x /= o.x; y /= o.y; z /= o.z; return *this; }
};
static inline float3 operator+(const float3& a, const float3& b) { float3 r; r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z; return r; }
static inline float3 operator-(const float3& a, const float3& b) { float3 r; r.x = a.x - b.x; r.y = a.y - b.y; r.z = a.z - b.z; return r; }
static inline float3 operator*(const float3& a, const float3& b) { float3 r; r.x = a.x * b.x; r.y = a.y * b.y; r.z = a.z * b.z; return r; }
static inline float3 operator/(const float3& a, const float3& b) { float3 r; r.x = a.x / b.x; r.y = a.y / b.y; r.z = a.z / b.z; return r; }
static inline bool3  operator<=(const float3& a, const float3& b) { return bool3{ a.x <= b.x, a.y <= b.y, a.z <= b.z }; }
static inline float3 pow(const float3& a, const float3& b) { float3 r; r.x = hlsl_pow(a.x, b.x); r.y = hlsl_pow(a.y, b.y); r.z = hlsl_pow(a.z, b.z); return r; }
static inline float3 max(const float3& a, const float3& b) { float3 r; r.x = a.x > b.x ? a.x : b.x; r.y = a.y > b.y ? a.y : b.y; r.z = a.z > b.z ? a.z : b.z; return r; }
static inline float3 exp(const float3& a) { float3 r; r.x = expf(a.x); r.y = expf(a.y); r.z = expf(a.z); return r; }
static inline float  dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // DP3: mul, mad, mad -- unfused here (LITERAL)
static inline float  pow(float a, float b) { return hlsl_pow(a, b); }
