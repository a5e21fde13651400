// glsl_shim.h — a minimal GLSL 4.50 compute "platform" in C++20 (TEST INFRASTRUCTURE).
//
// oracle/ref_glsl/build_ref.py transpiles the reference's OWN shader sources where they lie
// (/root/reference/assets/shaders/*.comp, *.glsl — never copied into this repo) into C++
// translation units under oracle/_ref/ and compiles them against this header into
// oracle/_ref/libddgi_ref.so.  The header supplies what a GLSL implementation supplies:
// vector types, built-in functions, images and the invocation id.  It contains no part of
// the reference's algorithm.
//
// Where GLSL leaves behaviour to the implementation this platform makes the choices the
// oracle pins (oracle/ddgi_oracle.c PINS): fp32 with no contraction; dot = (x*x + y*y) + z*z;
// normalize(v) = v * (1 / sqrt(dot(v,v))); min/max return the non-NaN operand; int(NaN) = 0,
// saturating; sin/cos/acos = the pinned fdlibm-style evaluations exported by the oracle
// library; rgba8 stores round(clamp(x,0,1)*255) with NaN -> 0.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef unsigned int uint;

// pinned transcendental functions live in the oracle library (one definition for both)
extern "C" void orc_pin_sincos(const float* x, int n, float* s, float* c);
extern "C" void orc_pin_acos(const float* x, int n, float* out);

namespace glsl {

// ------------------------------------------------------------------ scalars
inline float g_sin(float x) { float s, c; orc_pin_sincos(&x, 1, &s, &c); return s; }
inline float g_cos(float x) { float s, c; orc_pin_sincos(&x, 1, &s, &c); return c; }
inline float g_acos(float x) { float r; orc_pin_acos(&x, 1, &r); return r; }
inline float g_tan(float x) { return (float)::tan((double)x); }
inline float g_min(float a, float b) { if (b != b) return a; if (a != a) return b; return b < a ? b : a; }
inline float g_max(float a, float b) { if (b != b) return a; if (a != a) return b; return a < b ? b : a; }
inline int g_f2i(float x)
{
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x;
}
inline uint g_f2u(float x)
{
    if (x != x || x <= 0.0f) return 0u;
    if (x >= 4294967296.0f) return 0xffffffffu;
    return (uint)x;
}

}  // namespace glsl

// ------------------------------------------------------------------ vectors
struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec3; struct uvec3;

struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    explicit ivec2(int s) : x(s), y(s) {}
    ivec2(int a, int b) : x(a), y(b) {}
    ivec2(uint a, uint b) : x((int)a), y((int)b) {}
    explicit ivec2(const vec2& v);
    int& operator[](int i) { return i == 0 ? x : y; }
    int operator[](int i) const { return i == 0 ? x : y; }
};
struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    explicit ivec3(int s) : x(s), y(s), z(s) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
    explicit ivec3(const vec3& v);
    int& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    int operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct uvec3 {
    uint x, y, z;
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    struct uvec2v { uint x, y; };
    uvec2v xy() const { return {x, y}; }
};
struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    explicit vec2(int s) : x((float)s), y((float)s) {}
    explicit vec2(double s) : x((float)s), y((float)s) {}
    template <class A, class B> vec2(A a, B b) : x((float)a), y((float)b) {}
    explicit vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
    vec2(const uvec3::uvec2v& v) : x((float)v.x), y((float)v.y) {}
    float& operator[](int i) { return i == 0 ? x : y; }
    float operator[](int i) const { return i == 0 ? x : y; }
    vec2 xy() const { return *this; }
    vec2 yx() const { return vec2(y, x); }
    vec2 rg() const { return *this; }
};
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    explicit vec3(int s) : x((float)s), y((float)s), z((float)s) {}
    explicit vec3(double s) : x((float)s), y((float)s), z((float)s) {}
    template <class A, class B, class C> vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
    template <class C> vec3(const vec2& v, C c) : x(v.x), y(v.y), z((float)c) {}
    template <class A> vec3(A a, const vec2& v) : x((float)a), y(v.x), z(v.y) {}
    explicit vec3(const ivec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {}
    explicit vec3(const vec4& v);
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float r() const { return x; }
    vec2 xy() const { return vec2(x, y); }
    vec2 xz() const { return vec2(x, z); }
    vec2 yz() const { return vec2(y, z); }
    vec2 zy() const { return vec2(z, y); }
    vec2 rg() const { return vec2(x, y); }
    vec3 xyz() const { return *this; }
    vec3 rgb() const { return *this; }
    vec3 xzy() const { return vec3(x, z, y); }
    vec3 zyx() const { return vec3(z, y, x); }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    explicit vec4(int s) : x((float)s), y((float)s), z((float)s), w((float)s) {}
    explicit vec4(double s) : x((float)s), y((float)s), z((float)s), w((float)s) {}
    template <class A, class B, class C, class D> vec4(A a, B b, C c, D d) : x((float)a), y((float)b), z((float)c), w((float)d) {}
    template <class D> vec4(const vec3& v, D d) : x(v.x), y(v.y), z(v.z), w((float)d) {}
    template <class C, class D> vec4(const vec2& v, C c, D d) : x(v.x), y(v.y), z((float)c), w((float)d) {}
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    vec3 xyz() const { return vec3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
    vec2 xy() const { return vec2(x, y); }
    vec2 rg() const { return vec2(x, y); }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
inline ivec2::ivec2(const vec2& v) : x(glsl::g_f2i(v.x)), y(glsl::g_f2i(v.y)) {}
inline ivec3::ivec3(const vec3& v) : x(glsl::g_f2i(v.x)), y(glsl::g_f2i(v.y)), z(glsl::g_f2i(v.z)) {}

// componentwise arithmetic ------------------------------------------------------
#define GLSL_VEC_OPS(V, EXPR2, EXPRS, EXPRS_L)                                       \
    inline V operator+(const V& a, const V& b) { return EXPR2(+); }                  \
    inline V operator-(const V& a, const V& b) { return EXPR2(-); }                  \
    inline V operator*(const V& a, const V& b) { return EXPR2(*); }                  \
    inline V operator/(const V& a, const V& b) { return EXPR2(/); }                  \
    inline V operator+(const V& a, float s) { return EXPRS(+); }                     \
    inline V operator-(const V& a, float s) { return EXPRS(-); }                     \
    inline V operator*(const V& a, float s) { return EXPRS(*); }                     \
    inline V operator/(const V& a, float s) { return EXPRS(/); }                     \
    inline V operator+(float s, const V& a) { return EXPRS_L(+); }                   \
    inline V operator-(float s, const V& a) { return EXPRS_L(-); }                   \
    inline V operator*(float s, const V& a) { return EXPRS_L(*); }                   \
    inline V operator/(float s, const V& a) { return EXPRS_L(/); }                   \
    inline V& operator+=(V& a, const V& b) { a = a + b; return a; }                  \
    inline V& operator-=(V& a, const V& b) { a = a - b; return a; }                  \
    inline V& operator*=(V& a, const V& b) { a = a * b; return a; }                  \
    inline V& operator/=(V& a, const V& b) { a = a / b; return a; }                  \
    inline V& operator+=(V& a, float s) { a = a + s; return a; }                     \
    inline V& operator-=(V& a, float s) { a = a - s; return a; }                     \
    inline V& operator*=(V& a, float s) { a = a * s; return a; }                     \
    inline V& operator/=(V& a, float s) { a = a / s; return a; }

#define E2_2(op) vec2(a.x op b.x, a.y op b.y)
#define ES_2(op) vec2(a.x op s, a.y op s)
#define EL_2(op) vec2(s op a.x, s op a.y)
GLSL_VEC_OPS(vec2, E2_2, ES_2, EL_2)
#define E2_3(op) vec3(a.x op b.x, a.y op b.y, a.z op b.z)
#define ES_3(op) vec3(a.x op s, a.y op s, a.z op s)
#define EL_3(op) vec3(s op a.x, s op a.y, s op a.z)
GLSL_VEC_OPS(vec3, E2_3, ES_3, EL_3)
#define E2_4(op) vec4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w)
#define ES_4(op) vec4(a.x op s, a.y op s, a.z op s, a.w op s)
#define EL_4(op) vec4(s op a.x, s op a.y, s op a.z, s op a.w)
GLSL_VEC_OPS(vec4, E2_4, ES_4, EL_4)
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const vec3& a, const vec3& b) { return !(a == b); }
inline bool operator==(const vec2& a, const vec2& b) { return a.x == b.x && a.y == b.y; }

inline ivec2 operator+(const ivec2& a, const ivec2& b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(const ivec2& a, const ivec2& b) { return ivec2(a.x - b.x, a.y - b.y); }
inline ivec2 operator*(const ivec2& a, int s) { return ivec2(a.x * s, a.y * s); }
inline ivec2 operator*(int s, const ivec2& a) { return ivec2(a.x * s, a.y * s); }
inline ivec2 operator/(const ivec2& a, int s) { return ivec2(a.x / s, a.y / s); }
inline bool operator==(const ivec2& a, const ivec2& b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(const ivec2& a, const ivec2& b) { return !(a == b); }
inline ivec3 operator+(const ivec3& a, const ivec3& b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator-(const ivec3& a, const ivec3& b) { return ivec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline ivec3 operator*(const ivec3& a, const ivec3& b) { return ivec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline ivec3 operator*(const ivec3& a, int s) { return ivec3(a.x * s, a.y * s, a.z * s); }
inline ivec3 operator*(int s, const ivec3& a) { return ivec3(a.x * s, a.y * s, a.z * s); }
inline ivec3 operator/(const ivec3& a, int s) { return ivec3(a.x / s, a.y / s, a.z / s); }
inline ivec3 operator+(const ivec3& a, int s) { return ivec3(a.x + s, a.y + s, a.z + s); }
inline ivec3 operator-(const ivec3& a, int s) { return ivec3(a.x - s, a.y - s, a.z - s); }
inline ivec3 operator&(const ivec3& a, const ivec3& b) { return ivec3(a.x & b.x, a.y & b.y, a.z & b.z); }
inline ivec3 operator-(const ivec3& a) { return ivec3(-a.x, -a.y, -a.z); }
inline bool operator==(const ivec3& a, const ivec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline ivec2& operator+=(ivec2& a, const ivec2& b) { a = a + b; return a; }
inline ivec2& operator-=(ivec2& a, const ivec2& b) { a = a - b; return a; }
inline ivec3& operator+=(ivec3& a, const ivec3& b) { a = a + b; return a; }
// implicit int -> float promotion of GLSL: an integer vector meeting a float operand becomes a vec3
#define GLSL_IVEC3_PROMOTE(op)                                                                   \
    inline vec3 operator op(const ivec3& a, float s) { return vec3(a) op s; }                    \
    inline vec3 operator op(float s, const ivec3& a) { return s op vec3(a); }                    \
    inline vec3 operator op(const ivec3& a, const vec3& b) { return vec3(a) op b; }              \
    inline vec3 operator op(const vec3& a, const ivec3& b) { return a op vec3(b); }
GLSL_IVEC3_PROMOTE(+)
GLSL_IVEC3_PROMOTE(-)
GLSL_IVEC3_PROMOTE(*)
GLSL_IVEC3_PROMOTE(/)
inline vec2 operator*(const ivec2& a, float s) { return vec2(a) * s; }
inline vec2 operator/(const ivec2& a, float s) { return vec2(a) / s; }
inline vec2 operator+(const ivec2& a, const vec2& b) { return vec2(a) + b; }
inline vec2 operator/(const vec2& a, const ivec2& b) { return a / vec2(b); }
inline vec2 operator*(const vec2& a, const ivec2& b) { return a * vec2(b); }
inline vec2 operator+(const vec2& a, const ivec2& b) { return a + vec2(b); }

struct bvec3 {
    bool x, y, z;
    bvec3(bool a, bool b, bool c) : x(a), y(b), z(c) {}
};
inline bvec3 equal(const vec3& a, const vec3& b) { return bvec3(a.x == b.x, a.y == b.y, a.z == b.z); }
inline bvec3 notEqual(const vec3& a, const vec3& b) { return bvec3(a.x != b.x, a.y != b.y, a.z != b.z); }
inline bvec3 lessThan(const vec3& a, const vec3& b) { return bvec3(a.x < b.x, a.y < b.y, a.z < b.z); }
inline bvec3 greaterThan(const vec3& a, const vec3& b) { return bvec3(a.x > b.x, a.y > b.y, a.z > b.z); }
inline bvec3 lessThanEqual(const vec3& a, const vec3& b) { return bvec3(a.x <= b.x, a.y <= b.y, a.z <= b.z); }
inline bvec3 greaterThanEqual(const vec3& a, const vec3& b) { return bvec3(a.x >= b.x, a.y >= b.y, a.z >= b.z); }
inline bool all(const bvec3& b) { return b.x && b.y && b.z; }
inline bool any(const bvec3& b) { return b.x || b.y || b.z; }

// matrices (column major) ----------------------------------------------------------
struct mat4 {
    vec4 c[4];
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v)
{
    // GLSL mat * vec: linear combination of the columns, left to right
    return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w;
}
struct mat2 {
    vec2 c[2];
    mat2() {}
    mat2(float a, float b, float d, float e) { c[0] = vec2(a, b); c[1] = vec2(d, e); }
    vec2& operator[](int i) { return c[i]; }
    const vec2& operator[](int i) const { return c[i]; }
};
inline vec2 operator*(const mat2& m, const vec2& v) { return m.c[0] * v.x + m.c[1] * v.y; }
struct mat3 {
    vec3 c[3];
    mat3() {}
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline vec3 operator*(const mat3& m, const vec3& v) { return (m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z; }
inline mat3 inverse(const mat3& m)
{
    // cofactor expansion (only the dead triangle intersection of the inherited path tracer uses it)
    const vec3 &a = m.c[0], &b = m.c[1], &d = m.c[2];
    float det = a.x * (b.y * d.z - d.y * b.z) - b.x * (a.y * d.z - d.y * a.z) + d.x * (a.y * b.z - b.y * a.z);
    float id = 1.0f / det;
    mat3 r;
    r.c[0] = vec3((b.y * d.z - d.y * b.z) * id, -(a.y * d.z - d.y * a.z) * id, (a.y * b.z - b.y * a.z) * id);
    r.c[1] = vec3(-(b.x * d.z - d.x * b.z) * id, (a.x * d.z - d.x * a.z) * id, -(a.x * b.z - b.x * a.z) * id);
    r.c[2] = vec3((b.x * d.y - d.x * b.y) * id, -(a.x * d.y - d.x * a.y) * id, (a.x * b.y - b.x * a.y) * id);
    return r;
}

// ------------------------------------------------------------------ built-in functions
#define GLSL_MAP1(name, fn)                                                                  \
    inline vec2 name(const vec2& v) { return vec2(fn(v.x), fn(v.y)); }                       \
    inline vec3 name(const vec3& v) { return vec3(fn(v.x), fn(v.y), fn(v.z)); }              \
    inline vec4 name(const vec4& v) { return vec4(fn(v.x), fn(v.y), fn(v.z), fn(v.w)); }

namespace glsl {
inline float s_fract(float x) { return x - floorf(x); }
inline float s_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float s_round(float x) { return roundf(x); }
inline float s_isqrt(float x) { return 1.0f / sqrtf(x); }
inline float s_abs(float x) { return fabsf(x); }
}  // namespace glsl

inline float glsl_sin(float x) { return glsl::g_sin(x); }
inline float glsl_cos(float x) { return glsl::g_cos(x); }
inline float glsl_tan(float x) { return glsl::g_tan(x); }
inline float glsl_acos(float x) { return glsl::g_acos(x); }
inline float glsl_asin(float x) { return (float)::asin((double)x); }
inline float glsl_atan(float y, float x) { return (float)::atan2((double)y, (double)x); }
inline float glsl_atan(float x) { return (float)::atan((double)x); }
inline float glsl_sqrt(float x) { return sqrtf(x); }
inline float inversesqrt(float x) { return glsl::s_isqrt(x); }
inline float glsl_pow(float a, float b) { return powf(a, b); }
// pow with a small non-negative integer exponent (pow(2.f, i), pow(0.5, i), pow(w, 3)): the product
// in fp64 rounded once to fp32 (oracle PIN 11); exact for the powers of two the noise code asks for
inline float glsl_pow(float a, int b)
{
    if (b < 0 || b > 64) return powf(a, (float)b);
    double r = 1.0;
    for (int i = 0; i < b; i++) r = r * (double)a;
    return (float)r;
}
inline float glsl_exp(float x) { return expf(x); }
inline float glsl_log(float x) { return logf(x); }
inline float glsl_exp2(float x) { return exp2f(x); }
inline float glsl_log2(float x) { return log2f(x); }
inline float glsl_floor(float x) { return floorf(x); }
inline float glsl_ceil(float x) { return ceilf(x); }
inline float glsl_round(float x) { return glsl::s_round(x); }
inline float glsl_trunc(float x) { return truncf(x); }
inline float fract(float x) { return glsl::s_fract(x); }
inline float sign(float x) { return glsl::s_sign(x); }
inline float glsl_abs(float x) { return fabsf(x); }
inline int glsl_abs(int x) { return x < 0 ? -x : x; }
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }
inline float degrees(float r) { return r * 57.295779513082320876798154814105f; }
inline float mod(float x, float y) { return x - y * floorf(x / y); }
inline float min(float a, float b) { return glsl::g_min(a, b); }
inline float max(float a, float b) { return glsl::g_max(a, b); }
inline float min(float a, int b) { return glsl::g_min(a, (float)b); }
inline float max(float a, int b) { return glsl::g_max(a, (float)b); }
inline float min(int a, float b) { return glsl::g_min((float)a, b); }
inline float max(int a, float b) { return glsl::g_max((float)a, b); }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x)
{
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline bool glsl_isnan(float x) { return x != x; }
inline bool glsl_isinf(float x) { return x == INFINITY || x == -INFINITY; }

GLSL_MAP1(glsl_sin, glsl_sin)
GLSL_MAP1(glsl_cos, glsl_cos)
GLSL_MAP1(glsl_sqrt, sqrtf)
GLSL_MAP1(glsl_floor, floorf)
GLSL_MAP1(glsl_ceil, ceilf)
GLSL_MAP1(glsl_round, glsl::s_round)
GLSL_MAP1(fract, glsl::s_fract)
GLSL_MAP1(sign, glsl::s_sign)
GLSL_MAP1(glsl_abs, glsl::s_abs)
GLSL_MAP1(glsl_exp, expf)

#define GLSL_MAP2(name, fn)                                                                                  \
    inline vec2 name(const vec2& a, const vec2& b) { return vec2(fn(a.x, b.x), fn(a.y, b.y)); }              \
    inline vec3 name(const vec3& a, const vec3& b) { return vec3(fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z)); } \
    inline vec4 name(const vec4& a, const vec4& b) { return vec4(fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z), fn(a.w, b.w)); } \
    inline vec2 name(const vec2& a, float b) { return vec2(fn(a.x, b), fn(a.y, b)); }                        \
    inline vec3 name(const vec3& a, float b) { return vec3(fn(a.x, b), fn(a.y, b), fn(a.z, b)); }            \
    inline vec4 name(const vec4& a, float b) { return vec4(fn(a.x, b), fn(a.y, b), fn(a.z, b), fn(a.w, b)); }
GLSL_MAP2(min, glsl::g_min)
GLSL_MAP2(max, glsl::g_max)
GLSL_MAP2(mod, mod)
GLSL_MAP2(glsl_pow, powf)
inline vec2 clamp(const vec2& v, float lo, float hi) { return min(max(v, lo), hi); }
inline vec3 clamp(const vec3& v, float lo, float hi) { return min(max(v, lo), hi); }
inline vec4 clamp(const vec4& v, float lo, float hi) { return min(max(v, lo), hi); }
inline vec3 clamp(const vec3& v, const vec3& lo, const vec3& hi) { return min(max(v, lo), hi); }
inline vec2 clamp(const vec2& v, const vec2& lo, const vec2& hi) { return min(max(v, lo), hi); }
inline vec2 mix(const vec2& a, const vec2& b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return a * (1.0f - t) + b * t; }
inline vec4 mix(const vec4& a, const vec4& b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(const vec3& a, const vec3& b, const vec3& t) { return a * (vec3(1.0f) - t) + b * t; }
inline vec3 mix(const vec3& a, const vec3& b, const ivec3& t) { return mix(a, b, vec3(t)); }
inline vec3 glsl_floor(const ivec3& v) { return vec3(v); }  // floor(ivec3): promoted to vec3, already integral
inline vec2 mix(const vec2& a, const vec2& b, const vec2& t) { return a * (vec2(1.0f) - t) + b * t; }
inline vec3 step(float e, const vec3& v) { return vec3(step(e, v.x), step(e, v.y), step(e, v.z)); }
inline vec3 step(const vec3& e, const vec3& v) { return vec3(step(e.x, v.x), step(e.y, v.y), step(e.z, v.z)); }
inline vec3 smoothstep(float a, float b, const vec3& v) { return vec3(smoothstep(a, b, v.x), smoothstep(a, b, v.y), smoothstep(a, b, v.z)); }

inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline float length(const vec2& a) { return sqrtf(dot(a, a)); }
inline float length(const vec3& a) { return sqrtf(dot(a, a)); }
inline float length(const vec4& a) { return sqrtf(dot(a, a)); }
inline float length(float a) { return fabsf(a); }
inline float distance(const vec2& a, const vec2& b) { return length(a - b); }
inline float distance(const vec3& a, const vec3& b) { return length(a - b); }
inline vec2 normalize(const vec2& a) { return a * (1.0f / sqrtf(dot(a, a))); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / sqrtf(dot(a, a))); }
inline vec4 normalize(const vec4& a) { return a * (1.0f / sqrtf(dot(a, a))); }
inline vec3 cross(const vec3& a, const vec3& b)
{
    return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline vec3 reflect(const vec3& i, const vec3& n) { return i - 2.0f * dot(n, i) * n; }
inline vec3 refract(const vec3& i, const vec3& n, float eta)
{
    float k = 1.0f - eta * eta * (1.0f - dot(n, i) * dot(n, i));
    if (k < 0.0f) return vec3(0.0f);
    return eta * i - (eta * dot(n, i) + sqrtf(k)) * n;
}
inline vec3 faceforward(const vec3& n, const vec3& i, const vec3& nref) { return dot(nref, i) < 0.0f ? n : -n; }
inline bool any(bool b) { return b; }
inline bool all(bool b) { return b; }

// scalar constructors: int(x), uint(x) with the pinned conversion; int(vecN) = first component
inline int glsl_int(float x) { return glsl::g_f2i(x); }
inline int glsl_int(double x) { return glsl::g_f2i((float)x); }
inline int glsl_int(int x) { return x; }
inline int glsl_int(uint x) { return (int)x; }
inline int glsl_int(bool x) { return x ? 1 : 0; }
inline int glsl_int(const vec2& v) { return glsl::g_f2i(v.x); }
inline int glsl_int(const vec3& v) { return glsl::g_f2i(v.x); }
inline int glsl_int(const vec4& v) { return glsl::g_f2i(v.x); }
inline int glsl_int(const ivec3& v) { return v.x; }
inline uint glsl_uint(float x) { return glsl::g_f2u(x); }
inline uint glsl_uint(int x) { return (uint)x; }
inline uint glsl_uint(uint x) { return x; }

// ------------------------------------------------------------------ images (rgba8)
struct image2D {
    int width = 0, height = 0;
    uint32_t* data = nullptr;   // RGBA8, row major; out-of-range accesses are dropped / read as 0
    float* f32 = nullptr;       // optional: the value handed to imageStore before quantisation
};
inline ivec2 imageSize(const image2D& img) { return ivec2(img.width, img.height); }
inline vec4 imageLoad(const image2D& img, const ivec2& p)
{
    if (p.x < 0 || p.y < 0 || p.x >= img.width || p.y >= img.height || !img.data) return vec4(0.0f);
    uint32_t v = img.data[(size_t)p.y * img.width + p.x];
    return vec4((float)(v & 255u) / 255.0f, (float)((v >> 8) & 255u) / 255.0f, (float)((v >> 16) & 255u) / 255.0f,
                (float)(v >> 24) / 255.0f);
}
inline uint32_t glsl_unorm8(float x)
{
    if (x != x) return 0u;
    float c = min(max(x, 0.0f), 1.0f);
    return (uint32_t)floorf(c * 255.0f + 0.5f);
}
inline void imageStore(image2D& img, const ivec2& p, const vec4& v)
{
    if (p.x < 0 || p.y < 0 || p.x >= img.width || p.y >= img.height || !img.data) return;
    size_t at = (size_t)p.y * img.width + p.x;
    img.data[at] = glsl_unorm8(v.x) | (glsl_unorm8(v.y) << 8) | (glsl_unorm8(v.z) << 16) | (glsl_unorm8(v.w) << 24);
    if (img.f32) { img.f32[4 * at] = v.x; img.f32[4 * at + 1] = v.y; img.f32[4 * at + 2] = v.z; img.f32[4 * at + 3] = v.w; }
}

// ------------------------------------------------------------------ invocation state
extern uvec3 gl_GlobalInvocationID;
// voxel lookups of the current invocation: build_ref.py adds `glsl_count_lookup();` as the
// first statement of getBlockAt so the harness can report the reference's own step count
extern uint32_t glsl_lookup_counter;
inline void glsl_count_lookup() { glsl_lookup_counter++; }

// GLSL built-in names that collide with <math.h> / <stdlib.h> are defined as glsl_<name> above
// and mapped here; this header must therefore be included after every system header.
#define sin glsl_sin
#define cos glsl_cos
#define tan glsl_tan
#define acos glsl_acos
#define asin glsl_asin
#define atan glsl_atan
#define sqrt glsl_sqrt
#define pow glsl_pow
#define exp glsl_exp
#define log glsl_log
#define exp2 glsl_exp2
#define log2 glsl_log2
#define floor glsl_floor
#define ceil glsl_ceil
#define round glsl_round
#define trunc glsl_trunc
#define abs glsl_abs
#define isnan glsl_isnan
#define isinf glsl_isinf
