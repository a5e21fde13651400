// glm/glm.hpp — TEST-ONLY stand-in for the few glm 0.9.9.8 types and functions the reference's
// HOST code on the path uses (ray generator src/rvpt/rvpt.cpp:1145-1224, src/rvpt/probe.h, and the vec4 list
// RVPT::update hands to the camera uniform), so that its own text can be compiled where it lies
// (oracle/ref_glsl/build_ref.py: ref_generate_probe_rays; build_shim.py: rvpt_shim_main).  glm is not
// in this image (external/CMakeLists.txt:21-25 fetches it from the network).  Semantics follow
// glm's published definitions: component-wise operators, int -> float conversion on construction,
// normalize(v) = v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x),
// dot(a, b) = (a.x*b.x + a.y*b.y) + a.z*b.z.
#pragma once
#include <cmath>

namespace glm {

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    template <typename A, typename B>
    vec2(A a, B b) : x(static_cast<float>(a)), y(static_cast<float>(b))
    {
    }
};

struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    explicit ivec3(int s) : x(s), y(s), z(s) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
};
inline ivec3 operator-(const ivec3& a, const ivec3& b) { return ivec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline ivec3 operator/(const ivec3& a, int s) { return ivec3(a.x / s, a.y / s, a.z / s); }

struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    template <typename A, typename B, typename C>
    vec3(A a, B b, C c) : x(static_cast<float>(a)), y(static_cast<float>(b)), z(static_cast<float>(c))
    {
    }
    vec3(const ivec3& v) : x(static_cast<float>(v.x)), y(static_cast<float>(v.y)), z(static_cast<float>(v.z)) {}
    template <typename S>
    vec3& operator*=(S s)
    {
        x *= static_cast<float>(s);
        y *= static_cast<float>(s);
        z *= static_cast<float>(s);
        return *this;
    }
    vec3& operator+=(const vec3& o)
    {
        x += o.x;
        y += o.y;
        z += o.z;
        return *this;
    }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }

}  // namespace glm
