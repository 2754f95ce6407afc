// TEST INFRASTRUCTURE ONLY.  What the reference's ray-generation / hit / miss / intersection shaders need beyond
// glsl_prelude.h to compile as C++: the remaining vector types, image / buffer / acceleration-structure stand-ins and the
// gl_* built-in variables of the ray-tracing stages.  oracle/shader_ref.cpp owns the objects declared here.
#pragma once
#include "glsl_prelude.h"
#include <type_traits>
namespace glsl {
static inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
static inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
static inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
static inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
static inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
static inline vec2 operator*(float s, vec2 a) { return vec2(a.x * s, a.y * s); }
static inline vec2 operator/(vec2 a, float s) { return vec2(a.x / s, a.y / s); }
static inline vec2 operator*(vec2 a, int s) { return vec2(a.x * float(s), a.y * float(s)); }
static inline vec2 operator+(vec2 a, float s) { return vec2(a.x + s, a.y + s); }
template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> static inline vec3 operator/(vec3 a, T s) { return a / float(s); }
template <class T, class = typename std::enable_if<std::is_integral<T>::value>::type> static inline vec3 operator*(T s, vec3 a) { return a * float(s); }
static inline vec3 &operator/=(vec3 &a, float s) { a = a / s; return a; }
static inline vec3 &operator/=(vec3 &a, int s) { a = a / float(s); return a; }
static inline vec3 &operator*=(vec3 &a, vec3 b) { a = a * b; return a; }
static inline vec3 &operator-=(vec3 &a, vec3 b) { a = a - b; return a; }
static inline bool operator==(vec3 a, vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
static inline vec3 min(vec3 a, vec3 b) { return vec3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
static inline vec3 max(vec3 a, vec3 b) { return vec3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }           // GLSL.std.450 FMix: x * (1 - a) + y * a
static inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
template <class T> static inline float asin(T x) { return GLSL_M(asinF, asinf)(float(x)); }
template <class A, class B> static inline float atan(A y, B x) { return GLSL_M(atan2F, atan2f)(float(y), float(x)); }
static inline float length(vec2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
struct vec4 {
    union { struct { float x, y, z, w; }; xyz_t xyz; };
    vec4() = default;
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
};
struct ivec2 { int x, y; ivec2() = default; ivec2(int a, int b) : x(a), y(b) {} ivec2(uxy_t u) : x(int(u.x)), y(int(u.y)) {} };
struct ivec3 { int x, y, z; ivec3() = default; ivec3(uint a, uint b, uint c) : x(int(a)), y(int(b)), z(int(c)) {} };
struct uvec3 { union { struct { uint x, y, z; }; uxy_t xy; }; };
// mat4 (glsl_prelude.h) is column-major like GLSL: element (row r, column c) = m[4 * c + r]
static inline vec4 operator*(const mat4 &M, vec4 v) {
    vec4 r;
    r.x = M.m[0] * v.x + M.m[4] * v.y + M.m[8] * v.z + M.m[12] * v.w;
    r.y = M.m[1] * v.x + M.m[5] * v.y + M.m[9] * v.z + M.m[13] * v.w;
    r.z = M.m[2] * v.x + M.m[6] * v.y + M.m[10] * v.z + M.m[14] * v.w;
    r.w = M.m[3] * v.x + M.m[7] * v.y + M.m[11] * v.z + M.m[15] * v.w;
    return r;
}
template <class T> static inline uint atomicAdd(T &x, int v) { const uint old = uint(x); x = T(x + v); return old; }
#define nonuniformEXT(x) (x)
// images: RGBA32F, row-major
struct image2D { float *px; int width; };
static inline vec4 imageLoad(image2D img, ivec2 p) { const float *q = img.px + (size_t(p.y) * img.width + p.x) * 4; return vec4(q[0], q[1], q[2], q[3]); }
static inline void imageStore(image2D img, ivec2 p, vec4 v) { float *q = img.px + (size_t(p.y) * img.width + p.x) * 4; q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w; }
struct accelerationStructureEXT { int id; };
static const uint gl_RayFlagsOpaqueEXT = 1u;
}  // namespace glsl
