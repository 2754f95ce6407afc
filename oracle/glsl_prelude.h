// TEST INFRASTRUCTURE ONLY.  The GLSL the reference's include files are written in (shaders/random.glsl, transform.glsl,
// guiding.glsl, raycommon.glsl, wavefront.glsl, limits.glsl) is C-like enough to be compiled as C++ once the vector types and
// built-ins exist: this header supplies them (float semantics: every built-in on float calls the float libm function).
// oracle/Makefile pipes the shader files through sed (parameter qualifiers `in` / `out` / `inout` -> C++ references, #include
// lines dropped) into oracle/_ref/glsl/*.inc — build intermediates, deleted after the compile, never copied into this repository —
// and oracle/glsl_ref.cpp includes them where they are generated.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#undef M_PI
// The elementary functions are the driver's, not the reference's.  Default: the C library (an independent check of
// include/b200pt_detmath.h as well).  GLSL_DETMATH: the deterministic functions the oracle and the kernels use, so that a whole
// frame of the compiled shaders can be compared with the oracle bit for bit — any difference left is logic or operation order.
#ifdef GLSL_DETMATH
#include "../include/b200pt_detmath.h"
#define GLSL_M(det, libm) b200pt_dm::det
#else
#define GLSL_M(det, libm) libm
#endif
namespace glsl {      // the built-ins below must hide the C library's overloads of the same names, not compete with them
typedef uint32_t uint;
struct uxy_t { uint x, y; };            // gl_LaunchIDEXT.xy
struct vec2 {
    float x, y;
    vec2() = default; vec2(float a) : x(a), y(a) {} vec2(float a, float b) : x(a), y(b) {} vec2(uxy_t u) : x(float(u.x)), y(float(u.y)) {}
    // uint(vec2) takes the first component (raytrace.rahit:39).  Out-of-range float -> uint is undefined in GLSL; like the oracle and the
    // kernels it saturates here (negative and NaN -> 0)
    explicit operator uint() const { return x >= 4294967296.0f ? 0xFFFFFFFFu : (x > 0.0f ? uint(x) : 0u); }
};
struct vec3;
struct xyz_t { float x, y, z; inline operator vec3() const; };          // `.xyz` of a vec3 / vec4 / texel
struct vec3 {
    union { struct { float x, y, z; }; xyz_t xyz; };
    vec3() = default;      // trivial: a `switch` of the shaders may jump over `vec3 v;` (GLSL scoping), which C++ only allows for trivially constructible types
    vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3 &operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    vec3 &operator+=(vec3 o) { x += o.x; y += o.y; z += o.z; return *this; }
};
inline xyz_t::operator vec3() const { return vec3(x, y, z); }
struct mat4 { float m[16]; };
static inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
static inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
static inline vec3 operator*(float s, vec3 a) { return vec3(a.x * s, a.y * s, a.z * s); }
static inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
static inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
static inline float length(vec3 a) { return sqrtf(dot(a, a)); }
static inline vec3 normalize(vec3 a) { return a / sqrtf(dot(a, a)); }                   // built-ins are the driver's, not the reference's: same definition as the oracle and the kernels
static inline vec3 reflect(vec3 i, vec3 n) { return i - 2.0f * dot(n, i) * n; }
// built-ins: GLSL computes in float whatever mixture of float variables and (double-looking) literals the source writes
template <class T> static inline float abs(T x) { return fabsf(float(x)); }
template <class T> static inline float sqrt(T x) { return sqrtf(float(x)); }
template <class T> static inline float cos(T x) { return GLSL_M(cosF, cosf)(float(x)); }
template <class T> static inline float sin(T x) { return GLSL_M(sinF, sinf)(float(x)); }
template <class T> static inline float tan(T x) { return GLSL_M(tanF, tanf)(float(x)); }
template <class T> static inline float acos(T x) { return GLSL_M(acosF, acosf)(float(x)); }
template <class T> static inline float atan(T x) { return GLSL_M(atanF, atanf)(float(x)); }
template <class T> static inline float log(T x) { return GLSL_M(logF, logf)(float(x)); }
template <class T> static inline float exp(T x) { return GLSL_M(expF, expf)(float(x)); }
template <class A, class B> static inline float pow(A x, B y) { return GLSL_M(powF, powf)(float(x), float(y)); }
template <class A, class B> static inline float max(A x, B y) { return fmaxf(float(x), float(y)); }
template <class A, class B> static inline float min(A x, B y) { return fminf(float(x), float(y)); }
static inline bool isnan(float x) { return x != x; }
static inline bool isinf(float x) { return x == INFINITY || x == -INFINITY; }
static inline vec3 operator*(vec3 a, double s) { return a * float(s); }
static inline vec3 operator*(double s, vec3 a) { return a * float(s); }
static inline vec3 operator/(vec3 a, double s) { return a / float(s); }
static inline vec3 operator*(vec3 a, int s) { return a * float(s); }
static inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline vec3 refract(vec3 i, vec3 n, float eta) {           // GLSL.std.450 Refract
    const float d = dot(n, i), k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return vec3(0.0f);
    return eta * i - (eta * d + sqrtf(k)) * n;
}
// what raytrace.rgen's BSDF functions touch besides plain arithmetic: the sampler array (the unit tests use untextured
// materials: a texture fetch aborts), the ray flag constants of its header, and `.xyz` of a fetched texel
struct texel4 { xyz_t xyz; float w; };
struct sampler2D { int id; };
#ifndef GLSL_RT      // (oracle/shader_ref.cpp brings real textures)
static sampler2D textureSamplers[1];
static inline texel4 texture(sampler2D, vec2) { abort(); }
#endif
static const uint gl_RayFlagsNoneEXT = 0u, gl_RayFlagsTerminateOnFirstHitEXT = 4u, gl_RayFlagsSkipClosestHitShaderEXT = 8u;
static const vec3 GLSL_XXX(1.0f, 1.0f, 1.0f);                      // `scalar.xxx` is rewritten to `scalar * GLSL_XXX`
}  // namespace glsl
