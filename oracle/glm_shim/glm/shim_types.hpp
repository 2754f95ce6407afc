// Minimal stand-in for the parts of GLM that external/lightpmm and external/guiding of the reference use.
// GLM itself is a system dependency of the reference (not vendored under /root/reference and not installed in this
// image), so the oracle build puts this directory on the include path instead.  Test infrastructure only.
// Semantics that matter for parity: vec3::length() is GLM's *component count* (3), not the Euclidean norm — the
// reference calls it in incrementalcovariance2d.h:211 — and the default constructors leave the
// members zero (GLM leaves them uninitialised; the reference only reads such values before the first fit).
#pragma once
#include <cmath>
#include <cstddef>
namespace glm {
struct vec2 {
    float x = 0.0f, y = 0.0f;
    vec2() = default;
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(float s) : x(s), y(s) {}
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
    vec2 &operator*=(float s) { x *= s; y *= s; return *this; }
    static constexpr int length() { return 2; }
};
struct vec3 {
    float x = 0.0f, y = 0.0f, z = 0.0f;
    vec3() = default;
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
    vec3 &operator+=(const vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3 &operator-=(const vec3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3 &operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    vec3 &operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
    static constexpr int length() { return 3; }
};
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(const vec3 &a, const vec3 &b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3 &a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline bool operator==(const vec3 &a, const vec3 &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline vec2 operator*(const vec2 &a, float s) { return vec2(a.x * s, a.y * s); }
inline bool operator==(const vec2 &a, const vec2 &b) { return a.x == b.x && a.y == b.y; }
inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // GLM: (x*x + y*y) + z*z
inline vec3 cross(const vec3 &a, const vec3 &b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(const vec3 &a) { return std::sqrt(dot(a, a)); }
inline float length(const vec2 &a) { return std::sqrt(a.x * a.x + a.y * a.y); }
inline vec3 normalize(const vec3 &a) { return a * (1.0f / std::sqrt(dot(a, a))); }   // GLM: v * inversesqrt(dot(v, v))
struct mat2x2 {   // column-major like GLM: m[col][row]
    vec2 c[2];
    mat2x2() = default;
    mat2x2(float x0, float y0, float x1, float y1) { c[0] = vec2(x0, y0); c[1] = vec2(x1, y1); }
    vec2 &operator[](int i) { return c[i]; }
    const vec2 &operator[](int i) const { return c[i]; }
};
template <typename T> constexpr T pi() { return static_cast<T>(3.14159265358979323846264338327950288); }
}  // namespace glm
