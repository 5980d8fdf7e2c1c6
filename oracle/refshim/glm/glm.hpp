/*
 * glm/glm.hpp — TEST-ONLY shim of the few glm types the reference's hot-path headers touch
 * (src/trace_ray.hpp:32-57, src/camera.hpp:76-88, src/scene.hpp). glm is an unvendored third-party
 * dependency; operation order follows glm's documented component-wise definitions
 * (normalize = v * inversesqrt(dot(v, v)); mat3 * vec3 = m[0]*v.x + m[1]*v.y + m[2]*v.z).
 */
#pragma once
#include <cmath>
namespace glm {
struct vec2 {
    float x, y;
};
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct vec4 {
    float x, y, z, w;
};
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline float dot(const vec3 &a, const vec3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 normalize(const vec3 &a) { return a * (1.0f / std::sqrt(dot(a, a))); }
struct mat3 {
    vec3 c[3];
    mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    vec3 &operator[](int i) { return c[i]; }
    const vec3 &operator[](int i) const { return c[i]; }
};
inline vec3 operator*(const mat3 &m, const vec3 &v) { return (m[0] * v.x + m[1] * v.y) + m[2] * v.z; }
struct mat4 {
    float m[16];
    mat4() : m{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1} {}
    explicit mat4(float d) : m{d, 0, 0, 0, 0, d, 0, 0, 0, 0, d, 0, 0, 0, 0, d} {}
};
struct quat {
    float x = 0, y = 0, z = 0, w = 0;
};
} // namespace glm
