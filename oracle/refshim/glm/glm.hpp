/*
 * glm/glm.hpp — TEST-ONLY shim of the glm types and functions the reference touches
 * (src/trace_ray.hpp:32-57, src/camera.hpp:76-88, src/scene.hpp, src/scene.cpp). glm is an unvendored
 * third-party dependency; operation order follows glm's documented component-wise definitions
 * (normalize = v * inversesqrt(dot(v, v)); mat * vec = m[0]*v.x + m[1]*v.y + ...; mat3 inverse by
 * cofactors times one reciprocal determinant; quaternion <-> matrix by glm's mat3_cast / quat_cast).
 */
#pragma once
#include <cmath>
#include <cstring>
namespace glm {
struct vec2 {
    float x, y;
};
struct vec4;
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(const vec4 &v);
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3 &v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
};
inline vec3::vec3(const vec4 &v) : x(v.x), y(v.y), z(v.z) {}
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(const vec3 &a, const vec3 &b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float dot(const vec3 &a, const vec3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 cross(const vec3 &a, const vec3 &b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline vec3 normalize(const vec3 &a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec4 operator+(const vec4 &a, const vec4 &b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(const vec4 &a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float degrees(float r) { return r * 57.295779513082320876798154814105f; }
inline vec3 degrees(const vec3 &v) { return vec3(degrees(v.x), degrees(v.y), degrees(v.z)); }
inline float tan(float x) { return std::tan(x); }
struct mat4;
struct mat3 {
    vec3 c[3];
    mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    explicit mat3(const mat4 &m);
    vec3 &operator[](int i) { return c[i]; }
    const vec3 &operator[](int i) const { return c[i]; }
};
inline vec3 operator*(const mat3 &m, const vec3 &v) { return (m[0] * v.x + m[1] * v.y) + m[2] * v.z; }
struct quat;
struct mat4 {
    vec4 c[4];
    mat4() { c[0] = vec4(1, 0, 0, 0); c[1] = vec4(0, 1, 0, 0); c[2] = vec4(0, 0, 1, 0); c[3] = vec4(0, 0, 0, 1); }
    explicit mat4(float d) { c[0] = vec4(d, 0, 0, 0); c[1] = vec4(0, d, 0, 0); c[2] = vec4(0, 0, d, 0); c[3] = vec4(0, 0, 0, d); }
    explicit mat4(const quat &q);
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4 &m) { c[0] = vec3(m[0]); c[1] = vec3(m[1]); c[2] = vec3(m[2]); }
inline vec4 operator*(const mat4 &m, const vec4 &v) {
    /* glm: Mov0*m[0] + Mov1*m[1] + (Mov2*m[2] + Mov3*m[3]) */
    return (m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w);
}
inline mat4 operator*(const mat4 &a, const mat4 &b) {
    mat4 r;
    for (int j = 0; j < 4; j++) /* glm: SrcA0*SrcB_j[0] + SrcA1*SrcB_j[1] + SrcA2*SrcB_j[2] + SrcA3*SrcB_j[3], left to right */
        r[j] = ((a[0] * b[j][0] + a[1] * b[j][1]) + a[2] * b[j][2]) + a[3] * b[j][3];
    return r;
}
inline mat3 transpose(const mat3 &m) {
    mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i][j] = m[j][i];
    return r;
}
inline mat3 inverse(const mat3 &m) {
    const float det = (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])) +
                      m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
    const float od = 1.0f / det;
    mat3 r;
    r[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * od;
    r[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * od;
    r[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * od;
    r[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * od;
    r[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * od;
    r[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * od;
    r[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * od;
    r[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * od;
    r[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * od;
    return r;
}
/* memory order x, y, z, w (glm's default), constructor order (w, x, y, z); `quat{}` is all zeros */
struct quat {
    float x = 0, y = 0, z = 0, w = 0;
    quat() = default;
    quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
};
inline mat3 mat3_cast(const quat &q) {
    mat3 r;
    const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z, qwx = q.w * q.x,
                qwy = q.w * q.y, qwz = q.w * q.z;
    r[0][0] = 1.0f - 2.0f * (qyy + qzz);
    r[0][1] = 2.0f * (qxy + qwz);
    r[0][2] = 2.0f * (qxz - qwy);
    r[1][0] = 2.0f * (qxy - qwz);
    r[1][1] = 1.0f - 2.0f * (qxx + qzz);
    r[1][2] = 2.0f * (qyz + qwx);
    r[2][0] = 2.0f * (qxz + qwy);
    r[2][1] = 2.0f * (qyz - qwx);
    r[2][2] = 1.0f - 2.0f * (qxx + qyy);
    return r;
}
inline mat4::mat4(const quat &q) {
    const mat3 r = mat3_cast(q);
    c[0] = vec4(r[0], 0);
    c[1] = vec4(r[1], 0);
    c[2] = vec4(r[2], 0);
    c[3] = vec4(0, 0, 0, 1);
}
inline quat quat_cast(const mat3 &m) {
    const float fx = m[0][0] - m[1][1] - m[2][2], fy = m[1][1] - m[0][0] - m[2][2], fz = m[2][2] - m[0][0] - m[1][1],
                fw = m[0][0] + m[1][1] + m[2][2];
    int big = 0;
    float fb = fw;
    if (fx > fb) { fb = fx; big = 1; }
    if (fy > fb) { fb = fy; big = 2; }
    if (fz > fb) { fb = fz; big = 3; }
    const float bv = std::sqrt(fb + 1.0f) * 0.5f, mult = 0.25f / bv;
    switch (big) {
    case 0: return quat(bv, (m[1][2] - m[2][1]) * mult, (m[2][0] - m[0][2]) * mult, (m[0][1] - m[1][0]) * mult);
    case 1: return quat((m[1][2] - m[2][1]) * mult, bv, (m[0][1] + m[1][0]) * mult, (m[2][0] + m[0][2]) * mult);
    case 2: return quat((m[2][0] - m[0][2]) * mult, (m[0][1] + m[1][0]) * mult, bv, (m[1][2] + m[2][1]) * mult);
    default: return quat((m[0][1] - m[1][0]) * mult, (m[2][0] + m[0][2]) * mult, (m[1][2] + m[2][1]) * mult, bv);
    }
}
inline quat quat_cast(const mat4 &m) { return quat_cast(mat3(m)); }
inline vec3 operator*(const quat &q, const vec3 &v) {
    const vec3 qv(q.x, q.y, q.z), uv = cross(qv, v), uuv = cross(qv, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}
inline vec3 eulerAngles(const quat &q) { /* only printed by the reference; pitch, yaw, roll */
    const float y = 2.0f * (q.y * q.z + q.w * q.x), x = q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z;
    const float pitch = (x == 0.0f && y == 0.0f) ? 2.0f * std::atan2(q.x, q.w) : std::atan2(y, x);
    float s = -2.0f * (q.x * q.z - q.w * q.y);
    s = s < -1.0f ? -1.0f : (s > 1.0f ? 1.0f : s);
    const float roll = std::atan2(2.0f * (q.x * q.y + q.w * q.z), q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z);
    return vec3(pitch, std::asin(s), roll);
}
inline mat4 translate(const mat4 &m, const vec3 &v) {
    mat4 r = m;
    r[3] = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3];
    return r;
}
inline mat4 scale(const mat4 &m, const vec3 &v) {
    mat4 r;
    r[0] = m[0] * v.x;
    r[1] = m[1] * v.y;
    r[2] = m[2] * v.z;
    r[3] = m[3];
    return r;
}
template <class T> inline vec3 make_vec3(const T *p) { return vec3((float)p[0], (float)p[1], (float)p[2]); }
template <class T> inline quat make_quat(const T *p) { return quat((float)p[3], (float)p[0], (float)p[1], (float)p[2]); }
template <class T> inline mat4 make_mat4x4(const T *p) {
    mat4 r;
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++) r[j][i] = (float)p[j * 4 + i];
    return r;
}
} // namespace glm
