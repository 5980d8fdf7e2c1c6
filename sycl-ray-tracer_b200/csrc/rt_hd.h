/*
 * rt_hd.h — scalar/vector helpers shared by every kernel of the path.
 *
 * All per-item device logic in csrc/ is written as RT_HD inline functions and wrapped by thin
 * __global__ kernels. Compiled by nvcc for sm_100a this is the product; the same headers also
 * compile with plain g++ (tests/hostemu, TEST-ONLY) so that the kernel source itself can be
 * stepped on a machine without a GPU. The library (librt_b200.so) contains only the CUDA path.
 *
 * Arithmetic contract: device code is compiled with -fmad=false and without
 * fast-math, so a*b+c written here is two IEEE roundings, exactly like the oracle
 * (g++ -ffp-contract=off). Where a fused multiply-add is wanted (conservative box slabs) it is
 * spelled rt_fma().
 */
#ifndef RT_HD_H
#define RT_HD_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define RT_HD __host__ __device__ __forceinline__
#define RT_D __device__ __forceinline__
#else
#define RT_HD inline
#define RT_D inline
#include <string.h>
#endif

#if defined(__CUDA_ARCH__)
#define RT_DEVICE_CODE 1
#else
#define RT_DEVICE_CODE 0
#endif

/* ---------------------------------------------------------------- bit / intrinsic wrappers */
RT_HD uint32_t rt_f2u(float f) {
#if RT_DEVICE_CODE
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
RT_HD float rt_u2f(uint32_t u) {
#if RT_DEVICE_CODE
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
RT_HD int rt_clz32(uint32_t x) {
#if RT_DEVICE_CODE
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
RT_HD int rt_clz64(uint64_t x) {
#if RT_DEVICE_CODE
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
RT_HD int rt_popc(uint32_t x) {
#if RT_DEVICE_CODE
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
/* index of the highest set bit (x != 0) */
RT_HD int rt_bfind(uint32_t x) { return 31 - rt_clz32(x); }
/* index of the lowest set bit (x != 0) */
RT_HD int rt_ctz(uint32_t x) {
#if RT_DEVICE_CODE
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
RT_HD float rt_fma(float a, float b, float c) {
#if RT_DEVICE_CODE
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
/* two independent fused multiply-adds sharing the multiplier b: d0 = fma(a0, b, c0), d1 = fma(a1, b, c1).
 * sm_100a has a packed FFMA2 (fma.rn.f32x2): one issue slot for both, each half rounded exactly like
 * FFMA, so results stay bit-identical to the host emulation and the oracle. */
#ifndef RT_USE_FFMA2
#define RT_USE_FFMA2 0 /* measured neutral (C3 +0.7 %, C2/C4 -0.4 %): the node test is not issue-slot bound */
#endif
RT_HD void rt_fma2(float a0, float a1, float b, float c0, float c1, float &d0, float &d1) {
#if RT_DEVICE_CODE && RT_USE_FFMA2
    unsigned long long A, B, C, D;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(C));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
#else
    d0 = rt_fma(a0, b, c0);
    d1 = rt_fma(a1, b, c1);
#endif
}
RT_HD float rt_min(float a, float b) { return fminf(a, b); }
RT_HD float rt_max(float a, float b) { return fmaxf(a, b); }
/* 3-input min/max: one FMNMX3 on sm_100a */
RT_HD float rt_min3(float a, float b, float c) {
#if RT_DEVICE_CODE
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
#else
    return fminf(fminf(a, b), c);
#endif
}
RT_HD float rt_max3(float a, float b, float c) {
#if RT_DEVICE_CODE
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}
/* IEEE round-to-nearest divide / sqrt / u32->f32 (never the approximate forms) */
RT_HD float rt_div(float a, float b) {
#if RT_DEVICE_CODE
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
/* v / 255.0f for v = 0 .. 255 (texel byte -> colour, src/material.hpp via the unorm8 read), correctly rounded without the
 * IEEE-division sequence and its range check: q0 = v * RN(1/255), one FMA residual, one FMA correction (the last step of
 * Markstein's division). Equal to v / 255.0f for all 256 inputs (tests/test_hostemu_parity.py checks every one). */
RT_HD float rt_u8_to_unit(float v) {
    const float rc = 0.0039215688593685627f; /* 0x3b808081 */
    const float q0 = v * rc;
    return rt_fma(rt_fma(-255.0f, q0, v), rc, q0);
}
RT_HD float rt_sqrt(float a) {
#if RT_DEVICE_CODE
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}
RT_HD float rt_u32_to_float(uint32_t a) {
#if RT_DEVICE_CODE
    return __uint2float_rn(a);
#else
    return (float)a;
#endif
}
/* fp32 -> fp16 -> fp32, round to nearest even (sycl::half, src/camera.hpp:18-28) */
RT_HD float rt_round_half(float v) {
#if defined(__CUDACC__)
    return __half2float(__float2half_rn(v));
#else
    return (float)(_Float16)v;
#endif
}
RT_HD uint16_t rt_float_to_half_bits(float v) {
#if defined(__CUDACC__)
    return __half_as_ushort(__float2half_rn(v));
#else
    _Float16 h = (_Float16)v;
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
#endif
}
RT_HD float rt_half_bits_to_float(uint16_t b) {
#if defined(__CUDACC__)
    return __half2float(__ushort_as_half(b));
#else
    _Float16 h;
    memcpy(&h, &b, 2);
    return (float)h;
#endif
}

/* ---------------------------------------------------------------- float3 */
struct f3 {
    float x, y, z;
};
RT_HD f3 mk3(float x, float y, float z) {
    f3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
RT_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
RT_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
RT_HD f3 div3(f3 a, float s) { return mk3(rt_div(a.x, s), rt_div(a.y, s), rt_div(a.z, s)); }
RT_HD float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
RT_HD float length3(f3 a) { return rt_sqrt(dot3(a, a)); }
/* sycl::normalize / glm::normalize pinned as v * (1 / sqrt(dot(v, v))) */
RT_HD f3 normalize3(f3 a) {
    float inv = rt_div(1.0f, rt_sqrt(dot3(a, a)));
    return a * inv;
}
RT_HD f3 round_half3(f3 a) { return mk3(rt_round_half(a.x), rt_round_half(a.y), rt_round_half(a.z)); }
RT_HD float sel3(f3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

/* ---------------------------------------------------------------- xorshift32
 * src/xorshift.hpp:8-40. One 32-bit state per pixel; float = u32 -> f32 (RNE) * 2^-32, so the
 * draw lies in [0, 1] inclusive. */
struct XorShift32 {
    uint32_t a;
    RT_HD float next() {
        uint32_t x = a;
        x ^= x << 13;
        x ^= x >> 17;
        x ^= x << 5;
        a = x;
        return rt_u32_to_float(x) * 2.3283064365386963e-10f; /* 2^-32, exact */
    }
    RT_HD float next(float mn, float mx) { return mn + (mx - mn) * next(); }
    /* draws in x, y, z order; normalised without rejection (src/xorshift.hpp:30-40) */
    RT_HD f3 random_unit_vector() {
        float x = next(-1.0f, 1.0f);
        float y = next(-1.0f, 1.0f);
        float z = next(-1.0f, 1.0f);
        return normalize3(mk3(x, y, z));
    }
};

#endif /* RT_HD_H */
