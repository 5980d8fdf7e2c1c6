/*
 * rt_traverse.h — closest-hit traversal of the 8-wide compressed BVH + watertight triangle test.
 *
 * Replaces Embree's rtcIntersect1 as called at src/trace_ray.hpp:18-22 (Embree 4 is a
 * third-party dependency that is not under /root/reference; there is no source to follow, so
 * the algorithm is our own and the oracle states the same contract):
 *   - closest hit with tnear < t <= tfar, double sided, no filters;
 *   - u, v = barycentric weights of vertex 1 and 2 (Embree convention, F11);
 *   - ray/triangle test: Woop, Benthin, Wald, "Watertight Ray/Triangle Intersection"
 *     (JCGT 2013), unfused fp32 with the double-precision fallback on zero edge functions;
 *   - ties in t resolve to the lowest global triangle id, so the result does not depend on the
 *     traversal order or on the tree.
 *
 * Node format (80 bytes = 5 x 16-byte vector loads, after Ylitie, Karras, Laine, "Efficient
 * Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs", HPG 2017):
 *   n0 = { p.x, p.y, p.z, ex | ey<<8 | ez<<16 | imask<<24 }
 *   n1 = { child_base, tri_base, tmask, 0 }
 *   n2 = { qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7] }
 *   n3 = { qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7] }
 *   n4 = { qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7] }
 * child box k = p + q * 2^(e-127)  (e = biased fp32 exponent byte).
 * imask bit s = slot s is an inner child (child node = child_base + rank of s among the imask bits).
 * tmask bits 3s..3s+2 = unary triangle count (1..3) of leaf slot s; the node's leaf triangles are
 * stored contiguously in slot order from tri_base, so triangle position b (a set bit of tmask) is
 * tri_base + popc(tmask below b). Empty slots have neither bit and an inverted box.
 * Triangles: 3 x float4 in leaf order: {v0, 0}, {v1, 0}, {v2, global_id_bits}.
 */
#ifndef RT_TRAVERSE_H
#define RT_TRAVERSE_H

#include "rt_hd.h"

#define RT_STACK_SIZE 40 /* pending node groups (<= 1 per level); rt_scene_commit refuses deeper trees */
#define RT_MISS 0xffffffffu
#define RT_BOX_PAD 1.00000048f /* 1 + 2^-21: far planes / tmax are padded by ~4 ulp */

#if !defined(__CUDACC__)
struct rt_u4 {
    uint32_t x, y, z, w;
};
struct rt_f4 {
    float x, y, z, w;
};
struct rt_u2 {
    uint32_t x, y;
};
typedef rt_u4 rt_uint4;
typedef rt_f4 rt_float4;
typedef rt_u2 rt_uint2;
#else
typedef uint4 rt_uint4;
typedef float4 rt_float4;
typedef uint2 rt_uint2;
#endif
RT_HD rt_float4 rt_mk_float4(float x, float y, float z, float w) {
    rt_float4 r;
    r.x = x; r.y = y; r.z = z; r.w = w;
    return r;
}
RT_HD rt_uint2 rt_mk_uint2(uint32_t x, uint32_t y) {
    rt_uint2 r;
    r.x = x; r.y = y;
    return r;
}

struct RtHit {
    float t, u, v;
    uint32_t tri; /* slot in the leaf-ordered triangle array, RT_MISS when nothing was hit */
    uint32_t gid; /* global triangle id = first_tri[inst] + primID (tie-break key) */
};

struct RtBvh {
    const rt_uint4 *nodes;  /* RT_NODE_VEC4 per node, 32-byte aligned */
    const rt_float4 *tris;  /* RT_TRI_VEC4 per triangle, 32-byte aligned */
};

/* Record strides in 16-byte units. Records are padded to multiples of 32 bytes so that they can be
 * fetched with 256-bit loads (LDG.E.ENL2.256 on sm_100a): every lane of an incoherent warp touches
 * its own cache line, and the L1TEX data stage spends one wavefront per (instruction, line), so a
 * 96-byte node costs 3 wavefronts per lane instead of 5 with 128-bit loads (profiles/README.md). */
#ifndef RT_USE_LDG256
#define RT_USE_LDG256 0 /* measured on B200 (C3/C4): no gain over 128-bit loads, 20-33 % more memory */
#endif
#if RT_USE_LDG256
#define RT_NODE_VEC4 6 /* 80 bytes of node + 16 spare */
#define RT_TRI_VEC4 4  /* 48 bytes of vertices/id + 16 spare */
#else
#define RT_NODE_VEC4 5
#define RT_TRI_VEC4 3
#endif

/* read-only 16-byte loads (LDG.E.128.CONSTANT on device) */
RT_HD rt_uint4 rt_ldg(const rt_uint4 *p) {
#if RT_DEVICE_CODE
    return __ldg(p);
#else
    return *p;
#endif
}
RT_HD rt_float4 rt_ldg(const rt_float4 *p) {
#if RT_DEVICE_CODE
    return __ldg(p);
#else
    return *p;
#endif
}

/* Cache-policy hints on the record fetches (compile-time experiments, profiles/README.md): triangle and shading records
 * are read once per (ray, leaf) and compete in L1 with the nodes of the upper levels, which every ray of the warp re-reads.
 *   RT_TRI_LD_HINT / RT_SHADE_LD_HINT: 0 = plain LDG.CONSTANT, 1 = L1::no_allocate, 2 = L1::evict_first
 *   RT_NODE_LD_HINT: 0 = plain, 1 = L1::evict_last */
#ifndef RT_TRI_LD_HINT
#define RT_TRI_LD_HINT 0
#endif
#ifndef RT_SHADE_LD_HINT
#define RT_SHADE_LD_HINT 0
#endif
#ifndef RT_NODE_LD_HINT
#define RT_NODE_LD_HINT 0
#endif
template <int HINT>
RT_HD rt_float4 rt_ldg_hint(const rt_float4 *p) {
#if RT_DEVICE_CODE
    if (HINT == 1) {
        rt_float4 r;
        asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
        return r;
    }
    if (HINT == 2) {
        rt_float4 r;
        asm("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
        return r;
    }
#endif
    return rt_ldg(p);
}
RT_HD rt_uint4 rt_ldg_node(const rt_uint4 *p) {
#if RT_DEVICE_CODE && RT_NODE_LD_HINT == 1
    rt_uint4 r;
    asm("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#else
    return rt_ldg(p);
#endif
}

/* read-only 32-byte load of two consecutive 16-byte records (p must be 32-byte aligned) */
RT_HD void rt_ldg2(const rt_uint4 *p, rt_uint4 &a, rt_uint4 &b) {
#if RT_DEVICE_CODE && RT_USE_LDG256
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
        : "l"(p));
#else
    a = rt_ldg(p);
    b = rt_ldg(p + 1);
#endif
}
RT_HD void rt_ldg2(const rt_float4 *p, rt_float4 &a, rt_float4 &b) {
#if RT_DEVICE_CODE && RT_USE_LDG256
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
#else
    a = rt_ldg(p);
    b = rt_ldg(p + 1);
#endif
}

/* global id of the triangle in leaf slot `slot` (RT_MISS -> RT_MISS: no hit yet loses every tie) */
RT_HD uint32_t rt_tri_gid(const RtBvh &bvh, uint32_t slot) {
    return slot == RT_MISS ? RT_MISS : rt_f2u(rt_ldg(bvh.tris + (size_t)slot * RT_TRI_VEC4 + 2).w);
}

/* byte j of w -> a float that encodes the quantised plane b, with ONE instruction and no I2F (the
 * conversion pipe is the narrowest one on sm_100a; 48 conversions per node visit made it the busiest):
 *   PRMT (ALU pipe): 0x3F800000 | b << 8         = 1 + b * 2^-15   one byte permute
 *   IDP.4A (FMA pipe): 0x3F800000 + 128 * b      = 1 + b * 2^-16   dot product of the four bytes of w with
 *                                                                  (0, .., 128, .., 0), accumulated onto 1.0f
 * The node test folds the "1 +" and the 2^-15 / 2^-16 into its per-node constants. Which of the two an
 * axis uses is a compile-time choice (RT_BYTE_IDP_MASK, bit a = axis a through IDP.4A): a node visit has
 * ~150 ALU-pipe instructions (48 byte extractions, 32 FMNMX, selects, the hit mask) against ~75 on the FMA
 * pipe, and both pipes take a warp instruction every other cycle, so the byte extractions are what can
 * be moved to balance them (tools/microbench/pipe_bench.cu, profiles/README.md). */
#ifndef RT_BYTE_IDP_MASK
#define RT_BYTE_IDP_MASK 3 /* measured on B200 (slim state): x, y through IDP.4A, z through PRMT: C3 +0.6 / +2.5 %, C4 +0.8 / +2.7 % (megakernel / wavefront) over all-PRMT */
#endif
template <int J, int UNIQ, int IDP>
RT_HD float rt_byte_to_unit(uint32_t w, uint32_t one_bits /* 0x3F800000, held in a register */) {
#if RT_DEVICE_CODE
    uint32_t r;
    if (IDP) {
        asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "n"(128u << (8 * J)), "r"(one_bits));
    } else {
        /* selector: byte0 <- one.b0, byte1 <- w.bJ, byte2 <- one.b2, byte3 <- one.b3. Only the low 16 bits
         * of a PRMT selector are read; UNIQ makes every call site's constant different so that ptxas
         * encodes it as an immediate instead of hoisting four shared selectors into registers and
         * re-materialising them before each of the 48 permutes of a node test (measured: +48 moves). */
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(one_bits), "n"(0x7604 | (J << 4) | (UNIQ << 16)));
    }
    return __uint_as_float(r);
#else
    return IDP ? rt_u2f(one_bits + (((w >> (8 * J)) & 0xffu) << 7)) : rt_u2f(one_bits | (((w >> (8 * J)) & 0xffu) << 8));
#endif
}
/* PRMT takes one immediate: if the compiler sees both 0x3F800000 and the selector as constants it keeps
 * the SELECTOR in a register and re-materialises it before every one of the 48 permutes of a node
 * test (measured: +48 moves per node visit). Hiding the constant behind an empty asm pins it in a
 * register instead, so the selectors are encoded as immediates. */
RT_HD uint32_t rt_unit_bits() {
    uint32_t one_bits = 0x3F800000u;
#if RT_DEVICE_CODE
    asm volatile("" : "+r"(one_bits));
#endif
    return one_bits;
}

/* ---- watertight ray/triangle -------------------------------------------------------- */
/* Woop/Benthin/Wald shear, with the axis permutation folded into three per-ray vectors so that the
 * per-triangle work has no selects: for kz = dominant axis of dir, (kx, ky) the other two (swapped
 * when dir[kz] < 0),
 *     mx[kx] = 1, mx[kz] = -dir[kx]/dir[kz];  my[ky] = 1, my[kz] = -dir[ky]/dir[kz];  mz[kz] = 1/dir[kz]
 * (all other components 0) and the sheared coordinates of a vertex P are P.mx, P.my, P.mz. Products
 * with 0 and 1 are exact, so this IS the paper's  P[kx] - Sx*P[kz]  etc. The dot is evaluated as
 * fma(P.x, m.x, fma(P.y, m.y, P.z*m.z)) in both the CUDA path and the oracle. A vertex shared by two
 * triangles gets identical sheared coordinates, and the edge functions below stay unfused, so edge
 * functions of a shared edge are exact negatives of each other (watertightness). */
struct RtRayTri {
    f3 org;
    f3 mx, my, mz;
};

RT_HD float rt_shear(f3 p, f3 m) { return rt_fma(p.x, m.x, rt_fma(p.y, m.y, p.z * m.z)); }

/* The direction-only part of the per-ray set-up (six IEEE divisions), separated so that the persistent
 * megakernel can compute it while it shades (many lanes active) and park it with the ray: starting the
 * traversal later is then a handful of loads and selects. code = kx | ky << 2 | kz << 4 | neg << 6 |
 * oct_inv << 9 (see the wide-node test). */
struct RtRayPre {
    f3 rcp;             /* 1 / dir with |dir| clamped away from 0 (box slabs) */
    float nSx, nSy, Sz; /* -dir[kx]/dir[kz], -dir[ky]/dir[kz], 1/dir[kz] (triangle shear) */
    uint32_t code;
};

RT_HD RtRayPre rt_ray_pre(f3 dir) {
    RtRayPre r;
    float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    int kz = 0;
    float am = ax;
    if (ay > am) {
        kz = 1;
        am = ay;
    }
    if (az > am) kz = 2;
    int kx = kz + 1;
    if (kx == 3) kx = 0;
    int ky = kx + 1;
    if (ky == 3) ky = 0;
    const float dz = sel3(dir, kz);
    if (dz < 0.0f) {
        int tmp = kx;
        kx = ky;
        ky = tmp;
    }
    r.nSx = -rt_div(sel3(dir, kx), dz);
    r.nSy = -rt_div(sel3(dir, ky), dz);
    const float tiny = 1e-20f;
    float dx = fabsf(dir.x) > tiny ? dir.x : copysignf(tiny, dir.x);
    float dy = fabsf(dir.y) > tiny ? dir.y : copysignf(tiny, dir.y);
    float dzz = fabsf(dir.z) > tiny ? dir.z : copysignf(tiny, dir.z);
    r.rcp = mk3(rt_div(1.0f, dx), rt_div(1.0f, dy), rt_div(1.0f, dzz));
    /* 1 / dir[kz]: the dominant component is not clamped unless the direction vanishes altogether, so it IS the slab
     * reciprocal of that axis (one IEEE division less per ray) */
    r.Sz = fabsf(dz) > tiny ? sel3(r.rcp, kz) : rt_div(1.0f, dz);
    const uint32_t neg = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dzz < 0.0f ? 4u : 0u);
    /* slot s is visited in order of increasing (s ^ octant), octant = x<<2 | y<<1 | z sign bits */
    const uint32_t octant = (dx < 0.0f ? 4u : 0u) | (dy < 0.0f ? 2u : 0u) | (dzz < 0.0f ? 1u : 0u);
    r.code = (uint32_t)kx | ((uint32_t)ky << 2) | ((uint32_t)kz << 4) | (neg << 6) | ((7u - octant) << 9);
    return r;
}

RT_HD RtRayTri rt_ray_tri_from_pre(f3 org, const RtRayPre &q) {
    RtRayTri r;
    r.org = org;
    const int kx = (int)(q.code & 3u), ky = (int)((q.code >> 2) & 3u), kz = (int)((q.code >> 4) & 3u);
    const float nSx = q.nSx, nSy = q.nSy, Sz = q.Sz;
    r.mx = mk3(kx == 0 ? 1.0f : (kz == 0 ? nSx : 0.0f), kx == 1 ? 1.0f : (kz == 1 ? nSx : 0.0f),
               kx == 2 ? 1.0f : (kz == 2 ? nSx : 0.0f));
    r.my = mk3(ky == 0 ? 1.0f : (kz == 0 ? nSy : 0.0f), ky == 1 ? 1.0f : (kz == 1 ? nSy : 0.0f),
               ky == 2 ? 1.0f : (kz == 2 ? nSy : 0.0f));
    r.mz = mk3(kz == 0 ? Sz : 0.0f, kz == 1 ? Sz : 0.0f, kz == 2 ? Sz : 0.0f);
    return r;
}

/* Updates the closest hit (t, u, v, slot) when triangle (v0, v1, v2) is closer, or equally far with a lower
 * id. Exactly the oracle's operation order. The id of the current closest hit is only needed when the two
 * distances are EQUAL (rare: coplanar duplicates, shared edges), so it is not carried in registers but
 * re-read from the record of the current hit (third vertex, w). */
RT_HD void rt_tri_test(const RtBvh &bvh, const RtRayTri &r, f3 v0, f3 v1, f3 v2, uint32_t slot, uint32_t gid, float tnear,
                       float &best_t, float &best_u, float &best_v, uint32_t &best_tri) {
    const f3 A = v0 - r.org, B = v1 - r.org, C = v2 - r.org;
    const float Ax = rt_shear(A, r.mx), Ay = rt_shear(A, r.my);
    const float Bx = rt_shear(B, r.mx), By = rt_shear(B, r.my);
    const float Cx = rt_shear(C, r.mx), Cy = rt_shear(C, r.my);
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return;
    const float det = (U + V) + W;
    if (det == 0.0f) return;
    const float Az = rt_shear(A, r.mz), Bz = rt_shear(B, r.mz), Cz = rt_shear(C, r.mz);
    const float T = (U * Az + V * Bz) + W * Cz;
    const float rcp = rt_div(1.0f, det);
    const float t = T * rcp;
    if (!(t > tnear)) return;
    if (t < best_t || (t == best_t && gid < rt_tri_gid(bvh, best_tri))) {
        best_t = t;
        best_u = V * rcp;
        best_v = W * rcp;
        best_tri = slot;
    }
}

/* ---- wide-node test ------------------------------------------------------------------ */
/* The ray enters the node test as origin, 1 / dir (|dir| clamped away from 0) and oct_inv = 7 - octant,
 * octant = (dir.x < 0) << 2 | (dir.y < 0) << 1 | (dir.z < 0): the inner child in slot s has visiting
 * priority s ^ oct_inv, and a clear bit 2 / 1 / 0 of oct_inv means dir.x / y / z is negative (near and far
 * planes swap). */
/* bit i of x -> bit i ^ k (k = 0..7): the octant permutation of a byte of child-hit flags */
RT_HD uint32_t rt_xor_perm8(uint32_t x, uint32_t k) {
    if (k & 1u) x = ((x & 0x55u) << 1) | ((x >> 1) & 0x55u);
    if (k & 2u) x = ((x & 0x33u) << 2) | ((x >> 2) & 0x33u);
    if (k & 4u) x = ((x & 0x0fu) << 4) | ((x >> 4) & 0x0fu);
    return x;
}

/* hits |= BIT when the slab interval is not empty: FSETP + one predicated LOP3 with an immediate */
template <uint32_t BIT>
RT_HD void rt_or_if_le(float a, float b, uint32_t &hits) {
#if RT_DEVICE_CODE
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(hits) : "f"(a), "f"(b), "n"(BIT));
#else
    if (a <= b) hits |= BIT;
#endif
}

/* one child: slot 4 * H + J, planes = byte J of the six quantised-plane words of half H */
template <int J, int H, int MASK>
RT_HD void rt_child_test(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx, uint32_t fy, uint32_t fz, float Sx,
                         float Sy, float Sz, float onx, float ony, float onz, float ofx, float ofy, float ofz,
                         float tmin, float tmax_pad, uint32_t one, uint32_t &hits) {
    float tnx, tny, tnz, tfx, tfy, tfz; /* near and far plane of one axis share the multiplier: one FFMA2 */
    constexpr int IX = MASK & 1, IY = (MASK >> 1) & 1, IZ = (MASK >> 2) & 1;
    rt_fma2(rt_byte_to_unit<J, 1 + 0 + 6 * J + 24 * H, IX>(nx, one), rt_byte_to_unit<J, 1 + 3 + 6 * J + 24 * H, IX>(fx, one), Sx, onx, ofx, tnx, tfx);
    rt_fma2(rt_byte_to_unit<J, 1 + 1 + 6 * J + 24 * H, IY>(ny, one), rt_byte_to_unit<J, 1 + 4 + 6 * J + 24 * H, IY>(fy, one), Sy, ony, ofy, tny, tfy);
    rt_fma2(rt_byte_to_unit<J, 1 + 2 + 6 * J + 24 * H, IZ>(nz, one), rt_byte_to_unit<J, 1 + 5 + 6 * J + 24 * H, IZ>(fz, one), Sz, onz, ofz, tnz, tfz);
    const float cmin = rt_max(rt_max3(tnx, tny, tnz), tmin);
    const float cmax = rt_min(rt_min3(tfx, tfy, tfz), tmax_pad);
    rt_or_if_le<1u << (4 * H + J)>(cmin, cmax, hits);
}

/* four children (one 32-bit word of every quantised plane) */
template <int H, int MASK>
RT_HD void rt_half_test(uint32_t oct_inv, uint32_t qlox, uint32_t qloy, uint32_t qloz, uint32_t qhix,
                        uint32_t qhiy, uint32_t qhiz, float Sx, float Sy, float Sz, float onx, float ony, float onz,
                        float ofx, float ofy, float ofz, float tmin, float tmax_pad, uint32_t one, uint32_t &hits) {
    const uint32_t nx = (oct_inv & 4u) ? qlox : qhix, fx = (oct_inv & 4u) ? qhix : qlox;
    const uint32_t ny = (oct_inv & 2u) ? qloy : qhiy, fy = (oct_inv & 2u) ? qhiy : qloy;
    const uint32_t nz = (oct_inv & 1u) ? qloz : qhiz, fz = (oct_inv & 1u) ? qhiz : qloz;
    rt_child_test<0, H, MASK>(nx, ny, nz, fx, fy, fz, Sx, Sy, Sz, onx, ony, onz, ofx, ofy, ofz, tmin, tmax_pad, one, hits);
    rt_child_test<1, H, MASK>(nx, ny, nz, fx, fy, fz, Sx, Sy, Sz, onx, ony, onz, ofx, ofy, ofz, tmin, tmax_pad, one, hits);
    rt_child_test<2, H, MASK>(nx, ny, nz, fx, fy, fz, Sx, Sy, Sz, onx, ony, onz, ofx, ofy, ofz, tmin, tmax_pad, one, hits);
    rt_child_test<3, H, MASK>(nx, ny, nz, fx, fy, fz, Sx, Sy, Sz, onx, ony, onz, ofx, ofy, ofz, tmin, tmax_pad, one, hits);
}

/* returns the hit flags of one wide node in SLOT order: bit s = the ray's interval [tmin, tmax_pad]
 * overlaps the (conservatively padded) box of slot s. Empty slots never hit (inverted boxes). */
template <int MASK>
RT_HD uint32_t rt_node_test(f3 org, f3 rcp, uint32_t oct_inv, rt_uint4 n0, rt_uint4 n2, rt_uint4 n3, rt_uint4 n4, float tmin,
                            float tmax_pad) {
    const float sx = rt_u2f((n0.w & 0xffu) << 23), sy = rt_u2f(((n0.w >> 8) & 0xffu) << 23),
                sz = rt_u2f(((n0.w >> 16) & 0xffu) << 23);
    /* plane t = q*id + o with id = 2^e/d, o = (p - org)/d */
    const float idx = sx * rcp.x, idy = sy * rcp.y, idz = sz * rcp.z;
    const float ox = (rt_u2f(n0.x) - org.x) * rcp.x, oy = (rt_u2f(n0.y) - org.y) * rcp.y, oz = (rt_u2f(n0.z) - org.z) * rcp.z;
    /* q enters as u = 1 + q*2^-k (rt_byte_to_unit; k = 15 for PRMT, 16 for IDP.4A): t = u*S + (o - S), S = id * 2^k (exact). */
    constexpr float KX = (MASK & 1) ? 65536.0f : 32768.0f, KY = (MASK & 2) ? 65536.0f : 32768.0f, KZ = (MASK & 4) ? 65536.0f : 32768.0f;
    const float Sx = idx * KX, Sy = idy * KY, Sz = idz * KZ;
    /* Conservative slabs. The sum cancels when the ray starts next to a plane that lies far from
     * the node origin p (|q*id|, |o| >> |t|), e.g. a bounce ray leaving an axis-aligned wall. The
     * roundings of o, of o - S and of the fma are bounded by 2^-24 * (3|o| + 2|S|) plus the three
     * roundings of the plain formula, 3 * 2^-24 * (255|id| + |o|); near planes are moved back and
     * far planes forward by e = 2^-21 |o| + 2^(k-23) |id| (> the bound, 0.4 - 0.8 % of a quantisation
     * step), so rounding can never cull a box whose triangle the exact-difference triangle test
     * would accept. */
    constexpr float PX = KX * 1.1920929e-7f, PY = KY * 1.1920929e-7f, PZ = KZ * 1.1920929e-7f; /* 2^(k-23) */
    const float ex = rt_fma(fabsf(idx), PX, fabsf(ox) * 4.76837158e-7f),
                ey = rt_fma(fabsf(idy), PY, fabsf(oy) * 4.76837158e-7f),
                ez = rt_fma(fabsf(idz), PZ, fabsf(oz) * 4.76837158e-7f);
    const float onx = (ox - ex) - Sx, ony = (oy - ey) - Sy, onz = (oz - ez) - Sz;
    const float ofx = (ox + ex) - Sx, ofy = (oy + ey) - Sy, ofz = (oz + ez) - Sz;
    uint32_t hits = 0;
    const uint32_t one = rt_unit_bits();
    rt_half_test<0, MASK>(oct_inv, n2.x, n2.z, n3.x, n3.z, n4.x, n4.z, Sx, Sy, Sz, onx, ony, onz, ofx, ofy, ofz, tmin, tmax_pad, one, hits);
    rt_half_test<1, MASK>(oct_inv, n2.y, n2.w, n3.y, n3.w, n4.y, n4.w, Sx, Sy, Sz, onx, ony, onz, ofx, ofy, ofz, tmin, tmax_pad, one, hits);
    return hits;
}

/* ---- traversal ------------------------------------------------------------------------ */
#ifndef RT_COUNTERS
#define RT_COUNT_NODE()
#define RT_COUNT_TRI()
#endif

/* Resumable traversal state. The traversal is split into two kinds of unit work so that the
 * persistent kernels can run each kind with the warp converged:
 *   rt_trav_node_step : visit one wide node; leaf hits are only RECORDED (a small per-lane stack of
 *                       triangle groups), not tested;
 *   rt_trav_tri_step  : test one recorded triangle.
 * The kernels run node steps for all lanes until enough lanes have triangles pending (or have run out
 * of nodes), then drain the pending triangles together. Deferring the triangle tests only delays the
 * shrinking of tmax; the closest hit (min t, then min id) does not depend on the test order. */
#define RT_TSTACK_SIZE 8
/* RT_STACK_TOP = 1 keeps the top of the node stack in two registers: a pop then is a register move and the load of the
 * entry below it is off the critical path (popped group -> node address -> node fetch). Costs two registers. */
#ifndef RT_STACK_TOP
#define RT_STACK_TOP 0
#endif
/* Kept small on purpose: the node test needs ~40 registers of its own (20 of node data), and whatever of
 * this state does not fit next to it in the kernels' 64 registers is re-loaded from local memory on EVERY
 * node visit (round 1: ten spill loads per visit, a quarter of the L1TEX sectors of the node loop). So the
 * triangle shear is held as its three scalars + axis code (the nine-float form is rebuilt per triangle
 * batch, rt_trav_ray_tri), the padded tmax is recomputed from the hit distance, the sign bits are read off
 * oct_inv, and the closest hit's id (needed for exact-t ties only) is re-read from the triangle record. */
struct RtTravState {
    f3 org;
    f3 rcp;             /* 1 / dir with |dir| clamped away from 0 */
    float nSx, nSy, Sz; /* triangle shear (RtRayPre) */
    uint32_t code;      /* RtRayPre::code: kx | ky << 2 | kz << 4 | neg << 6 | oct_inv << 9 */
    float tnear;
    float t, u, v;      /* closest hit so far (t = tfar: none) */
    uint32_t tri;       /* its slot in the leaf-ordered triangle array, RT_MISS when nothing was hit */
    uint32_t ng_x, ng_y; /* current node group: (child base, child hit bits << 24 | imask) */
#if RT_STACK_TOP
    uint32_t top_x, top_y; /* the most recently deferred group, held in registers (top_y <= 0x00ffffff: none) */
#endif
    int sp, tsp;
};
RT_HD uint32_t rt_trav_oct_inv(const RtTravState &s) { return (s.code >> 9) & 7u; }
RT_HD RtRayTri rt_trav_ray_tri(const RtTravState &s) {
    RtRayPre q;
    q.rcp = s.rcp;
    q.nSx = s.nSx;
    q.nSy = s.nSy;
    q.Sz = s.Sz;
    q.code = s.code;
    return rt_ray_tri_from_pre(s.org, q);
}

/* The dynamically indexed storage is kept apart from RtTravState so that the scalar state is
 * promoted to registers; entries are (base, bits) pairs packed in 64 bits (one access each).
 *   node entry: (child_base, inner hit flags in PRIORITY order << 24 | imask)
 *   tri entry : (tri_base, leaf hit flags in slot order << 24 | tmask)
 * perm() is the octant permutation of a byte of hit flags (bit s -> bit s ^ k); the persistent
 * kernels replace it by a 2 KB shared-memory table (render.cu). */
struct RtTravStacks {
    uint64_t node[RT_STACK_SIZE];  /* pending node groups */
    uint64_t tri[RT_TSTACK_SIZE];  /* pending triangle groups */
    RT_HD uint64_t node_get(int i) const { return node[i]; }
    RT_HD void node_put(int i, uint64_t v) { node[i] = v; }
    RT_HD uint64_t tri_get(int i) const { return tri[i]; }
    RT_HD void tri_put(int i, uint64_t v) { tri[i] = v; }
    RT_HD uint32_t perm(uint32_t k, uint32_t x) const { return rt_xor_perm8(x, k); }
};
RT_HD uint64_t rt_pack2(uint32_t x, uint32_t y) { return (uint64_t)x | ((uint64_t)y << 32); }

RT_HD void rt_trav_init_pre(RtTravState &s, f3 org, const RtRayPre &pre, float tnear, float tfar) {
    s.org = org;
    s.rcp = pre.rcp;
    s.nSx = pre.nSx;
    s.nSy = pre.nSy;
    s.Sz = pre.Sz;
    s.code = pre.code;
    s.tnear = tnear;
    s.t = tfar;
    s.u = 0.0f;
    s.v = 0.0f;
    s.tri = RT_MISS;
    s.sp = 0;
    s.tsp = 0;
#if RT_STACK_TOP
    s.top_x = 0;
    s.top_y = 0;
#endif
    s.ng_x = 0;
    s.ng_y = 0x80000000u; /* the root, as the only child of a virtual group (imask 0: rank 0 for any slot) */
}

RT_HD void rt_trav_init(RtTravState &s, f3 org, f3 dir, float tnear, float tfar) {
    rt_trav_init_pre(s, org, rt_ray_pre(dir), tnear, tfar);
}

RT_HD bool rt_trav_has_node(const RtTravState &s) { return s.ng_y > 0x00ffffffu; }
RT_HD bool rt_trav_has_tri(const RtTravState &s) { return s.tsp > 0; }
RT_HD bool rt_trav_tri_full(const RtTravState &s) { return s.tsp >= RT_TSTACK_SIZE; }

/* precondition: rt_trav_has_node(s) && !rt_trav_tri_full(s). MASK = which axes extract their plane bytes with IDP.4A
 * (rt_byte_to_unit): a per-kernel choice, the best split depends on what else the kernel keeps the two pipes busy with */
template <int MASK, class Stacks>
RT_HD void rt_trav_node_step(const RtBvh &bvh, RtTravState &s, Stacks &k) {
    const uint32_t imask = s.ng_y & 0xffu;
    const int bit = rt_bfind(s.ng_y);
    s.ng_y &= ~(1u << bit);
    if (s.ng_y > 0x00ffffffu) { /* siblings still pending: keep the group for later */
#if RT_STACK_TOP
        if (s.top_y > 0x00ffffffu) {
            k.node_put(s.sp, rt_pack2(s.top_x, s.top_y));
            s.sp++;
        }
        s.top_x = s.ng_x;
        s.top_y = s.ng_y;
#else
        k.node_put(s.sp, rt_pack2(s.ng_x, s.ng_y));
        s.sp++;
#endif
    }
    const uint32_t oct_inv = rt_trav_oct_inv(s);
    const uint32_t slot = ((uint32_t)bit - 24u) ^ oct_inv;
    const uint32_t rel = (uint32_t)rt_popc(imask & ~(0xffffffffu << slot));
    const rt_uint4 *np = bvh.nodes + (size_t)(s.ng_x + rel) * RT_NODE_VEC4;
    rt_uint4 n0, n1, n2, n3, n4;
#if RT_USE_LDG256
    rt_uint4 n5;
    rt_ldg2(np, n0, n1);
    rt_ldg2(np + 2, n2, n3);
    rt_ldg2(np + 4, n4, n5);
    (void)n5;
#else
    n0 = rt_ldg_node(np);
    n1 = rt_ldg_node(np + 1);
    n2 = rt_ldg_node(np + 2);
    n3 = rt_ldg_node(np + 3);
    n4 = rt_ldg_node(np + 4);
#endif
    RT_COUNT_NODE();
    const uint32_t hits = rt_node_test<MASK>(s.org, s.rcp, oct_inv, n0, n2, n3, n4, s.tnear, s.t * RT_BOX_PAD);
    const uint32_t im = n0.w >> 24;
    const uint32_t leaf = hits & ~im;
    s.ng_x = n1.x;
    s.ng_y = (k.perm(oct_inv, hits & im) << 24) | im;
    if (leaf) {
        k.tri_put(s.tsp, rt_pack2(n1.y, (leaf << 24) | n1.z));
        s.tsp++;
    }
#if RT_STACK_TOP
    if (s.ng_y <= 0x00ffffffu && s.top_y > 0x00ffffffu) { /* no child hit: next pending group, from registers */
        s.ng_x = s.top_x;
        s.ng_y = s.top_y;
        s.top_y = 0;
        if (s.sp > 0) { /* refill the register copy; nobody waits for this load */
            s.sp--;
            const uint64_t e = k.node_get(s.sp);
            s.top_x = (uint32_t)e;
            s.top_y = (uint32_t)(e >> 32);
        }
    }
#else
    if (s.ng_y <= 0x00ffffffu && s.sp > 0) { /* no child hit: next pending group */
        s.sp--;
        const uint64_t e = k.node_get(s.sp);
        s.ng_x = (uint32_t)e;
        s.ng_y = (uint32_t)(e >> 32);
    }
#endif
}

template <class Stacks>
RT_HD void rt_trav_node_step(const RtBvh &bvh, RtTravState &s, Stacks &k) {
    rt_trav_node_step<RT_BYTE_IDP_MASK>(bvh, s, k);
}

/* precondition: rt_trav_has_tri(s). Tests ONE triangle of the top group: the lowest remaining
 * triangle position of the lowest hit leaf slot. rt = rt_trav_ray_tri(s), built once per batch of tests. */
template <class Stacks>
RT_HD void rt_trav_tri_step(const RtBvh &bvh, RtTravState &s, Stacks &k, const RtRayTri &rt) {
    const uint64_t e = k.tri_get(s.tsp - 1);
    const uint32_t base = (uint32_t)e;
    uint32_t w = (uint32_t)(e >> 32);
    const uint32_t j = (uint32_t)rt_ctz(w >> 24), sh = 3u * j;  /* lowest hit leaf slot, first bit of its unary count */
    const uint32_t pos = sh + (uint32_t)rt_ctz((w >> sh) & 7u); /* its lowest untested triangle */
    const uint32_t tslot = base + (uint32_t)rt_popc(w & ((1u << pos) - 1u)); /* pos < 24: only tmask bits are counted */
    /* the tested position leaves the mask; every later position of this group lies above it, so
     * the base moves up by one to keep  slot = base + popc(mask below position)  true */
    w &= ~(1u << pos);
    if (((w >> sh) & 7u) == 0u) w &= ~(0x01000000u << j);
    if (w > 0x00ffffffu) k.tri_put(s.tsp - 1, rt_pack2(base + 1u, w));
    else s.tsp--;
    const rt_float4 *tp = bvh.tris + (size_t)tslot * RT_TRI_VEC4;
    rt_float4 a, b, c;
#if RT_USE_LDG256
    rt_float4 d;
    rt_ldg2(tp, a, b);
    rt_ldg2(tp + 2, c, d);
    (void)d;
#else
    a = rt_ldg_hint<RT_TRI_LD_HINT>(tp);
    b = rt_ldg_hint<RT_TRI_LD_HINT>(tp + 1);
    c = rt_ldg_hint<RT_TRI_LD_HINT>(tp + 2);
#endif
    RT_COUNT_TRI();
    rt_tri_test(bvh, rt, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), tslot, rt_f2u(c.w), s.tnear, s.t, s.u, s.v,
                s.tri);
}

/* the closest hit without its id (all that shading needs) */
RT_HD RtHit rt_trav_hit_noid(const RtTravState &s) {
    RtHit h;
    h.t = s.t;
    h.u = s.u;
    h.v = s.v;
    h.tri = s.tri;
    h.gid = 0;
    return h;
}
/* the closest hit of a finished traversal; the id is read back from the triangle record */
RT_HD RtHit rt_trav_hit(const RtBvh &bvh, const RtTravState &s) {
    RtHit h;
    h.t = s.t;
    h.u = s.u;
    h.v = s.v;
    h.tri = s.tri;
    h.gid = s.tri == RT_MISS ? RT_MISS : rt_f2u(rt_ldg(bvh.tris + (size_t)s.tri * RT_TRI_VEC4 + 2).w);
    return h;
}

/* simple front-to-back traversal: triangles are tested right after the node that produced them */
RT_HD RtHit rt_traverse(const RtBvh &bvh, f3 org, f3 dir, float tnear, float tfar) {
    RtTravState s;
    RtTravStacks k;
    rt_trav_init(s, org, dir, tnear, tfar);
    const RtRayTri rt = rt_trav_ray_tri(s);
    while (rt_trav_has_node(s)) {
        rt_trav_node_step(bvh, s, k);
        while (rt_trav_has_tri(s)) rt_trav_tri_step(bvh, s, k, rt);
    }
    return rt_trav_hit(bvh, s);
}

#endif /* RT_TRAVERSE_H */
