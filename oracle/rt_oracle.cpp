/*
 * rt_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE). See rt_oracle.h.
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -fopenmp -shared -fPIC  (oracle/Makefile)
 * Every float expression below is written in the exact association order that is pinned in
 * DESIGN.md ("Arithmetic contract"); -ffp-contract=off forbids FMA fusion so that the CUDA
 * path (nvcc -fmad=false, IEEE div/sqrt) can be compared bit for bit.
 */
#include "rt_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 {
    float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
/* sycl::normalize / glm::normalize, pinned as v * (1 / sqrt(dot(v,v))) (glm: v * inversesqrt) */
inline V3 normalize(V3 a) {
    float inv = 1.0f / std::sqrt(dot(a, a));
    return a * inv;
}
inline float at(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
inline V3 ld3(const float *p) { return v3(p[0], p[1], p[2]); }
inline void st3(float *p, V3 a) {
    p[0] = a.x;
    p[1] = a.y;
    p[2] = a.z;
}

/* sycl::half round trip, round-to-nearest-even (src/camera.hpp:18-28,35-43) */
inline float rh(float v) { return (float)(_Float16)v; }
inline V3 rh(V3 a) { return v3(rh(a.x), rh(a.y), rh(a.z)); }

/* src/xorshift.hpp:11-20 */
struct Rng {
    uint32_t a;
    inline float next() {
        uint32_t x = a;
        x ^= x << 13;
        x ^= x >> 17;
        x ^= x << 5;
        a = x;
        const float scale = 1.f / 4294967296.0f; /* 1.f / (uint64_t{1} << 32) */
        return (float)a * scale;                 /* u32 -> f32 is RNE */
    }
    /* src/xorshift.hpp:22-24 */
    inline float next(float mn, float mx) { return mn + (mx - mn) * next(); }
    /* src/xorshift.hpp:30-36; evaluation order pinned x, y, z (F5) */
    inline V3 vec(float mn, float mx) {
        float x = next(mn, mx);
        float y = next(mn, mx);
        float z = next(mn, mx);
        return v3(x, y, z);
    }
    /* src/xorshift.hpp:38-40: no rejection, cube -> sphere projection */
    inline V3 random_unit_vector() { return normalize(vec(-1.0f, 1.0f)); }
};

/* src/util.hpp:103-107 */
inline bool near_zero(V3 e) {
    const float s = 1e-8f;
    return (std::fabs(e.x) < s) && (std::fabs(e.y) < s) && (std::fabs(e.z) < s);
}
/* src/util.hpp:109-112: length() squared, not dot */
inline float length_squared(V3 v) {
    float l = length(v);
    return l * l;
}
/* src/util.hpp:114-116 */
inline V3 reflect(V3 v, V3 n) { return v - (2.0f * dot(v, n)) * n; }
/* src/util.hpp:118-125 */
inline V3 refract(V3 uv, V3 n, float etai_over_etat) {
    float cos_theta = std::fmin(dot(-uv, n), 1.0f);
    V3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    V3 r_out_parallel = (-std::sqrt(std::fabs(1.0f - length_squared(r_out_perp)))) * n;
    return r_out_perp + r_out_parallel;
}

/* Nearest, repeat, normalised coordinates (sampler at src/render_megakernel.cpp:99-103).
 * Texel selection follows the OpenCL 3.0 spec section 8.2 rule the SYCL sampler maps to:
 * u = (s - floor(s)) * w;  i = (int)floor(u);  if (i > w - 1) i -= w. */
inline int wrap_texel(float s) {
    float u = (s - std::floor(s)) * (float)ORC_TEX_SIZE;
    int i = (int)std::floor(u);
    if (i > ORC_TEX_SIZE - 1) i -= ORC_TEX_SIZE;
    if (i < 0) i = 0; /* NaN guard only */
    return i;
}
inline V3 texture_sample(const uint8_t *tex, uint32_t n_layers, int layer, float su, float sv) {
    if (!tex || layer < 0 || (uint32_t)layer >= n_layers) return v3(0, 0, 0);
    int ix = wrap_texel(su), iy = wrap_texel(sv);
    const uint8_t *p =
        tex + (((size_t)layer * ORC_TEX_SIZE + (size_t)iy) * ORC_TEX_SIZE + (size_t)ix) * 4;
    /* unorm_int8 -> float (src/image_manager.hpp:93); alpha ignored (src/material.hpp:51) */
    return v3((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f);
}

/* src/material.hpp:45-53 */
inline V3 albedo_sample(const orc_material &m, const uint8_t *tex, uint32_t n_layers, float su,
                        float sv) {
    if (m.albedo_image >= 0) return texture_sample(tex, n_layers, m.albedo_image, su, sv);
    return ld3(m.albedo_color);
}

/* src/material.hpp:124-129; pow(x, 5) pinned as ((x*x)*(x*x))*x */
inline float reflectance(float cosine, float ref_idx) {
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    float x = 1.0f - cosine;
    float x2 = x * x;
    float x5 = (x2 * x2) * x;
    return r0 + (1.0f - r0) * x5;
}

/* src/material.hpp:72-86, 99-110, 131-160, 211-227 */
inline bool scatter(const orc_material &m, const uint8_t *tex, uint32_t n_layers, Rng &rng,
                    V3 dir, V3 normal, float su, float sv, V3 &out_dir, V3 &out_att) {
    switch (m.type) {
    case ORC_MAT_DIFFUSE: {
        out_dir = normal + rng.random_unit_vector();
        if (near_zero(dir)) out_dir = normal; /* F8: tests the INCOMING dir */
        out_att = albedo_sample(m, tex, n_layers, su, sv);
        return true;
    }
    case ORC_MAT_METALLIC: {
        V3 reflected = reflect(dir, normal);
        out_dir = reflected + m.roughness * rng.random_unit_vector();
        out_att = albedo_sample(m, tex, n_layers, su, sv);
        return dot(out_dir, normal) > 0.0f;
    }
    case ORC_MAT_DIELECTRIC: {
        out_att = v3(1, 1, 1);
        bool front_face = dot(dir, normal) < 0.0f;
        V3 n = front_face ? normal : -normal;
        float ratio = front_face ? (1.0f / m.ior) : m.ior;
        V3 unit_direction = normalize(dir);
        float cos_theta = std::fmin(dot(-unit_direction, n), 1.0f);
        float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
        bool cannot_refract = ratio * sin_theta > 1.0f;
        /* short-circuit: no draw on total internal reflection (F5). rng(0,1) = 0 + 1*r = r */
        if (cannot_refract || reflectance(cos_theta, ratio) > rng.next(0.0f, 1.0f)) {
            out_dir = reflect(unit_direction, n);
        } else {
            out_dir = refract(unit_direction, n, ratio);
        }
        return true;
    }
    default: return false; /* eNone */
    }
}
inline V3 emitted(const orc_material &m) {
    if (m.type == ORC_MAT_DIFFUSE || m.type == ORC_MAT_METALLIC) return ld3(m.emissive);
    return v3(0, 0, 0);
}

/* ------------------------------------------------------------------ intersection ---- */

struct Hit {
    float t, u, v;
    uint32_t tri; /* global triangle id, 0xffffffff = miss */
};

struct RayPre {
    V3 org;
    V3 mx, my, mz;
};

/* Woop, Benthin, Wald, "Watertight Ray/Triangle Intersection", JCGT 2013, section 3. The axis
 * permutation (kx, ky, kz) is folded into three vectors: sheared coordinate of P = P.m with
 * mx[kx] = 1, mx[kz] = -Sx, my[ky] = 1, my[kz] = -Sy, mz[kz] = Sz, other components 0; products
 * with 0 and 1 are exact, so P.mx is the paper's P[kx] - Sx*P[kz]. The dot is pinned as
 * fma(P.x, m.x, fma(P.y, m.y, P.z*m.z)) (std::fmaf = one IEEE rounding per fma). */
inline float shear(V3 p, V3 m) { return std::fmaf(p.x, m.x, std::fmaf(p.y, m.y, p.z * m.z)); }

inline RayPre ray_precompute(V3 org, V3 dir) {
    RayPre r;
    r.org = org;
    float ax = std::fabs(dir.x), ay = std::fabs(dir.y), az = std::fabs(dir.z);
    int kz = 0;
    float am = ax;
    if (ay > am) {
        kz = 1;
        am = ay;
    }
    if (az > am) { kz = 2; }
    int kx = kz + 1;
    if (kx == 3) kx = 0;
    int ky = kx + 1;
    if (ky == 3) ky = 0;
    if (at(dir, kz) < 0.0f) std::swap(kx, ky);
    const float nSx = -(at(dir, kx) / at(dir, kz));
    const float nSy = -(at(dir, ky) / at(dir, kz));
    const float Sz = 1.0f / at(dir, kz);
    float mx[3] = {0, 0, 0}, my[3] = {0, 0, 0}, mz[3] = {0, 0, 0};
    mx[kx] = 1.0f;
    mx[kz] = nSx;
    my[ky] = 1.0f;
    my[kz] = nSy;
    mz[kz] = Sz;
    r.mx = v3(mx[0], mx[1], mx[2]);
    r.my = v3(my[0], my[1], my[2]);
    r.mz = v3(mz[0], mz[1], mz[2]);
    return r;
}

/* closest-hit update with the deterministic tie-break (min t, then min triangle id) */
inline void tri_test(const RayPre &r, const float *tv, uint32_t tri_id, float tnear, Hit &best) {
    V3 A = ld3(tv) - r.org, B = ld3(tv + 3) - r.org, C = ld3(tv + 6) - r.org;
    float Ax = shear(A, r.mx), Ay = shear(A, r.my);
    float Bx = shear(B, r.mx), By = shear(B, r.my);
    float Cx = shear(C, r.mx), Cy = shear(C, r.my);
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
    float det = (U + V) + W;
    if (det == 0.0f) return;
    float Az = shear(A, r.mz), Bz = shear(B, r.mz), Cz = shear(C, r.mz);
    float T = (U * Az + V * Bz) + W * Cz;
    float rcp = 1.0f / det;
    float t = T * rcp;
    if (!(t > tnear)) return; /* Embree: tnear < t <= tfar */
    if (t < best.t || (t == best.t && tri_id < best.tri)) {
        best.t = t;
        best.u = V * rcp; /* Embree convention: u weights v1, v weights v2 (F11) */
        best.v = W * rcp;
        best.tri = tri_id;
    }
}

struct BvhNode {
    float lo[3], hi[3];
    uint32_t left;  /* inner: index of left child (right = left+1); leaf: first ref */
    uint32_t count; /* 0 = inner, else number of refs */
};

} // namespace

struct orc_scene {
    struct Inst {
        std::vector<float> normals, uvs;
        std::vector<uint32_t> indices;
        float nmat[9]; /* column-major transpose(inverse(mat3(T))) */
        orc_material mat;
        uint32_t first_tri;
    };
    std::vector<Inst> insts;
    std::vector<float> tris;        /* 9 floats per world-space triangle */
    std::vector<uint32_t> tri_inst; /* instance of each global triangle */
    std::vector<uint8_t> tex;
    uint32_t n_layers = 0;
    V3 sky{0.5f, 0.7f, 1.0f};
    /* CPU BVH (our stand-in for Embree's; validated against brute force in tests) */
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> refs;
};

namespace {

/* ---- CPU BVH: binned SAH, built lazily ---- */
struct BuildCtx {
    const std::vector<float> &tris;
    std::vector<uint32_t> &refs;
    std::vector<BvhNode> &nodes;
    std::vector<float> cen; /* centroids */
    std::vector<float> blo, bhi;
};

inline void tri_bounds(const float *tv, float lo[3], float hi[3]) {
    for (int k = 0; k < 3; k++) {
        lo[k] = std::min(tv[k], std::min(tv[3 + k], tv[6 + k]));
        hi[k] = std::max(tv[k], std::max(tv[3 + k], tv[6 + k]));
    }
}
inline float half_area(const float lo[3], const float hi[3]) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

void build_rec(BuildCtx &c, uint32_t node_idx, uint32_t first, uint32_t count) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = first; i < first + count; i++) {
        uint32_t r = c.refs[i];
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], c.blo[r * 3 + k]);
            hi[k] = std::max(hi[k], c.bhi[r * 3 + k]);
            clo[k] = std::min(clo[k], c.cen[r * 3 + k]);
            chi[k] = std::max(chi[k], c.cen[r * 3 + k]);
        }
    }
    {
        BvhNode &n = c.nodes[node_idx];
        for (int k = 0; k < 3; k++) {
            n.lo[k] = lo[k];
            n.hi[k] = hi[k];
        }
    }
    auto make_leaf = [&]() {
        c.nodes[node_idx].left = first;
        c.nodes[node_idx].count = count;
    };
    if (count <= 2) {
        make_leaf();
        return;
    }
    const int NB = 16;
    int best_axis = -1, best_split = -1;
    float best_cost = INFINITY;
    for (int ax = 0; ax < 3; ax++) {
        float ext = chi[ax] - clo[ax];
        if (!(ext > 0.0f)) continue;
        float bl[NB][3], bh[NB][3];
        uint32_t bc[NB];
        for (int b = 0; b < NB; b++) {
            bc[b] = 0;
            for (int k = 0; k < 3; k++) {
                bl[b][k] = INFINITY;
                bh[b][k] = -INFINITY;
            }
        }
        float scale = (float)NB / ext;
        for (uint32_t i = first; i < first + count; i++) {
            uint32_t r = c.refs[i];
            int b = std::min(NB - 1, (int)((c.cen[r * 3 + ax] - clo[ax]) * scale));
            bc[b]++;
            for (int k = 0; k < 3; k++) {
                bl[b][k] = std::min(bl[b][k], c.blo[r * 3 + k]);
                bh[b][k] = std::max(bh[b][k], c.bhi[r * 3 + k]);
            }
        }
        float ra[NB];
        uint32_t rc[NB];
        float tl[3] = {INFINITY, INFINITY, INFINITY}, th[3] = {-INFINITY, -INFINITY, -INFINITY};
        uint32_t cnt = 0;
        for (int b = NB - 1; b > 0; b--) {
            for (int k = 0; k < 3; k++) {
                tl[k] = std::min(tl[k], bl[b][k]);
                th[k] = std::max(th[k], bh[b][k]);
            }
            cnt += bc[b];
            ra[b] = cnt ? half_area(tl, th) : 0.0f;
            rc[b] = cnt;
        }
        for (int k = 0; k < 3; k++) {
            tl[k] = INFINITY;
            th[k] = -INFINITY;
        }
        cnt = 0;
        for (int b = 0; b < NB - 1; b++) {
            for (int k = 0; k < 3; k++) {
                tl[k] = std::min(tl[k], bl[b][k]);
                th[k] = std::max(th[k], bh[b][k]);
            }
            cnt += bc[b];
            if (cnt == 0 || rc[b + 1] == 0) continue;
            float cost = half_area(tl, th) * (float)cnt + ra[b + 1] * (float)rc[b + 1];
            if (cost < best_cost) {
                best_cost = cost;
                best_axis = ax;
                best_split = b;
            }
        }
    }
    uint32_t mid;
    if (best_axis < 0) {
        if (count <= 8) {
            make_leaf();
            return;
        }
        mid = first + count / 2; /* coincident centroids: split in the middle */
    } else {
        float leaf_cost = half_area(lo, hi) * (float)count;
        if (count <= 4 && leaf_cost <= best_cost) {
            make_leaf();
            return;
        }
        float ext = chi[best_axis] - clo[best_axis];
        float scale = (float)NB / ext;
        uint32_t *b0 = c.refs.data() + first, *b1 = b0 + count;
        uint32_t *m = std::partition(b0, b1, [&](uint32_t r) {
            int b = std::min(NB - 1, (int)((c.cen[r * 3 + best_axis] - clo[best_axis]) * scale));
            return b <= best_split;
        });
        mid = (uint32_t)(m - c.refs.data());
        if (mid == first || mid == first + count) mid = first + count / 2;
    }
    uint32_t left = (uint32_t)c.nodes.size();
    c.nodes.push_back(BvhNode{});
    c.nodes.push_back(BvhNode{});
    c.nodes[node_idx].left = left;
    c.nodes[node_idx].count = 0;
    build_rec(c, left, first, mid - first);
    build_rec(c, left + 1, mid, first + count - mid);
}

void ensure_bvh(orc_scene *s) {
    if (!s->nodes.empty()) return;
    uint32_t n = (uint32_t)(s->tris.size() / 9);
    s->refs.resize(n);
    for (uint32_t i = 0; i < n; i++) s->refs[i] = i;
    s->nodes.reserve(2 * (size_t)n + 1);
    s->nodes.push_back(BvhNode{});
    if (n == 0) {
        s->nodes[0].count = 0;
        s->nodes[0].left = 0;
        for (int k = 0; k < 3; k++) {
            s->nodes[0].lo[k] = INFINITY;
            s->nodes[0].hi[k] = -INFINITY;
        }
        return;
    }
    BuildCtx c{s->tris, s->refs, s->nodes, {}, {}, {}};
    c.cen.resize((size_t)n * 3);
    c.blo.resize((size_t)n * 3);
    c.bhi.resize((size_t)n * 3);
    for (uint32_t i = 0; i < n; i++) {
        tri_bounds(&s->tris[(size_t)i * 9], &c.blo[(size_t)i * 3], &c.bhi[(size_t)i * 3]);
        for (int k = 0; k < 3; k++)
            c.cen[(size_t)i * 3 + k] = 0.5f * (c.blo[(size_t)i * 3 + k] + c.bhi[(size_t)i * 3 + k]);
    }
    build_rec(c, 0, 0, n);
}

/* conservative slab test: NaN-safe min/max. The triangle test computes t and "inside" with its own roundings, which grow
 * with the distance of the triangle's VERTICES from the ray origin: next to a 100 m quad it reports t = 1.04e-4 for a
 * bounce ray whose exact crossing lies at 9.0e-5 (< tnear), and accepts points a few 1e-6 outside the triangle's exact box.
 * The brute-force loop IS the definition of the closest hit, so the box test must never cull what the triangle test would
 * accept: every box plane is moved outward by 2^-18 of the largest coordinate of box and origin (all axes), and the far
 * side keeps a relative pad for the slab arithmetic itself. Found by the stadium scene: two pixels of an 80 x 40 crop
 * differed between this BVH and the brute-force loop (the CUDA path agreed with brute force). */
inline bool box_hit(const BvhNode &n, V3 org, V3 inv, float tnear, float tfar, float &tentry) {
    float tmin = tnear, tmax = tfar;
    const float o[3] = {org.x, org.y, org.z}, iv[3] = {inv.x, inv.y, inv.z};
    float m = 0.0f;
    for (int k = 0; k < 3; k++) m = std::fmax(m, std::fmax(std::fmax(std::fabs(n.lo[k]), std::fabs(n.hi[k])), std::fabs(o[k])));
    const float e = m < 3.0e38f ? m * 3.8146973e-6f : 0.0f;
    for (int k = 0; k < 3; k++) {
        float t0 = ((n.lo[k] - e) - o[k]) * iv[k];
        float t1 = ((n.hi[k] + e) - o[k]) * iv[k];
        float a = std::fmin(t0, t1), b = std::fmax(t0, t1);
        tmin = std::fmax(tmin, a); /* fmax ignores NaN (0*inf) */
        tmax = std::fmin(tmax, b);
    }
    tentry = tmin;
    /* tmin >= tnear > 0 here, so a relative pad on the far side is a pad on the interval */
    return tmin <= tmax * 1.0000005f;
}

inline Hit intersect_bvh(const orc_scene *s, V3 org, V3 dir, float tnear, float tfar) {
    Hit best{tfar, 0, 0, 0xffffffffu};
    if (s->tris.empty()) return best;
    RayPre pre = ray_precompute(org, dir);
    V3 inv = v3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        uint32_t ni = stack[--sp];
        const BvhNode &n = s->nodes[ni];
        float te;
        if (!box_hit(n, org, inv, tnear, best.t, te)) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count; i++) {
                uint32_t tri = s->refs[n.left + i];
                tri_test(pre, &s->tris[(size_t)tri * 9], tri, tnear, best);
            }
        } else {
            float t0, t1;
            bool h0 = box_hit(s->nodes[n.left], org, inv, tnear, best.t, t0);
            bool h1 = box_hit(s->nodes[n.left + 1], org, inv, tnear, best.t, t1);
            if (h0 && h1) {
                if (sp + 2 > 128) continue; /* cannot happen for sane trees */
                if (t0 <= t1) {
                    stack[sp++] = n.left + 1;
                    stack[sp++] = n.left;
                } else {
                    stack[sp++] = n.left;
                    stack[sp++] = n.left + 1;
                }
            } else if (h0) {
                stack[sp++] = n.left;
            } else if (h1) {
                stack[sp++] = n.left + 1;
            }
        }
    }
    return best;
}

inline Hit intersect_brute(const orc_scene *s, V3 org, V3 dir, float tnear, float tfar) {
    Hit best{tfar, 0, 0, 0xffffffffu};
    RayPre pre = ray_precompute(org, dir);
    uint32_t n = (uint32_t)(s->tris.size() / 9);
    for (uint32_t i = 0; i < n; i++) tri_test(pre, &s->tris[(size_t)i * 9], i, tnear, best);
    return best;
}

inline Hit intersect(const orc_scene *s, bool use_bvh, V3 org, V3 dir, float tnear, float tfar) {
    return use_bvh ? intersect_bvh(s, org, dir, tnear, tfar)
                   : intersect_brute(s, org, dir, tnear, tfar);
}

/* ------------------------------------------------------------------ shading ---- */

struct RayState { /* RayData, src/camera.hpp:12-63: fp32 origin, fp16-rounded dir/att/rad */
    V3 org, dir, att, rad;
};

/* src/camera.hpp:109-131 */
inline RayState camera_get_ray(const orc_camera &c, int x, int y, Rng &rng) {
    V3 p00 = ld3(c.pixel00_loc), du = ld3(c.pixel_delta_u), dv = ld3(c.pixel_delta_v);
    V3 pixel_center = (p00 + ((float)x * du)) + ((float)y * dv);
    float px = -0.5f + rng.next(); /* pixel_sample_square: px then py */
    float py = -0.5f + rng.next();
    V3 pixel_sample = pixel_center + ((px * du) + (py * dv));
    V3 origin = ld3(c.center);
    V3 direction = pixel_sample - origin;
    RayState r;
    r.org = origin;
    r.dir = rh(direction); /* RayData ctor quantises the direction (F6) */
    r.att = v3(1, 1, 1);
    r.rad = v3(0, 0, 0);
    return r;
}

/* src/trace_ray.hpp:11-82. Returns true when the path terminated with colour `result`. */
inline bool trace_ray(const orc_scene *s, bool use_bvh, Rng &rng, V3 &org, V3 &dir, V3 &att,
                      V3 &rad, V3 &result) {
    Hit h = intersect(s, use_bvh, org, dir, 0.0001f, std::numeric_limits<float>::infinity());
    if (h.tri == 0xffffffffu) {
        result = att * (s->sky + rad); /* :25-27 */
        return true;
    }
    const orc_scene::Inst &g = s->insts[s->tri_inst[h.tri]]; /* instID[0] -> GeometryData */
    uint32_t prim = h.tri - g.first_tri;
    const uint32_t *idx = &g.indices[(size_t)prim * 3];
    float bx = h.u, by = h.v;
    float bw = (1.0f - bx) - by; /* (1 - bary.x - bary.y) */
    /* :43-44 */
    const float *t0 = &g.uvs[(size_t)idx[0] * 2], *t1 = &g.uvs[(size_t)idx[1] * 2],
                *t2 = &g.uvs[(size_t)idx[2] * 2];
    float su = (bw * t0[0] + bx * t1[0]) + by * t2[0];
    float sv = (bw * t0[1] + bx * t1[1]) + by * t2[1];
    /* :47-50 glm::normalize */
    V3 n0 = ld3(&g.normals[(size_t)idx[0] * 3]), n1 = ld3(&g.normals[(size_t)idx[1] * 3]),
       n2 = ld3(&g.normals[(size_t)idx[2] * 3]);
    V3 vn = normalize(((bw * n0) + (bx * n1)) + (by * n2));
    /* :52 glm mat3 * vec3 = m[0]*v.x + m[1]*v.y + m[2]*v.z */
    const float *m = g.nmat;
    V3 gn = v3((m[0] * vn.x + m[3] * vn.y) + m[6] * vn.z, (m[1] * vn.x + m[4] * vn.y) + m[7] * vn.z,
               (m[2] * vn.x + m[5] * vn.y) + m[8] * vn.z);
    V3 normal = normalize(gn);       /* :53-54 */
    V3 ndir = normalize(dir);        /* :56-57 */
    rad = rad + emitted(g.mat);      /* :59, F7: not multiplied by attenuation */
    V3 sdir, satt;
    if (scatter(g.mat, s->tex.empty() ? nullptr : s->tex.data(), s->n_layers, rng, ndir, normal,
                su, sv, sdir, satt)) {
        org = org + dir * h.t; /* :62-64 org + dir * tfar (unfused) */
        dir = sdir;
        att = att * satt;
        return false;
    }
    result = att * rad; /* :73 */
    return true;
}

/* src/render_megakernel.cpp:20-63 (the wavefront path runs the identical per-ray sequence,
 * src/render_wavefront.cpp:244-296, because every pixel has at most one ray in flight) */
inline V3 render_sample(const orc_scene *s, const orc_camera &cam, bool use_bvh, Rng &rng, int x,
                        int y, uint32_t max_depth, uint64_t &ray_count, bool roulette = false) {
    RayState r = camera_get_ray(cam, x, y, rng);
    for (uint32_t i = 0; i < max_depth; i++) {
        ray_count++;
        V3 att = r.att, rad = r.rad; /* fp16 -> fp32 */
        V3 org = r.org, dir = r.dir;
        V3 res;
        bool done = trace_ray(s, use_bvh, rng, org, dir, att, rad, res);
        if (!done && roulette && i + 1 >= 3 && i + 1 < max_depth) {
            /* optional Russian roulette (not in the reference: PLAN.md:23-24 lists it as open). From the third bounce
             * on the path survives with probability q = clamp(max(att), 0.05, 1) — one extra draw — and its
             * attenuation is divided by q; a killed path contributes black like a path cut at max_depth. */
            float q = std::fmin(std::fmax(std::fmax(std::fmax(att.x, att.y), att.z), 0.05f), 1.0f);
            if (rng.next() > q) return v3(0, 0, 0);
            float inv = 1.0f / q;
            att = att * inv;
        }
        r.org = org;
        r.dir = rh(dir);
        r.att = rh(att);
        r.rad = rh(rad);
        if (done) return res;
    }
    return v3(0, 0, 0); /* F7: survivors contribute black */
}

inline float clamp01(float v) { return std::fmin(std::fmax(v, 0.0f), 1.0f); }

double g_last_render_seconds = 0.0;

} // namespace

/* =========================================================== extern "C" surface ==== */

extern "C" {

float orc_xorshift_next(uint32_t *state) {
    Rng r{*state};
    float f = r.next();
    *state = r.a;
    return f;
}

void orc_random_unit_vector(uint32_t *state, float out[3]) {
    Rng r{*state};
    st3(out, r.random_unit_vector());
    *state = r.a;
}

float orc_round_half(float v) { return rh(v); }

uint8_t orc_output_byte(float g) {
    /* image write to unorm_int8: convert_uchar_sat_rte(f * 255.0f) (OpenCL 3.0, 8.3.1.1) */
    float c = g * 255.0f;
    float q;
    if (!(c > 0.0f)) q = 0.0f; /* also NaN */
    else if (c >= 255.0f) q = 255.0f;
    else q = std::nearbyint(c); /* default rounding mode = RNE */
    /* src/util.hpp:16-22: read back normalised, *255.0f, truncating cast */
    float back = (q / 255.0f) * 255.0f;
    return (uint8_t)back;
}

uint32_t orc_pixel_seed(int32_t mode, int32_t x, int32_t y, int32_t width, int32_t height) {
    if (mode == ORC_MODE_MEGAKERNEL) {
        /* src/render_megakernel.cpp:90-93,117,145: get_global_linear_id of a (W_pad, H_pad)
         * nd_range = x * H_pad + y, H_pad = ceil(H/8)*8; std::hash<size_t> is the identity */
        uint64_t h_pad = (uint64_t)((height + 7) / 8) * 8;
        return (uint32_t)((uint64_t)x * h_pad + (uint64_t)y);
    }
    /* src/render_wavefront.cpp:69-73 */
    (void)height;
    return (uint32_t)((uint64_t)x + (uint64_t)y * (uint64_t)width);
}

void orc_camera_init(orc_camera *cam, int32_t width, int32_t height, const float pos[3],
                     const float dirv[3], float focal_length) {
    cam->img_size[0] = width;
    cam->img_size[1] = height;
    V3 center = ld3(pos);
    V3 dir = normalize(ld3(dirv));
    V3 world_up = v3(0, 1, 0);
    V3 right = normalize(cross(dir, world_up));
    V3 up = normalize(cross(right, dir));
    float vp0 = 1.0f * ((float)width / (float)height), vp1 = 1.0f;
    V3 viewport_u = (-right) * vp0;
    V3 viewport_v = up * vp1;
    V3 p00 = ((center + viewport_u) + viewport_v) + dir * focal_length;
    V3 du = right / ((float)width / (vp0 * 2.0f));
    V3 dv = (-up) / ((float)height / (vp1 * 2.0f));
    st3(cam->center, center);
    st3(cam->pixel00_loc, p00);
    st3(cam->pixel_delta_u, du);
    st3(cam->pixel_delta_v, dv);
}

void orc_camera_get_ray(const orc_camera *cam, int32_t x, int32_t y, uint32_t *rng_state,
                        float org[3], float dir[3]) {
    Rng r{*rng_state};
    RayState rs = camera_get_ray(*cam, x, y, r);
    *rng_state = r.a;
    st3(org, rs.org);
    st3(dir, rs.dir);
}

int orc_material_scatter(const orc_material *m, const uint8_t *textures, uint32_t n_layers,
                         uint32_t *rng_state, const float dir[3], const float normal[3],
                         const float uv[2], float out_dir[3], float out_att[3]) {
    Rng r{*rng_state};
    V3 od = v3(0, 0, 0), oa = v3(0, 0, 0);
    bool ok = scatter(*m, textures, n_layers, r, ld3(dir), ld3(normal), uv[0], uv[1], od, oa);
    *rng_state = r.a;
    st3(out_dir, od);
    st3(out_att, oa);
    return ok ? 1 : 0;
}

void orc_texture_sample(const uint8_t *textures, uint32_t n_layers, int32_t layer,
                        const float uv[2], float out_rgb[3]) {
    st3(out_rgb, texture_sample(textures, n_layers, layer, uv[0], uv[1]));
}

void orc_normal_matrix(const float T[16], float out[9]) {
    /* glm::mat3(T): upper-left 3x3, m[c][r] = T[c*4+r]; glm::inverse(mat3) cofactor form
     * (glm/detail/func_matrix.inl, compute_inverse<3,3>), then transpose. glm is a
     * third-party dependency absent from /root/reference (unpinned version). */
    float m[3][3];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) m[c][r] = T[c * 4 + r];
    float ood = 1.0f / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) -
                        m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) +
                        m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    float inv[3][3];
    inv[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * ood;
    inv[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * ood;
    inv[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * ood;
    inv[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * ood;
    inv[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * ood;
    inv[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * ood;
    inv[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * ood;
    inv[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * ood;
    inv[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * ood;
    /* transpose: out[c][r] = inv[r][c] */
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) out[c * 3 + r] = inv[r][c];
}

orc_scene *orc_scene_create(const orc_instance *instances, uint32_t n_instances,
                            const uint8_t *textures, uint32_t n_layers, const float sky[3]) {
    orc_scene *s = new orc_scene();
    if (sky) s->sky = ld3(sky);
    if (textures && n_layers) {
        s->n_layers = n_layers;
        s->tex.assign(textures, textures + (size_t)n_layers * ORC_TEX_SIZE * ORC_TEX_SIZE * 4);
    }
    s->insts.resize(n_instances);
    uint32_t tri_base = 0;
    for (uint32_t i = 0; i < n_instances; i++) {
        const orc_instance &in = instances[i];
        orc_scene::Inst &g = s->insts[i];
        g.normals.assign(in.normals, in.normals + (size_t)in.vertex_count * 3);
        g.uvs.assign(in.uvs, in.uvs + (size_t)in.vertex_count * 2);
        g.indices.assign(in.indices, in.indices + in.index_count);
        g.mat = in.material;
        g.first_tri = tri_base;
        orc_normal_matrix(in.transform, g.nmat);
        /* world-space flatten: glm mat4 * vec4(v,1) = m[0]*x + m[1]*y + m[2]*z + m[3]
         * (the instance transform of src/scene.cpp:491-494 applied to the geometry instead of
         * to the ray; t, u, v are invariant under the affine map) */
        const float *T = in.transform;
        std::vector<float> wp((size_t)in.vertex_count * 3);
        for (uint32_t v = 0; v < in.vertex_count; v++) {
            float x = in.positions[(size_t)v * 3], y = in.positions[(size_t)v * 3 + 1],
                  z = in.positions[(size_t)v * 3 + 2];
            for (int r = 0; r < 3; r++)
                wp[(size_t)v * 3 + r] = ((T[r] * x + T[4 + r] * y) + T[8 + r] * z) + T[12 + r];
        }
        uint32_t ntri = in.index_count / 3;
        for (uint32_t t = 0; t < ntri; t++) {
            for (int k = 0; k < 3; k++) {
                uint32_t vi = in.indices[(size_t)t * 3 + k];
                s->tris.push_back(wp[(size_t)vi * 3]);
                s->tris.push_back(wp[(size_t)vi * 3 + 1]);
                s->tris.push_back(wp[(size_t)vi * 3 + 2]);
            }
            s->tri_inst.push_back(i);
        }
        tri_base += ntri;
    }
    return s;
}

void orc_scene_destroy(orc_scene *s) { delete s; }
uint64_t orc_scene_triangle_count(const orc_scene *s) { return s->tris.size() / 9; }
const float *orc_scene_world_triangles(const orc_scene *s) { return s->tris.data(); }

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_intersect(const orc_scene *cs, int32_t use_bvh, int32_t threads, uint64_t n,
                   const float *org, const float *dir, float tnear, float tfar, int32_t *inst,
                   int32_t *prim, float *u, float *v, float *t) {
    orc_scene *s = const_cast<orc_scene *>(cs);
    if (use_bvh) ensure_bvh(s);
    int nt = threads > 0 ? threads : orc_max_threads();
    (void)nt;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Hit h = intersect(s, use_bvh != 0, ld3(org + i * 3), ld3(dir + i * 3), tnear, tfar);
        if (h.tri == 0xffffffffu) {
            inst[i] = -1;
            prim[i] = -1;
            u[i] = 0;
            v[i] = 0;
            t[i] = tfar;
        } else {
            uint32_t gi = s->tri_inst[h.tri];
            inst[i] = (int32_t)gi;
            prim[i] = (int32_t)(h.tri - s->insts[gi].first_tri);
            u[i] = h.u;
            v[i] = h.v;
            t[i] = h.t;
        }
    }
}

uint64_t orc_render(const orc_scene *cs, const orc_camera *cam, const orc_render_params *p,
                    float *accum, uint8_t *rgba8, uint32_t *rng_out) {
    orc_scene *s = const_cast<orc_scene *>(cs);
    bool use_bvh = p->use_bvh != 0;
    if (use_bvh) ensure_bvh(s);
    const int W = cam->img_size[0], H = cam->img_size[1];
    int x0 = p->x0, y0 = p->y0, x1 = p->x1, y1 = p->y1;
    if (x1 <= x0 || y1 <= y0) {
        x0 = 0;
        y0 = 0;
        x1 = W;
        y1 = H;
    }
    const int cw = x1 - x0;
    const bool wave = p->mode == ORC_MODE_WAVEFRONT;
    const uint32_t spp = p->sample_count;
    uint64_t total_rays = 0;
    int nt = p->threads > 0 ? p->threads : orc_max_threads();
    (void)nt;
    auto t_begin = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_rays) num_threads(nt)
    for (int y = y0; y < y1; y++) {
        for (int x = x0; x < x1; x++) {
            Rng rng{orc_pixel_seed(p->mode, x, y, W, H) ^ p->seed_salt};
            uint64_t rays = 0;
            V3 sum = v3(0, 0, 0);
            for (uint32_t sidx = 0; sidx < spp; sidx++) {
                V3 c = render_sample(s, *cam, use_bvh, rng, x, y, p->max_depth, rays, p->roulette != 0);
                if (wave) /* src/render_wavefront.cpp:277: per-sample clamp (F9) */
                    c = v3(clamp01(c.x), clamp01(c.y), clamp01(c.z));
                sum = sum + c; /* megakernel :151-152; wavefront merge_samples :350-352 */
            }
            total_rays += rays;
            size_t o = (size_t)(y - y0) * cw + (size_t)(x - x0);
            if (accum) {
                accum[o * 4 + 0] = sum.x;
                accum[o * 4 + 1] = sum.y;
                accum[o * 4 + 2] = sum.z;
                accum[o * 4 + 3] = (float)spp;
            }
            if (rgba8) {
                /* mean, sqrt gamma (src/render_megakernel.cpp:154-156,
                 * src/render_wavefront.cpp:387-389), then F10 */
                V3 mean = sum / (float)spp;
                rgba8[o * 4 + 0] = orc_output_byte(std::sqrt(mean.x));
                rgba8[o * 4 + 1] = orc_output_byte(std::sqrt(mean.y));
                rgba8[o * 4 + 2] = orc_output_byte(std::sqrt(mean.z));
                rgba8[o * 4 + 3] = orc_output_byte(1.0f);
            }
            if (rng_out) rng_out[o] = rng.a;
        }
    }
    auto t_end = std::chrono::steady_clock::now();
    g_last_render_seconds = std::chrono::duration<double>(t_end - t_begin).count();
    return total_rays;
}

double orc_last_render_seconds(void) { return g_last_render_seconds; }

} /* extern "C" */
