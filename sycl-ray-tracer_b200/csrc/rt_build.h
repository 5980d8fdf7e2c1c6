/*
 * rt_build.h — per-item logic of the GPU BVH build (the rtcCommitScene replacement,
 * src/scene.cpp:101-107,406-439,483-507):
 *
 *   flatten      instance transform applied to the geometry (one world-space triangle soup;
 *                Embree instead transforms the ray per instance, t/u/v are invariant)
 *   split        large triangles become several REFERENCES with the tight boxes of their pieces
 *                (longest-edge bisection): the tree's leaves are references, the leaf records are copies
 *                of the whole triangle, so hits are unchanged
 *   morton       63-bit Morton code of the reference-AABB centroid
 *   (sort)       cub::DeviceRadixSort in bvh_build.cu
 *   karras       binary radix tree over the sorted codes (Karras, HPG 2012)
 *   fit          bottom-up AABB fit, one atomic flag per inner node
 *   wide_select  greedy surface-area collapse of the binary tree into <= 8 children
 *   wide_emit    quantise child boxes, assign octant-ordered slots, write the 80-byte node (imask / tmask),
 *                leaf-ordered triangles and the pre-gathered shading records
 *
 * Each function handles ONE item and is wrapped by a thin kernel in bvh_build.cu.
 */
#ifndef RT_BUILD_H
#define RT_BUILD_H

#include "rt_shade.h"

#define RT_LEAF_MAX 3 /* triangles per leaf child (unary count in the slot's 3 tmask bits) */

struct RtInstanceGeom {
    float transform[16]; /* column-major */
    uint32_t first_vertex, first_index, first_tri, tri_count;
};

struct RtBuild {
    uint32_t n_tris, n_inst;
    uint32_t n_items;       /* leaves of the tree: n_tris, or the number of references when triangles were split */
    /* triangle splitting: null when no triangle was split (reference i = triangle i, box = the triangle's) */
    const uint32_t *ref_tri;            /* n_items: global triangle id of a reference */
    const rt_float4 *ref_lo, *ref_hi;   /* n_items: its box */
    /* concatenated instance geometry (object space) */
    const float *positions, *normals, *uvs;
    const uint32_t *indices;
    const RtInstanceGeom *geom;
    /* flatten */
    rt_float4 *wtris;       /* 3 per triangle in global-id order; .w of the third = gid bits */
    int32_t *cen_bounds;    /* 6 ordered-int floats: min xyz, max xyz of AABB centroids */
    /* morton + sort */
    uint64_t *keys;
    uint32_t *vals;         /* sorted: vals[k] = global id of the k-th triangle in Morton order */
    /* binary radix tree; unified node ids: inner i in [0, n-2], leaf j = (n-1) + j */
    uint32_t *left, *right; /* n-1 */
    uint32_t *parent;       /* 2n-1 */
    uint32_t *range_first, *range_last; /* n-1 */
    rt_float4 *box_lo, *box_hi;         /* 2n-1 */
    uint32_t *flags;                    /* n-1 */
    /* SAH-optimal wide collapse (Ylitie et al. 2017, section 4.1), filled bottom-up by the fit pass */
    rt_float4 *dp_cost;                 /* 2 per node: C(n,1..4), C(n,5..7) */
    uint32_t *dp_dec;                   /* 2 per node: 8 decision bytes, see rt_dp_node */
    /* wide collapse (per level) */
    const uint32_t *items;  /* binary node id of every wide node of this level */
    uint32_t *next_items;
    uint32_t *sel;          /* 8 per item: unified ids of the chosen children, RT_MISS = none */
    uint64_t *counts;       /* per item: inner children << 32 | leaf triangles */
    const uint64_t *offsets;/* exclusive scan of counts */
    uint32_t level_first_node, next_level_first_node, tri_cursor;
    /* outputs */
    rt_uint4 *nodes;        /* RT_NODE_VEC4 per wide node */
    rt_float4 *tris;        /* RT_TRI_VEC4 per triangle, leaf order */
    rt_float4 *shade;       /* 4 per triangle, leaf order */
};

/* order-preserving float <-> int map for atomicMin/Max */
RT_HD int32_t rt_float_to_ordered(float f) {
    int32_t i = (int32_t)rt_f2u(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
RT_HD float rt_ordered_to_float(int32_t i) { return rt_u2f((uint32_t)(i >= 0 ? i : i ^ 0x7fffffff)); }

RT_HD uint32_t rt_find_instance(const RtBuild &b, uint32_t gid) {
    uint32_t lo = 0, hi = b.n_inst; /* last instance with first_tri <= gid and tri_count > 0 */
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (b.geom[mid].first_tri <= gid) lo = mid;
        else hi = mid;
    }
    return lo;
}

/* ---- flatten: world-space vertices of global triangle gid ------------------------------- */
RT_HD void rt_flatten_tri(const RtBuild &b, uint32_t gid, f3 &lo, f3 &hi) {
    const uint32_t inst = rt_find_instance(b, gid);
    const RtInstanceGeom &g = b.geom[inst];
    const uint32_t prim = gid - g.first_tri;
    const float *T = g.transform;
    f3 w[3];
    for (int k = 0; k < 3; k++) {
        const uint32_t vi = b.indices[(size_t)g.first_index + (size_t)prim * 3 + k] + g.first_vertex;
        const float x = b.positions[(size_t)vi * 3], y = b.positions[(size_t)vi * 3 + 1],
                    z = b.positions[(size_t)vi * 3 + 2];
        /* glm mat4 * vec4(v, 1) = m[0]*x + m[1]*y + m[2]*z + m[3] */
        w[k] = mk3(((T[0] * x + T[4] * y) + T[8] * z) + T[12], ((T[1] * x + T[5] * y) + T[9] * z) + T[13],
                   ((T[2] * x + T[6] * y) + T[10] * z) + T[14]);
    }
    rt_float4 a, c, d;
    a.x = w[0].x; a.y = w[0].y; a.z = w[0].z; a.w = 0.0f;
    c.x = w[1].x; c.y = w[1].y; c.z = w[1].z; c.w = 0.0f;
    d.x = w[2].x; d.y = w[2].y; d.z = w[2].z; d.w = rt_u2f(gid);
    b.wtris[(size_t)gid * 3] = a;
    b.wtris[(size_t)gid * 3 + 1] = c;
    b.wtris[(size_t)gid * 3 + 2] = d;
    lo = mk3(rt_min3(w[0].x, w[1].x, w[2].x), rt_min3(w[0].y, w[1].y, w[2].y), rt_min3(w[0].z, w[1].z, w[2].z));
    hi = mk3(rt_max3(w[0].x, w[1].x, w[2].x), rt_max3(w[0].y, w[1].y, w[2].y), rt_max3(w[0].z, w[1].z, w[2].z));
}

RT_HD void rt_tri_box(const RtBuild &b, uint32_t gid, f3 &lo, f3 &hi) {
    const rt_float4 a = b.wtris[(size_t)gid * 3], c = b.wtris[(size_t)gid * 3 + 1], d = b.wtris[(size_t)gid * 3 + 2];
    lo = mk3(rt_min3(a.x, c.x, d.x), rt_min3(a.y, c.y, d.y), rt_min3(a.z, c.z, d.z));
    hi = mk3(rt_max3(a.x, c.x, d.x), rt_max3(a.y, c.y, d.y), rt_max3(a.z, c.z, d.z));
}

/* ---- split: references of large triangles --------------------------------------------------------------------
 * A Morton-ordered tree files a triangle under the cell of its box centre, so ONE long or large triangle (an
 * architectural quad, a cable) drags a box as big as itself into a subtree of small neighbours, and every ray that
 * crosses that box visits the subtree ("teapot in a stadium"; SAH builders such as Embree's, src/scene.cpp:406-439,
 * avoid it by construction). Remedy: early split. A triangle whose longest edge exceeds `len` is bisected at the
 * midpoint of its longest edge, recursively, and every piece becomes a REFERENCE (triangle id + box of the piece,
 * clipped to the triangle's own box, padded by 8 ulp for the rounded midpoints). The tree is built over references;
 * a leaf stores a copy of the WHOLE triangle (same id), so intersection results — closest t, ties by id — cannot
 * change, a duplicate is merely tested twice. `len` = 2^-RT_SPLIT_TAU_LOG2 of the diagonal of the centroid bounds,
 * doubled by the host until the extra references fit the budget (scenes made of large triangles only).
 * OPT-IN (RT_SPLIT=1 at rt_scene_commit): on the scenes measured so far (tools/tree_quality.py) every threshold buys fewer
 * triangle tests with MORE node visits — stadium 7.4 / 7.3 (visits / tests per ray) -> 7.5 / 6.5 at 1/4 of the diagonal,
 * 8.9 / 5.7 at 1/8, 10.8 / 4.1 at 1/32; Cornell 3.8 / 2.8 -> 5.3 / 2.3 — and a node visit costs 2.5 triangle tests.
 * KNOWN LIMIT: the exact, flat box of a piece of an axis-aligned 100 m wall culls the self-hit of a bounce ray whose exact
 * crossing lies below tnear while the triangle test's rounding reports t just above it (a few paths per million differ from
 * the brute-force definition, tests/test_gpu_parity.py); the unsplit wall sits in a coarsely quantised slot and is found. */
#ifndef RT_SPLIT_TAU_LOG2
#define RT_SPLIT_TAU_LOG2 2
#endif
#define RT_SPLIT_MAX_DEPTH 14

RT_HD void rt_load_wtri(const RtBuild &b, uint32_t gid, f3 v[3]) {
    const rt_float4 a = b.wtris[(size_t)gid * 3], c = b.wtris[(size_t)gid * 3 + 1], d = b.wtris[(size_t)gid * 3 + 2];
    v[0] = mk3(a.x, a.y, a.z);
    v[1] = mk3(c.x, c.y, c.z);
    v[2] = mk3(d.x, d.y, d.z);
}
RT_HD float rt_len2(f3 a, f3 c) {
    const f3 d = a - c;
    return (d.x * d.x + d.y * d.y) + d.z * d.z;
}

/* visits the pieces of triangle gid for edge-length threshold len2 (squared): emit(lo, hi) per piece; returns their
 * number (1 = not split; then the one piece is the triangle's own box, unpadded) */
template <class Emit>
RT_HD uint32_t rt_split_visit(const RtBuild &b, uint32_t gid, float len2, Emit emit) {
    f3 v[3];
    rt_load_wtri(b, gid, v);
    const f3 tlo = mk3(rt_min3(v[0].x, v[1].x, v[2].x), rt_min3(v[0].y, v[1].y, v[2].y), rt_min3(v[0].z, v[1].z, v[2].z));
    const f3 thi = mk3(rt_max3(v[0].x, v[1].x, v[2].x), rt_max3(v[0].y, v[1].y, v[2].y), rt_max3(v[0].z, v[1].z, v[2].z));
    const float e0 = rt_len2(v[0], v[1]), e1 = rt_len2(v[1], v[2]), e2 = rt_len2(v[2], v[0]);
    if (!(len2 > 0.0f) || !(rt_max3(e0, e1, e2) > len2) || !(rt_max3(e0, e1, e2) < 3.0e38f)) { /* small, or not finite */
        emit(tlo, thi);
        return 1u;
    }
    const float pad = rt_max(rt_max3(fabsf(tlo.x), fabsf(tlo.y), fabsf(tlo.z)), rt_max3(fabsf(thi.x), fabsf(thi.y), fabsf(thi.z))) * 9.5367432e-7f;
    f3 st[RT_SPLIT_MAX_DEPTH + 1][3];
    int depth[RT_SPLIT_MAX_DEPTH + 1];
    int sp = 0;
    st[0][0] = v[0]; st[0][1] = v[1]; st[0][2] = v[2];
    depth[0] = 0;
    sp = 1;
    uint32_t count = 0;
    while (sp > 0) {
        sp--;
        const f3 a = st[sp][0], c = st[sp][1], d = st[sp][2];
        const int dep = depth[sp];
        const float l0 = rt_len2(a, c), l1 = rt_len2(c, d), l2 = rt_len2(d, a);
        const float lm = rt_max3(l0, l1, l2);
        if (lm > len2 && dep < RT_SPLIT_MAX_DEPTH) { /* bisect the longest edge (the first one on ties) */
            f3 p, q, o; /* edge p-q, opposite vertex o */
            if (l0 >= l1 && l0 >= l2) { p = a; q = c; o = d; }
            else if (l1 >= l2) { p = c; q = d; o = a; }
            else { p = d; q = a; o = c; }
            const f3 m = mk3(0.5f * (p.x + q.x), 0.5f * (p.y + q.y), 0.5f * (p.z + q.z));
            st[sp][0] = p; st[sp][1] = m; st[sp][2] = o; depth[sp] = dep + 1;
            sp++;
            st[sp][0] = m; st[sp][1] = q; st[sp][2] = o; depth[sp] = dep + 1;
            sp++;
            continue;
        }
        f3 lo = mk3(rt_min3(a.x, c.x, d.x) - pad, rt_min3(a.y, c.y, d.y) - pad, rt_min3(a.z, c.z, d.z) - pad);
        f3 hi = mk3(rt_max3(a.x, c.x, d.x) + pad, rt_max3(a.y, c.y, d.y) + pad, rt_max3(a.z, c.z, d.z) + pad);
        lo = mk3(rt_max(lo.x, tlo.x), rt_max(lo.y, tlo.y), rt_max(lo.z, tlo.z)); /* the triangle lies inside its own box */
        hi = mk3(rt_min(hi.x, thi.x), rt_min(hi.y, thi.y), rt_min(hi.z, thi.z));
        emit(lo, hi);
        count++;
    }
    return count;
}

struct RtSplitCount {
    RT_HD void operator()(f3, f3) const {}
};
RT_HD uint32_t rt_split_count(const RtBuild &b, uint32_t gid, float len2) { return rt_split_visit(b, gid, len2, RtSplitCount()); }

/* writes the references of triangle gid from slot `first` on; returns the bounds of their centres in clo / chi */
struct RtSplitEmit {
    uint32_t *ref_tri;
    rt_float4 *ref_lo, *ref_hi;
    uint32_t gid, next;
    f3 clo, chi;
    RT_HD void operator()(f3 lo, f3 hi) {
        ref_tri[next] = gid;
        ref_lo[next] = rt_mk_float4(lo.x, lo.y, lo.z, 0.0f);
        ref_hi[next] = rt_mk_float4(hi.x, hi.y, hi.z, 0.0f);
        next++;
        const f3 c = mk3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
        clo = mk3(rt_min(clo.x, c.x), rt_min(clo.y, c.y), rt_min(clo.z, c.z));
        chi = mk3(rt_max(chi.x, c.x), rt_max(chi.y, c.y), rt_max(chi.z, c.z));
    }
};
struct RtSplitEmitRef { /* by reference, so that the visitor's state survives the by-value template argument */
    RtSplitEmit *e;
    RT_HD void operator()(f3 lo, f3 hi) const { (*e)(lo, hi); }
};
RT_HD void rt_split_emit(const RtBuild &b, uint32_t gid, float len2, uint32_t first, uint32_t *ref_tri, rt_float4 *ref_lo, rt_float4 *ref_hi,
                         f3 &clo, f3 &chi) {
    RtSplitEmit e;
    e.ref_tri = ref_tri;
    e.ref_lo = ref_lo;
    e.ref_hi = ref_hi;
    e.gid = gid;
    e.next = first;
    e.clo = mk3(INFINITY, INFINITY, INFINITY);
    e.chi = mk3(-INFINITY, -INFINITY, -INFINITY);
    RtSplitEmitRef r;
    r.e = &e;
    rt_split_visit(b, gid, len2, r);
    clo = e.clo;
    chi = e.chi;
}
/* edge length (squared) from which triangles are split, for doubling step `shift`: 0 when the scene has no extent */
RT_HD float rt_split_len2(const RtBuild &b, int shift) {
    float d2 = 0.0f;
    for (int a = 0; a < 3; a++) {
        const float e = rt_ordered_to_float(b.cen_bounds[3 + a]) - rt_ordered_to_float(b.cen_bounds[a]);
        if (e > 0.0f && e < 3.0e38f) d2 += e * e;
    }
    const float k = rt_u2f((uint32_t)(127 - 2 * RT_SPLIT_TAU_LOG2 + 2 * shift) << 23); /* 4^(shift - tau) */
    return d2 * k;
}

/* leaf item (reference) -> triangle id / box */
RT_HD uint32_t rt_item_tri(const RtBuild &b, uint32_t item) { return b.ref_tri ? b.ref_tri[item] : item; }
RT_HD void rt_item_box(const RtBuild &b, uint32_t item, f3 &lo, f3 &hi) {
    if (b.ref_tri) {
        const rt_float4 l = b.ref_lo[item], h = b.ref_hi[item];
        lo = mk3(l.x, l.y, l.z);
        hi = mk3(h.x, h.y, h.z);
    } else {
        rt_tri_box(b, item, lo, hi);
    }
}

/* ---- morton ------------------------------------------------------------------------------ */
/* 63-bit Morton code with an ADAPTIVE axis order (after Vinkler, Bittner, Havran, "Extended Morton
 * Codes for High Performance Bounding Volume Hierarchy Construction", HPG 2017): every bit splits the
 * axis along which the current cell is longest, instead of cycling x, y, z. For the flat,
 * anisotropic scenes of the benchmark (height fields: x, z extent >> y) the plain interleave wastes a
 * third of the bits on an axis that does not separate anything. The axis sequence depends only on the
 * scene's centroid bounds, so every triangle derives the same sequence. */
/* SIZE CLASS: the top RT_SIZE_CLASS_BITS bits of the key hold a size class, so that items whose box is large against the
 * scene (longest side > RT_SIZE_CLASS_T0 of the largest extent of the centroid bounds, next class RT_SIZE_CLASS_STEP times
 * that) sort into their OWN subtree right under the root instead of inflating the boxes of a subtree of small neighbours —
 * the "teapot in a stadium" case that SAH builders (Embree, src/scene.cpp:406-439) handle by construction. One bit at 1/16:
 * node visits per ray 7.36 -> 6.09 on the stadium scene (a SAH-binned tree: 5.17), 3.77 -> 3.35 on Cornell (SAH 3.55),
 * unchanged on C3 / C4 where no triangle is that large (tools/tree_quality.py, profiles/README.md). */
#ifndef RT_SIZE_CLASS_BITS
#define RT_SIZE_CLASS_BITS 1
#endif
#ifndef RT_SIZE_CLASS_T0
#define RT_SIZE_CLASS_T0 0.0625f
#endif
#ifndef RT_SIZE_CLASS_STEP
#define RT_SIZE_CLASS_STEP 4.0f
#endif
RT_HD void rt_morton_tri(const RtBuild &b, uint32_t gid /* leaf item */) {
    f3 lo, hi;
    rt_item_box(b, gid, lo, hi);
    const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    float ext[3];
    uint32_t u[3];
    int pos[3] = {31, 31, 31};
    for (int a = 0; a < 3; a++) {
        const float mn = rt_ordered_to_float(b.cen_bounds[a]), mx = rt_ordered_to_float(b.cen_bounds[3 + a]);
        ext[a] = mx - mn;
        double f = ext[a] > 0.0f ? ((double)c[a] - (double)mn) / (double)ext[a] : 0.0;
        f = f < 0.0 ? 0.0 : (f > 0.99999999 ? 0.99999999 : f);
        u[a] = (uint32_t)(f * 4294967296.0);
    }
    uint64_t code = 0;
    int cls_bits = RT_SIZE_CLASS_BITS;
    float big = rt_max3(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z), scene = rt_max3(ext[0], ext[1], ext[2]);
    uint32_t cls = 0;
    if (cls_bits > 0 && scene > 0.0f) {
        float thr = scene * RT_SIZE_CLASS_T0;
        const uint32_t top = (1u << cls_bits) - 1u;
        while (cls < top && big > thr) {
            cls++;
            thr *= RT_SIZE_CLASS_STEP;
        }
    }
    for (int bit = 0; bit < 63 - cls_bits; bit++) {
        int a = 0;
        if (ext[1] > ext[a]) a = 1;
        if (ext[2] > ext[a]) a = 2;
        code = (code << 1) | (uint64_t)((u[a] >> pos[a]) & 1u);
        ext[a] *= 0.5f;
        if (--pos[a] < 0) ext[a] = -1.0f; /* 32 bits of this axis used up */
    }
    code |= (uint64_t)cls << (63 - cls_bits);
    b.keys[gid] = code;
    b.vals[gid] = gid;
}

/* ---- Karras 2012 radix tree ---------------------------------------------------------------- */
RT_HD int rt_delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], c = keys[j];
    if (a == c) return 64 + rt_clz32((uint32_t)i ^ (uint32_t)j);
    return rt_clz64(a ^ c);
}

RT_HD void rt_karras_node(const RtBuild &b, uint32_t iu) {
    const int n = (int)b.n_items, i = (int)iu;
    const uint64_t *keys = b.keys;
    const int d = (rt_delta(keys, n, i, i + 1) - rt_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = rt_delta(keys, n, i, i - d);
    int lmax = 2;
    while (rt_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (rt_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = rt_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (rt_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int first = i < j ? i : j, last = i < j ? j : i;
    const uint32_t leaf0 = (uint32_t)(n - 1);
    const uint32_t lc = (first == gamma) ? leaf0 + (uint32_t)gamma : (uint32_t)gamma;
    const uint32_t rc = (last == gamma + 1) ? leaf0 + (uint32_t)(gamma + 1) : (uint32_t)(gamma + 1);
    b.left[i] = lc;
    b.right[i] = rc;
    b.parent[lc] = (uint32_t)i;
    b.parent[rc] = (uint32_t)i;
    b.range_first[i] = (uint32_t)first;
    b.range_last[i] = (uint32_t)last;
    if (i == 0) b.parent[0] = RT_MISS;
}

/* The fit reads boxes written by other SMs during the same launch: go through L2 (.cg), the
 * per-SM L1 is not coherent and may hold a stale line that a neighbouring box pulled in. */
RT_HD rt_float4 rt_ld_cg(const rt_float4 *p) {
#if RT_DEVICE_CODE
    return __ldcg(p);
#else
    return *p;
#endif
}
RT_HD void rt_st_cg(rt_float4 *p, rt_float4 v) {
#if RT_DEVICE_CODE
    __stcg(p, v);
#else
    *p = v;
#endif
}

/* ---- SAH-optimal collapse: dynamic program over the binary tree ---------------------------------
 * C(n, i) = cheapest way to represent the subtree of binary node n by at most i wide-tree children:
 *   C(n,1) = min( A_n * P_n * c_prim            if P_n <= RT_LEAF_MAX   (one leaf child),
 *                 A_n * c_node + D(n,8) )                               (one inner wide node)
 *   C(n,i) = min( D(n,i), C(n,i-1) ),  D(n,j) = min_{0<k<j} C(left,k) + C(right,j-k)
 * Decision bytes per node: d[0] = 1 when C(n,1) is the leaf; d[1] = k of D(n,8);
 * d[i] (i = 2..7) = k of D(n,i), or 0 when C(n,i) = C(n,i-1). */
#ifndef RT_SAH_C_NODE
#define RT_SAH_C_NODE 1.0f
#endif
#ifndef RT_SAH_C_PRIM
#define RT_SAH_C_PRIM 0.6f
#endif

RT_HD float rt_half_area(rt_float4 lo, rt_float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}
RT_HD float rt_dp_get(rt_float4 a, rt_float4 c, int i) { /* i = 1..7 */
    return i == 1 ? a.x : i == 2 ? a.y : i == 3 ? a.z : i == 4 ? a.w : i == 5 ? c.x : i == 6 ? c.y : c.z;
}

RT_HD void rt_dp_leaf(const RtBuild &b, uint32_t id, rt_float4 lo, rt_float4 hi) {
    const float c = rt_half_area(lo, hi) * RT_SAH_C_PRIM;
    rt_st_cg(&b.dp_cost[(size_t)id * 2], rt_mk_float4(c, c, c, c));
    rt_st_cg(&b.dp_cost[(size_t)id * 2 + 1], rt_mk_float4(c, c, c, c));
    b.dp_dec[(size_t)id * 2] = 1u; /* d[0] = leaf */
    b.dp_dec[(size_t)id * 2 + 1] = 0u;
}

RT_HD void rt_dp_node(const RtBuild &b, uint32_t node, uint32_t lc, uint32_t rc, rt_float4 lo, rt_float4 hi) {
    const rt_float4 l0 = rt_ld_cg(&b.dp_cost[(size_t)lc * 2]), l1 = rt_ld_cg(&b.dp_cost[(size_t)lc * 2 + 1]);
    const rt_float4 r0 = rt_ld_cg(&b.dp_cost[(size_t)rc * 2]), r1 = rt_ld_cg(&b.dp_cost[(size_t)rc * 2 + 1]);
    float D[9];
    uint32_t K[9];
    for (int j = 2; j <= 8; j++) {
        float best = 3.0e38f;
        uint32_t bk = 1;
        for (int k = 1; k < j; k++) {
            if (k > 7 || j - k > 7) continue;
            const float c = rt_dp_get(l0, l1, k) + rt_dp_get(r0, r1, j - k);
            if (c < best) {
                best = c;
                bk = (uint32_t)k;
            }
        }
        D[j] = best;
        K[j] = bk;
    }
    const float area = rt_half_area(lo, hi);
    const uint32_t count = b.range_last[node] - b.range_first[node] + 1u;
    const float c_inner = area * RT_SAH_C_NODE + D[8];
    const float c_leaf = count <= RT_LEAF_MAX ? area * (float)count * RT_SAH_C_PRIM : 3.0e38f;
    float C[8];
    uint32_t d[8];
    d[0] = c_leaf <= c_inner ? 1u : 0u;
    d[1] = K[8];
    C[1] = d[0] ? c_leaf : c_inner;
    for (int i = 2; i <= 7; i++) {
        if (D[i] < C[i - 1]) {
            C[i] = D[i];
            d[i] = K[i];
        } else {
            C[i] = C[i - 1];
            d[i] = 0u;
        }
    }
    rt_st_cg(&b.dp_cost[(size_t)node * 2], rt_mk_float4(C[1], C[2], C[3], C[4]));
    rt_st_cg(&b.dp_cost[(size_t)node * 2 + 1], rt_mk_float4(C[5], C[6], C[7], 0.0f));
    b.dp_dec[(size_t)node * 2] = d[0] | (d[1] << 8) | (d[2] << 16) | (d[3] << 24);
    b.dp_dec[(size_t)node * 2 + 1] = d[4] | (d[5] << 8) | (d[6] << 16) | (d[7] << 24);
}
RT_HD uint32_t rt_dp_decision(const RtBuild &b, uint32_t id, int i) { /* byte i of the node's decisions */
    return (b.dp_dec[(size_t)id * 2 + (i >> 2)] >> ((i & 3) * 8)) & 0xffu;
}

/* ---- bottom-up fit + DP: called once per leaf; `arrive` returns the previous flag value -------- */
template <class Arrive>
RT_HD void rt_fit_leaf(const RtBuild &b, uint32_t j, Arrive arrive) {
    const uint32_t leaf0 = b.n_items - 1;
    f3 lo, hi;
    rt_item_box(b, b.vals[j], lo, hi);
    rt_float4 l4, h4;
    l4.x = lo.x; l4.y = lo.y; l4.z = lo.z; l4.w = 0.0f;
    h4.x = hi.x; h4.y = hi.y; h4.z = hi.z; h4.w = 0.0f;
    rt_st_cg(&b.box_lo[leaf0 + j], l4);
    rt_st_cg(&b.box_hi[leaf0 + j], h4);
    rt_dp_leaf(b, leaf0 + j, l4, h4);
    if (b.n_items == 1) return;
    uint32_t node = b.parent[leaf0 + j];
    while (node != RT_MISS) {
        if (arrive(&b.flags[node]) == 0) return; /* first child to arrive: the sibling finishes */
        const uint32_t lc = b.left[node], rc = b.right[node];
        const rt_float4 la = rt_ld_cg(&b.box_lo[lc]), lb = rt_ld_cg(&b.box_lo[rc]);
        const rt_float4 ha = rt_ld_cg(&b.box_hi[lc]), hb = rt_ld_cg(&b.box_hi[rc]);
        l4.x = rt_min(la.x, lb.x); l4.y = rt_min(la.y, lb.y); l4.z = rt_min(la.z, lb.z);
        h4.x = rt_max(ha.x, hb.x); h4.y = rt_max(ha.y, hb.y); h4.z = rt_max(ha.z, hb.z);
        rt_st_cg(&b.box_lo[node], l4);
        rt_st_cg(&b.box_hi[node], h4);
        rt_dp_node(b, node, lc, rc, l4, h4);
        node = b.parent[node];
    }
}

/* ---- wide collapse ------------------------------------------------------------------------- */
RT_HD bool rt_is_leaf_id(const RtBuild &b, uint32_t id) { return id >= b.n_items - 1; }
RT_HD uint32_t rt_subtree_count(const RtBuild &b, uint32_t id) {
    return rt_is_leaf_id(b, id) ? 1u : b.range_last[id] - b.range_first[id] + 1u;
}
RT_HD uint32_t rt_subtree_first(const RtBuild &b, uint32_t id) {
    return rt_is_leaf_id(b, id) ? id - (b.n_items - 1) : b.range_first[id];
}
RT_HD float rt_box_area(const RtBuild &b, uint32_t id) {
    const rt_float4 lo = b.box_lo[id], hi = b.box_hi[id];
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

/* a child of a wide node is a leaf when it is a single triangle or the DP chose "leaf" for it */
RT_HD bool rt_child_is_leaf(const RtBuild &b, uint32_t id) {
    return rt_is_leaf_id(b, id) || rt_dp_decision(b, id, 0) != 0u;
}

/* children of the wide node rooted at binary node items[item]: follow the DP decisions
 * (roots(n, i) = the <= i subtrees that represent n) */
RT_HD void rt_wide_select(const RtBuild &b, uint32_t item) {
    const uint32_t root = b.items[item];
    uint32_t c[8];
    int n = 0;
    if (rt_is_leaf_id(b, root) || rt_subtree_count(b, root) <= 1u) {
        c[n++] = root;
    } else {
        uint32_t st_id[16];
        int st_i[16], sp = 0;
        const int k8 = (int)rt_dp_decision(b, root, 1);
        st_id[sp] = b.right[root]; st_i[sp++] = 8 - k8;
        st_id[sp] = b.left[root]; st_i[sp++] = k8;
        while (sp > 0) {
            sp--;
            const uint32_t m = st_id[sp];
            int i = st_i[sp];
            if (rt_is_leaf_id(b, m)) {
                c[n++] = m;
                continue;
            }
            if (i > 7) i = 7;
            while (i > 1 && rt_dp_decision(b, m, i) == 0u) i--; /* C(m,i) = C(m,i-1) */
            if (i <= 1) {
                c[n++] = m;
                continue;
            }
            const int k = (int)rt_dp_decision(b, m, i);
            st_id[sp] = b.right[m]; st_i[sp++] = i - k;
            st_id[sp] = b.left[m]; st_i[sp++] = k;
        }
    }
    uint32_t n_inner = 0, n_leaf_tris = 0;
    for (int k = 0; k < 8; k++) {
        const uint32_t id = k < n ? c[k] : RT_MISS;
        b.sel[(size_t)item * 8 + k] = id;
        if (k < n) {
            if (rt_child_is_leaf(b, id)) n_leaf_tris += rt_subtree_count(b, id);
            else n_inner++;
        }
    }
    b.counts[item] = ((uint64_t)n_inner << 32) | n_leaf_tris;
}

/* pre-gather the shading attributes of global triangle gid (src/trace_ray.hpp:29-41) */
RT_HD void rt_write_shade_record(const RtBuild &b, uint32_t gid, uint32_t slot) {
    const uint32_t inst = rt_find_instance(b, gid);
    const RtInstanceGeom &g = b.geom[inst];
    const uint32_t prim = gid - g.first_tri;
    float nrm[9], uv[6];
    for (int k = 0; k < 3; k++) {
        const uint32_t vi = b.indices[(size_t)g.first_index + (size_t)prim * 3 + k] + g.first_vertex;
        nrm[k * 3] = b.normals[(size_t)vi * 3];
        nrm[k * 3 + 1] = b.normals[(size_t)vi * 3 + 1];
        nrm[k * 3 + 2] = b.normals[(size_t)vi * 3 + 2];
        uv[k * 2] = b.uvs[(size_t)vi * 2];
        uv[k * 2 + 1] = b.uvs[(size_t)vi * 2 + 1];
    }
    rt_float4 s0, s1, s2, s3;
    s0.x = nrm[0]; s0.y = nrm[1]; s0.z = nrm[2]; s0.w = nrm[3];
    s1.x = nrm[4]; s1.y = nrm[5]; s1.z = nrm[6]; s1.w = nrm[7];
    s2.x = nrm[8]; s2.y = uv[0]; s2.z = uv[1]; s2.w = uv[2];
    s3.x = uv[3]; s3.y = uv[4]; s3.z = uv[5]; s3.w = rt_u2f(inst);
    b.shade[(size_t)slot * 4] = s0;
    b.shade[(size_t)slot * 4 + 1] = s1;
    b.shade[(size_t)slot * 4 + 2] = s2;
    b.shade[(size_t)slot * 4 + 3] = s3;
}

/* smallest biased exponent e with p + 255 * 2^(e-127) >= hi */
RT_HD uint32_t rt_quant_exponent(float lo, float hi) {
    const float ext = hi - lo;
    uint32_t e = 1;
    if (ext > 0.0f) {
        const uint32_t bits = rt_f2u(ext / 255.0f);
        e = (bits >> 23) & 0xffu;
        if (bits & 0x7fffffu) e++;
        if (e < 1) e = 1;
    }
    while (e < 254 && lo + 255.0f * rt_u2f(e << 23) < hi) e++;
    return e;
}
RT_HD uint32_t rt_quant_lo(float p, float scale, float v) {
    float q = floorf((v - p) / scale);
    q = rt_min(rt_max(q, 0.0f), 255.0f);
    while (q > 0.0f && p + q * scale > v) q -= 1.0f;
    return (uint32_t)q;
}
RT_HD uint32_t rt_quant_hi(float p, float scale, float v) {
    float q = ceilf((v - p) / scale);
    q = rt_min(rt_max(q, 0.0f), 255.0f);
    while (q < 255.0f && p + q * scale < v) q += 1.0f;
    return (uint32_t)q;
}

/* write wide node `level_first_node + item` and everything it owns */
RT_HD void rt_wide_emit(const RtBuild &b, uint32_t item) {
    const uint32_t root = b.items[item];
    const uint64_t off = b.offsets[item];
    const uint32_t child_base = b.next_level_first_node + (uint32_t)(off >> 32);
    const uint32_t tri_base = b.tri_cursor + (uint32_t)(off & 0xffffffffu);

    uint32_t c[8];
    int n = 0;
    for (int k = 0; k < 8; k++) {
        const uint32_t id = b.sel[(size_t)item * 8 + k];
        if (id != RT_MISS) c[n++] = id;
    }
    const rt_float4 plo = b.box_lo[root], phi = b.box_hi[root];
    const f3 pc = mk3(0.5f * (plo.x + phi.x), 0.5f * (plo.y + phi.y), 0.5f * (plo.z + phi.z));

    /* octant-ordered slot assignment: slot s is visited first by rays whose direction signs are
     * D_s = (s&4 ? - : +, s&2 ? - : +, s&1 ? - : +); give it the child that is nearest along D_s.
     * Greedy minimum over the 8x8 cost table. */
    float cost[8][8];
    for (int k = 0; k < n; k++) {
        const rt_float4 lo = b.box_lo[c[k]], hi = b.box_hi[c[k]];
        const f3 cc = mk3(0.5f * (lo.x + hi.x) - pc.x, 0.5f * (lo.y + hi.y) - pc.y, 0.5f * (lo.z + hi.z) - pc.z);
        for (int s = 0; s < 8; s++)
            cost[k][s] = ((s & 4) ? -cc.x : cc.x) + ((s & 2) ? -cc.y : cc.y) + ((s & 1) ? -cc.z : cc.z);
    }
    int slot_child[8];
    for (int s = 0; s < 8; s++) slot_child[s] = -1;
    uint32_t child_done = 0, slot_done = 0;
    for (int it = 0; it < n; it++) {
        int bk = -1, bs = -1;
        float bc = 3.0e38f;
        for (int k = 0; k < n; k++) {
            if (child_done & (1u << k)) continue;
            for (int s = 0; s < 8; s++) {
                if (slot_done & (1u << s)) continue;
                if (cost[k][s] < bc) {
                    bc = cost[k][s];
                    bk = k;
                    bs = s;
                }
            }
        }
        if (bk < 0) { /* NaN boxes: fall back to first free pair */
            for (int k = 0; k < n && bk < 0; k++)
                if (!(child_done & (1u << k))) bk = k;
            for (int s = 0; s < 8 && bs < 0; s++)
                if (!(slot_done & (1u << s))) bs = s;
        }
        slot_child[bs] = bk;
        child_done |= 1u << bk;
        slot_done |= 1u << bs;
    }

    const uint32_t ex = rt_quant_exponent(plo.x, phi.x), ey = rt_quant_exponent(plo.y, phi.y),
                   ez = rt_quant_exponent(plo.z, phi.z);
    const float sx = rt_u2f(ex << 23), sy = rt_u2f(ey << 23), sz = rt_u2f(ez << 23);

    uint32_t imask = 0, tmask = 0;
    uint32_t q[6][2]; /* qlo x,y,z, qhi x,y,z ; two words of four bytes */
    for (int a = 0; a < 6; a++) {
        q[a][0] = 0;
        q[a][1] = 0;
    }
    uint32_t inner_rank = 0, tri_off = 0;
    for (int s = 0; s < 8; s++) {
        const int w = s >> 2, sh = (s & 3) * 8;
        const int k = slot_child[s];
        if (k < 0) { /* empty slot: inverted box never hits */
            q[0][w] |= 255u << sh;
            q[1][w] |= 255u << sh;
            q[2][w] |= 255u << sh;
            continue;
        }
        const uint32_t id = c[k];
        const rt_float4 lo = b.box_lo[id], hi = b.box_hi[id];
        q[0][w] |= rt_quant_lo(plo.x, sx, lo.x) << sh;
        q[1][w] |= rt_quant_lo(plo.y, sy, lo.y) << sh;
        q[2][w] |= rt_quant_lo(plo.z, sz, lo.z) << sh;
        q[3][w] |= rt_quant_hi(plo.x, sx, hi.x) << sh;
        q[4][w] |= rt_quant_hi(plo.y, sy, hi.y) << sh;
        q[5][w] |= rt_quant_hi(plo.z, sz, hi.z) << sh;
        const uint32_t cnt = rt_subtree_count(b, id);
        if (!rt_child_is_leaf(b, id)) {
            imask |= 1u << s;
            b.next_items[(child_base - b.next_level_first_node) + inner_rank] = id;
            inner_rank++;
        } else {
            tmask |= ((1u << cnt) - 1u) << (3 * s); /* unary count; triangles follow in slot order */
            const uint32_t first = rt_subtree_first(b, id);
            for (uint32_t t = 0; t < cnt; t++) {
                const uint32_t gid = rt_item_tri(b, b.vals[first + t]);
                const uint32_t slot = tri_base + tri_off + t;
                b.tris[(size_t)slot * RT_TRI_VEC4] = b.wtris[(size_t)gid * 3];
                b.tris[(size_t)slot * RT_TRI_VEC4 + 1] = b.wtris[(size_t)gid * 3 + 1];
                b.tris[(size_t)slot * RT_TRI_VEC4 + 2] = b.wtris[(size_t)gid * 3 + 2];
#if RT_TRI_VEC4 > 3
                b.tris[(size_t)slot * RT_TRI_VEC4 + 3] = rt_mk_float4(0.0f, 0.0f, 0.0f, 0.0f);
#endif
                rt_write_shade_record(b, gid, slot);
            }
            tri_off += cnt;
        }
    }
    rt_uint4 n0, n1, n2, n3, n4;
    n0.x = rt_f2u(plo.x); n0.y = rt_f2u(plo.y); n0.z = rt_f2u(plo.z);
    n0.w = ex | (ey << 8) | (ez << 16) | (imask << 24);
    n1.x = child_base; n1.y = tri_base; n1.z = tmask; n1.w = 0u;
    n2.x = q[0][0]; n2.y = q[0][1]; n2.z = q[1][0]; n2.w = q[1][1];
    n3.x = q[2][0]; n3.y = q[2][1]; n3.z = q[3][0]; n3.w = q[3][1];
    n4.x = q[4][0]; n4.y = q[4][1]; n4.z = q[5][0]; n4.w = q[5][1];
    rt_uint4 *np = b.nodes + (size_t)(b.level_first_node + item) * RT_NODE_VEC4;
    np[0] = n0; np[1] = n1; np[2] = n2; np[3] = n3; np[4] = n4;
#if RT_NODE_VEC4 > 5
    np[5] = n0; /* spare */
#endif
}

#endif /* RT_BUILD_H */
