/*
 * hostemu.cpp — TEST-ONLY host emulation of the CUDA kernels' per-item source.
 *
 * NOT a CPU fallback and not part of the product: librt_b200.so never loads or links this, and
 * the package fails loudly without a GPU. The container that builds this repo has no GPU, so this
 * harness compiles the SAME headers the kernels are made of (csrc/rt_build.h, rt_traverse.h,
 * rt_shade.h, rt_wavefront.h) with plain g++ and drives them with sequential loops in place of
 * the launch grid (std::stable_sort for the radix sort, a running sum for the scan). It lets the
 * `-m "not gpu"` suite check the kernel logic (tree build, compressed-node traversal, shading,
 * wavefront queueing) against the oracle before any GPU time is spent.
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

/* traversal statistics (nodes visited / triangles tested), host emulation only */
static unsigned long long g_count_node = 0, g_count_tri = 0;
#define RT_COUNTERS 1
#define RT_COUNT_NODE() (g_count_node++)
#define RT_COUNT_TRI() (g_count_tri++)

#include "../../include/rt_api.h"
#include "../../sycl-ray-tracer_b200/csrc/rt_build.h"
#include "../../sycl-ray-tracer_b200/csrc/rt_wavefront.h"
#include "../../sycl-ray-tracer_b200/csrc/rt_blocks.h"

namespace {

void normal_matrix(const float T[16], float out[9]) {
    float m[3][3];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) m[c][r] = T[c * 4 + r];
    const float ood = 1.0f / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) -
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
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) out[c * 3 + r] = inv[r][c];
}

struct HostArrive {
    uint32_t operator()(uint32_t *flag) const { return (*flag)++; }
};

} // namespace

struct emu_scene {
    std::vector<RtInstance> inst;
    std::vector<uint8_t> tex;
    std::vector<rt_uint4> nodes;
    std::vector<rt_float4> tris, shade;
    uint32_t n_tris = 0, n_items = 0, n_nodes = 0, depth = 0;
    RtScene view;
};

extern "C" {

emu_scene *emu_scene_create(const rt_scene_desc *desc) {
    emu_scene *s = new emu_scene();
    std::vector<float> pos, nrm, uv;
    std::vector<uint32_t> idx;
    std::vector<RtInstanceGeom> geom(desc->instance_count ? desc->instance_count : 1);
    s->inst.resize(desc->instance_count ? desc->instance_count : 1);
    uint32_t v0 = 0, i0 = 0;
    for (uint32_t i = 0; i < desc->instance_count; i++) {
        const rt_instance &in = desc->instances[i];
        pos.insert(pos.end(), in.positions, in.positions + (size_t)in.vertex_count * 3);
        nrm.insert(nrm.end(), in.normals, in.normals + (size_t)in.vertex_count * 3);
        uv.insert(uv.end(), in.uvs, in.uvs + (size_t)in.vertex_count * 2);
        idx.insert(idx.end(), in.indices, in.indices + in.index_count);
        RtInstanceGeom &g = geom[i];
        memcpy(g.transform, in.transform, sizeof(g.transform));
        g.first_vertex = v0;
        g.first_index = i0;
        g.first_tri = i0 / 3;
        g.tri_count = in.index_count / 3;
        RtInstance &m = s->inst[i];
        normal_matrix(in.transform, m.nmat);
        m.type = in.material.type;
        m.albedo_image = in.material.albedo_image;
        memcpy(m.albedo, in.material.albedo_color, sizeof(m.albedo));
        m.roughness = in.material.roughness;
        m.ior = in.material.ior;
        memcpy(m.emissive, in.material.emissive, sizeof(m.emissive));
        m.first_tri = g.first_tri;
        v0 += in.vertex_count;
        i0 += in.index_count;
    }
    uint32_t n = i0 / 3;
    s->n_tris = n;
    if (desc->texture_layer_count)
        s->tex.assign(desc->texture_layers,
                      desc->texture_layers + (size_t)desc->texture_layer_count * RT_TEX_SIZE * RT_TEX_SIZE * 4);

    if (n == 0) {
        const uint32_t ff = 0xffffffffu;
        s->nodes = {{0, 0, 0, 1u | (1u << 8) | (1u << 16)}, {0, 0, 0, 0}, {ff, ff, ff, ff}, {ff, ff, 0, 0}, {0, 0, 0, 0}};
        s->nodes.resize(RT_NODE_VEC4);
        s->tris.resize(RT_TRI_VEC4);
        s->shade.resize(4);
        s->n_nodes = 1;
        s->depth = 1;
    } else {
        RtBuild b = {};
        b.n_tris = n;
        b.n_inst = desc->instance_count;
        b.positions = pos.data();
        b.normals = nrm.data();
        b.uvs = uv.data();
        b.indices = idx.data();
        b.geom = geom.data();
        std::vector<rt_float4> wtris((size_t)n * 3);
        std::vector<int32_t> cb(8);
        b.wtris = wtris.data();
        b.cen_bounds = cb.data();
        auto reset_bounds = [&]() {
            for (int k = 0; k < 3; k++) {
                cb[k] = rt_float_to_ordered(INFINITY);
                cb[3 + k] = rt_float_to_ordered(-INFINITY);
            }
        };
        auto grow_bounds = [&](f3 lo, f3 hi) {
            const float c0[3] = {lo.x, lo.y, lo.z}, c1[3] = {hi.x, hi.y, hi.z};
            for (int k = 0; k < 3; k++) {
                cb[k] = std::min(cb[k], rt_float_to_ordered(c0[k]));
                cb[3 + k] = std::max(cb[3 + k], rt_float_to_ordered(c1[k]));
            }
        };
        reset_bounds();
        for (uint32_t g = 0; g < n; g++) {
            f3 lo, hi;
            rt_flatten_tri(b, g, lo, hi);
            const f3 c = mk3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
            grow_bounds(c, c);
        }
        /* references for large triangles: the same steps as build_bvh (csrc/bvh_build.cu, 1b) */
        std::vector<uint32_t> ref_tri;
        std::vector<rt_float4> ref_lo, ref_hi;
        uint32_t m = n;
        if (getenv("RT_SPLIT") && atoi(getenv("RT_SPLIT")) > 0 && n > 1) {
            std::vector<uint32_t> split_counts(n);
            const unsigned long long budget = std::max<unsigned long long>(n / 2, 65536ull);
            unsigned long long extra = 0;
            float len2 = 0.0f;
            for (int shift = 0; shift <= 8; shift++) {
                len2 = rt_split_len2(b, shift);
                extra = 0;
                for (uint32_t g = 0; g < n; g++) {
                    split_counts[g] = rt_split_count(b, g, len2);
                    extra += split_counts[g] - 1u;
                }
                if (extra <= budget) break;
                if (shift == 8) extra = 0;
            }
            if (extra > 0 && (unsigned long long)n + extra < 0x7fffffffull) {
                m = n + (uint32_t)extra;
                ref_tri.resize(m);
                ref_lo.resize(m);
                ref_hi.resize(m);
                reset_bounds();
                uint32_t first = 0;
                for (uint32_t g = 0; g < n; g++) {
                    f3 lo, hi;
                    rt_split_emit(b, g, len2, first, ref_tri.data(), ref_lo.data(), ref_hi.data(), lo, hi);
                    grow_bounds(lo, hi);
                    first += split_counts[g];
                }
                b.ref_tri = ref_tri.data();
                b.ref_lo = ref_lo.data();
                b.ref_hi = ref_hi.data();
            }
        }
        b.n_items = m;
        s->n_items = m;
        const uint32_t n_tris_in = n;
        (void)n_tris_in;
        n = m; /* from here on n = leaves of the tree (references) */
        std::vector<rt_float4> box_lo((size_t)2 * n), box_hi((size_t)2 * n);
        std::vector<uint64_t> keys(n), keys_s(n);
        std::vector<uint32_t> vals(n), vals_s(n), left(n), right(n), parent((size_t)2 * n), rf(n), rl(n), flags(n, 0);
        b.keys = keys.data();
        b.vals = vals.data();
        for (uint32_t g = 0; g < n; g++) rt_morton_tri(b, g);
        std::vector<uint32_t> order(n);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return keys[a] < keys[c]; });
        for (uint32_t k = 0; k < n; k++) {
            keys_s[k] = keys[order[k]];
            vals_s[k] = vals[order[k]];
        }
        b.keys = keys_s.data();
        b.vals = vals_s.data();
        b.left = left.data();
        b.right = right.data();
        b.parent = parent.data();
        b.range_first = rf.data();
        b.range_last = rl.data();
        b.box_lo = box_lo.data();
        b.box_hi = box_hi.data();
        b.flags = flags.data();
        std::vector<rt_float4> dp_cost((size_t)4 * n);
        std::vector<uint32_t> dp_dec((size_t)4 * n);
        b.dp_cost = dp_cost.data();
        b.dp_dec = dp_dec.data();
        if (getenv("EMU_SAH") && n > 1) {
            /* EXPERIMENT ONLY: quality headroom of a SAH binary tree under the same wide collapse */
            std::vector<float> clo((size_t)n * 3), chi((size_t)n * 3);
            for (uint32_t g = 0; g < n; g++) {
                f3 lo, hi;
                rt_item_box(b, g, lo, hi);
                clo[g * 3] = lo.x; clo[g * 3 + 1] = lo.y; clo[g * 3 + 2] = lo.z;
                chi[g * 3] = hi.x; chi[g * 3 + 1] = hi.y; chi[g * 3 + 2] = hi.z;
            }
            std::iota(vals_s.begin(), vals_s.end(), 0u);
            uint32_t next_inner = 0;
            struct Job { uint32_t first, count, parent; int side; };
            std::vector<Job> jobs = {{0, n, RT_MISS, 0}};
            while (!jobs.empty()) {
                Job j = jobs.back();
                jobs.pop_back();
                uint32_t id;
                if (j.count == 1) id = (n - 1) + j.first;
                else id = next_inner++;
                if (j.parent != RT_MISS) { (j.side ? right : left)[j.parent] = id; }
                parent[id] = j.parent;
                if (j.count == 1) continue;
                rf[id] = j.first; rl[id] = j.first + j.count - 1;
                float cl[3] = {1e30f, 1e30f, 1e30f}, ch[3] = {-1e30f, -1e30f, -1e30f};
                for (uint32_t i = j.first; i < j.first + j.count; i++)
                    for (int a = 0; a < 3; a++) {
                        float c = 0.5f * (clo[vals_s[i] * 3 + a] + chi[vals_s[i] * 3 + a]);
                        cl[a] = std::min(cl[a], c); ch[a] = std::max(ch[a], c);
                    }
                int bax = -1, bsp = -1; float bcost = 1e30f; const int NB = 16;
                for (int a = 0; a < 3; a++) {
                    if (!(ch[a] > cl[a])) continue;
                    float bl[NB][3], bh[NB][3]; uint32_t bc[NB];
                    for (int q = 0; q < NB; q++) { bc[q] = 0; for (int k = 0; k < 3; k++) { bl[q][k] = 1e30f; bh[q][k] = -1e30f; } }
                    float sc = NB / (ch[a] - cl[a]);
                    for (uint32_t i = j.first; i < j.first + j.count; i++) {
                        uint32_t g = vals_s[i];
                        int q = std::min(NB - 1, (int)((0.5f * (clo[g * 3 + a] + chi[g * 3 + a]) - cl[a]) * sc));
                        bc[q]++;
                        for (int k = 0; k < 3; k++) { bl[q][k] = std::min(bl[q][k], clo[g * 3 + k]); bh[q][k] = std::max(bh[q][k], chi[g * 3 + k]); }
                    }
                    float ra[NB]; uint32_t rc[NB]; float tl[3] = {1e30f, 1e30f, 1e30f}, th[3] = {-1e30f, -1e30f, -1e30f}; uint32_t cnt = 0;
                    auto area = [](float *l, float *h) { float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2]; return dx * dy + dy * dz + dz * dx; };
                    for (int q = NB - 1; q > 0; q--) { for (int k = 0; k < 3; k++) { tl[k] = std::min(tl[k], bl[q][k]); th[k] = std::max(th[k], bh[q][k]); } cnt += bc[q]; ra[q] = cnt ? area(tl, th) : 0; rc[q] = cnt; }
                    for (int k = 0; k < 3; k++) { tl[k] = 1e30f; th[k] = -1e30f; } cnt = 0;
                    for (int q = 0; q < NB - 1; q++) { for (int k = 0; k < 3; k++) { tl[k] = std::min(tl[k], bl[q][k]); th[k] = std::max(th[k], bh[q][k]); } cnt += bc[q]; if (!cnt || !rc[q + 1]) continue; float c = area(tl, th) * cnt + ra[q + 1] * rc[q + 1]; if (c < bcost) { bcost = c; bax = a; bsp = q; } }
                }
                uint32_t mid = j.first + j.count / 2;
                if (bax >= 0) {
                    float sc = NB / (ch[bax] - cl[bax]);
                    auto m = std::partition(vals_s.begin() + j.first, vals_s.begin() + j.first + j.count, [&](uint32_t g) {
                        return std::min(NB - 1, (int)((0.5f * (clo[g * 3 + bax] + chi[g * 3 + bax]) - cl[bax]) * sc)) <= bsp; });
                    uint32_t mm = (uint32_t)(m - vals_s.begin());
                    if (mm > j.first && mm < j.first + j.count) mid = mm;
                }
                jobs.push_back({mid, j.first + j.count - mid, id, 1});
                jobs.push_back({j.first, mid - j.first, id, 0});
            }
        } else
        for (uint32_t i = 0; i + 1 < n; i++) rt_karras_node(b, i);
        for (uint32_t j = 0; j < n; j++) rt_fit_leaf(b, j, HostArrive());

        const size_t max_items = (size_t)n / 2 + 8;
        std::vector<uint32_t> items_a(max_items), items_b(max_items), sel(max_items * 8);
        std::vector<uint64_t> counts(max_items), offsets(max_items);
        s->nodes.resize((size_t)n * RT_NODE_VEC4);
        s->tris.resize((size_t)n * RT_TRI_VEC4);
        s->shade.resize((size_t)n * 4);
        b.sel = sel.data();
        b.counts = counts.data();
        b.offsets = offsets.data();
        b.nodes = s->nodes.data();
        b.tris = s->tris.data();
        b.shade = s->shade.data();
        items_a[0] = 0;
        uint32_t level_items = 1, level_first = 0, tri_cursor = 0, depth = 0;
        uint32_t *cur = items_a.data(), *nxt = items_b.data();
        while (level_items > 0) {
            b.items = cur;
            b.next_items = nxt;
            b.level_first_node = level_first;
            b.next_level_first_node = level_first + level_items;
            b.tri_cursor = tri_cursor;
            for (uint32_t i = 0; i < level_items; i++) rt_wide_select(b, i);
            uint64_t run = 0;
            for (uint32_t i = 0; i < level_items; i++) {
                offsets[i] = run;
                run += counts[i];
            }
            for (uint32_t i = 0; i < level_items; i++) rt_wide_emit(b, i);
            level_first += level_items;
            level_items = (uint32_t)(run >> 32);
            tri_cursor += (uint32_t)(run & 0xffffffffu);
            depth++;
            std::swap(cur, nxt);
        }
        s->n_nodes = level_first;
        s->depth = depth;
        s->nodes.resize((size_t)s->n_nodes * RT_NODE_VEC4);
        if (tri_cursor != n) fprintf(stderr, "hostemu: triangle count mismatch %u != %u\n", tri_cursor, n);
    }
    s->view.bvh.nodes = s->nodes.data();
    s->view.bvh.tris = s->tris.data();
    s->view.shade = s->shade.data();
    s->view.inst = s->inst.data();
    s->view.tex_raw = s->tex.empty() ? nullptr : s->tex.data();
    s->view.tex = 0;
    s->view.n_layers = desc->texture_layer_count;
    s->view.sky = mk3(desc->sky_color[0], desc->sky_color[1], desc->sky_color[2]);
    return s;
}

void emu_scene_destroy(emu_scene *s) { delete s; }
uint32_t emu_scene_node_count(const emu_scene *s) { return s->n_nodes; }
uint32_t emu_scene_depth(const emu_scene *s) { return s->depth; }

/* structural check of the emitted tree. 0 = ok */
int emu_scene_validate(const emu_scene *s) {
    const uint32_t n = s->n_tris, m = s->n_items; /* triangles, leaf records (references of split triangles repeat the triangle) */
    if (n == 0) return 0;
    std::vector<uint8_t> seen_tri(m, 0), seen_gid(n, 0), seen_node(s->n_nodes, 0);
    std::vector<uint32_t> refs_of(n, 0);
    for (uint32_t i = 0; i < m; i++) {
        const uint32_t gid = rt_f2u(s->tris[(size_t)i * RT_TRI_VEC4 + 2].w);
        if (gid >= n) return 7;
        refs_of[gid]++;
    }
    std::vector<uint32_t> stack = {0};
    seen_node[0] = 1;
    while (!stack.empty()) {
        const uint32_t ni = stack.back();
        stack.pop_back();
        const rt_uint4 *np = &s->nodes[(size_t)ni * RT_NODE_VEC4];
        const float p[3] = {rt_u2f(np[0].x), rt_u2f(np[0].y), rt_u2f(np[0].z)};
        const float sc[3] = {rt_u2f((np[0].w & 0xffu) << 23), rt_u2f(((np[0].w >> 8) & 0xffu) << 23),
                             rt_u2f(((np[0].w >> 16) & 0xffu) << 23)};
        const uint32_t imask = np[0].w >> 24;
        const uint32_t q[6][2] = {{np[2].x, np[2].y}, {np[2].z, np[2].w}, {np[3].x, np[3].y},
                                  {np[3].z, np[3].w}, {np[4].x, np[4].y}, {np[4].z, np[4].w}};
        uint32_t rank = 0;
        for (int slot = 0; slot < 8; slot++) {
            const uint32_t tmask = np[1].z, unary = (tmask >> (3 * slot)) & 7u;
            const bool inner = (imask >> slot) & 1u;
            if (tmask >> 24) return 1;
            if (inner && unary) return 2;
            if (!inner && !unary) continue; /* empty slot */
            float lo[3], hi[3];
            for (int a = 0; a < 3; a++) {
                lo[a] = p[a] + (float)((q[a][slot >> 2] >> ((slot & 3) * 8)) & 0xffu) * sc[a];
                hi[a] = p[a] + (float)((q[3 + a][slot >> 2] >> ((slot & 3) * 8)) & 0xffu) * sc[a];
            }
            if (inner) {
                const uint32_t child = np[1].x + rank++;
                if (child >= s->n_nodes || seen_node[child]) return 3;
                seen_node[child] = 1;
                /* the child's own origin must lie inside the decoded box */
                const rt_uint4 *cp = &s->nodes[(size_t)child * RT_NODE_VEC4];
                const float cpv[3] = {rt_u2f(cp[0].x), rt_u2f(cp[0].y), rt_u2f(cp[0].z)};
                for (int a = 0; a < 3; a++)
                    if (cpv[a] < lo[a] || cpv[a] > hi[a]) return 4;
                stack.push_back(child);
            } else {
                const uint32_t cnt = rt_popc(unary), off = rt_popc(tmask & ((1u << (3 * slot)) - 1u));
                if (cnt < 1 || cnt > 3 || unary != (1u << cnt) - 1u) return 5;
                for (uint32_t t = 0; t < cnt; t++) {
                    const uint32_t tri = np[1].y + off + t;
                    if (tri >= m || seen_tri[tri]) return 6;
                    seen_tri[tri] = 1;
                    const rt_float4 *tp = &s->tris[(size_t)tri * RT_TRI_VEC4];
                    const uint32_t gid = rt_f2u(tp[2].w);
                    if (gid >= n || (seen_gid[gid] && refs_of[gid] == 1)) return 7;
                    seen_gid[gid] = 1;
                    float tl[3] = {INFINITY, INFINITY, INFINITY}, th[3] = {-INFINITY, -INFINITY, -INFINITY};
                    for (int k = 0; k < 3; k++) {
                        const float v[3] = {tp[k].x, tp[k].y, tp[k].z};
                        for (int a = 0; a < 3; a++) {
                            if (refs_of[gid] == 1 && (v[a] < lo[a] || v[a] > hi[a])) return 8; /* an unsplit triangle lies inside its slot's box */
                            tl[a] = std::min(tl[a], v[a]);
                            th[a] = std::max(th[a], v[a]);
                        }
                    }
                    for (int a = 0; a < 3; a++) /* a reference's box is a piece of the triangle's box */
                        if (th[a] < lo[a] || tl[a] > hi[a]) return 8;
                }
            }
        }
    }
    for (uint32_t i = 0; i < m; i++)
        if (!seen_tri[i]) return 9;
    for (uint32_t i = 0; i < n; i++)
        if (!seen_gid[i]) return 9;
    for (uint32_t i = 0; i < s->n_nodes; i++)
        if (!seen_node[i]) return 10;
    return 0;
}

void emu_intersect(const emu_scene *s, uint64_t n, const float *org, const float *dir, float tnear, float tfar,
                   int32_t *inst, int32_t *prim, float *u, float *v, float *t) {
    for (uint64_t i = 0; i < n; i++) {
        const RtHit h = rt_traverse(s->view.bvh, mk3(org[i * 3], org[i * 3 + 1], org[i * 3 + 2]),
                                    mk3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2]), tnear, tfar);
        if (h.tri == RT_MISS) {
            inst[i] = -1;
            prim[i] = -1;
            u[i] = 0;
            v[i] = 0;
            t[i] = tfar;
        } else {
            const uint32_t ii = rt_f2u(s->shade[(size_t)h.tri * 4 + 3].w);
            inst[i] = (int32_t)ii;
            prim[i] = (int32_t)(h.gid - s->inst[ii].first_tri);
            u[i] = h.u;
            v[i] = h.v;
            t[i] = h.t;
        }
    }
}

/* emulates rt_render_frame: kind 0 = k_megakernel, 1 = generate / extend / shade loop */
uint64_t emu_render(const emu_scene *s, int kind, const rt_camera *camera, const rt_render_params *params,
                    float *accum, uint8_t *rgba8, uint32_t *rng_out) {
    RtFrameParams p;
    memset(&p, 0, sizeof(p));
    p.roulette = (params->flags & RT_RENDER_ROULETTE) ? 1 : 0;
    p.cam.center = mk3(camera->center[0], camera->center[1], camera->center[2]);
    p.cam.pixel00 = mk3(camera->pixel00_loc[0], camera->pixel00_loc[1], camera->pixel00_loc[2]);
    p.cam.du = mk3(camera->pixel_delta_u[0], camera->pixel_delta_u[1], camera->pixel_delta_u[2]);
    p.cam.dv = mk3(camera->pixel_delta_v[0], camera->pixel_delta_v[1], camera->pixel_delta_v[2]);
    p.cam.w = camera->img_size[0];
    p.cam.h = camera->img_size[1];
    p.max_depth = params->max_depth;
    p.spp = params->sample_count;
    p.seed_salt = params->shard.seed_salt;
    p.rank = params->shard.rank;
    p.world = params->shard.world;
    p.tile_size = params->shard.tile_size;
    p.wavefront_seed = kind == 1;
    p.clamp_samples = kind == 1;
    p.resume = (params->flags & RT_RENDER_RESUME) ? 1 : 0; /* accum / rng_out then hold the previous frame (in/out) */
    p.keep_foreign = 0;
    p.tune_refill = 8;
    const uint32_t n_pix = (uint32_t)p.cam.w * (uint32_t)p.cam.h;
    std::vector<rt_float4> acc(n_pix, rt_mk_float4(0, 0, 0, 0));
    std::vector<uint32_t> bytes(n_pix, 0), rng_final(n_pix, 0);
    if (p.resume) {
        memcpy(acc.data(), accum, (size_t)n_pix * 16);
        memcpy(rng_final.data(), rng_out, (size_t)n_pix * 4);
    }
    unsigned long long rays = 0;
    if (kind == 0) {
        for (int y = 0; y < p.cam.h; y++)
            for (int x = 0; x < p.cam.w; x++) {
                if (!rt_owns_pixel(p, x, y)) continue;
                const size_t pix = (size_t)y * p.cam.w + x;
                XorShift32 rng;
                rng.a = rng_final[pix];
                const float count = (p.resume ? acc[pix].w : 0.0f) + (float)p.spp;
                const f3 sum = rt_megakernel_pixel(s->view, p, x, y, rng, rays, mk3(acc[pix].x, acc[pix].y, acc[pix].z), p.resume != 0);
                acc[pix] = rt_mk_float4(sum.x, sum.y, sum.z, count);
                bytes[pix] = rt_resolve_pixel(sum.x, sum.y, sum.z, count);
                rng_final[pix] = rng.a;
            }
    } else {
        std::vector<rt_float4> org(n_pix), hit(n_pix);
        std::vector<rt_uint2> dir(n_pix), att(n_pix), rad(n_pix), prog(n_pix);
        std::vector<uint32_t> rng(rng_final), q0(n_pix), q1(n_pix); /* resume: the renderer's own rng buffer lives on */
        uint32_t counts[2] = {0, 0};
        RtWavefrontState w;
        w.org = org.data();
        w.dir = dir.data();
        w.att = att.data();
        w.rad = rad.data();
        w.hit = hit.data();
        w.prog = prog.data();
        w.rng = rng.data();
        w.queue[0] = q0.data();
        w.queue[1] = q1.data();
        w.count[0] = &counts[0];
        w.count[1] = &counts[1];
        RtFrameOut out;
        out.accum = acc.data();
        out.rgba8 = bytes.data();
        out.rng = rng_final.data();
        out.gather = nullptr;
        for (uint32_t pix = 0; pix < n_pix; pix++)
            if (rt_wf_generate_pixel(p, w, out, pix)) w.queue[0][counts[0]++] = pix;
        int cur = 0;
        while (counts[cur]) {
            counts[cur ^ 1] = 0;
            rays += counts[cur];
            for (uint32_t i = 0; i < counts[cur]; i++) rt_wf_extend_pixel(s->view, w, w.queue[cur][i]);
            for (uint32_t i = 0; i < counts[cur]; i++) {
                const uint32_t pix = w.queue[cur][i];
                if (rt_wf_shade_pixel(s->view, p, w, out, pix)) w.queue[cur ^ 1][counts[cur ^ 1]++] = pix;
            }
            cur ^= 1;
        }
        for (uint32_t pix = 0; pix < n_pix; pix++) {
            const int x = (int)(pix % (uint32_t)p.cam.w), y = (int)(pix / (uint32_t)p.cam.w);
            if (rt_owns_pixel(p, x, y)) bytes[pix] = rt_resolve_pixel(acc[pix].x, acc[pix].y, acc[pix].z, acc[pix].w);
            rng_final[pix] = rng[pix];
        }
    }
    if (accum) memcpy(accum, acc.data(), (size_t)n_pix * 16);
    if (rgba8) memcpy(rgba8, bytes.data(), (size_t)n_pix * 4);
    if (rng_out) memcpy(rng_out, rng_final.data(), (size_t)n_pix * 4);
    return rays;
}

} /* extern "C" */

/* debugging aid: record the (org, dir) of every segment of pixel (x, y) (megakernel formulation) */
extern "C" uint32_t emu_trace_pixel(const emu_scene *s, int wavefront_seed, const rt_camera *camera,
                                    const rt_render_params *params, int x, int y, float *rays6, uint32_t max_rays) {
    RtCamera cam;
    cam.center = mk3(camera->center[0], camera->center[1], camera->center[2]);
    cam.pixel00 = mk3(camera->pixel00_loc[0], camera->pixel00_loc[1], camera->pixel00_loc[2]);
    cam.du = mk3(camera->pixel_delta_u[0], camera->pixel_delta_u[1], camera->pixel_delta_u[2]);
    cam.dv = mk3(camera->pixel_delta_v[0], camera->pixel_delta_v[1], camera->pixel_delta_v[2]);
    cam.w = camera->img_size[0];
    cam.h = camera->img_size[1];
    XorShift32 rng;
    rng.a = rt_pixel_seed(wavefront_seed, x, y, cam.w, cam.h) ^ params->shard.seed_salt;
    uint32_t n = 0;
    for (uint32_t sidx = 0; sidx < params->sample_count; sidx++) {
        RtRayState r = rt_camera_ray(cam, x, y, rng);
        for (uint32_t depth = 0; depth < params->max_depth; depth++) {
            if (n < max_rays) {
                float *o = rays6 + (size_t)n * 6;
                o[0] = r.org.x; o[1] = r.org.y; o[2] = r.org.z; o[3] = r.dir.x; o[4] = r.dir.y; o[5] = r.dir.z;
            }
            n++;
            const RtHit h = rt_traverse(s->view.bvh, r.org, r.dir, 0.0001f, INFINITY);
            f3 org = r.org, dir = r.dir, att = r.att, rad = r.rad, res;
            const bool done = rt_shade_segment(s->view, h, rng, org, dir, att, rad, res);
            r.org = org;
            r.dir = round_half3(dir);
            r.att = round_half3(att);
            r.rad = round_half3(rad);
            if (done) break;
        }
    }
    return n;
}

/* rt_u8_to_unit (rt_hd.h) and the direction-only ray set-up (rt_traverse.h) for the exhaustive / randomised checks */
extern "C" float emu_u8_to_unit(int v) { return rt_u8_to_unit((float)v); }
extern "C" void emu_ray_pre(const float *dir, float *out7, uint32_t *code) {
    const RtRayPre q = rt_ray_pre(mk3(dir[0], dir[1], dir[2]));
    out7[0] = q.rcp.x; out7[1] = q.rcp.y; out7[2] = q.rcp.z;
    out7[3] = q.nSx; out7[4] = q.nSy; out7[5] = q.Sz;
    *code = q.code;
}
extern "C" void emu_counters(unsigned long long *nodes, unsigned long long *tris, int reset) {
    *nodes = g_count_node;
    *tris = g_count_tri;
    if (reset) g_count_node = g_count_tri = 0;
}

/* tree statistics: out[0..8] = histogram of children per node, out[9] = leaf children, out[10] = inner children,
 * out[11..13] = leaves holding 1/2/3 triangles */
extern "C" void emu_scene_tree_stats(const emu_scene *s, uint64_t *out) {
    for (int i = 0; i < 14; i++) out[i] = 0;
    for (uint32_t ni = 0; ni < s->n_nodes; ni++) {
        const rt_uint4 *np = &s->nodes[(size_t)ni * RT_NODE_VEC4];
        int n = 0;
        for (int slot = 0; slot < 8; slot++) {
            const uint32_t unary = (np[1].z >> (3 * slot)) & 7u;
            const bool inner = (np[0].w >> 24) & (1u << slot);
            if (!inner && !unary) continue;
            n++;
            if (inner) out[10]++;
            else {
                out[9]++;
                out[10 + rt_popc(unary)]++;
            }
        }
        out[n]++;
    }
}

/* megakernel pixel-slot enumeration (rt_blocks.h): counts[y*w+x] += 1 for every in-image pixel of every
 * block this rank enumerates; returns the number of blocks */
extern "C" uint32_t emu_enumerate_blocks(int w, int h, uint32_t rank, uint32_t world, uint32_t tile_size, uint32_t *counts,
                                         uint32_t *foreign_out) {
    RtFrameParams p;
    memset(&p, 0, sizeof(p));
    p.cam.w = w;
    p.cam.h = h;
    p.rank = rank;
    p.world = world;
    p.tile_size = tile_size;
    const RtBlockGeom g = rt_block_geom(p);
    uint32_t foreign = 0;
    for (uint32_t blk = 0; blk < g.n_blocks; blk++) {
        uint32_t x0, y0;
        rt_block_origin(p, g, blk, x0, y0);
        for (uint32_t in = 0; in < 32; in++) {
            const int x = (int)(x0 + (in & 7u)), y = (int)(y0 + (in >> 3));
            if (x >= w || y >= h) continue;
            if (!rt_owns_pixel(p, x, y)) foreign++;
            counts[(size_t)y * w + x]++;
        }
    }
    *foreign_out = foreign;
    return g.n_blocks;
}

/* scheduling studies (tools/simd_model.py): per ray segment of pixel (x, y) the node steps and triangle tests of
 * the simple traversal, and whether the segment is the first of a sample. Returns the number of segments. */
extern "C" uint32_t emu_pixel_costs(const emu_scene *s, int wavefront_seed, const rt_camera *camera, const rt_render_params *params,
                                    int x, int y, uint16_t *node_steps, uint16_t *tri_steps, uint8_t *first_of_sample, uint32_t max_rays) {
    RtCamera cam;
    cam.center = mk3(camera->center[0], camera->center[1], camera->center[2]);
    cam.pixel00 = mk3(camera->pixel00_loc[0], camera->pixel00_loc[1], camera->pixel00_loc[2]);
    cam.du = mk3(camera->pixel_delta_u[0], camera->pixel_delta_u[1], camera->pixel_delta_u[2]);
    cam.dv = mk3(camera->pixel_delta_v[0], camera->pixel_delta_v[1], camera->pixel_delta_v[2]);
    cam.w = camera->img_size[0];
    cam.h = camera->img_size[1];
    XorShift32 rng;
    rng.a = rt_pixel_seed(wavefront_seed, x, y, cam.w, cam.h) ^ params->shard.seed_salt;
    uint32_t n = 0;
    for (uint32_t sidx = 0; sidx < params->sample_count; sidx++) {
        RtRayState r = rt_camera_ray(cam, x, y, rng);
        for (uint32_t depth = 0; depth < params->max_depth; depth++) {
            RtTravState tv;
            RtTravStacks ks;
            rt_trav_init(tv, r.org, r.dir, 0.0001f, INFINITY);
            uint32_t nn = 0, nt = 0;
            const RtRayTri rtri = rt_trav_ray_tri(tv);
            while (rt_trav_has_node(tv)) {
                rt_trav_node_step(s->view.bvh, tv, ks);
                nn++;
                while (rt_trav_has_tri(tv)) {
                    rt_trav_tri_step(s->view.bvh, tv, ks, rtri);
                    nt++;
                }
            }
            if (n < max_rays) {
                node_steps[n] = (uint16_t)std::min(nn, 65535u);
                tri_steps[n] = (uint16_t)std::min(nt, 65535u);
                first_of_sample[n] = depth == 0;
            }
            n++;
            f3 org = r.org, dir = r.dir, att = r.att, rad = r.rad, res;
            const bool done = rt_shade_segment(s->view, rt_trav_hit_noid(tv), rng, org, dir, att, rad, res);
            r.org = org;
            r.dir = round_half3(dir);
            r.att = round_half3(att);
            r.rad = round_half3(rad);
            if (done) break;
        }
    }
    return n;
}
