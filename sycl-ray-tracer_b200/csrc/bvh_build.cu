/*
 * bvh_build.cu — GPU BVH build for sm_100a: the rtcCommitScene replacement
 * (src/scene.cpp:101-107; per-primitive scenes :406-439; instances :483-507).
 *
 *   k_flatten -> [k_split_count -> cub exclusive scan -> k_split_emit] -> k_morton -> cub radix sort (63-bit keys)
 *   -> k_karras -> k_fit -> per level { k_wide_select, cub exclusive scan, k_wide_emit }
 *
 * The per-item logic lives in rt_build.h. Output: 80-byte compressed 8-wide nodes in BFS order
 * (children of a node contiguous), triangles (48 B) and shading records (64 B) in leaf order.
 */
#include <cub/cub.cuh>

#include "rt_internal.h"

namespace {

constexpr int kBlock = 256;

__global__ void k_flatten(RtBuild b) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    float cx = INFINITY, cy = INFINITY, cz = INFINITY, dx = -INFINITY, dy = -INFINITY, dz = -INFINITY;
    if (gid < b.n_tris) {
        f3 lo, hi;
        rt_flatten_tri(b, gid, lo, hi);
        cx = dx = 0.5f * (lo.x + hi.x);
        cy = dy = 0.5f * (lo.y + hi.y);
        cz = dz = 0.5f * (lo.z + hi.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx = fminf(cx, __shfl_xor_sync(0xffffffffu, cx, o));
        cy = fminf(cy, __shfl_xor_sync(0xffffffffu, cy, o));
        cz = fminf(cz, __shfl_xor_sync(0xffffffffu, cz, o));
        dx = fmaxf(dx, __shfl_xor_sync(0xffffffffu, dx, o));
        dy = fmaxf(dy, __shfl_xor_sync(0xffffffffu, dy, o));
        dz = fmaxf(dz, __shfl_xor_sync(0xffffffffu, dz, o));
    }
    if ((threadIdx.x & 31) == 0 && cx <= dx) {
        atomicMin(&b.cen_bounds[0], rt_float_to_ordered(cx));
        atomicMin(&b.cen_bounds[1], rt_float_to_ordered(cy));
        atomicMin(&b.cen_bounds[2], rt_float_to_ordered(cz));
        atomicMax(&b.cen_bounds[3], rt_float_to_ordered(dx));
        atomicMax(&b.cen_bounds[4], rt_float_to_ordered(dy));
        atomicMax(&b.cen_bounds[5], rt_float_to_ordered(dz));
    }
}

/* every index of every instance must address one of that instance's vertices (checked where the indices already are:
 * the host loop over config 4's 30 M indices cost 15 ms per scene, more than their upload) */
__global__ void k_validate_indices(const uint32_t *indices, const RtInstanceGeom *geom, uint32_t n_inst, uint32_t n_verts, uint64_t n_idx,
                                   uint32_t *bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_idx) return;
    uint32_t lo = 0, hi = n_inst; /* last instance with first_index <= i (empty instances share an offset: take the last) */
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t)geom[mid].first_index <= i) lo = mid;
        else hi = mid;
    }
    const uint32_t v_end = lo + 1 < n_inst ? geom[lo + 1].first_vertex : n_verts;
    if (indices[i] >= v_end - geom[lo].first_vertex) atomicOr(bad, 1u);
}

__global__ void k_init_bounds(int32_t *cb) {
    if (threadIdx.x < 3) cb[threadIdx.x] = rt_float_to_ordered(INFINITY);
    else if (threadIdx.x < 6) cb[threadIdx.x] = rt_float_to_ordered(-INFINITY);
}

/* triangle splitting (rt_build.h, "split"): edge-length threshold for doubling step `shift` */
__global__ void k_split_len(RtBuild b, int shift, float *len2) { *len2 = rt_split_len2(b, shift); }

/* references per triangle for that threshold, and how many MORE references than triangles that makes */
__global__ void k_split_count(RtBuild b, const float *len2, uint32_t *counts, unsigned long long *extra) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (gid < b.n_tris) {
        c = rt_split_count(b, gid, *len2);
        counts[gid] = c;
        c -= 1u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(extra, (unsigned long long)c);
}

/* writes every triangle's references at its scanned offset; the centroid bounds become those of the references */
__global__ void k_split_emit(RtBuild b, const float *len2, const uint32_t *offsets, uint32_t *ref_tri, rt_float4 *ref_lo, rt_float4 *ref_hi) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    f3 lo = mk3(INFINITY, INFINITY, INFINITY), hi = mk3(-INFINITY, -INFINITY, -INFINITY);
    if (gid < b.n_tris) rt_split_emit(b, gid, *len2, offsets[gid], ref_tri, ref_lo, ref_hi, lo, hi);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0 && lo.x <= hi.x) {
        atomicMin(&b.cen_bounds[0], rt_float_to_ordered(lo.x));
        atomicMin(&b.cen_bounds[1], rt_float_to_ordered(lo.y));
        atomicMin(&b.cen_bounds[2], rt_float_to_ordered(lo.z));
        atomicMax(&b.cen_bounds[3], rt_float_to_ordered(hi.x));
        atomicMax(&b.cen_bounds[4], rt_float_to_ordered(hi.y));
        atomicMax(&b.cen_bounds[5], rt_float_to_ordered(hi.z));
    }
}

__global__ void k_morton(RtBuild b) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < b.n_items) rt_morton_tri(b, gid);
}

__global__ void k_karras(RtBuild b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < b.n_items) rt_karras_node(b, i);
}

struct DeviceArrive {
    __device__ uint32_t operator()(uint32_t *flag) const {
        __threadfence(); /* publish this subtree's box before signalling */
        const uint32_t old = atomicAdd(flag, 1u);
        __threadfence(); /* and do not read the sibling's box before the flag */
        return old;
    }
};

__global__ void k_fit(RtBuild b) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < b.n_items) rt_fit_leaf(b, j, DeviceArrive());
}

__global__ void k_wide_select(RtBuild b, uint32_t n_items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_items) rt_wide_select(b, i);
}

__global__ void k_wide_emit(RtBuild b, uint32_t n_items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_items) rt_wide_emit(b, i);
}

/* a root with eight empty slots (inverted boxes, imask = tmask = 0): every ray misses */
__global__ void k_empty_root(rt_uint4 *nodes) {
    const uint32_t ff = 0xffffffffu;
    nodes[0] = make_uint4(0, 0, 0, 1u | (1u << 8) | (1u << 16));
    nodes[1] = make_uint4(0, 0, 0, 0);
    nodes[2] = make_uint4(ff, ff, ff, ff); /* qlo x, y = 255 */
    nodes[3] = make_uint4(ff, ff, 0, 0);   /* qlo z = 255, qhi x = 0 */
    nodes[4] = make_uint4(0, 0, 0, 0);
#if RT_NODE_VEC4 > 5
    nodes[5] = make_uint4(0, 0, 0, 0);
#endif
}

inline uint32_t grid_for(uint64_t n) { return (uint32_t)((n + kBlock - 1) / kBlock); }

struct Scratch {
    rt_context *ctx;
    std::vector<void *> ptrs;
    explicit Scratch(rt_context *c) : ctx(c) {}
    ~Scratch() {
        for (void *p : ptrs) rt_pool_free(ctx, p);
    }
    template <class T>
    cudaError_t alloc(T **out, size_t count) {
        void *p = nullptr;
        cudaError_t e = rt_pool_alloc(ctx, &p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = (T *)p;
        return e;
    }
};

} // namespace

cudaError_t rt_launch_validate_indices(cudaStream_t st, const uint32_t *indices, const RtInstanceGeom *geom, uint32_t n_inst, uint32_t n_verts,
                                       uint64_t n_idx, uint32_t *bad) {
    if (!n_idx) return cudaSuccess;
    k_validate_indices<<<(unsigned)((n_idx + 255) / 256), 256, 0, st>>>(indices, geom, n_inst, n_verts, n_idx, bad);
    return cudaGetLastError();
}

static rt_status build_bvh(rt_scene *s) {
    rt_context *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t n = s->n_tris;

    RT_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, st));

    if (n == 0) { /* empty scene: a root with no children, every ray misses */
        RT_CUDA_TRY(ctx, rt_pool_alloc(ctx, (void **)&s->d_nodes, RT_NODE_VEC4 * sizeof(rt_uint4)));
        RT_CUDA_TRY(ctx, rt_pool_alloc(ctx, (void **)&s->d_tris, RT_TRI_VEC4 * sizeof(rt_float4)));
        RT_CUDA_TRY(ctx, rt_pool_alloc(ctx, (void **)&s->d_shade, 4 * sizeof(rt_float4)));
        k_empty_root<<<1, 1, 0, st>>>(s->d_nodes);
        RT_CUDA_TRY(ctx, cudaGetLastError());
        RT_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        s->stats.node_count = 1;
        s->stats.wide_depth = 1;
        s->n_items = 0;
        return RT_OK;
    }

    Scratch scratch(ctx);
    RtBuild b = {};
    b.n_tris = n;
    b.n_inst = s->n_inst;
    b.positions = s->d_positions;
    b.normals = s->d_normals;
    b.uvs = s->d_uvs;
    b.indices = s->d_indices;
    b.geom = s->d_geom;

    uint64_t *keys_in = nullptr, *keys_out = nullptr;
    uint32_t *vals_in = nullptr, *vals_out = nullptr;
    uint32_t *items_a = nullptr, *items_b = nullptr;
    uint64_t *counts = nullptr, *offsets = nullptr;
    rt_uint4 *nodes_tmp = nullptr;

    /* 1. flatten + centroid bounds */
    RT_CUDA_TRY(ctx, scratch.alloc(&b.wtris, (size_t)n * 3));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.cen_bounds, 8));
    k_init_bounds<<<1, 32, 0, st>>>(b.cen_bounds);
    k_flatten<<<grid_for(n), kBlock, 0, st>>>(b);
    RT_CUDA_TRY(ctx, cudaGetLastError());

    /* 1b. OPTIONAL (RT_SPLIT=1 in the environment of rt_scene_commit): references for large triangles (rt_build.h "split").
     * The threshold doubles until the extra references fit the budget. Off by default: measured with tools/tree_quality.py,
     * early splitting trades triangle tests for MORE node visits on every scene tried (stadium, Cornell), while the size-class
     * bit of the Morton key (rt_morton_tri) removes most of the LBVH's handicap on large triangles at no cost. */
    uint32_t n_items = n;
    if (getenv("RT_SPLIT") && atoi(getenv("RT_SPLIT")) > 0 && n > 1) {
        uint32_t *split_counts = nullptr, *split_offsets = nullptr;
        unsigned long long *d_extra = nullptr;
        float *d_len2 = nullptr;
        RT_CUDA_TRY(ctx, scratch.alloc(&split_counts, n));
        RT_CUDA_TRY(ctx, scratch.alloc(&d_extra, 1));
        RT_CUDA_TRY(ctx, scratch.alloc(&d_len2, 1));
        const unsigned long long budget = std::max<unsigned long long>(n / 2, 65536ull);
        unsigned long long extra = 0;
        for (int shift = 0; shift <= 8; shift++) {
            RT_CUDA_TRY(ctx, cudaMemsetAsync(d_extra, 0, sizeof(unsigned long long), st));
            k_split_len<<<1, 1, 0, st>>>(b, shift, d_len2);
            k_split_count<<<grid_for(n), kBlock, 0, st>>>(b, d_len2, split_counts, d_extra);
            RT_CUDA_TRY(ctx, cudaGetLastError());
            RT_CUDA_TRY(ctx, cudaMemcpyAsync(&extra, d_extra, sizeof(extra), cudaMemcpyDeviceToHost, st));
            RT_CUDA_TRY(ctx, cudaStreamSynchronize(st));
            if (extra <= budget) break;
            if (shift == 8) extra = 0; /* give up: no references */
        }
        if (extra > 0 && (unsigned long long)n + extra < 0x7fffffffull) {
            n_items = n + (uint32_t)extra;
            uint32_t *ref_tri = nullptr;
            rt_float4 *ref_lo = nullptr, *ref_hi = nullptr;
            RT_CUDA_TRY(ctx, scratch.alloc(&split_offsets, n));
            RT_CUDA_TRY(ctx, scratch.alloc(&ref_tri, n_items));
            RT_CUDA_TRY(ctx, scratch.alloc(&ref_lo, n_items));
            RT_CUDA_TRY(ctx, scratch.alloc(&ref_hi, n_items));
            size_t tmp_bytes = 0;
            RT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, split_counts, split_offsets, (int)n, st));
            void *tmp = nullptr;
            RT_CUDA_TRY(ctx, scratch.alloc((uint8_t **)&tmp, tmp_bytes));
            RT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, split_counts, split_offsets, (int)n, st));
            k_init_bounds<<<1, 32, 0, st>>>(b.cen_bounds);
            k_split_emit<<<grid_for(n), kBlock, 0, st>>>(b, d_len2, split_offsets, ref_tri, ref_lo, ref_hi);
            RT_CUDA_TRY(ctx, cudaGetLastError());
            b.ref_tri = ref_tri;
            b.ref_lo = ref_lo;
            b.ref_hi = ref_hi;
        }
    }
    b.n_items = n_items;
    s->n_items = n_items;
    const uint32_t m = n_items; /* leaves of the tree */
    const size_t max_level_items = (size_t)m / 2 + 8; /* an inner wide node covers >= 2 leaves */

    RT_CUDA_TRY(ctx, scratch.alloc(&keys_in, m));
    RT_CUDA_TRY(ctx, scratch.alloc(&keys_out, m));
    RT_CUDA_TRY(ctx, scratch.alloc(&vals_in, m));
    RT_CUDA_TRY(ctx, scratch.alloc(&vals_out, m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.left, m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.right, m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.parent, (size_t)2 * m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.range_first, m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.range_last, m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.box_lo, (size_t)2 * m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.box_hi, (size_t)2 * m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.flags, m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.dp_cost, (size_t)4 * m));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.dp_dec, (size_t)4 * m));
    RT_CUDA_TRY(ctx, scratch.alloc(&items_a, max_level_items));
    RT_CUDA_TRY(ctx, scratch.alloc(&items_b, max_level_items));
    RT_CUDA_TRY(ctx, scratch.alloc(&b.sel, max_level_items * 8));
    RT_CUDA_TRY(ctx, scratch.alloc(&counts, max_level_items));
    RT_CUDA_TRY(ctx, scratch.alloc(&offsets, max_level_items));
    RT_CUDA_TRY(ctx, scratch.alloc(&nodes_tmp, (size_t)m * RT_NODE_VEC4));
    RT_CUDA_TRY(ctx, rt_pool_alloc(ctx, (void **)&s->d_tris, (size_t)m * RT_TRI_VEC4 * sizeof(rt_float4)));
    RT_CUDA_TRY(ctx, rt_pool_alloc(ctx, (void **)&s->d_shade, (size_t)m * 4 * sizeof(rt_float4)));
    b.tris = s->d_tris;
    b.shade = s->d_shade;
    b.nodes = nodes_tmp;
    b.counts = counts;
    b.offsets = offsets;

    /* 2. Morton codes */
    b.keys = keys_in;
    b.vals = vals_in;
    k_morton<<<grid_for(m), kBlock, 0, st>>>(b);
    RT_CUDA_TRY(ctx, cudaGetLastError());

    /* 3. radix sort by Morton code */
    {
        size_t tmp_bytes = 0;
        RT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in,
                                                         vals_out, (int)m, 0, 63, st));
        void *tmp = nullptr;
        RT_CUDA_TRY(ctx, scratch.alloc((uint8_t **)&tmp, tmp_bytes));
        RT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out,
                                                         (int)m, 0, 63, st));
    }
    b.keys = keys_out;
    b.vals = vals_out;

    /* 4. radix tree, 5. bottom-up boxes */
    RT_CUDA_TRY(ctx, cudaMemsetAsync(b.flags, 0, (size_t)m * sizeof(uint32_t), st));
    if (m > 1) k_karras<<<grid_for(m - 1), kBlock, 0, st>>>(b);
    k_fit<<<grid_for(m), kBlock, 0, st>>>(b);
    RT_CUDA_TRY(ctx, cudaGetLastError());

    /* 6. collapse to the 8-wide compressed tree, one BFS level at a time */
    size_t scan_tmp_bytes = 0;
    RT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp_bytes, counts, offsets,
                                                   (int)max_level_items, st));
    void *scan_tmp = nullptr;
    RT_CUDA_TRY(ctx, scratch.alloc((uint8_t **)&scan_tmp, scan_tmp_bytes));

    const uint32_t root_id = 0; /* inner node 0, or leaf 0 when n == 1 (leaf0 = n-1 = 0) */
    RT_CUDA_TRY(ctx, cudaMemcpyAsync(items_a, &root_id, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    uint32_t level_items = 1, level_first = 0, tri_cursor = 0, depth = 0;
    uint32_t *cur = items_a, *nxt = items_b;
    while (level_items > 0) {
        b.items = cur;
        b.next_items = nxt;
        b.level_first_node = level_first;
        b.next_level_first_node = level_first + level_items;
        b.tri_cursor = tri_cursor;
        k_wide_select<<<grid_for(level_items), kBlock, 0, st>>>(b, level_items);
        RT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(scan_tmp, scan_tmp_bytes, counts, offsets,
                                                       (int)level_items, st));
        k_wide_emit<<<grid_for(level_items), kBlock, 0, st>>>(b, level_items);
        RT_CUDA_TRY(ctx, cudaGetLastError());
        uint64_t last_count = 0, last_off = 0;
        RT_CUDA_TRY(ctx, cudaMemcpyAsync(&last_count, counts + (level_items - 1), 8, cudaMemcpyDeviceToHost, st));
        RT_CUDA_TRY(ctx, cudaMemcpyAsync(&last_off, offsets + (level_items - 1), 8, cudaMemcpyDeviceToHost, st));
        RT_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        const uint64_t total = last_count + last_off;
        level_first += level_items;
        level_items = (uint32_t)(total >> 32);
        tri_cursor += (uint32_t)(total & 0xffffffffu);
        depth++;
        if (level_items > max_level_items)
            return rt_set_error(ctx, RT_ERR_STATE, "rt_build_bvh", "level overflow");
        uint32_t *t = cur;
        cur = nxt;
        nxt = t;
    }
    if (tri_cursor != m) return rt_set_error(ctx, RT_ERR_STATE, "rt_build_bvh", "triangle count mismatch");
    if (depth > RT_STACK_SIZE - 2)
        return rt_set_error(ctx, RT_ERR_STATE, "rt_build_bvh", "wide tree deeper than the traversal stack");

    const uint32_t node_count = level_first;
    RT_CUDA_TRY(ctx, rt_pool_alloc(ctx, (void **)&s->d_nodes, (size_t)node_count * RT_NODE_VEC4 * sizeof(rt_uint4)));
    RT_CUDA_TRY(ctx, cudaMemcpyAsync(s->d_nodes, nodes_tmp, (size_t)node_count * RT_NODE_VEC4 * sizeof(rt_uint4),
                                     cudaMemcpyDeviceToDevice, st));
    RT_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, st));
    RT_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    float ms = 0.0f;
    RT_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    s->stats.build_ms = ms;
    s->stats.node_count = node_count;
    s->stats.wide_depth = depth;
    return RT_OK;
}

rt_status rt_build_bvh(rt_scene *s) {
    const rt_status st = build_bvh(s);
    if (st != RT_OK) { /* a failed commit leaves nothing behind: the scene can be destroyed or committed again */
        rt_pool_free(s->ctx, s->d_nodes);
        rt_pool_free(s->ctx, s->d_tris);
        rt_pool_free(s->ctx, s->d_shade);
        s->d_nodes = nullptr;
        s->d_tris = nullptr;
        s->d_shade = nullptr;
    }
    return st;
}
