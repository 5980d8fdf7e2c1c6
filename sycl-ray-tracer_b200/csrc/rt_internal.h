/*
 * rt_internal.h — host-side objects behind the opaque handles of include/rt_api.h.
 */
#ifndef RT_INTERNAL_H
#define RT_INTERNAL_H

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/rt_api.h"
#include "rt_build.h"

/* App (src/app.hpp:31-58): device + in-order stream owner */
struct rt_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr; /* created by rt_context_create; `stream` may be borrowed */
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0;
    std::string name;
    std::string err;
    /* layered texture arrays of destroyed scenes, kept for the next scene with the same layer count:
     * cudaMalloc3DArray / cudaFreeArray are synchronous driver calls worth tens of milliseconds */
    std::vector<std::pair<uint32_t, cudaArray_t>> tex_cache;
    /* grow-only device scratch of rt_resolve / rt_intersect (no cudaMalloc / cudaFree per call) */
    void *scratch[2] = {nullptr, nullptr};
    size_t scratch_bytes[2] = {0, 0};
    /* pinned staging ring + copy streams of rt_upload (pageable host memory -> device at PCIe speed) */
    void *stage[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t stage_stream[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_event[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_begin = nullptr; /* the copy streams start after the work already on ctx->stream */
};

/* device scratch slot `k` of at least `bytes` (kept for the next call) */
cudaError_t rt_scratch(rt_context *ctx, int k, size_t bytes, void **out);
/* host (pageable / pinned) or device memory -> device, ordered on ctx->stream. Large pageable sources go through a
 * ring of pinned chunks filled by a few host threads, each with its own copy stream: a pageable cudaMemcpyAsync runs at
 * ~7 GB/s and stalls the calling thread, the staged path at PCIe speed (C4's 280 MB scene: ~40 ms -> ~12 ms). */
cudaError_t rt_upload(rt_context *ctx, void *dst, const void *src, size_t bytes);

/* Scene's device side (src/scene.hpp:64-91) */
struct rt_scene {
    rt_context *ctx = nullptr;
    bool committed = false;
    uint32_t n_inst = 0, n_tris = 0, n_verts = 0, n_indices = 0;
    uint32_t n_items = 0; /* leaf records of the tree: n_tris + the extra references of split triangles (rt_build.h) */
    /* object-space geometry as uploaded (inputs of the build) */
    float *d_positions = nullptr, *d_normals = nullptr, *d_uvs = nullptr;
    uint32_t *d_indices = nullptr;
    RtInstanceGeom *d_geom = nullptr;
    std::vector<RtInstanceGeom> h_geom;
    /* shading tables */
    RtInstance *d_inst = nullptr;
    std::vector<RtInstance> h_inst;
    /* textures: layered CUDA array bound as a texture object (ImageManager, F13) */
    cudaArray_t tex_array = nullptr;
    cudaTextureObject_t tex = 0;
    uint32_t n_layers = 0;
    float sky[3] = {0.5f, 0.7f, 1.0f};
    /* committed acceleration structure */
    rt_uint4 *d_nodes = nullptr;
    rt_float4 *d_tris = nullptr;
    rt_float4 *d_shade = nullptr;
    rt_scene_stats stats = {};
    RtScene view = {};
};

rt_status rt_set_error(rt_context *ctx, rt_status st, const char *what, const char *detail);

/* Scene-side device memory comes from the device's stream-ordered pool (release threshold raised in
 * rt_context_create), so building scene after scene reuses memory without a driver round trip:
 * cudaMalloc/cudaFree cost 20-700 ms per scene build on a fresh box. */
inline cudaError_t rt_pool_alloc(rt_context *ctx, void **p, size_t bytes) {
    return cudaMallocAsync(p, bytes ? bytes : 16, ctx->stream);
}
inline void rt_pool_free(rt_context *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}
rt_status rt_build_bvh(rt_scene *s); /* bvh_build.cu */
cudaError_t rt_launch_validate_indices(cudaStream_t st, const uint32_t *indices, const RtInstanceGeom *geom, uint32_t n_inst, uint32_t n_verts,
                                       uint64_t n_idx, uint32_t *bad); /* bvh_build.cu */
void rt_renderer_mark_exported(rt_renderer *r); /* rt_api.cu, for rt_group.cu */

#define RT_CUDA_TRY(ctx, expr)                                                     \
    do {                                                                           \
        cudaError_t rt_e_ = (expr);                                                \
        if (rt_e_ != cudaSuccess)                                                  \
            return rt_set_error((ctx), RT_ERR_CUDA, #expr, cudaGetErrorString(rt_e_)); \
    } while (0)

#endif
