/*
 * rt_api.cu — the C ABI of include/rt_api.h over the CUDA kernels (librt_b200.so).
 * There is no CPU path in this library: every entry point that computes launches kernels on
 * the context's stream.
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <chrono>
#include <cstdio>
#include <thread>

#include "rt_internal.h"
#include "rt_render.h"

static thread_local std::string g_create_error;

rt_status rt_set_error(rt_context *ctx, rt_status st, const char *what, const char *detail) {
    std::string msg = std::string(what ? what : "") + ": " + (detail ? detail : "");
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return st;
}

cudaError_t rt_scratch(rt_context *ctx, int k, size_t bytes, void **out) {
    if (ctx->scratch_bytes[k] < bytes) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream); /* nothing may still use the old block */
        if (e != cudaSuccess) return e;
        cudaFree(ctx->scratch[k]);
        ctx->scratch[k] = nullptr;
        ctx->scratch_bytes[k] = 0;
        if ((e = cudaMalloc(&ctx->scratch[k], bytes)) != cudaSuccess) return e;
        ctx->scratch_bytes[k] = bytes;
    }
    *out = ctx->scratch[k];
    return cudaSuccess;
}

static bool is_pageable_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

static bool is_device_pointer(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

cudaError_t rt_upload(rt_context *ctx, void *dst, const void *src, size_t bytes) {
    constexpr size_t kChunk = 8u << 20; /* 8 MiB per pinned chunk */
    constexpr int kLanes = 4;           /* host copy threads = copy streams = pinned chunks */
    if (bytes == 0) return cudaSuccess;
    if (bytes < (4u << 20) || !is_pageable_host(src)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream);
    for (int k = 0; k < kLanes; k++) {
        cudaError_t e = cudaSuccess;
        if (!ctx->stage[k] && (e = cudaMallocHost(&ctx->stage[k], kChunk)) != cudaSuccess) return e;
        if (!ctx->stage_stream[k] && (e = cudaStreamCreateWithFlags(&ctx->stage_stream[k], cudaStreamNonBlocking)) != cudaSuccess) return e;
        if (!ctx->stage_event[k] && (e = cudaEventCreateWithFlags(&ctx->stage_event[k], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    /* the destination may still be in use by earlier work on the context stream (pool re-use): order the copy streams after it */
    cudaError_t e = cudaSuccess;
    if (!ctx->stage_begin && (e = cudaEventCreateWithFlags(&ctx->stage_begin, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(ctx->stage_begin, ctx->stream)) != cudaSuccess) return e;
    const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
    cudaError_t errs[kLanes] = {cudaSuccess, cudaSuccess, cudaSuccess, cudaSuccess};
    auto lane = [&](int k) {
        cudaError_t le = cudaSetDevice(ctx->device);
        if (le == cudaSuccess) le = cudaStreamWaitEvent(ctx->stage_stream[k], ctx->stage_begin, 0);
        for (size_t c = (size_t)k; c < n_chunks && le == cudaSuccess; c += kLanes) {
            const size_t off = c * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
            le = cudaEventSynchronize(ctx->stage_event[k]); /* the chunk's previous copy has left the pinned buffer */
            if (le != cudaSuccess) break;
            memcpy(ctx->stage[k], (const char *)src + off, len);
            le = cudaMemcpyAsync((char *)dst + off, ctx->stage[k], len, cudaMemcpyHostToDevice, ctx->stage_stream[k]);
            if (le == cudaSuccess) le = cudaEventRecord(ctx->stage_event[k], ctx->stage_stream[k]);
        }
        errs[k] = le;
    };
    {
        std::thread t1(lane, 1), t2(lane, 2), t3(lane, 3);
        lane(0);
        t1.join();
        t2.join();
        t3.join();
    }
    for (int k = 0; k < kLanes; k++) {
        if (errs[k] != cudaSuccess) return errs[k];
        if ((e = cudaStreamWaitEvent(ctx->stream, ctx->stage_event[k], 0)) != cudaSuccess) return e; /* the context stream continues after the last chunk of every lane */
    }
    return cudaSuccess;
}

struct rt_renderer {
    rt_context *ctx = nullptr;
    rt_renderer_kind kind = RT_MEGAKERNEL;
    int32_t w = 0, h = 0;
    float4 *d_accum = nullptr;
    uint32_t *d_rgba8 = nullptr;
    uint32_t *d_rng = nullptr;
    uint32_t *d_work = nullptr;             /* megakernel tile counter */
    unsigned long long *d_rays = nullptr;   /* ray segment counter */
    RtWavefrontState wf = {};
    uint32_t *d_counts = nullptr;           /* 2 queue lengths */
    uint32_t *h_counts = nullptr;           /* pinned mirror */
    unsigned long long *h_rays = nullptr;   /* pinned */
    int grid_mega = 0, grid_extend = 0, grid_shade = 0, grid_persist = 0, grid_flow = 0, flow_warps = 8;
    size_t queue_capacity = 0;    /* slots per wavefront id queue */
    int wf_persist = 2;           /* wavefront: 2 = queue-driven warps (one launch), 1 = per-CTA iterations (one launch), 0 = streaming kernels (RT_WF_PERSIST) */
    cudaEvent_t ev_batch[2] = {nullptr, nullptr}; /* wavefront: per-batch queue-length read-back */
    bool has_frame = false; /* a frame has been rendered: RT_RENDER_RESUME is allowed ... */
    const rt_scene *last_scene = nullptr; /* ... for the same scene, camera, depth and shard only */
    rt_camera last_camera = {};
    rt_shard last_shard = {};
    uint32_t last_depth = 0;
    uint32_t *d_order[4] = {nullptr, nullptr, nullptr, nullptr}; /* block order: keys in/out, values in/out */
    void *d_order_temp = nullptr;
    uint32_t *d_region_cost = nullptr;
    size_t order_temp_bytes = 0;
    uint32_t order_capacity = 0;
    int block_order = 1;          /* megakernel: hand blocks out by decreasing probed cost (RT_BLOCK_ORDER=0 disables) */
    int block_order_min_spp = 32; /* ... from this many samples per pixel (RT_BLOCK_ORDER_MIN_SPP) */
    int order_region = 32, order_probes = 1; /* cost-ordered hand-out: region side in pixels (RT_ORDER_REGION: 8 .. 256; 64 / 32 / 16 / 8 measured: 32 is +2 % on C2, neutral elsewhere), probe paths per block (RT_ORDER_PROBES: more than one never paid) */
    int sample_parts = 3;         /* megakernel: a pixel's samples are handed out in up to this many parts (RtFrameParams.n_parts; RT_SAMPLE_PARTS=1 disables) */
    uint32_t *d_part_done = nullptr; /* per pixel: parts finished in the current frame */
    uint32_t *gather = nullptr;   /* tile shards: owned pixels are also stored here (peer memory) */
    bool gather_ipc = false;      /* gather was opened from an IPC handle (close it) */
    bool exported = false;        /* d_rgba8 is a gather destination: never clear foreign pixels */
    float4 *d_chain_accum = nullptr; /* sample chains: one accumulation plane per chain ... */
    uint32_t *d_chain_rng = nullptr; /* ... and (megakernel) one stream-state plane per chain */
    uint32_t chain_planes = 0;       /* planes allocated (also the planes of the wavefront's per-pixel state) */
    uint32_t last_chains = 1;
    const float4 *peer_accum[16] = {}; /* spp slices across processes: every rank's accumulation buffer (IPC mappings) */
    uint32_t peer_world = 0, peer_rank = 0;
    int tune_refill = 16; /* lanes that must run dry before a warp refills (RT_TUNE_REFILL overrides; measured together with tune_carry,
                             profiles/r02_ab_refill_carry.log: megakernel 16 / 2, queue-driven wavefront 12 / 3) */
    int tune_carry = 2;   /* drain: lanes (of 32) that may carry unfinished triangles into the next drain instead of holding the
                             warp (RT_TUNE_CARRY; render.cu traverse_phase). 0 / 1 / 2 / 3 / 4 / 8 on C3: 3502 / 3571 / 3580 / 3559 / 3524 /
                             3359 Mrays/s, bit-identical images (profiles/r02_ab_drain_carry.log) */
    int tune_ctx = 0;     /* megakernel: parked ray contexts per lane (RT_MEGA_CTX 1-4 = k_megakernel_ctx; measured slower than the
                             one-pixel-in-registers kernel on C2/C3/C4, profiles/README.md) */
    int tune_inflight = 64; /* wavefront, queue-driven warps: pixels in flight per warp (RT_TUNE_INFLIGHT; measured 32 / 64 / 96 / 128 / 256 / 512:
                               64 is best on C2, C3 and C4 — the warp's pixel state stays in L1) */
    int tune_shade = 24, tune_idle = 4; /* megakernel contexts: shade-pass triggers (RT_TUNE_SHADE, RT_TUNE_IDLE) */
};

namespace {

template <class T>
cudaError_t dev_alloc(T **p, size_t count) {
    return cudaMalloc((void **)p, (count ? count : 1) * sizeof(T));
}

void normal_matrix(const float T[16], float out[9]) {
    /* transpose(inverse(mat3(T))) with glm's cofactor formula (src/scene.cpp:502) */
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

RtCamera to_device_camera(const rt_camera &c) {
    RtCamera d;
    d.center = mk3(c.center[0], c.center[1], c.center[2]);
    d.pixel00 = mk3(c.pixel00_loc[0], c.pixel00_loc[1], c.pixel00_loc[2]);
    d.du = mk3(c.pixel_delta_u[0], c.pixel_delta_u[1], c.pixel_delta_u[2]);
    d.dv = mk3(c.pixel_delta_v[0], c.pixel_delta_v[1], c.pixel_delta_v[2]);
    d.w = c.img_size[0];
    d.h = c.img_size[1];
    return d;
}

} // namespace

extern "C" {

uint32_t rt_api_version(void) { return RT_API_VERSION; }

const char *rt_last_error(rt_context *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

rt_status rt_context_create(int device, rt_context **out) {
    if (!out) return rt_set_error(nullptr, RT_ERR_INVALID, "rt_context_create", "out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return rt_set_error(nullptr, RT_ERR_NO_DEVICE, "rt_context_create",
                            e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device (there is no CPU path)");
    if (device < 0 || device >= n_dev) return rt_set_error(nullptr, RT_ERR_INVALID, "rt_context_create", "bad device index");
    rt_context *ctx = new (std::nothrow) rt_context();
    if (!ctx) return rt_set_error(nullptr, RT_ERR_INVALID, "rt_context_create", "out of memory");
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        rt_set_error(nullptr, RT_ERR_CUDA, "rt_context_create", cudaGetErrorString(e));
        delete ctx;
        return RT_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    {
        cudaMemPool_t pool = nullptr;
        unsigned long long keep = ~0ull; /* keep freed scene memory in the pool for the next build */
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        (void)cudaGetLastError();
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->name = prop.name;
    /* arithmetic-contract self test: a*b+c must not be contracted into an FMA */
    float *d = nullptr, h[2] = {0, 0};
    /* (1+2^-12)^2 = 1 + 2^-11 + 2^-24 rounds to 1 + 2^-11: unfused a*b+c == 0, fused == 2^-24 */
    const float a = 1.0f + 1.0f / 4096.0f, b = a, c = -(1.0f + 1.0f / 2048.0f);
    if ((e = cudaMalloc(&d, 2 * sizeof(float))) != cudaSuccess || (e = rt_launch_selftest(ctx->stream, a, b, c, d)) != cudaSuccess ||
        (e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
        rt_set_error(nullptr, RT_ERR_CUDA, "rt_context_create(selftest)", cudaGetErrorString(e));
        if (d) cudaFree(d);
        rt_context_destroy(ctx);
        return RT_ERR_CUDA;
    }
    cudaFree(d);
    if (!(h[0] == 0.0f && h[1] != 0.0f)) {
        rt_set_error(nullptr, RT_ERR_STATE, "rt_context_create",
                     "library was built with FMA contraction enabled (needs -fmad=false)");
        rt_context_destroy(ctx);
        return RT_ERR_STATE;
    }
    *out = ctx;
    return RT_OK;
}

void rt_context_destroy(rt_context *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto &t : ctx->tex_cache) cudaFreeArray(t.second);
    for (int k = 0; k < 2; k++) cudaFree(ctx->scratch[k]);
    for (int k = 0; k < 4; k++) {
        if (ctx->stage_stream[k]) {
            cudaStreamSynchronize(ctx->stage_stream[k]);
            cudaStreamDestroy(ctx->stage_stream[k]);
        }
        if (ctx->stage_event[k]) cudaEventDestroy(ctx->stage_event[k]);
        if (ctx->stage[k]) cudaFreeHost(ctx->stage[k]);
    }
    if (ctx->stage_begin) cudaEventDestroy(ctx->stage_begin);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

void *rt_context_stream(rt_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
rt_status rt_context_set_stream(rt_context *ctx, void *cuda_stream) {
    if (!ctx) return RT_ERR_INVALID;
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RT_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)cuda_stream;
    return RT_OK;
}
const char *rt_context_device_name(rt_context *ctx) { return ctx ? ctx->name.c_str() : ""; }

void rt_camera_init(rt_camera *cam, int32_t width, int32_t height, const float position[3], const float direction[3],
                    float focal_length) {
    /* Camera::Camera, src/camera.hpp:74-106 */
    cam->img_size[0] = width;
    cam->img_size[1] = height;
    const f3 center = mk3(position[0], position[1], position[2]);
    const f3 dir = normalize3(mk3(direction[0], direction[1], direction[2]));
    const f3 world_up = mk3(0.0f, 1.0f, 0.0f);
    auto cross = [](f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); };
    const f3 right = normalize3(cross(dir, world_up));
    const f3 up = normalize3(cross(right, dir));
    const float vp0 = 1.0f * ((float)width / (float)height), vp1 = 1.0f;
    const f3 viewport_u = (-right) * vp0;
    const f3 viewport_v = up * vp1;
    const f3 p00 = ((center + viewport_u) + viewport_v) + dir * focal_length;
    const f3 du = div3(right, (float)width / (vp0 * 2.0f));
    const f3 dv = div3(-up, (float)height / (vp1 * 2.0f));
    const f3 v[4] = {center, p00, du, dv};
    float *dst[4] = {cam->center, cam->pixel00_loc, cam->pixel_delta_u, cam->pixel_delta_v};
    for (int i = 0; i < 4; i++) {
        dst[i][0] = v[i].x;
        dst[i][1] = v[i].y;
        dst[i][2] = v[i].z;
    }
}

/* ------------------------------------------------------------------------------- scene */
rt_status rt_scene_create(rt_context *ctx, const rt_scene_desc *desc, rt_scene **out) {
    if (!ctx) return RT_ERR_INVALID;
    if (!desc || !out) return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "NULL argument");
    *out = nullptr;
    if (desc->instance_count && !desc->instances) return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "instances is NULL");
    if (desc->texture_layer_count > RT_MAX_IMAGES)
        return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "too many texture layers (MAX_IMAGES = 128)");
    if (desc->texture_layer_count && !desc->texture_layers)
        return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "texture_layers is NULL");
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const bool trace = getenv("RT_TRACE") != nullptr; /* development: host-side phase times on stderr */
    const auto t_begin = std::chrono::steady_clock::now();

    uint64_t n_verts = 0, n_idx = 0;
    for (uint32_t i = 0; i < desc->instance_count; i++) {
        const rt_instance &in = desc->instances[i];
        if (in.index_count % 3 != 0) return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "index_count not a multiple of 3");
        if (in.index_count && (!in.indices || !in.positions || !in.normals || !in.uvs))
            return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "instance buffer is NULL (POSITION, NORMAL, TEXCOORD_0 and indices are required)");
        if (in.material.type < RT_MAT_NONE || in.material.type > RT_MAT_DIELECTRIC)
            return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "bad material type");
        if (in.material.albedo_image >= (int32_t)desc->texture_layer_count)
            return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "albedo_image out of range");
        if (in.index_count && !in.vertex_count) return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "index out of range");
        n_verts += in.vertex_count;
        n_idx += in.index_count;
    }
    if (n_idx >= 0xffffffffull || n_verts >= 0xffffffffull) /* index / vertex offsets are 32-bit (first_index, first_vertex) */
        return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "scene too large (2^32 indices or vertices)");

    const auto t_checked = std::chrono::steady_clock::now();
    rt_scene *s = new (std::nothrow) rt_scene();
    if (!s) return rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "out of memory");
    s->ctx = ctx;
    s->n_inst = desc->instance_count;
    s->n_verts = (uint32_t)n_verts;
    s->n_indices = (uint32_t)n_idx;
    s->n_tris = (uint32_t)(n_idx / 3);
    memcpy(s->sky, desc->sky_color, sizeof(s->sky));

    s->h_geom.resize(desc->instance_count ? desc->instance_count : 1);
    s->h_inst.resize(desc->instance_count ? desc->instance_count : 1);
    uint32_t v0 = 0, i0 = 0;
    for (uint32_t i = 0; i < desc->instance_count; i++) {
        const rt_instance &in = desc->instances[i];
        RtInstanceGeom &g = s->h_geom[i];
        memcpy(g.transform, in.transform, sizeof(g.transform));
        g.first_vertex = v0;
        g.first_index = i0;
        g.first_tri = i0 / 3;
        g.tri_count = in.index_count / 3;
        RtInstance &m = s->h_inst[i];
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

    rt_status st = RT_OK;
    auto fail = [&](cudaError_t e, const char *what) {
        st = rt_set_error(ctx, RT_ERR_CUDA, what, cudaGetErrorString(e));
    };
    cudaError_t e;
    cudaStream_t stream = ctx->stream;
    uint32_t *d_bad = nullptr, h_bad = 0;
    do {
        if ((e = rt_pool_alloc(ctx, (void **)&s->d_positions, n_verts * 3 * sizeof(float))) != cudaSuccess) { fail(e, "alloc positions"); break; }
        if ((e = rt_pool_alloc(ctx, (void **)&s->d_normals, n_verts * 3 * sizeof(float))) != cudaSuccess) { fail(e, "alloc normals"); break; }
        if ((e = rt_pool_alloc(ctx, (void **)&s->d_uvs, n_verts * 2 * sizeof(float))) != cudaSuccess) { fail(e, "alloc uvs"); break; }
        if ((e = rt_pool_alloc(ctx, (void **)&s->d_indices, n_idx * sizeof(uint32_t))) != cudaSuccess) { fail(e, "alloc indices"); break; }
        if ((e = rt_pool_alloc(ctx, (void **)&s->d_geom, s->h_geom.size() * sizeof(*s->d_geom))) != cudaSuccess) { fail(e, "alloc geom"); break; }
        if ((e = rt_pool_alloc(ctx, (void **)&s->d_inst, s->h_inst.size() * sizeof(*s->d_inst))) != cudaSuccess) { fail(e, "alloc inst"); break; }
        /* every instance's arrays go straight from the caller's memory to their offset in the
         * concatenated device arrays (small or pinned sources directly, large pageable ones through rt_upload's
         * pinned ring) */
        for (uint32_t i = 0; i < desc->instance_count && e == cudaSuccess; i++) {
            const rt_instance &in = desc->instances[i];
            const RtInstanceGeom &g = s->h_geom[i];
            if (in.vertex_count) {
                e = rt_upload(ctx, s->d_positions + (size_t)g.first_vertex * 3, in.positions, (size_t)in.vertex_count * 12);
                if (e == cudaSuccess) e = rt_upload(ctx, s->d_normals + (size_t)g.first_vertex * 3, in.normals, (size_t)in.vertex_count * 12);
                if (e == cudaSuccess) e = rt_upload(ctx, s->d_uvs + (size_t)g.first_vertex * 2, in.uvs, (size_t)in.vertex_count * 8);
            }
            if (e == cudaSuccess && in.index_count)
                e = rt_upload(ctx, s->d_indices + g.first_index, in.indices, (size_t)in.index_count * 4);
        }
        if (e != cudaSuccess) { fail(e, "copy geometry"); break; }
        if ((e = cudaMemcpyAsync(s->d_geom, s->h_geom.data(), s->h_geom.size() * sizeof(RtInstanceGeom), cudaMemcpyHostToDevice, stream)) != cudaSuccess) { fail(e, "copy geom"); break; }
        /* index range check on the device, where the indices now are; the verdict is read after the upload has drained */
        if ((e = rt_scratch(ctx, 1, 16, (void **)&d_bad)) != cudaSuccess) { fail(e, "alloc"); break; }
        if ((e = cudaMemsetAsync(d_bad, 0, 4, stream)) != cudaSuccess) { fail(e, "memset"); break; }
        if ((e = rt_launch_validate_indices(stream, s->d_indices, s->d_geom, s->n_inst, s->n_verts, n_idx, d_bad)) != cudaSuccess) { fail(e, "validate indices"); break; }
        if ((e = cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) { fail(e, "validate indices"); break; }
        if ((e = cudaMemcpyAsync(s->d_inst, s->h_inst.data(), s->h_inst.size() * sizeof(RtInstance), cudaMemcpyHostToDevice, stream)) != cudaSuccess) { fail(e, "copy inst"); break; }
        /* texture array: RGBA8 512x512xN layered, point sampled, raw element reads (F13) */
        s->n_layers = desc->texture_layer_count;
        if (s->n_layers) {
            cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
            cudaExtent ext = make_cudaExtent(RT_TEX_SIZE, RT_TEX_SIZE, s->n_layers);
            for (size_t k = 0; k < ctx->tex_cache.size(); k++)
                if (ctx->tex_cache[k].first == s->n_layers) {
                    s->tex_array = ctx->tex_cache[k].second;
                    ctx->tex_cache.erase(ctx->tex_cache.begin() + (long)k);
                    break;
                }
            if (!s->tex_array && (e = cudaMalloc3DArray(&s->tex_array, &cd, ext, cudaArrayLayered)) != cudaSuccess) { fail(e, "alloc texture array"); break; }
            cudaMemcpy3DParms cp = {};
            cp.srcPtr = make_cudaPitchedPtr((void *)desc->texture_layers, RT_TEX_SIZE * 4, RT_TEX_SIZE, RT_TEX_SIZE);
            cp.dstArray = s->tex_array;
            cp.extent = ext;
            cp.kind = cudaMemcpyHostToDevice;
            if ((e = cudaMemcpy3DAsync(&cp, stream)) != cudaSuccess) { fail(e, "copy textures"); break; }
            cudaResourceDesc rd = {};
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = s->tex_array;
            cudaTextureDesc td = {};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            if ((e = cudaCreateTextureObject(&s->tex, &rd, &td, nullptr)) != cudaSuccess) { fail(e, "create texture object"); break; }
        }
        const auto t_queued = std::chrono::steady_clock::now();
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) { fail(e, "upload"); break; }
        if (trace) {
            const auto t_done = std::chrono::steady_clock::now();
            auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[rt trace] rt_scene_create: validate %.2f ms, stage + enqueue %.2f ms, drain %.2f ms (%.1f MB)\n", ms(t_begin, t_checked),
                    ms(t_checked, t_queued), ms(t_queued, t_done), (double)(n_verts * 32 + n_idx * 4) / 1e6);
        }
    } while (0);
    if (st == RT_OK && h_bad) st = rt_set_error(ctx, RT_ERR_INVALID, "rt_scene_create", "index out of range");
    if (st != RT_OK) {
        rt_scene_destroy(s);
        return st;
    }
    s->stats.triangle_count = s->n_tris;
    *out = s;
    return RT_OK;
}

rt_status rt_scene_commit(rt_scene *s) {
    if (!s) return RT_ERR_INVALID;
    rt_context *ctx = s->ctx;
    if (s->committed) return rt_set_error(ctx, RT_ERR_STATE, "rt_scene_commit", "scene already committed (immutable)");
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const auto t_commit = std::chrono::steady_clock::now();
    rt_status st = rt_build_bvh(s);
    if (st != RT_OK) return st;
    if (getenv("RT_TRACE")) /* development: host time of the build against its device time (the difference is allocation and the per-level read-backs) */
        fprintf(stderr, "[rt trace] rt_scene_commit: host %.2f ms, device %.2f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_commit).count(), (double)s->stats.build_ms);
    /* the object-space inputs are no longer needed */
    rt_pool_free(ctx, s->d_positions); s->d_positions = nullptr;
    rt_pool_free(ctx, s->d_normals); s->d_normals = nullptr;
    rt_pool_free(ctx, s->d_uvs); s->d_uvs = nullptr;
    rt_pool_free(ctx, s->d_indices); s->d_indices = nullptr;
    s->view.bvh.nodes = s->d_nodes;
    s->view.bvh.tris = s->d_tris;
    s->view.shade = s->d_shade;
    s->view.inst = s->d_inst;
    s->view.tex_raw = nullptr;
    s->view.tex = s->tex;
    s->view.n_layers = s->n_layers;
    s->view.sky = mk3(s->sky[0], s->sky[1], s->sky[2]);
    s->stats.bvh_bytes = s->stats.node_count * (16ull * RT_NODE_VEC4) + (uint64_t)s->n_items * (16ull * RT_TRI_VEC4);
    s->stats.shading_bytes = (uint64_t)s->n_items * 64ull + (uint64_t)s->n_inst * sizeof(RtInstance);
    s->stats.max_leaf_tris = RT_LEAF_MAX;
    s->committed = true;
    return RT_OK;
}

rt_status rt_scene_get_stats(const rt_scene *s, rt_scene_stats *out) {
    if (!s || !out) return RT_ERR_INVALID;
    *out = s->stats;
    return RT_OK;
}

void rt_scene_destroy(rt_scene *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    if (s->tex) cudaDestroyTextureObject(s->tex);
    if (s->tex_array) {
        if (s->ctx->tex_cache.size() < 2) {
            cudaStreamSynchronize(s->ctx->stream); /* no kernel may still sample it when it is reused */
            s->ctx->tex_cache.emplace_back(s->n_layers, s->tex_array);
        } else {
            cudaFreeArray(s->tex_array);
        }
    }
    void *bufs[] = {s->d_positions, s->d_normals, s->d_uvs, s->d_indices, s->d_geom, s->d_inst, s->d_nodes, s->d_tris, s->d_shade};
    for (void *b : bufs) rt_pool_free(s->ctx, b);
    delete s;
}

/* ------------------------------------------------------------------------------- intersect */
rt_status rt_intersect(rt_context *ctx, const rt_scene *scene, uint64_t n, const float *org, const float *dir,
                       float tnear, float tfar, int32_t *inst, int32_t *prim, float *u, float *v, float *t,
                       float *device_ms) {
    if (!ctx || !scene) return RT_ERR_INVALID;
    if (scene->ctx != ctx) return rt_set_error(ctx, RT_ERR_INVALID, "rt_intersect", "scene belongs to another context");
    if (!scene->committed) return rt_set_error(ctx, RT_ERR_STATE, "rt_intersect", "scene not committed");
    if (n && (!org || !dir || !inst || !prim || !u || !v || !t))
        return rt_set_error(ctx, RT_ERR_INVALID, "rt_intersect", "NULL argument");
    if (device_ms) *device_ms = 0.0f;
    if (n == 0) return RT_OK;
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    float *d_org = nullptr, *d_dir = nullptr, *d_f = nullptr;
    int32_t *d_i = nullptr;
    rt_status rs = RT_OK;
    cudaError_t e = cudaSuccess;
    do {
        if ((e = dev_alloc(&d_org, n * 3)) != cudaSuccess) break;
        if ((e = dev_alloc(&d_dir, n * 3)) != cudaSuccess) break;
        if ((e = dev_alloc(&d_f, n * 3)) != cudaSuccess) break;
        if ((e = dev_alloc(&d_i, n * 2)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(d_org, org, n * 12, cudaMemcpyDefault, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(d_dir, dir, n * 12, cudaMemcpyDefault, st)) != cudaSuccess) break;
        if ((e = cudaEventRecord(ctx->ev0, st)) != cudaSuccess) break;
        if ((e = rt_launch_intersect(st, scene->view, scene->d_inst, n, d_org, d_dir, tnear, tfar, d_i, d_i + n, d_f,
                                     d_f + n, d_f + 2 * n)) != cudaSuccess) break;
        if ((e = cudaEventRecord(ctx->ev1, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(inst, d_i, n * 4, cudaMemcpyDefault, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(prim, d_i + n, n * 4, cudaMemcpyDefault, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(u, d_f, n * 4, cudaMemcpyDefault, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(v, d_f + n, n * 4, cudaMemcpyDefault, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(t, d_f + 2 * n, n * 4, cudaMemcpyDefault, st)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
        float ms = 0.0f;
        if ((e = cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1)) != cudaSuccess) break;
        if (device_ms) *device_ms = ms;
    } while (0);
    if (e != cudaSuccess) rs = rt_set_error(ctx, RT_ERR_CUDA, "rt_intersect", cudaGetErrorString(e));
    cudaFree(d_org);
    cudaFree(d_dir);
    cudaFree(d_f);
    cudaFree(d_i);
    return rs;
}

/* ------------------------------------------------------------------------------- renderers */
rt_status rt_renderer_create(rt_context *ctx, rt_renderer_kind kind, int32_t width, int32_t height, rt_renderer **out) {
    if (!ctx) return RT_ERR_INVALID;
    if (!out) return rt_set_error(ctx, RT_ERR_INVALID, "rt_renderer_create", "out is NULL");
    *out = nullptr;
    if (width <= 0 || height <= 0 || (uint64_t)width * (uint64_t)height > 0x7fffffffull)
        return rt_set_error(ctx, RT_ERR_INVALID, "rt_renderer_create", "bad image size");
    if (kind != RT_MEGAKERNEL && kind != RT_WAVEFRONT) return rt_set_error(ctx, RT_ERR_INVALID, "rt_renderer_create", "bad renderer kind");
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    rt_renderer *r = new (std::nothrow) rt_renderer();
    if (!r) return rt_set_error(ctx, RT_ERR_INVALID, "rt_renderer_create", "out of memory");
    r->ctx = ctx;
    r->kind = kind;
    r->w = width;
    r->h = height;
    if (const char *e = getenv("RT_TUNE_REFILL")) r->tune_refill = atoi(e) > 0 ? atoi(e) : r->tune_refill;
    if (const char *e = getenv("RT_BLOCK_ORDER")) r->block_order = atoi(e);
    if (const char *e = getenv("RT_ORDER_REGION")) { const int v = atoi(e); if (v >= 8 && v <= 256 && (v & (v - 1)) == 0) r->order_region = v; }
    if (const char *e = getenv("RT_ORDER_PROBES")) { const int v = atoi(e); if (v >= 1 && v <= 16) r->order_probes = v; }
    if (const char *e = getenv("RT_SAMPLE_PARTS")) r->sample_parts = atoi(e) >= 1 && atoi(e) <= 3 ? atoi(e) : r->sample_parts;
    if (const char *e = getenv("RT_BLOCK_ORDER_MIN_SPP")) r->block_order_min_spp = atoi(e) > 0 ? atoi(e) : r->block_order_min_spp;
    if (const char *e = getenv("RT_WF_PERSIST")) r->wf_persist = atoi(e);
    if (const char *e = getenv("RT_TUNE_CARRY")) r->tune_carry = atoi(e) >= 0 && atoi(e) <= 32 ? atoi(e) : r->tune_carry;
    if (const char *e = getenv("RT_TUNE_INFLIGHT")) r->tune_inflight = atoi(e) >= 32 && atoi(e) <= 65536 ? (atoi(e) + 31) / 32 * 32 : r->tune_inflight;
    if (const char *e = getenv("RT_MEGA_CTX")) r->tune_ctx = atoi(e) >= 0 && atoi(e) <= 4 ? atoi(e) : r->tune_ctx;
    if (const char *e = getenv("RT_TUNE_SHADE")) r->tune_shade = atoi(e) > 0 ? atoi(e) : r->tune_shade;
    if (const char *e = getenv("RT_TUNE_IDLE")) r->tune_idle = atoi(e) > 0 ? atoi(e) : r->tune_idle;
    if (kind == RT_MEGAKERNEL && r->tune_ctx > 0 && !getenv("RT_TUNE_REFILL")) r->tune_refill = 8;
    if (kind == RT_WAVEFRONT && !getenv("RT_TUNE_REFILL")) r->tune_refill = r->wf_persist >= 2 ? 12 : 14;
    if (kind == RT_WAVEFRONT && !getenv("RT_TUNE_CARRY")) r->tune_carry = r->wf_persist >= 2 ? 3 : 2;
    const size_t n = (size_t)width * (size_t)height;
    cudaError_t e = cudaSuccess;
    do {
        if ((e = dev_alloc(&r->d_accum, n)) != cudaSuccess) break;
        if ((e = dev_alloc(&r->d_rgba8, n)) != cudaSuccess) break;
        if ((e = dev_alloc(&r->d_rng, n)) != cudaSuccess) break;
        if ((e = dev_alloc(&r->d_work, 1)) != cudaSuccess) break;
        if ((e = dev_alloc(&r->d_rays, 1)) != cudaSuccess) break;
        if ((e = dev_alloc(&r->d_counts, 4)) != cudaSuccess) break;
        if ((e = cudaMallocHost((void **)&r->h_counts, 2 * sizeof(uint32_t))) != cudaSuccess) break;
        if ((e = cudaMallocHost((void **)&r->h_rays, sizeof(unsigned long long))) != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&r->ev_batch[0], cudaEventDisableTiming)) != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&r->ev_batch[1], cudaEventDisableTiming)) != cudaSuccess) break;
        if (kind == RT_MEGAKERNEL) {
            if ((e = rt_megakernel_grid(ctx->sm_count, r->tune_ctx, &r->grid_mega)) != cudaSuccess) break;
        } else {
            /* Buffers + rng_buffer, src/render_wavefront.hpp:18-37, .cpp:52-54 */
            if ((e = dev_alloc(&r->wf.org, n)) != cudaSuccess) break;
            if ((e = dev_alloc(&r->wf.dir, n)) != cudaSuccess) break;
            if ((e = dev_alloc(&r->wf.att, n)) != cudaSuccess) break;
            if ((e = dev_alloc(&r->wf.rad, n)) != cudaSuccess) break;
            if ((e = dev_alloc(&r->wf.hit, n)) != cudaSuccess) break;
            if ((e = dev_alloc(&r->wf.prog, n)) != cudaSuccess) break;
            if ((e = dev_alloc(&r->wf.rng, n)) != cudaSuccess) break;
            if ((e = rt_wavefront_grid(ctx->sm_count, &r->grid_extend, &r->grid_shade)) != cudaSuccess) break;
            if ((e = rt_wf_persistent_grid(ctx->sm_count, &r->grid_persist)) != cudaSuccess) break;
            if ((e = rt_wf_flow_grid(ctx->sm_count, &r->grid_flow, &r->flow_warps)) != cudaSuccess) break;
            /* persistent form: one private queue segment per CTA, whole 8x4 blocks (rt_blocks.h) */
            r->queue_capacity = ((size_t)((width + 7) / 8) * (size_t)((height + 3) / 4) + (size_t)r->grid_persist) * 32u;
            if ((e = dev_alloc(&r->wf.queue[0], r->queue_capacity)) != cudaSuccess) break;
            if ((e = dev_alloc(&r->wf.queue[1], r->queue_capacity)) != cudaSuccess) break;
            r->wf.count[0] = r->d_counts;
            r->wf.count[1] = r->d_counts + 1;
            r->wf.head = r->d_counts + 2;
        }
    } while (0);
    if (e != cudaSuccess) {
        rt_status st = rt_set_error(ctx, RT_ERR_CUDA, "rt_renderer_create", cudaGetErrorString(e));
        rt_renderer_destroy(r);
        return st;
    }
    *out = r;
    return RT_OK;
}

static void peers_detach(rt_renderer *r) {
    for (uint32_t i = 0; i < r->peer_world; i++)
        if (i != r->peer_rank && r->peer_accum[i]) cudaIpcCloseMemHandle((void *)r->peer_accum[i]);
    memset(r->peer_accum, 0, sizeof(r->peer_accum));
    r->peer_world = r->peer_rank = 0;
}

static void gather_detach(rt_renderer *r) {
    if (r->gather && r->gather_ipc) cudaIpcCloseMemHandle(r->gather);
    r->gather = nullptr;
    r->gather_ipc = false;
}

void rt_renderer_destroy(rt_renderer *r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    gather_detach(r);
    peers_detach(r);
    for (uint32_t *q : r->d_order) cudaFree(q);
    cudaFree(r->d_order_temp);
    cudaFree(r->d_region_cost);
    cudaFree(r->d_chain_accum);
    cudaFree(r->d_chain_rng);
    cudaFree(r->d_accum);
    cudaFree(r->d_rgba8);
    cudaFree(r->d_rng);
    cudaFree(r->d_work);
    cudaFree(r->d_part_done);
    cudaFree(r->d_rays);
    cudaFree(r->d_counts);
    if (r->h_counts) cudaFreeHost(r->h_counts);
    if (r->h_rays) cudaFreeHost(r->h_rays);
    if (r->ev_batch[0]) cudaEventDestroy(r->ev_batch[0]);
    if (r->ev_batch[1]) cudaEventDestroy(r->ev_batch[1]);
    cudaFree(r->wf.org);
    cudaFree(r->wf.dir);
    cudaFree(r->wf.att);
    cudaFree(r->wf.rad);
    cudaFree(r->wf.hit);
    cudaFree(r->wf.prog);
    cudaFree(r->wf.rng);
    cudaFree(r->wf.queue[0]);
    cudaFree(r->wf.queue[1]);
    delete r;
}

} /* extern "C" */
void rt_renderer_mark_exported(rt_renderer *r) { r->exported = true; } /* rt_group: device 0's image is the gather destination */
extern "C" {

rt_status rt_renderer_export_image(rt_renderer *r, rt_ipc_handle *out) {
    if (!r) return RT_ERR_INVALID;
    rt_context *ctx = r->ctx;
    if (!out) return rt_set_error(ctx, RT_ERR_INVALID, "rt_renderer_export_image", "out is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(rt_ipc_handle), "rt_ipc_handle must hold a cudaIpcMemHandle_t");
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    RT_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, r->d_rgba8));
    memcpy(out->bytes, &h, sizeof(h));
    r->exported = true;
    return RT_OK;
}

rt_status rt_renderer_set_gather(rt_renderer *r, const rt_ipc_handle *handle, void *device_rgba8) {
    if (!r) return RT_ERR_INVALID;
    rt_context *ctx = r->ctx;
    if (handle && device_rgba8) return rt_set_error(ctx, RT_ERR_INVALID, "rt_renderer_set_gather", "pass a handle or a pointer, not both");
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RT_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); /* no frame may still be storing to the old target */
    gather_detach(r);
    if (handle) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle->bytes, sizeof(h));
        void *p = nullptr;
        RT_CUDA_TRY(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        r->gather = (uint32_t *)p;
        r->gather_ipc = true;
    } else if (device_rgba8) {
        r->gather = (uint32_t *)device_rgba8;
    }
    return RT_OK;
}

rt_status rt_renderer_export_accum(rt_renderer *r, rt_ipc_handle *out) {
    if (!r) return RT_ERR_INVALID;
    rt_context *ctx = r->ctx;
    if (!out) return rt_set_error(ctx, RT_ERR_INVALID, "rt_renderer_export_accum", "out is NULL");
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    RT_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, r->d_accum));
    memcpy(out->bytes, &h, sizeof(h));
    return RT_OK;
}

rt_status rt_renderer_set_peers(rt_renderer *r, const rt_ipc_handle *handles, uint32_t world, uint32_t rank) {
    if (!r) return RT_ERR_INVALID;
    rt_context *ctx = r->ctx;
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RT_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    peers_detach(r);
    if (!handles || world == 0) return RT_OK;
    if (world > 16 || rank >= world) return rt_set_error(ctx, RT_ERR_INVALID, "rt_renderer_set_peers", "world <= 16 and rank < world");
    r->peer_world = world;
    r->peer_rank = rank;
    for (uint32_t i = 0; i < world; i++) {
        if (i == rank) {
            r->peer_accum[i] = r->d_accum;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles[i].bytes, sizeof(h));
        void *p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            peers_detach(r);
            return rt_set_error(ctx, RT_ERR_CUDA, "cudaIpcOpenMemHandle", cudaGetErrorString(e));
        }
        r->peer_accum[i] = (const float4 *)p;
    }
    return RT_OK;
}

rt_status rt_renderer_reduce_resolve(rt_renderer *r) {
    if (!r) return RT_ERR_INVALID;
    rt_context *ctx = r->ctx;
    if (!r->peer_world) return rt_set_error(ctx, RT_ERR_STATE, "rt_renderer_reduce_resolve", "no peers attached (rt_renderer_set_peers)");
    uint32_t *dst = r->peer_rank == 0 ? r->d_rgba8 : r->gather;
    if (!dst) return rt_set_error(ctx, RT_ERR_STATE, "rt_renderer_reduce_resolve", "no destination image (rt_renderer_set_gather with rank 0's handle)");
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)r->w * (size_t)r->h;
    const uint32_t per = (uint32_t)((n + r->peer_world - 1) / r->peer_world), first = r->peer_rank * per;
    const uint32_t count = first < n ? (uint32_t)(n - first < per ? n - first : per) : 0u;
    RT_CUDA_TRY(ctx, rt_launch_reduce_resolve_peer(ctx->stream, r->peer_accum, r->peer_world, first, count, nullptr, dst));
    return RT_OK;
}

float *rt_renderer_device_accum(rt_renderer *r) { return r ? (float *)r->d_accum : nullptr; }
uint8_t *rt_renderer_device_rgba8(rt_renderer *r) { return r ? (uint8_t *)r->d_rgba8 : nullptr; }
uint32_t *rt_renderer_device_rng(rt_renderer *r) { return r ? r->d_rng : nullptr; }

rt_status rt_render_frame(rt_renderer *r, const rt_scene *scene, const rt_camera *camera,
                          const rt_render_params *params, rt_frame *frame) {
    if (!r) return RT_ERR_INVALID;
    rt_context *ctx = r->ctx;
    if (!scene || !camera || !params || !frame) return rt_set_error(ctx, RT_ERR_INVALID, "rt_render_frame", "NULL argument");
    if (scene->ctx != ctx) return rt_set_error(ctx, RT_ERR_INVALID, "rt_render_frame", "scene belongs to another context");
    if (!scene->committed) return rt_set_error(ctx, RT_ERR_STATE, "rt_render_frame", "scene not committed");
    if (camera->img_size[0] != r->w || camera->img_size[1] != r->h)
        return rt_set_error(ctx, RT_ERR_INVALID, "rt_render_frame", "camera image size differs from the renderer's");
    const rt_shard &sh = params->shard;
    if (sh.world > 1 && sh.rank >= sh.world) return rt_set_error(ctx, RT_ERR_INVALID, "rt_render_frame", "shard rank >= world");
    if (sh.world > 1 && sh.tile_size % 8 != 0) return rt_set_error(ctx, RT_ERR_INVALID, "rt_render_frame", "tile_size must be a multiple of 8");
    if (sh.world > 1 && sh.tile_size > 16384u) return rt_set_error(ctx, RT_ERR_INVALID, "rt_render_frame", "tile_size must be <= 16384");
    if (params->flags & ~(RT_RENDER_RESUME | RT_RENDER_ROULETTE)) return rt_set_error(ctx, RT_ERR_INVALID, "rt_render_frame", "unknown bits in rt_render_params.flags");
    const uint32_t chains = params->sample_chains > 1u ? params->sample_chains : 1u;
    if (chains > 16u) return rt_set_error(ctx, RT_ERR_INVALID, "rt_render_frame", "sample_chains must be <= 16");
    if ((uint64_t)r->w * (uint64_t)r->h * chains > 0x7fffffffull) return rt_set_error(ctx, RT_ERR_INVALID, "rt_render_frame", "image size x sample_chains too large");
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    RtFrameParams p;
    p.cam = to_device_camera(*camera);
    p.max_depth = params->max_depth;
    p.spp = params->sample_count;
    p.seed_salt = sh.seed_salt;
    p.rank = sh.rank;
    p.world = sh.world;
    p.tile_size = sh.tile_size;
    p.wavefront_seed = r->kind == RT_WAVEFRONT ? 1 : 0;
    p.clamp_samples = r->kind == RT_WAVEFRONT ? 1 : 0;
    p.tune_refill = (r->tune_refill & 0xff) | ((r->tune_ctx == 0 ? r->tune_carry : 0) << 8); /* packed: both reach traverse_phase as one kernel argument */
    p.tune_ctx = r->tune_ctx;
    p.tune_shade = r->tune_shade;
    p.tune_idle = r->tune_idle;
    p.tune_inflight = r->tune_inflight;
    p.resume = (params->flags & RT_RENDER_RESUME) ? 1 : 0;
    p.roulette = (params->flags & RT_RENDER_ROULETTE) ? 1 : 0;
    p.chains = chains;
    p.n_parts = 1;
    p.order_region = (uint32_t)r->order_region;
    p.order_probes = (uint32_t)r->order_probes;
    p.part_end[0] = p.part_end[1] = p.part_end[2] = p.spp;
    if (p.resume && !r->has_frame) return rt_set_error(ctx, RT_ERR_STATE, "rt_render_frame", "RT_RENDER_RESUME without a previous frame");
    if (p.resume && (r->last_scene != scene || memcmp(&r->last_camera, camera, sizeof(rt_camera)) != 0 || r->last_depth != params->max_depth ||
                     r->last_shard.rank != sh.rank || r->last_shard.world != sh.world || r->last_shard.tile_size != sh.tile_size ||
                     r->last_shard.seed_salt != sh.seed_salt || r->last_chains != chains))
        return rt_set_error(ctx, RT_ERR_STATE, "rt_render_frame", "RT_RENDER_RESUME with a different scene, camera, depth or shard than the previous frame");
    RtFrameOut out;
    out.accum = r->d_accum;
    out.rgba8 = r->d_rgba8;
    out.rng = r->d_rng;
    out.gather = (sh.world > 1 && sh.tile_size) ? r->gather : nullptr;
    out.part_done = nullptr;
    p.keep_foreign = r->exported ? 1 : 0;
    const size_t n = (size_t)r->w * (size_t)r->h;
    uint32_t launches = 0;
    /* sample chains: the kernels work on `chains` planes of per-pixel state and k_combine_chains writes the frame */
    RtFrameOut out_frame = out;
    if (chains > 1u) {
        if (r->chain_planes < chains) {
            RT_CUDA_TRY(ctx, cudaStreamSynchronize(st));
            cudaFree(r->d_chain_accum);
            cudaFree(r->d_chain_rng);
            r->d_chain_accum = nullptr;
            r->d_chain_rng = nullptr;
            r->chain_planes = 0;
            r->has_frame = false;
            if (p.resume) return rt_set_error(ctx, RT_ERR_STATE, "rt_render_frame", "RT_RENDER_RESUME with more sample chains than the previous frame");
            RT_CUDA_TRY(ctx, dev_alloc(&r->d_chain_accum, n * chains));
            RT_CUDA_TRY(ctx, dev_alloc(&r->d_chain_rng, n * chains));
            if (r->kind == RT_WAVEFRONT) { /* the per-pixel ray state gets one plane per chain too */
                cudaFree(r->wf.org); cudaFree(r->wf.dir); cudaFree(r->wf.att); cudaFree(r->wf.rad); cudaFree(r->wf.hit); cudaFree(r->wf.prog); cudaFree(r->wf.rng);
                r->wf.org = nullptr; r->wf.dir = nullptr; r->wf.att = nullptr; r->wf.rad = nullptr; r->wf.hit = nullptr; r->wf.prog = nullptr; r->wf.rng = nullptr;
                RT_CUDA_TRY(ctx, dev_alloc(&r->wf.org, n * chains));
                RT_CUDA_TRY(ctx, dev_alloc(&r->wf.dir, n * chains));
                RT_CUDA_TRY(ctx, dev_alloc(&r->wf.att, n * chains));
                RT_CUDA_TRY(ctx, dev_alloc(&r->wf.rad, n * chains));
                RT_CUDA_TRY(ctx, dev_alloc(&r->wf.hit, n * chains));
                RT_CUDA_TRY(ctx, dev_alloc(&r->wf.prog, n * chains));
                RT_CUDA_TRY(ctx, dev_alloc(&r->wf.rng, n * chains));
            }
            r->chain_planes = chains;
        }
        out.accum = r->d_chain_accum;
        out.rng = r->d_chain_rng;
    }

    /* block order (both persistent kernels hand pixels out in 8x4 blocks from a global counter): the probe is a fixed cost
     * (one low-occupancy path per block), the tail it removes grows with the length of a pixel's sequential chain:
     * measured +10 % on C2 (64 spp), +1.5 % on C3, -1.5 % on C4 at 16 spp -> only from 32 spp per chain */
    auto block_order = [&](const uint32_t **order) -> rt_status {
        *order = nullptr;
        const uint32_t n_blocks = rt_block_count(p);
        if (!(r->block_order && p.spp >= (uint32_t)r->block_order_min_spp * chains && p.max_depth >= 2 && n_blocks >= 1024 &&
              (uint64_t)r->w + sh.tile_size <= 65536u && (uint64_t)r->h + sh.tile_size <= 65536u)) /* block origins are packed 16 + 16 bits, partial edge tiles included */
            return RT_OK;
        if (n_blocks > r->order_capacity) { /* first frame (or a coarser tiling): (re)allocate */
            RT_CUDA_TRY(ctx, cudaStreamSynchronize(st));
            for (uint32_t *&q : r->d_order) {
                cudaFree(q);
                q = nullptr;
            }
            cudaFree(r->d_order_temp);
            r->d_order_temp = nullptr;
            r->order_capacity = 0;
            for (uint32_t *&q : r->d_order) RT_CUDA_TRY(ctx, dev_alloc(&q, n_blocks));
            RT_CUDA_TRY(ctx, rt_block_order_temp_bytes(n_blocks, &r->order_temp_bytes));
            RT_CUDA_TRY(ctx, cudaMalloc(&r->d_order_temp, r->order_temp_bytes ? r->order_temp_bytes : 1));
            if (!r->d_region_cost) RT_CUDA_TRY(ctx, dev_alloc(&r->d_region_cost, rt_region_count(r->w, r->h)));
            r->order_capacity = n_blocks;
        }
        RT_CUDA_TRY(ctx, rt_launch_block_order(st, scene->view, p, r->d_region_cost, r->d_order[0], r->d_order[1], r->d_order[2], r->d_order[3],
                                               r->d_order_temp, r->order_temp_bytes));
        *order = r->d_order[3];
        launches += 2;
        return RT_OK;
    };

    RT_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, st));
    RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_rays, 0, sizeof(unsigned long long), st));
    if (r->kind == RT_MEGAKERNEL) {
        RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_work, 0, sizeof(uint32_t), st));
        if (sh.world > 1 && sh.tile_size && !p.resume) { /* pixels of other ranks stay zero */
            RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_accum, 0, n * sizeof(float4), st));
            if (!r->exported) RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_rgba8, 0, n * 4, st));
            RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_rng, 0, n * 4, st));
        }
        if (chains > 1u) p.tune_ctx = 0; /* chains are implemented by the one-pixel-in-registers kernel */
        /* sample parts (rt_shade.h, RtFrameParams): the frame ends with the longest chain handed out last, ~1.4 pixel chains
         * after the work counter runs dry — 10 % of a 1080p frame at any spp, because a B200 keeps 151 k lanes busy and the
         * frame has only 13.7 pixels per lane. Each part's tail is filled by the next part's work when
         * (next part) / (this part) >= 1.43 * lanes / pixels, so the parts shrink geometrically by rho = 1.6 * lanes / pixels */
        if (chains <= 1u && p.tune_ctx == 0 && r->sample_parts > 1 && p.spp >= 2u && p.max_depth >= 1u) {
            const double lanes = (double)r->grid_mega * 128.0, pixels = (double)rt_block_count(p) * 32.0;
            const double rho = std::min(0.5, std::max(1.0 / 64.0, 1.6 * lanes / std::max(pixels, 1.0)));
            const uint32_t want = p.spp >= 3u && r->sample_parts >= 3 ? 3u : 2u;
            const double norm = want == 3u ? 1.0 + rho + rho * rho : 1.0 + rho;
            uint32_t c2 = want == 3u ? std::max(1u, (uint32_t)llround(p.spp * rho * rho / norm)) : 0u;
            uint32_t c1 = std::max(1u, (uint32_t)llround(p.spp * rho / norm));
            while (c1 + c2 >= p.spp) { /* tiny spp: keep at least one sample in the first part */
                if (c1 > 1u) c1--;
                else if (c2 > 1u) c2--;
                else break;
            }
            if (c1 + c2 < p.spp) {
                p.n_parts = want;
                p.part_end[0] = p.spp - c1 - c2;
                p.part_end[1] = p.spp - c2;
                p.part_end[2] = p.spp;
                if (want == 2u) p.part_end[1] = p.spp;
                if (!r->d_part_done) RT_CUDA_TRY(ctx, dev_alloc(&r->d_part_done, n));
                RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_part_done, 0, n * sizeof(uint32_t), st));
                out.part_done = r->d_part_done;
            }
        }
        const uint32_t *order = nullptr;
        {
            const rt_status os = block_order(&order);
            if (os != RT_OK) return os;
        }
        RT_CUDA_TRY(ctx, rt_launch_megakernel(st, r->grid_mega, scene->view, p, out, r->d_work, r->d_rays, order));
        launches++;
        if (chains > 1u) {
            RT_CUDA_TRY(ctx, rt_launch_combine_chains(st, p, (const float *)r->d_chain_accum, r->d_chain_rng, out_frame));
            launches++;
        }
    } else if (r->wf_persist || chains > 1u) {
        const uint32_t n_blocks = rt_block_count(p);
        uint32_t grid, cap;
        const bool flow = r->wf_persist >= 2 || chains > 1u; /* chains are implemented by the queue-driven form */
        if (flow) { /* one ray ring + one hit ring per warp, a power of two >= the pixels a warp keeps in flight */
            grid = (uint32_t)r->grid_flow;
            cap = 32u;
            while (cap < (uint32_t)r->tune_inflight) cap <<= 1;
            cap *= (uint32_t)r->flow_warps; /* per CTA, so that the allocation check below covers both forms */
        } else {
            grid = (uint32_t)r->grid_persist;
            cap = ((n_blocks + grid - 1) / grid) * 32u;
        }
        /* the whole frame in one launch: every CTA runs generate / {extend, shade}* / resolve for its own lattice of
         * 8x4 pixel blocks with CTA-local queues (k_wf_persistent) */
        if ((size_t)grid * cap > r->queue_capacity) { /* a tiling whose partial edge tiles enumerate more blocks than the image has */
            RT_CUDA_TRY(ctx, cudaStreamSynchronize(st));
            cudaFree(r->wf.queue[0]);
            cudaFree(r->wf.queue[1]);
            r->wf.queue[0] = r->wf.queue[1] = nullptr;
            r->queue_capacity = 0;
            RT_CUDA_TRY(ctx, dev_alloc(&r->wf.queue[0], (size_t)grid * cap));
            RT_CUDA_TRY(ctx, dev_alloc(&r->wf.queue[1], (size_t)grid * cap));
            r->queue_capacity = (size_t)grid * cap;
        }
        if (sh.world > 1 && sh.tile_size && !p.resume) { /* pixels of other ranks stay zero (they are never enumerated) */
            RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_accum, 0, n * sizeof(float4), st));
            if (!r->exported) RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_rgba8, 0, n * 4, st));
            RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_rng, 0, n * 4, st));
            RT_CUDA_TRY(ctx, cudaMemsetAsync(r->wf.rng, 0, n * 4, st));
        }
        if (flow) {
            RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_work, 0, sizeof(uint32_t), st));
            const uint32_t *order = nullptr;
            const rt_status os = block_order(&order);
            if (os != RT_OK) return os;
            RT_CUDA_TRY(ctx, rt_launch_wf_flow(st, (int)grid, cap / (uint32_t)r->flow_warps, scene->view, p, r->wf, out, r->d_work, r->d_rays, order));
            if (chains > 1u) {
                RT_CUDA_TRY(ctx, rt_launch_combine_chains(st, p, (const float *)r->d_chain_accum, r->wf.rng, out_frame));
                launches++;
            }
        }
        else RT_CUDA_TRY(ctx, rt_launch_wf_persistent(st, (int)grid, cap, scene->view, p, r->wf, out, r->d_rays));
        launches++;
    } else {
        RT_CUDA_TRY(ctx, cudaMemsetAsync(r->d_counts, 0, 4 * sizeof(uint32_t), st));
        RT_CUDA_TRY(ctx, rt_launch_wf_generate(st, r->grid_shade, p, r->wf, out));
        launches++;
        /* every pixel traces at most spp * max_depth segments, one per bounce iteration */
        const uint64_t max_iters = (uint64_t)p.spp * (uint64_t)p.max_depth;
        /* One batch = kBatch (even) bounce iterations = 2*kBatch kernels, captured once per frame in
         * a CUDA graph and replayed (the queues ping-pong, so every batch starts at queue 0). The
         * queue length is copied back after every batch, but the host only looks at the PREVIOUS
         * batch's value after enqueueing the next one, so the GPU never waits for the host. Batches
         * issued after the queue ran empty are no-ops (count 0). */
        const int kBatch = 8;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        auto issue_batch = [&]() -> cudaError_t {
            int cur = 0;
            for (int k = 0; k < kBatch; k++) {
                cudaError_t e = rt_launch_wf_extend(st, r->grid_extend, scene->view, r->wf, cur, r->d_rays, p);
                if (e != cudaSuccess) return e;
                e = rt_launch_wf_shade(st, r->grid_shade, scene->view, p, r->wf, out, cur);
                if (e != cudaSuccess) return e;
                cur ^= 1;
            }
            return cudaSuccess;
        };
        if (max_iters > (uint64_t)kBatch && cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            cudaError_t e1 = issue_batch();
            cudaError_t e2 = cudaStreamEndCapture(st, &graph);
            if (e1 != cudaSuccess || e2 != cudaSuccess || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                graph = nullptr;
                exec = nullptr;
            }
        }
        (void)cudaGetLastError(); /* capture is unsupported on the legacy default stream: plain launches */
        cudaError_t werr = cudaSuccess;
        bool finished = false;
        uint64_t b = 0;
        for (uint64_t it = 0; it < max_iters && !finished && werr == cudaSuccess; it += kBatch, b++) {
            werr = exec ? cudaGraphLaunch(exec, st) : issue_batch();
            launches += 2 * kBatch;
            if (werr == cudaSuccess) werr = cudaMemcpyAsync(&r->h_counts[b & 1], r->d_counts, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
            if (werr == cudaSuccess) werr = cudaEventRecord(r->ev_batch[b & 1], st);
            if (werr == cudaSuccess && b >= 1) {
                werr = cudaEventSynchronize(r->ev_batch[(b - 1) & 1]);
                if (r->h_counts[(b - 1) & 1] == 0) finished = true; /* every pixel has finished its samples */
            }
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        if (werr != cudaSuccess) return rt_set_error(ctx, RT_ERR_CUDA, "rt_render_frame(wavefront)", cudaGetErrorString(werr));
        RT_CUDA_TRY(ctx, rt_launch_resolve_owned(st, p, (const float *)r->d_accum, r->wf.rng, out));
        launches++;
    }
    RT_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, st));
    RT_CUDA_TRY(ctx, cudaMemcpyAsync(r->h_rays, r->d_rays, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (frame->rgba8) RT_CUDA_TRY(ctx, cudaMemcpyAsync(frame->rgba8, r->d_rgba8, n * 4, cudaMemcpyDefault, st));
    if (frame->accum) RT_CUDA_TRY(ctx, cudaMemcpyAsync(frame->accum, r->d_accum, n * sizeof(float4), cudaMemcpyDefault, st));
    if (frame->rng_state) RT_CUDA_TRY(ctx, cudaMemcpyAsync(frame->rng_state, r->d_rng, n * 4, cudaMemcpyDefault, st));
    RT_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    float ms = 0.0f;
    RT_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    frame->device_ms = ms;
    frame->ray_count = *r->h_rays;
    frame->kernel_launches = launches;
    r->has_frame = true;
    r->last_scene = scene;
    r->last_camera = *camera;
    r->last_shard = sh;
    r->last_depth = params->max_depth;
    r->last_chains = chains;
#ifdef RT_GPU_COUNTERS
    {
        unsigned long long c[2];
        rt_counters_read(c, true);
        fprintf(stderr, "[rt counters] %s: %llu rays, %.3f node visits/ray, %.3f triangle tests/ray\n", r->kind == RT_MEGAKERNEL ? "megakernel" : "wavefront",
                (unsigned long long)frame->ray_count, (double)c[0] / (double)frame->ray_count, (double)c[1] / (double)frame->ray_count);
    }
#endif
    return RT_OK;
}

rt_status rt_resolve(rt_context *ctx, const float *accum, uint32_t sample_count, int32_t width, int32_t height,
                     uint8_t *rgba8) {
    if (!ctx) return RT_ERR_INVALID;
    if (!accum || !rgba8 || width <= 0 || height <= 0) return rt_set_error(ctx, RT_ERR_INVALID, "rt_resolve", "bad argument");
    RT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)width * (size_t)height;
    /* device-resident arguments are used in place; host ones go through the context's grow-only scratch (no
     * cudaMalloc / cudaFree / staging copy per call: after a cross-GPU reduction both usually are device memory) */
    const bool a_dev = is_device_pointer(accum), o_dev = is_device_pointer(rgba8);
    void *d_a = (void *)accum, *d_o = (void *)rgba8;
    cudaError_t e = cudaSuccess;
    do {
        if (!a_dev) {
            if ((e = rt_scratch(ctx, 0, n * 16, &d_a)) != cudaSuccess) break;
            if ((e = rt_upload(ctx, d_a, accum, n * 16)) != cudaSuccess) break;
        }
        if (!o_dev && (e = rt_scratch(ctx, 1, n * 4, &d_o)) != cudaSuccess) break;
        if ((e = rt_launch_resolve(st, (const float *)d_a, (uint32_t *)d_o, (uint32_t)n, (float)sample_count)) != cudaSuccess) break;
        if (!o_dev && (e = cudaMemcpyAsync(rgba8, d_o, n * 4, cudaMemcpyDefault, st)) != cudaSuccess) break;
        e = cudaStreamSynchronize(st);
    } while (0);
    if (e != cudaSuccess) return rt_set_error(ctx, RT_ERR_CUDA, "rt_resolve", cudaGetErrorString(e));
    return RT_OK;
}

} /* extern "C" */
