/*
 * rt_group.cu — several B200s of one node behind one handle (rt_group_* of include/rt_api.h).
 *
 * The reference is single-device (src/app.hpp:43-55 picks one GPU); the path shards with no exchange until the
 * end of a frame (pixels own their xorshift streams), so a group is N ordinary contexts driven by one host thread
 * each, plus the one real exchange step, done here over NVLink peer memory instead of a library collective:
 *
 *   image tiles : every device renders its tiles and its render kernel stores each finished RGBA8 pixel straight
 *                 into device 0's image (rt_renderer_set_gather: peer stores spread over the whole frame). The
 *                 result — image, accumulation, streams, ray count — is bit-identical to one device.
 *   spp slices  : every device renders all pixels with its share of the samples and its own seed salt; then ONE
 *                 kernel per device (k_reduce_resolve_peer) does reduce-scatter + resolve + gather: it sums its
 *                 1/N slice of the pixels over all N accumulation buffers (peer loads, fixed rank order, so the
 *                 sum is deterministic), and stores the sum and the resolved RGBA8 pixel into device 0's buffers
 *                 (peer stores). 33 MB of fp32 per device at 1080p cross NVLink exactly once.
 */
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "rt_internal.h"
#include "rt_render.h"

#define RT_GROUP_MAX 16

struct rt_group {
    std::vector<rt_context *> ctx;
    std::string err;
};
struct rt_group_scene {
    rt_group *g = nullptr;
    std::vector<rt_scene *> scene;
};
struct rt_group_renderer {
    rt_group *g = nullptr;
    rt_renderer_kind kind = RT_MEGAKERNEL;
    int32_t w = 0, h = 0;
    std::vector<rt_renderer *> r;
    float4 *d_sum = nullptr;      /* device 0: accumulation summed over the group (spp slices) / merged (tiles) */
    uint32_t *d_rng = nullptr;    /* device 0: merged final stream states (tiles) */
    bool gather_attached = false; /* devices 1.. store their tiles into device 0's image */
};

namespace {

struct PeerPtrs {
    const float4 *accum[RT_GROUP_MAX];
    const uint32_t *rng[RT_GROUP_MAX];
};

/* spp slices: this device owns pixels [first, first + count): sum over the group in rank order, resolve, and store
 * both into device 0's buffers (all pointers may be peer memory) */
__global__ void k_reduce_resolve_peer(PeerPtrs p, uint32_t n_dev, uint32_t first, uint32_t count, float4 *sum_out, uint32_t *rgba8_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t pix = first + i;
    float4 s = p.accum[0][pix];
    for (uint32_t d = 1; d < n_dev; d++) {
        const float4 a = p.accum[d][pix];
        s.x += a.x;
        s.y += a.y;
        s.z += a.z;
        s.w += a.w;
    }
    if (sum_out) sum_out[pix] = s;
    rgba8_out[pix] = rt_resolve_pixel(s.x, s.y, s.z, s.w);
}

/* image tiles: collect accumulation and stream state of every pixel from the device that owns it */
__global__ void k_merge_tiles_peer(PeerPtrs p, uint32_t n_dev, uint32_t tile_size, uint32_t w, uint32_t h, float4 *accum_out, uint32_t *rng_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const uint32_t x = i % w, y = i / w;
    const uint32_t tiles_x = (w + tile_size - 1) / tile_size;
    const uint32_t owner = ((y / tile_size) * tiles_x + x / tile_size) % n_dev; /* rt_owns_pixel */
    accum_out[i] = p.accum[owner][i];
    rng_out[i] = p.rng[owner][i];
}

rt_status group_error(rt_group *g, rt_status st, const std::string &msg) {
    g->err = msg;
    return st;
}

/* run fn(i) for every device of the group on its own host thread; returns the first failure */
template <class F>
rt_status for_each_device(rt_group *g, const char *what, F fn) {
    const size_t n = g->ctx.size();
    std::vector<rt_status> st(n, RT_OK);
    std::vector<std::thread> th;
    for (size_t i = 1; i < n; i++) th.emplace_back([&, i]() { st[i] = fn((uint32_t)i); });
    st[0] = fn(0u);
    for (auto &t : th) t.join();
    for (size_t i = 0; i < n; i++)
        if (st[i] != RT_OK) return group_error(g, st[i], std::string(what) + " (device " + std::to_string(i) + "): " + rt_last_error(g->ctx[i]));
    return RT_OK;
}

} // namespace

cudaError_t rt_launch_reduce_resolve_peer(cudaStream_t st, const float4 *const *accum, uint32_t world, uint32_t first, uint32_t count,
                                          float4 *sum_out, uint32_t *rgba8_out) {
    if (!count) return cudaSuccess;
    PeerPtrs pp = {};
    for (uint32_t i = 0; i < world && i < RT_GROUP_MAX; i++) pp.accum[i] = accum[i];
    k_reduce_resolve_peer<<<(count + 255) / 256, 256, 0, st>>>(pp, world, first, count, sum_out, rgba8_out);
    return cudaGetLastError();
}

extern "C" {

rt_status rt_group_create(const int *devices, uint32_t n_devices, rt_group **out) {
    if (!out) return RT_ERR_INVALID;
    *out = nullptr;
    if (n_devices == 0 || n_devices > RT_GROUP_MAX) return rt_set_error(nullptr, RT_ERR_INVALID, "rt_group_create", "1 to 16 devices");
    rt_group *g = new (std::nothrow) rt_group();
    if (!g) return RT_ERR_INVALID;
    for (uint32_t i = 0; i < n_devices; i++) {
        rt_context *c = nullptr;
        const rt_status st = rt_context_create(devices ? devices[i] : (int)i, &c);
        if (st != RT_OK) {
            rt_group_destroy(g);
            return st; /* rt_last_error(NULL) has the message */
        }
        g->ctx.push_back(c);
    }
    /* every device loads from / stores to every other device's buffers (NVLink / NVSwitch peer memory) */
    for (uint32_t i = 0; i < n_devices; i++)
        for (uint32_t j = 0; j < n_devices; j++) {
            if (i == j || g->ctx[i]->device == g->ctx[j]->device) continue;
            int can = 0;
            cudaSetDevice(g->ctx[i]->device);
            if (cudaDeviceCanAccessPeer(&can, g->ctx[i]->device, g->ctx[j]->device) != cudaSuccess || !can) {
                rt_group_destroy(g);
                return rt_set_error(nullptr, RT_ERR_STATE, "rt_group_create", "the devices cannot access each other's memory (no peer access)");
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(g->ctx[j]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                rt_group_destroy(g);
                return rt_set_error(nullptr, RT_ERR_CUDA, "cudaDeviceEnablePeerAccess", cudaGetErrorString(e));
            }
            (void)cudaGetLastError();
        }
    *out = g;
    return RT_OK;
}

void rt_group_destroy(rt_group *g) {
    if (!g) return;
    for (rt_context *c : g->ctx) rt_context_destroy(c);
    delete g;
}

uint32_t rt_group_size(const rt_group *g) { return g ? (uint32_t)g->ctx.size() : 0u; }
rt_context *rt_group_context(rt_group *g, uint32_t i) { return g && i < g->ctx.size() ? g->ctx[i] : nullptr; }
const char *rt_group_last_error(rt_group *g) { return g ? g->err.c_str() : rt_last_error(nullptr); }

rt_status rt_group_scene_create(rt_group *g, const rt_scene_desc *desc, rt_group_scene **out) {
    if (!g || !desc || !out) return RT_ERR_INVALID;
    *out = nullptr;
    rt_group_scene *s = new (std::nothrow) rt_group_scene();
    if (!s) return RT_ERR_INVALID;
    s->g = g;
    s->scene.assign(g->ctx.size(), nullptr);
    /* the scene is replicated (the path shards pixels and samples, not geometry): upload + GPU BVH build on every
     * device at the same time */
    const rt_status st = for_each_device(g, "rt_group_scene_create", [&](uint32_t i) {
        rt_status e = rt_scene_create(g->ctx[i], desc, &s->scene[i]);
        return e != RT_OK ? e : rt_scene_commit(s->scene[i]);
    });
    if (st != RT_OK) {
        rt_group_scene_destroy(s);
        return st;
    }
    *out = s;
    return RT_OK;
}

void rt_group_scene_destroy(rt_group_scene *s) {
    if (!s) return;
    for (rt_scene *x : s->scene) rt_scene_destroy(x);
    delete s;
}

rt_scene *rt_group_scene_get(rt_group_scene *s, uint32_t i) { return s && i < s->scene.size() ? s->scene[i] : nullptr; }

rt_status rt_group_renderer_create(rt_group *g, rt_renderer_kind kind, int32_t width, int32_t height, rt_group_renderer **out) {
    if (!g || !out) return RT_ERR_INVALID;
    *out = nullptr;
    rt_group_renderer *r = new (std::nothrow) rt_group_renderer();
    if (!r) return RT_ERR_INVALID;
    r->g = g;
    r->kind = kind;
    r->w = width;
    r->h = height;
    r->r.assign(g->ctx.size(), nullptr);
    rt_status st = for_each_device(g, "rt_group_renderer_create", [&](uint32_t i) { return rt_renderer_create(g->ctx[i], kind, width, height, &r->r[i]); });
    if (st == RT_OK && g->ctx.size() > 1) {
        cudaSetDevice(g->ctx[0]->device);
        const size_t n = (size_t)width * (size_t)height;
        cudaError_t e = cudaMalloc((void **)&r->d_sum, n * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc((void **)&r->d_rng, n * sizeof(uint32_t));
        if (e != cudaSuccess) st = group_error(g, RT_ERR_CUDA, std::string("rt_group_renderer_create: ") + cudaGetErrorString(e));
    }
    if (st != RT_OK) {
        rt_group_renderer_destroy(r);
        return st;
    }
    *out = r;
    return RT_OK;
}

void rt_group_renderer_destroy(rt_group_renderer *r) {
    if (!r) return;
    for (size_t i = 0; i < r->r.size(); i++) {
        if (r->r[i] && i > 0 && r->gather_attached) rt_renderer_set_gather(r->r[i], nullptr, nullptr);
        rt_renderer_destroy(r->r[i]);
    }
    if (!r->g->ctx.empty()) cudaSetDevice(r->g->ctx[0]->device);
    cudaFree(r->d_sum);
    cudaFree(r->d_rng);
    delete r;
}

rt_renderer *rt_group_renderer_get(rt_group_renderer *r, uint32_t i) { return r && i < r->r.size() ? r->r[i] : nullptr; }

rt_status rt_group_render_frame(rt_group_renderer *r, const rt_group_scene *scene, const rt_camera *camera, const rt_group_params *params,
                                rt_frame *frame) {
    if (!r) return RT_ERR_INVALID;
    rt_group *g = r->g;
    if (!scene || !camera || !params || !frame) return group_error(g, RT_ERR_INVALID, "rt_group_render_frame: NULL argument");
    if (scene->g != g) return group_error(g, RT_ERR_INVALID, "rt_group_render_frame: scene belongs to another group");
    const uint32_t n_dev = (uint32_t)g->ctx.size();
    const size_t n = (size_t)r->w * (size_t)r->h;
    if (n_dev == 1) { /* one device: the plain call */
        rt_render_params p = {};
        p.max_depth = params->max_depth;
        p.sample_count = params->sample_count;
        p.flags = params->flags;
        p.sample_chains = params->sample_chains;
        const rt_status st = rt_render_frame(r->r[0], scene->scene[0], camera, &p, frame);
        return st == RT_OK ? st : group_error(g, st, std::string("rt_render_frame: ") + rt_last_error(g->ctx[0]));
    }
    if (params->mode != RT_GROUP_TILES && params->mode != RT_GROUP_SPP) return group_error(g, RT_ERR_INVALID, "rt_group_render_frame: bad mode");
    const bool tiles = params->mode == RT_GROUP_TILES;
    const uint32_t tile_size = tiles ? (params->tile_size ? params->tile_size : 64u) : 0u;
    if (!tiles && params->sample_count < n_dev && params->sample_count != 0)
        return group_error(g, RT_ERR_INVALID, "rt_group_render_frame: spp slices need at least one sample per device");

    /* image tiles: devices 1.. store finished pixels into device 0's image while they render */
    if (tiles && !r->gather_attached) {
        uint8_t *dst = rt_renderer_device_rgba8(r->r[0]);
        for (uint32_t i = 1; i < n_dev; i++) {
            const rt_status st = rt_renderer_set_gather(r->r[i], nullptr, dst);
            if (st != RT_OK) return group_error(g, st, std::string("rt_renderer_set_gather: ") + rt_last_error(g->ctx[i]));
        }
        rt_renderer_mark_exported(r->r[0]); /* device 0 keeps the pixels it does not own */
        r->gather_attached = true;
    }

    std::vector<rt_frame> f(n_dev);
    const rt_status st = for_each_device(g, "rt_render_frame", [&](uint32_t i) {
        rt_render_params p = {};
        p.max_depth = params->max_depth;
        p.flags = params->flags;
        p.sample_chains = params->sample_chains;
        p.shard.rank = i;
        p.shard.world = n_dev;
        if (tiles) {
            p.sample_count = params->sample_count;
            p.shard.tile_size = tile_size;
        } else { /* spp / N each, the first spp % N devices take one more; salt 0 on device 0 is the reference stream */
            p.sample_count = params->sample_count / n_dev + (i < params->sample_count % n_dev ? 1u : 0u);
            p.shard.seed_salt = i * 0x9E3779B9u;
        }
        f[i] = rt_frame{};
        return rt_render_frame(r->r[i], scene->scene[i], camera, &p, &f[i]); /* synchronous: the device is idle on return */
    });
    if (st != RT_OK) return st;

    frame->ray_count = 0;
    frame->device_ms = 0.0f;
    frame->kernel_launches = 0;
    for (uint32_t i = 0; i < n_dev; i++) {
        frame->ray_count += f[i].ray_count;
        frame->device_ms = f[i].device_ms > frame->device_ms ? f[i].device_ms : frame->device_ms;
        frame->kernel_launches += f[i].kernel_launches;
    }

    PeerPtrs pp = {};
    for (uint32_t i = 0; i < n_dev; i++) {
        pp.accum[i] = (const float4 *)rt_renderer_device_accum(r->r[i]);
        pp.rng[i] = rt_renderer_device_rng(r->r[i]);
    }
    rt_context *c0 = g->ctx[0];
    cudaError_t e = cudaSuccess;
    float exchange_ms = 0.0f;
    if (!tiles) {
        /* reduce-scatter + resolve + gather in one kernel per device over peer memory */
        uint32_t *img0 = (uint32_t *)rt_renderer_device_rgba8(r->r[0]);
        const uint32_t per = (uint32_t)((n + n_dev - 1) / n_dev);
        for (uint32_t i = 0; i < n_dev && e == cudaSuccess; i++) {
            const uint32_t first = i * per, count = first < n ? (uint32_t)(n - first < per ? n - first : per) : 0u;
            if (!count) continue;
            e = cudaSetDevice(g->ctx[i]->device);
            if (e == cudaSuccess && i == 0) e = cudaEventRecord(c0->ev0, c0->stream);
            if (e == cudaSuccess) {
                k_reduce_resolve_peer<<<(count + 255) / 256, 256, 0, g->ctx[i]->stream>>>(pp, n_dev, first, count, r->d_sum, img0);
                e = cudaGetLastError();
            }
        }
        for (uint32_t i = 0; i < n_dev && e == cudaSuccess; i++) {
            e = cudaSetDevice(g->ctx[i]->device);
            if (e == cudaSuccess) e = cudaStreamSynchronize(g->ctx[i]->stream);
        }
        frame->kernel_launches += n_dev;
        if (e == cudaSuccess) e = cudaSetDevice(c0->device);
        if (e == cudaSuccess) e = cudaEventRecord(c0->ev1, c0->stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(c0->ev1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&exchange_ms, c0->ev0, c0->ev1);
    } else if (frame->accum || frame->rng_state) {
        e = cudaSetDevice(c0->device);
        if (e == cudaSuccess) {
            k_merge_tiles_peer<<<(unsigned)((n + 255) / 256), 256, 0, c0->stream>>>(pp, n_dev, tile_size, (uint32_t)r->w, (uint32_t)r->h, r->d_sum, r->d_rng);
            e = cudaGetLastError();
        }
        frame->kernel_launches += 1;
    }
    frame->device_ms += exchange_ms;
    if (e == cudaSuccess) e = cudaSetDevice(c0->device);
    if (e == cudaSuccess && frame->rgba8) e = cudaMemcpyAsync(frame->rgba8, rt_renderer_device_rgba8(r->r[0]), n * 4, cudaMemcpyDefault, c0->stream);
    if (e == cudaSuccess && frame->accum) e = cudaMemcpyAsync(frame->accum, r->d_sum, n * sizeof(float4), cudaMemcpyDefault, c0->stream);
    if (e == cudaSuccess && frame->rng_state)
        e = cudaMemcpyAsync(frame->rng_state, tiles ? r->d_rng : rt_renderer_device_rng(r->r[0]), n * 4, cudaMemcpyDefault, c0->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c0->stream);
    if (e != cudaSuccess) return group_error(g, RT_ERR_CUDA, std::string("rt_group_render_frame: ") + cudaGetErrorString(e));
    return RT_OK;
}

} /* extern "C" */
