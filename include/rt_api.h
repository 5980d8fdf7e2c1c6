/*
 * rt_api.h — C ABI of the B200-native path-tracing hot path (librt_b200.so).
 *
 * This is the drop-in boundary for felipeagc/sycl-ray-tracer's
 *   IRenderer::render_frame(const Camera&, const Scene&)        (src/render.hpp:11-18)
 * and for the pieces of App / Scene / Camera that feed it. Plain pointers and sizes only:
 * no C++ types, no torch types. The C++ mirror of the reference classes
 * (sycl-ray-tracer_b200/host/raytracer.hpp) and the Python ctypes binding are thin layers
 * over exactly these entry points. Each entry point names the reference interface it replaces.
 *
 * Conventions
 *  - every call returns rt_status; on failure rt_last_error(ctx) has the message
 *    (reference: exceptions ending in std::terminate, src/app.hpp:8-19, src/main.cpp:71-74).
 *  - "any-space pointer": output/input pointers marked ANY may be host (pageable or pinned) or
 *    device memory; the library copies with cudaMemcpyDefault on the context stream.
 *  - all render calls are synchronous on return, like the reference's submit + wait
 *    (src/render_megakernel.cpp:171).
 *  - there is no CPU fallback: without a CUDA device rt_context_create fails.
 */
#ifndef RT_API_H
#define RT_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_API_VERSION 2
#define RT_TEX_SIZE 512   /* src/image_manager.hpp:14  IMAGE_SIZE  */
#define RT_MAX_IMAGES 128 /* src/image_manager.hpp:12  MAX_IMAGES  */

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID = 1,  /* bad argument */
    RT_ERR_CUDA = 2,     /* CUDA runtime error (message in rt_last_error) */
    RT_ERR_STATE = 3,    /* call order violated (e.g. render before commit) */
    RT_ERR_NO_DEVICE = 4 /* no usable CUDA device: there is no CPU path */
} rt_status;

/* src/material.hpp:61-66 MaterialType */
typedef enum rt_material_type {
    RT_MAT_NONE = 0,
    RT_MAT_DIFFUSE = 1,
    RT_MAT_METALLIC = 2,
    RT_MAT_DIELECTRIC = 3
} rt_material_type;

/* src/main.cpp:57-68: the -m / -w switch */
typedef enum rt_renderer_kind { RT_MEGAKERNEL = 0, RT_WAVEFRONT = 1 } rt_renderer_kind;

/* src/material.hpp:17-54 (Texture) + :68-238 (Material), flattened to a POD */
typedef struct rt_material {
    int32_t type;          /* rt_material_type */
    int32_t albedo_image;  /* >= 0: ImageRef.index (src/image_manager.hpp:27-29); < 0: albedo_color */
    float albedo_color[3];
    float roughness;       /* MaterialMetallic::roughness */
    float ior;             /* MaterialDielectric::ior */
    float emissive[3];
} rt_material;

/* One Embree instance = one glTF node x mesh primitive with its GeometryData
 * (src/scene.hpp:17-24, src/scene.cpp:483-507). Instance ids are array positions, i.e. the
 * reference's attach order (src/scene.cpp:101-106). All pointers are HOST memory and are
 * copied during rt_scene_create. */
typedef struct rt_instance {
    const float *positions;  /* 3 * vertex_count   (glm::vec3, src/scene.cpp:286-293) */
    const float *normals;    /* 3 * vertex_count   (src/scene.cpp:311-318) */
    const float *uvs;        /* 2 * vertex_count   (src/scene.cpp:337-343) */
    const uint32_t *indices; /* index_count, u32   (src/scene.cpp:359-401) */
    uint32_t vertex_count;
    uint32_t index_count;    /* multiple of 3 */
    float transform[16];     /* column-major 4x4, node_global_matrix (src/scene.cpp:137-146,491-494) */
    rt_material material;
} rt_instance;

/* What Scene hands to the kernels (src/scene.hpp:64-91): instances, sky colour, image array */
typedef struct rt_scene_desc {
    const rt_instance *instances;
    uint32_t instance_count;
    const uint8_t *texture_layers; /* layer_count * 512 * 512 * 4 RGBA8 (baked array,
                                      src/image_manager.hpp:76-100); may be NULL */
    uint32_t texture_layer_count;  /* <= RT_MAX_IMAGES */
    float sky_color[3];            /* src/scene.hpp:76 default (0.5, 0.7, 1.0) */
} rt_scene_desc;

/* src/camera.hpp:65-72, by value into the kernels (src/render_megakernel.cpp:95-96) */
typedef struct rt_camera {
    float center[3];
    float pixel00_loc[3];
    float pixel_delta_u[3];
    float pixel_delta_v[3];
    int32_t img_size[2];
} rt_camera;

/* Multi-GPU sharding of one frame (no reference equivalent: it is single-device). */
typedef struct rt_shard {
    uint32_t rank;       /* this context's shard */
    uint32_t world;      /* number of shards; 0 or 1 = unsharded */
    uint32_t tile_size;  /* > 0: image-tile sharding (a multiple of 8, <= 16384). Tiles of tile_size^2 pixels, numbered row
                            major, tile t belongs to rank t % world. Pixels of other ranks are left
                            untouched (zero) in accum/rgba8, so a SUM over ranks is the full image,
                            bit-identical to the unsharded render. */
    uint32_t seed_salt;  /* XORed into every pixel seed. spp sharding: all ranks render all pixels
                            with sample_count = spp / world and distinct salts (salt 0 on rank 0
                            reproduces the reference stream). */
} rt_shard;

/* rt_render_params.flags */
#define RT_RENDER_RESUME 1u /* progressive rendering: continue the previous rt_render_frame of this
                              renderer (same scene, camera, shard) — per-pixel xorshift streams and the
                              accumulation carry on, so k frames of n samples are bit-identical to one
                              frame of k*n samples (BASELINE config 5: 4096 spp in batches) */

/* The two items the reference's roadmap leaves open (PLAN.md:23-27), as OPTIONS: both change the sampling order, so
 * images are no longer the reference's bit for bit (the oracle implements the same options, tests/test_gpu_options.py). */
#define RT_RENDER_ROULETTE 2u /* Russian roulette: from the third bounce on a path survives with probability
                                q = clamp(max(attenuation), 0.05, 1) (one extra draw) and its attenuation is divided by q */

typedef struct rt_render_params {
    uint32_t max_depth;    /* -d, src/main.cpp:11 */
    uint32_t sample_count; /* -s, src/main.cpp:13; with RT_RENDER_RESUME: samples ADDED by this call */
    rt_shard shard;
    uint32_t flags;        /* RT_RENDER_RESUME | RT_RENDER_ROULETTE */
    uint32_t sample_chains; /* 0 or 1: one xorshift stream per pixel, samples strictly in sequence (the reference, F4).
                               k = 2..16 ("splats", PLAN.md:26-27): k independent sample chains per pixel are in flight at
                               once — chain c takes sample_count / k samples (the first sample_count % k chains one more)
                               on the stream seed ^ c * 0x9E3779B9 and accumulates into its own plane; the planes are summed
                               in chain order, so the result is deterministic and equals the sum of k seed-salted renders.
                               Shortens a pixel's sequential chain k-fold (the tail of a frame with few pixels per GPU). */
} rt_render_params;

/* Result of one render_frame. Every pointer is optional (NULL = not wanted) and ANY-space. */
typedef struct rt_frame {
    uint8_t *rgba8;   /* W*H*4: what the reference writes to out.png (mean, sqrt gamma, unorm8
                         store + *255 truncating read-back; src/util.hpp:16-22) */
    float *accum;     /* W*H*4 fp32: linear SUM over samples in rgb, sample count in a
                         (wavefront: sum of per-sample clamped values, src/render_wavefront.cpp:277) */
    uint32_t *rng_state; /* W*H: final xorshift32 state per pixel (bit-exact stream check) */
    uint64_t ray_count;  /* out: ray segments traced = rtcIntersect1 calls
                            (src/render_megakernel.cpp:32, src/render_wavefront.cpp:407) */
    float device_ms;     /* out: CUDA-event time of the render on the context stream */
    uint32_t kernel_launches; /* out: kernels launched by this call */
} rt_frame;

typedef struct rt_scene_stats {
    uint64_t triangle_count;
    uint64_t node_count;      /* 80-byte wide-BVH nodes */
    uint64_t bvh_bytes;       /* nodes + leaf-ordered triangle records */
    uint64_t shading_bytes;   /* per-triangle shading records + instance table */
    float build_ms;           /* device time of rt_scene_commit */
    uint32_t max_leaf_tris;
    uint32_t wide_depth;      /* levels of the wide tree */
} rt_scene_stats;

typedef struct rt_context rt_context;   /* App (src/app.hpp:31-58): device + stream owner */
typedef struct rt_scene rt_scene;       /* Scene (src/scene.hpp:64-104) */
typedef struct rt_renderer rt_renderer; /* MegakernelRenderer / WavefrontRenderer */

/* ---- context: replaces raytracer::App (src/app.hpp:43-55) -------------------------------- */
rt_status rt_context_create(int device, rt_context **out);
void rt_context_destroy(rt_context *ctx);
/* the cudaStream_t all work is enqueued on (replaces App::queue, src/app.hpp:33) */
void *rt_context_stream(rt_context *ctx);
/* run on a caller-owned cudaStream_t instead (e.g. the stream an NCCL collective is issued on,
 * so that render + all-reduce are ordered and timed on one stream). The stream is borrowed. */
rt_status rt_context_set_stream(rt_context *ctx, void *cuda_stream);
const char *rt_context_device_name(rt_context *ctx);
/* last error message of this context (ctx may be NULL for creation failures) */
const char *rt_last_error(rt_context *ctx);
uint32_t rt_api_version(void);

/* ---- scene: replaces Scene's device side (src/scene.cpp:101-107,406-439,483-507) --------- */
/* uploads instances (geometry, GeometryData) and the baked texture array */
rt_status rt_scene_create(rt_context *ctx, const rt_scene_desc *desc, rt_scene **out);
/* the rtcCommitScene analogue (src/scene.cpp:107): GPU flatten -> Morton codes -> radix sort
 * -> LBVH -> 8-wide compressed BVH. The scene is immutable afterwards. */
rt_status rt_scene_commit(rt_scene *scene);
rt_status rt_scene_get_stats(const rt_scene *scene, rt_scene_stats *out);
void rt_scene_destroy(rt_scene *scene);

/* ---- camera: replaces Camera::Camera (src/camera.hpp:74-106); pure host arithmetic ------- */
void rt_camera_init(rt_camera *cam, int32_t width, int32_t height, const float position[3],
                    const float direction[3], float focal_length);

/* ---- batch closest hit: replaces rtcIntersect1 (src/trace_ray.hpp:18-22) ------------------
 * n rays; org/dir are 3 floats per ray (ANY-space); outputs ANY-space, inst/prim = -1 on miss,
 * u,v in Embree's convention (weights of vertex 1 and 2), t in units of |dir|.
 * Hit accepted iff tnear < t <= tfar; ties resolve to the lowest (inst, prim). */
rt_status rt_intersect(rt_context *ctx, const rt_scene *scene, uint64_t n, const float *org,
                       const float *dir, float tnear, float tfar, int32_t *inst, int32_t *prim,
                       float *u, float *v, float *t, float *device_ms);

/* ---- renderers: replace MegakernelRenderer / WavefrontRenderer ----------------------------
 * ctor shape (App&, img_size, image&, max_depth, sample_count): src/render_megakernel.hpp:13-19,
 * src/render_wavefront.hpp:55-61. The wavefront renderer owns its ray queues and rng buffer
 * (src/render_wavefront.hpp:18-37,52). */
rt_status rt_renderer_create(rt_context *ctx, rt_renderer_kind kind, int32_t width, int32_t height,
                             rt_renderer **out);
void rt_renderer_destroy(rt_renderer *r);
/* IRenderer::render_frame (src/render.hpp:12-15). */
rt_status rt_render_frame(rt_renderer *r, const rt_scene *scene, const rt_camera *camera,
                          const rt_render_params *params, rt_frame *frame);
/* device-resident results of the last rt_render_frame (valid until the next call / destroy):
 * fp32 RGBA accumulation (W*H*4 floats) and RGBA8 image. For zero-copy hand-off to NCCL. */
float *rt_renderer_device_accum(rt_renderer *r);
uint8_t *rt_renderer_device_rgba8(rt_renderer *r);
uint32_t *rt_renderer_device_rng(rt_renderer *r); /* W*H final xorshift32 states */
/* ---- image-tile shards: gather over peer memory (no collective) --------------------------
 * Under image-tile sharding every pixel is finished by exactly one rank, so its final RGBA8 value
 * can be stored straight into ONE destination image by the render kernel itself (4-byte stores
 * over NVLink / NVSwitch peer memory, overlapped with the rendering) instead of all-reducing the
 * fp32 accumulation buffers afterwards:
 *   destination rank : rt_renderer_export_image(r, &h)   -> send h to the other ranks (64 opaque bytes)
 *   every other rank : rt_renderer_set_gather(r, &h, NULL)   (another process: CUDA IPC handle)
 *                  or  rt_renderer_set_gather(r, NULL, ptr)  (same process: a device pointer this
 *                                                             device can store to, W*H*4 bytes)
 * From then on each tile-sharded rt_render_frame of r also stores its owned pixels to the target;
 * the destination renderer stops clearing the pixels it does not own. The caller orders the
 * frames: all ranks must have returned from rt_render_frame (+ a barrier) before the destination
 * image (rt_renderer_device_rgba8 of the exporting renderer) is read, and it must not be read
 * while peers render the next frame. rt_renderer_set_gather(r, NULL, NULL) detaches. */
typedef struct rt_ipc_handle {
    uint8_t bytes[64];
} rt_ipc_handle;
rt_status rt_renderer_export_image(rt_renderer *r, rt_ipc_handle *out);
rt_status rt_renderer_set_gather(rt_renderer *r, const rt_ipc_handle *handle, void *device_rgba8);

/* ---- spp slices across PROCESSES: fused reduce-scatter + resolve + gather over peer memory -----------------
 * One process per GPU (torchrun, MPI): every rank exports its accumulation buffer, collects all ranks' handles and
 * attaches them; rank 0 also exports its image and the others attach it with rt_renderer_set_gather. After every
 * rank has finished its frame (a barrier), rt_renderer_reduce_resolve ENQUEUES one kernel on the context stream
 * that sums this rank's 1/world slice of the pixels over all ranks' buffers (peer loads, rank order), resolves it
 * with the total sample count and stores the RGBA8 pixels into rank 0's image (peer stores); a second barrier
 * tells rank 0 that the image is complete. No library collective touches the 16 bytes per pixel.
 *   rt_renderer_export_accum(r, &h)                      -> all-gather h
 *   rt_renderer_set_peers(r, handles, world, rank)       (handles[rank] is ignored; NULL / 0 detaches)
 *   rt_renderer_reduce_resolve(r)                        (asynchronous on rt_context_stream) */
rt_status rt_renderer_export_accum(rt_renderer *r, rt_ipc_handle *out);
rt_status rt_renderer_set_peers(rt_renderer *r, const rt_ipc_handle *accum_handles, uint32_t world, uint32_t rank);
rt_status rt_renderer_reduce_resolve(rt_renderer *r);

/* (re)compute the RGBA8 image from an accumulation buffer holding the sum over `sample_count`
 * samples — used after a cross-GPU reduction (src/render_wavefront.cpp:360-394 + F10).
 * accum, rgba8: ANY-space. */
rt_status rt_resolve(rt_context *ctx, const float *accum, uint32_t sample_count, int32_t width,
                     int32_t height, uint8_t *rgba8);

/* ---- several GPUs of one node behind one handle ---------------------------------------------
 * No reference equivalent (App picks ONE device, src/app.hpp:43-55); this is what makes the reference's call
 * sequence (src/main.cpp:57-70: scene, renderer, render_frame) run on N B200s from one process. The scene is
 * replicated; a frame is sharded by image tiles or by samples and the one exchange step runs over NVLink peer
 * memory inside the library's own kernels (csrc/rt_group.cu):
 *   RT_GROUP_TILES  tile t of tile_size^2 pixels -> device t % N; finished RGBA8 pixels are stored by the render
 *                   kernel straight into device 0's image. Image, accumulation, final stream states and ray
 *                   count are BIT-IDENTICAL to one device rendering the frame.
 *   RT_GROUP_SPP    device i renders every pixel with sample_count / N samples (the first sample_count % N devices
 *                   one more) and seed salt i * 0x9E3779B9 (device 0 = the reference stream); one kernel per device
 *                   then sums its 1/N slice of the pixels over all accumulation buffers in device order, resolves
 *                   and stores into device 0's buffers (reduce-scatter + resolve + gather, fused). accum = that sum.
 * All devices must be able to access each other's memory (NVLink / NVSwitch). Calls are synchronous. */
typedef enum rt_group_mode { RT_GROUP_TILES = 0, RT_GROUP_SPP = 1 } rt_group_mode;
typedef struct rt_group rt_group;
typedef struct rt_group_scene rt_group_scene;
typedef struct rt_group_renderer rt_group_renderer;
typedef struct rt_group_params {
    uint32_t max_depth;    /* -d */
    uint32_t sample_count; /* -s: samples per pixel of the WHOLE frame (with RT_RENDER_RESUME: added by this call) */
    uint32_t mode;         /* rt_group_mode */
    uint32_t tile_size;    /* RT_GROUP_TILES: a multiple of 8; 0 = 64 */
    uint32_t flags;        /* RT_RENDER_RESUME | RT_RENDER_ROULETTE */
    uint32_t sample_chains; /* as in rt_render_params */
} rt_group_params;

/* devices = CUDA device indices (NULL: 0 .. n_devices-1), 1 <= n_devices <= 16 */
rt_status rt_group_create(const int *devices, uint32_t n_devices, rt_group **out);
void rt_group_destroy(rt_group *g);
uint32_t rt_group_size(const rt_group *g);
rt_context *rt_group_context(rt_group *g, uint32_t i); /* the i-th device's context (owned by the group) */
const char *rt_group_last_error(rt_group *g);
/* rt_scene_create + rt_scene_commit on every device, concurrently */
rt_status rt_group_scene_create(rt_group *g, const rt_scene_desc *desc, rt_group_scene **out);
void rt_group_scene_destroy(rt_group_scene *s);
rt_scene *rt_group_scene_get(rt_group_scene *s, uint32_t i);
rt_status rt_group_renderer_create(rt_group *g, rt_renderer_kind kind, int32_t width, int32_t height,
                                   rt_group_renderer **out);
void rt_group_renderer_destroy(rt_group_renderer *r);
rt_renderer *rt_group_renderer_get(rt_group_renderer *r, uint32_t i);
/* IRenderer::render_frame on the whole group. frame->ray_count is the sum over devices, device_ms the slowest
 * device's render time plus the exchange, kernel_launches the sum; the pointers of `frame` are ANY-space. */
rt_status rt_group_render_frame(rt_group_renderer *r, const rt_group_scene *scene, const rt_camera *camera,
                                const rt_group_params *params, rt_frame *frame);

#ifdef __cplusplus
}
#endif
#endif /* RT_API_H */
