/*
 * rt_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A from-scratch CPU restatement of the per-pixel path-tracing hot path of
 * felipeagc/sycl-ray-tracer, used ONLY as the checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference).
 * Nothing under sycl-ray-tracer_b200/ includes, links or loads this.
 *
 * Parity status: the reference ships no tests, golden images or KATs, and its
 * own sources cannot be compiled as-is here (icpx/SYCL, Embree 4 and glm are
 * absent). The shading half of this restatement IS pinned against the
 * reference's own, unmodified headers compiled through API shims
 * (oracle/refshim -> oracle/_ref/, see oracle/Makefile and tests/golden/).
 * The intersection half (Embree's rtcIntersect1, third-party, unpinned version
 * "embree 4", CMakeLists.txt:13) has no source to follow: it is restated as a
 * brute-force closest-hit search with the Woop/Benthin/Wald watertight test,
 * honouring Embree's documented conventions (u,v weights of v1,v2;
 * tnear < t <= tfar; double sided; instID/primID); its closest-hit answer is
 * checked against exact rational arithmetic (tests/test_oracle_bvh.py). The
 * reference's whole program (loader + renderers + write_image, compiled in
 * place: oracle/_ref/libfullref.so) renders .glb files to the bytes and ray
 * counts this oracle produces from the product's loader output
 * (tests/test_glb_loader.py).
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_TEX_SIZE 512 /* src/image_manager.hpp:14 IMAGE_SIZE */

enum { ORC_MAT_NONE = 0, ORC_MAT_DIFFUSE = 1, ORC_MAT_METALLIC = 2, ORC_MAT_DIELECTRIC = 3 };
enum { ORC_MODE_MEGAKERNEL = 0, ORC_MODE_WAVEFRONT = 1 };

/* src/material.hpp:17-238 (Texture + Material tagged unions flattened) */
typedef struct orc_material {
    int32_t type;          /* ORC_MAT_* */
    int32_t albedo_image;  /* >=0: texture layer (TextureType::eImage); <0: use albedo_color */
    float albedo_color[3];
    float roughness;       /* metallic */
    float ior;             /* dielectric */
    float emissive[3];
} orc_material;

/* one Embree instance = one glTF node x primitive (src/scene.cpp:483-507) */
typedef struct orc_instance {
    const float *positions; /* 3 * vertex_count, object space */
    const float *normals;   /* 3 * vertex_count */
    const float *uvs;       /* 2 * vertex_count */
    const uint32_t *indices;/* index_count (multiple of 3) */
    uint32_t vertex_count;
    uint32_t index_count;
    float transform[16];    /* column-major 4x4 global transform (src/scene.cpp:491-494) */
    orc_material material;
} orc_instance;

/* src/camera.hpp:65-72 */
typedef struct orc_camera {
    float center[3];
    float pixel00_loc[3];
    float pixel_delta_u[3];
    float pixel_delta_v[3];
    int32_t img_size[2];
} orc_camera;

typedef struct orc_scene orc_scene;

typedef struct orc_render_params {
    int32_t mode;        /* ORC_MODE_*: selects seed mapping (F3) and per-sample clamp (F9) */
    uint32_t max_depth;
    uint32_t sample_count;
    uint32_t seed_salt;  /* XORed into every pixel seed (0 = reference behaviour) */
    int32_t use_bvh;     /* 0: brute-force intersector (ground truth); 1: CPU SAH BVH */
    int32_t threads;     /* OpenMP threads; <=0: all */
    /* crop window inside the full image (seeds always use the full image size) */
    int32_t x0, y0, x1, y1;
    int32_t roulette;    /* 0 = the reference's path loop; 1 = Russian roulette from the third bounce on (PLAN.md:23-24, the
                            RT_RENDER_ROULETTE option of the CUDA path): changes the number of draws, hence the streams */
} orc_render_params;

/* ---- primitives (for KATs) ---- */
/* src/xorshift.hpp:11-20: advances *state, returns the float draw */
float orc_xorshift_next(uint32_t *state);
/* src/xorshift.hpp:38-40 */
void orc_random_unit_vector(uint32_t *state, float out[3]);
/* src/camera.hpp:74-106 */
void orc_camera_init(orc_camera *cam, int32_t width, int32_t height, const float pos[3],
                     const float dir[3], float focal_length);
/* src/camera.hpp:109-131 + RayData ctor :30-45 (direction rounded through fp16) */
void orc_camera_get_ray(const orc_camera *cam, int32_t x, int32_t y, uint32_t *rng_state,
                        float org[3], float dir[3]);
/* seed for pixel (x,y): src/render_megakernel.cpp:144-146 / src/render_wavefront.cpp:69-73 */
uint32_t orc_pixel_seed(int32_t mode, int32_t x, int32_t y, int32_t width, int32_t height);
/* fp16 RNE round trip used between bounces (src/camera.hpp:18-28) */
float orc_round_half(float v);
/* F10 byte: unorm8 image write (sat, rte) then read-back *255 truncation (src/util.hpp:16-22) */
uint8_t orc_output_byte(float gamma_value);
/* src/material.hpp:68-238; returns 1 if scattered. textures = n_layers*512*512*4 RGBA8 or NULL */
int orc_material_scatter(const orc_material *m, const uint8_t *textures, uint32_t n_layers,
                         uint32_t *rng_state, const float dir[3], const float normal[3],
                         const float uv[2], float out_dir[3], float out_att[3]);
/* src/material.hpp:45-53 with the sampler of src/render_megakernel.cpp:99-103 */
void orc_texture_sample(const uint8_t *textures, uint32_t n_layers, int32_t layer,
                        const float uv[2], float out_rgb[3]);
/* transpose(inverse(mat3(T))) as glm computes it (src/scene.cpp:502) */
void orc_normal_matrix(const float transform[16], float out9_colmajor[9]);

/* ---- scene ---- */
orc_scene *orc_scene_create(const orc_instance *instances, uint32_t n_instances,
                            const uint8_t *textures, uint32_t n_layers, const float sky_color[3]);
void orc_scene_destroy(orc_scene *s);
uint64_t orc_scene_triangle_count(const orc_scene *s);
/* world-space triangle soup in (instance, primitive) order: 9 floats per triangle */
const float *orc_scene_world_triangles(const orc_scene *s);

/* closest hit for n rays (the rtcIntersect1 stand-in, src/trace_ray.hpp:18-22).
 * inst/prim = -1 on miss. */
void orc_intersect(const orc_scene *s, int32_t use_bvh, int32_t threads, uint64_t n,
                   const float *org, const float *dir, float tnear, float tfar,
                   int32_t *inst, int32_t *prim, float *u, float *v, float *t);

/* full render of the crop window. Outputs are crop-sized, row-major (y-major):
 *  accum   : 4 floats per pixel, linear SUM over samples (rgb), a = sample_count
 *  rgba8   : 4 bytes per pixel after F10
 *  rng_out : final xorshift state per pixel
 * any output may be NULL. Returns the number of ray segments (rtcIntersect1 calls). */
uint64_t orc_render(const orc_scene *s, const orc_camera *cam, const orc_render_params *p,
                    float *accum, uint8_t *rgba8, uint32_t *rng_out);

/* seconds spent inside the last orc_render's render loop (steady_clock) */
double orc_last_render_seconds(void);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
