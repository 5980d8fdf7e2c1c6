/*
 * refshim.cpp — TEST-ONLY: compiles the reference's OWN, UNMODIFIED sources from where they lie under
 * /root/reference (nothing is copied into this repo) against the API shims in this directory and
 * exposes them through a small C interface, so that the oracle's restatement can be pinned against
 * the reference itself running here:
 *
 *   src/xorshift.hpp, camera.hpp, util.hpp, material.hpp, image_manager.hpp, render_context.hpp,
 *   trace_ray.hpp, render_megakernel.cpp (MegakernelRenderer::render_frame, K1),
 *   render_wavefront.cpp (WavefrontRenderer: K2..K6 and the host loop)
 *
 * What is NOT the reference: Embree (rtcIntersect1 -> the oracle's brute-force closest hit with the
 * watertight test, Embree's hit conventions), the SYCL runtime (serial execution, fibers for
 * barriers, OpenCL's texel / unorm8 rules) and glm (component-wise definitions). The glTF loader
 * (scene.cpp) is not compiled; the scene tables the kernels read (GeometryData per instance, sky
 * colour, image array) are filled here from the same arrays the oracle gets.
 *
 * Built only where /root/reference exists (oracle/Makefile `ref`) into oracle/_ref/librefshim.so;
 * tests/tools/make_golden.py turns its outputs into the committed fixtures under tests/golden/.
 */
#define FMT_HEADER_ONLY 1
#include <cstdio>
#include <string>
#include <vector>

#include "../rt_oracle.h"

/* the reference's sources, verbatim from /root/reference/src (include path set by the Makefile) */
#include "render_megakernel.cpp"
#include "render_wavefront.cpp"

using namespace raytracer;

/* ---- pieces of the reference that live in files we do not compile ------------------------------ */
raytracer::Scene::Scene(App &, const std::string &, glm::vec3 gs) : global_scale(gs) {} /* scene.cpp:54 (loader) */
raytracer::Scene::~Scene() {}
raytracer::Node::~Node() {}

/* out.png capture: src/util.hpp:27 hands the final bytes to stbi_write_png */
static std::vector<uint8_t> g_png_bytes;
extern "C" int stbi_write_png(char const *, int w, int h, int comp, const void *data, int stride) {
    g_png_bytes.resize((size_t)w * h * comp);
    for (int y = 0; y < h; y++) memcpy(&g_png_bytes[(size_t)y * w * comp], (const uint8_t *)data + (size_t)y * stride, (size_t)w * comp);
    return 1;
}

/* ---- Embree stand-in -------------------------------------------------------------------------- */
struct RefScene {
    orc_scene *orc = nullptr;
    int use_bvh = 0;
    std::vector<GeometryData *> user_data; /* per instance, what rtcSetGeometryUserData got (scene.cpp:497-505) */
    std::vector<std::vector<float>> normals, uvs;
    std::vector<std::vector<uint32_t>> indices;
    std::vector<uint8_t> tex;
    uint32_t n_layers = 0;
    float sky[3];
};
static RefScene *g_scene = nullptr;
static unsigned long long g_intersect_calls = 0;

void rtcIntersect1(RTCScene scene, RTCRayHit *rh) {
    RefScene *s = (RefScene *)scene;
    g_intersect_calls++;
    int32_t inst = -1, prim = -1;
    float u = 0, v = 0, t = 0;
    const float org[3] = {rh->ray.org_x, rh->ray.org_y, rh->ray.org_z}, dir[3] = {rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z};
    orc_intersect(s->orc, s->use_bvh, 1, 1, org, dir, rh->ray.tnear, rh->ray.tfar, &inst, &prim, &u, &v, &t);
    if (inst >= 0) {
        rh->ray.tfar = t;
        rh->hit.u = u;
        rh->hit.v = v;
        rh->hit.primID = (unsigned)prim;
        rh->hit.geomID = 0;
        rh->hit.instID[0] = (unsigned)inst;
    }
}
void *rtcGetGeometryUserDataFromScene(RTCScene scene, unsigned int id) { return ((RefScene *)scene)->user_data[id]; }

static Material to_material(const orc_material &m) {
    Texture tex = m.albedo_image >= 0 ? Texture(ImageRef{(uint32_t)m.albedo_image})
                                      : Texture(sycl::float3(m.albedo_color[0], m.albedo_color[1], m.albedo_color[2]));
    const sycl::float3 em(m.emissive[0], m.emissive[1], m.emissive[2]);
    switch (m.type) {
    case ORC_MAT_DIFFUSE: return Material(MaterialDiffuse{tex, em});
    case ORC_MAT_METALLIC: return Material(MaterialMetallic{tex, m.roughness, em});
    case ORC_MAT_DIELECTRIC: return Material(MaterialDielectric{m.ior});
    default: return Material();
    }
}

extern "C" {

void *ref_scene_create(const orc_instance *inst, uint32_t n, const uint8_t *textures, uint32_t n_layers,
                       const float sky[3], int use_bvh) {
    RefScene *s = new RefScene();
    s->orc = orc_scene_create(inst, n, textures, n_layers, sky);
    s->use_bvh = use_bvh;
    memcpy(s->sky, sky, 12);
    s->n_layers = n_layers;
    if (n_layers) s->tex.assign(textures, textures + (size_t)n_layers * 512 * 512 * 4);
    s->normals.resize(n);
    s->uvs.resize(n);
    s->indices.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        s->normals[i].assign(inst[i].normals, inst[i].normals + (size_t)inst[i].vertex_count * 3);
        s->uvs[i].assign(inst[i].uvs, inst[i].uvs + (size_t)inst[i].vertex_count * 2);
        s->indices[i].assign(inst[i].indices, inst[i].indices + inst[i].index_count);
        float nm[9];
        orc_normal_matrix(inst[i].transform, nm); /* glm::transpose(glm::inverse(glm::mat3(T))), scene.cpp:502 */
        glm::mat3 m;
        for (int c = 0; c < 3; c++) m[c] = glm::vec3(nm[c * 3], nm[c * 3 + 1], nm[c * 3 + 2]);
        GeometryData *g = (GeometryData *)malloc(sizeof(GeometryData));
        g->vertex_buffer = nullptr;
        g->normal_buffer = (glm::vec3 *)s->normals[i].data();
        g->uv_buffer = (sycl::float2 *)s->uvs[i].data();
        g->index_buffer = s->indices[i].data();
        g->obj_to_world = m;
        new (&g->material) Material(to_material(inst[i].material));
        s->user_data.push_back(g);
    }
    return s;
}
void ref_scene_destroy(void *p) {
    RefScene *s = (RefScene *)p;
    for (auto g : s->user_data) free(g);
    orc_scene_destroy(s->orc);
    delete s;
}

/* src/xorshift.hpp */
float ref_xorshift_next(uint32_t *state) {
    XorShift32State r{*state};
    float f = r();
    *state = r.a;
    return f;
}
void ref_random_unit_vector(uint32_t *state, float out[3]) {
    XorShift32State r{*state};
    sycl::float3 v = r.random_unit_vector();
    *state = r.a;
    for (int k = 0; k < 3; k++) out[k] = v[k];
}
/* src/camera.hpp: out = center, pixel00_loc, pixel_delta_u, pixel_delta_v */
void ref_camera(int w, int h, const float pos[3], const float dir[3], float focal, float out12[12]) {
    Camera c(sycl::range<2>(w, h), glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(dir[0], dir[1], dir[2]), focal);
    for (int k = 0; k < 3; k++) {
        out12[k] = c.center[k];
        out12[3 + k] = c.pixel00_loc[k];
        out12[6 + k] = c.pixel_delta_u[k];
        out12[9 + k] = c.pixel_delta_v[k];
    }
}
void ref_camera_get_ray(int w, int h, const float pos[3], const float dir[3], float focal, int x, int y,
                        uint32_t *state, float org[3], float d[3]) {
    Camera c(sycl::range<2>(w, h), glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(dir[0], dir[1], dir[2]), focal);
    XorShift32State r{*state};
    RayData rd = c.get_ray(sycl::int2(x, y), r);
    *state = r.a;
    org[0] = rd.org_x; org[1] = rd.org_y; org[2] = rd.org_z;
    d[0] = rd.dir_x; d[1] = rd.dir_y; d[2] = rd.dir_z;
}

static RenderContext make_ctx(RefScene *s, const Camera &cam, sycl::image<3> &img, sycl::handler &cgh) {
    return RenderContext{cam, sycl::float3(s->sky[0], s->sky[1], s->sky[2]), (RTCScene)s,
                         sycl::sampler(sycl::coordinate_normalization_mode::normalized, sycl::addressing_mode::repeat,
                                       sycl::filtering_mode::nearest),
                         ImageReadAccessor(img, cgh)};
}

/* src/material.hpp Material::scatter / emitted on its own */
int ref_material_scatter(void *scene, const orc_material *m, uint32_t *state, const float dir[3], const float normal[3],
                         const float uv[2], float out_dir[3], float out_att[3]) {
    RefScene *s = (RefScene *)scene;
    static uint8_t dummy[4];
    sycl::image<3> img(s->n_layers ? (void *)s->tex.data() : (void *)dummy, sycl::image_channel_order::rgba,
                       sycl::image_channel_type::unorm_int8, sycl::range<3>(512, 512, s->n_layers ? s->n_layers : 1));
    sycl::handler cgh;
    Camera cam(sycl::range<2>(8, 8), glm::vec3(0, 0, 0), glm::vec3(0, 0, -1), 1.0f);
    RenderContext ctx = make_ctx(s, cam, img, cgh);
    Material mat = to_material(*m);
    XorShift32State r{*state};
    ScatterResult res;
    res.dir = sycl::float3(0.0f);
    res.attenuation = sycl::float3(0.0f);
    bool ok = mat.scatter(ctx, r, sycl::float3(dir[0], dir[1], dir[2]), sycl::float3(normal[0], normal[1], normal[2]),
                          sycl::float2(uv[0], uv[1]), res);
    *state = r.a;
    for (int k = 0; k < 3; k++) {
        out_dir[k] = res.dir[k];
        out_att[k] = res.attenuation[k];
    }
    return ok ? 1 : 0;
}

/* src/trace_ray.hpp: one segment. returns 1 when the path terminated (result valid) */
int ref_trace_ray(void *scene, uint32_t *state, float org[3], float dir[3], float att[3], float rad[3], float result[3]) {
    RefScene *s = (RefScene *)scene;
    static uint8_t dummy[4];
    sycl::image<3> img(s->n_layers ? (void *)s->tex.data() : (void *)dummy, sycl::image_channel_order::rgba,
                       sycl::image_channel_type::unorm_int8, sycl::range<3>(512, 512, s->n_layers ? s->n_layers : 1));
    sycl::handler cgh;
    Camera cam(sycl::range<2>(8, 8), glm::vec3(0, 0, 0), glm::vec3(0, 0, -1), 1.0f);
    RenderContext ctx = make_ctx(s, cam, img, cgh);
    XorShift32State r{*state};
    RTCRay ray = {org[0], org[1], org[2], 0.0001f, dir[0], dir[1], dir[2], 0.0f, std::numeric_limits<float>::infinity(),
                  UINT32_MAX, 0, 0};
    sycl::float3 a(att[0], att[1], att[2]), e(rad[0], rad[1], rad[2]);
    auto res = trace_ray(ctx, r, ray, a, e);
    *state = r.a;
    org[0] = ray.org_x; org[1] = ray.org_y; org[2] = ray.org_z;
    dir[0] = ray.dir_x; dir[1] = ray.dir_y; dir[2] = ray.dir_z;
    for (int k = 0; k < 3; k++) {
        att[k] = a[k];
        rad[k] = e[k];
        result[k] = res ? (*res)[k] : 0.0f;
    }
    return res ? 1 : 0;
}

/* The reference's renderers, run for real: MegakernelRenderer / WavefrontRenderer ::render_frame
 * (src/main.cpp:36-70 reproduced around them). rgba8 = the bytes handed to stbi_write_png.
 * Returns the number of rtcIntersect1 calls. */
unsigned long long ref_render(void *scene, int wavefront, int w, int h, const float pos[3], const float dir[3], float focal,
                              uint32_t max_depth, uint32_t spp, uint8_t *rgba8) {
    RefScene *s = (RefScene *)scene;
    g_scene = s;
    App app;
    sycl::range<2> img_size(w, h);
    uint8_t *image_buf = sycl::malloc_shared<uint8_t>(img_size[0] * img_size[1] * 4, app.queue);
    sycl::image<2> image(image_buf, sycl::image_channel_order::rgba, sycl::image_channel_type::unorm_int8, img_size);
    Scene sc(app, "", glm::vec3(1.0f, 1.0f, 1.0f));
    sc.scene = (RTCScene)s;
    sc.sky_color = sycl::float3(s->sky[0], s->sky[1], s->sky[2]);
    static uint8_t dummy[4];
    sc.image_array.emplace(s->n_layers ? (void *)s->tex.data() : (void *)dummy, sycl::image_channel_order::rgba,
                           sycl::image_channel_type::unorm_int8, sycl::range<3>(512, 512, s->n_layers ? s->n_layers : 1));
    Camera camera(img_size, glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(dir[0], dir[1], dir[2]), focal);
    g_intersect_calls = 0;
    g_png_bytes.clear();
    std::unique_ptr<IRenderer> renderer;
    if (wavefront) renderer.reset(new WavefrontRenderer(app, img_size, image, max_depth, spp));
    else renderer.reset(new MegakernelRenderer(app, img_size, image, max_depth, spp));
    renderer->render_frame(camera, sc);
    if (rgba8 && g_png_bytes.size() == (size_t)w * h * 4) memcpy(rgba8, g_png_bytes.data(), g_png_bytes.size());
    sycl::free(image_buf, app.queue);
    return g_intersect_calls;
}

} /* extern "C" */
