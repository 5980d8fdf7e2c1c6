/*
 * fullref.cpp — TEST-ONLY: the reference's WHOLE program on the CPU. Compiled in place and unmodified:
 * src/scene.cpp (the glTF loader, with deps/tiny_gltf.cpp, deps/stb_image.cpp, deps/stb_image_resize2.cpp) AND
 * src/render_megakernel.cpp, src/render_wavefront.cpp with every header they include — i.e. everything
 * src/main.cpp:30-70 runs, reproduced around them here with the image size as a parameter (main.cpp hardcodes
 * 1920x1080). Not the reference: Embree (scene construction is recorded, rtcIntersect1 is forwarded to the
 * oracle's brute-force closest hit over the recorded instances), the SYCL runtime and glm (oracle/refshim).
 * tests/test_glb_loader.py compares "reference program on a .glb" with "our loader + the oracle's renderer".
 */
#define FMT_HEADER_ONLY 1
#include <cstdio>
#include <string>
#include <vector>

#include "../rt_oracle.h"

#include "render_megakernel.cpp" /* /root/reference/src */
#include "render_wavefront.cpp"

#include "embree_record.hpp"

using namespace raytracer;

/* out.png capture: src/util.hpp:27 hands the final bytes to stbi_write_png */
static std::vector<uint8_t> g_png_bytes;
extern "C" {
int stbi_write_png(char const *, int w, int h, int comp, const void *data, int stride) {
    g_png_bytes.resize((size_t)w * h * comp);
    for (int y = 0; y < h; y++) memcpy(&g_png_bytes[(size_t)y * w * comp], (const uint8_t *)data + (size_t)y * stride, (size_t)w * comp);
    return 1;
}
/* tinygltf's image WRITER references these; never called */
int stbi_write_png_to_func(stbi_write_func *, void *, int, int, int, const void *, int) { return 0; }
int stbi_write_jpg_to_func(stbi_write_func *, void *, int, int, int, const void *, int) { return 0; }
int stbi_write_bmp_to_func(stbi_write_func *, void *, int, int, int, const void *) { return 0; }
}

/* ---- rtcIntersect1 over the recorded scene ---------------------------------------------------- */
namespace {
struct Intersector {
    RTCScene top = nullptr;
    orc_scene *orc = nullptr;
    unsigned long long calls = 0;
} g_isect;

void build_intersector(RTCScene scene) {
    Scn *s = (Scn *)scene;
    std::vector<orc_instance> inst(s->geoms.size());
    for (size_t i = 0; i < s->geoms.size(); i++) {
        const Geom *g = s->geoms[i];
        const Geom *tri = ((Scn *)g->instanced)->geoms[0];
        const GeometryData *gd = (const GeometryData *)g->user;
        orc_instance &o = inst[i];
        memset(&o, 0, sizeof(o));
        o.positions = (const float *)tri->vertices;
        o.normals = (const float *)gd->normal_buffer;
        o.uvs = (const float *)gd->uv_buffer;
        o.indices = (const uint32_t *)tri->indices;
        o.vertex_count = (uint32_t)tri->n_vertices;
        o.index_count = (uint32_t)tri->n_triangles * 3;
        memcpy(o.transform, g->xfm, 64);
        o.material.type = ORC_MAT_DIFFUSE; /* only geometry is used */
        o.material.albedo_image = -1;
    }
    const float sky[3] = {0, 0, 0};
    if (g_isect.orc) orc_scene_destroy(g_isect.orc);
    g_isect.orc = orc_scene_create(inst.data(), (uint32_t)inst.size(), nullptr, 0, sky);
    g_isect.top = scene;
}
} // namespace

void rtcIntersect1(RTCScene scene, RTCRayHit *rh) {
    if (g_isect.top != scene) build_intersector(scene);
    g_isect.calls++;
    int32_t inst = -1, prim = -1;
    float u = 0, v = 0, t = 0;
    const float org[3] = {rh->ray.org_x, rh->ray.org_y, rh->ray.org_z}, dir[3] = {rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z};
    orc_intersect(g_isect.orc, 0, 1, 1, org, dir, rh->ray.tnear, rh->ray.tfar, &inst, &prim, &u, &v, &t);
    if (inst >= 0) {
        rh->ray.tfar = t;
        rh->hit.u = u;
        rh->hit.v = v;
        rh->hit.primID = (unsigned)prim;
        rh->hit.geomID = 0;
        rh->hit.instID[0] = (unsigned)inst;
    }
}
void *rtcGetGeometryUserDataFromScene(RTCScene scene, unsigned int id) { return ((Scn *)scene)->geoms[id]->user; }

extern "C" {
static thread_local std::string g_full_err;
const char *fullref_last_error() { return g_full_err.c_str(); }
/* src/main.cpp:30-70 with img_size as a parameter. rgba8 <- the bytes handed to stbi_write_png (w*h*4);
 * returns the number of rtcIntersect1 calls, or ~0 on an exception. */
unsigned long long fullref_main(const char *scene_path, int use_megakernel, int w, int h, uint32_t max_depth, uint32_t sample_count,
                                uint8_t *rgba8) {
    try {
        raytracer::App app;
        sycl::range<2> img_size = sycl::range<2>(w, h);
        uint8_t *image_buf = sycl::malloc_shared<uint8_t>(img_size[0] * img_size[1] * 4, app.queue);
        sycl::image<2> image(image_buf, sycl::image_channel_order::rgba, sycl::image_channel_type::unorm_int8, img_size);
        raytracer::Scene scene(app, scene_path);
        raytracer::Camera camera(img_size, scene.camera_position, scene.camera_direction, scene.camera_focal_length);
        std::unique_ptr<raytracer::IRenderer> renderer;
        if (use_megakernel) renderer.reset(new raytracer::MegakernelRenderer(app, img_size, image, max_depth, sample_count));
        else renderer.reset(new raytracer::WavefrontRenderer(app, img_size, image, max_depth, sample_count));
        g_isect.calls = 0;
        g_png_bytes.clear();
        renderer->render_frame(camera, scene);
        if (rgba8 && g_png_bytes.size() == (size_t)w * h * 4) memcpy(rgba8, g_png_bytes.data(), g_png_bytes.size());
        return g_isect.calls;
    } catch (const std::exception &e) {
        g_full_err = e.what();
        return ~0ull;
    }
}
}
