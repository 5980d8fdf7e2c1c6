/*
 * scenref.cpp — TEST-ONLY: runs the reference's OWN glTF loader. `make -C oracle ref` compiles, in place
 * from /root/reference and unmodified, src/scene.cpp (Scene::Scene, load_images, load_primitives, load_node,
 * node_global_matrix) together with the reference's vendored deps/tiny_gltf.cpp, deps/stb_image.cpp,
 * deps/stb_image_resize2.cpp, deps/stb_image_write.cpp against the API shims of this directory, plus this
 * file, into oracle/_ref/libscenref.so. The Embree scene-construction calls (which only store pointers) are
 * recorded here, so that everything the loader hands to the renderer can be read back: per instance, in
 * attach order, the vertex / normal / uv / index buffers, the 4x4 transform given to Embree, the normal
 * matrix and Material of its GeometryData; sky colour; camera; the baked 512x512 image layers.
 * tests/test_glb_loader.py compares sycl-ray-tracer_b200/host/glb_loader.hpp with it on the same .glb files.
 * What is NOT the reference here: glm (oracle/refshim/glm: component-wise definitions) and the USM allocator.
 */
#define FMT_HEADER_ONLY 1
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "scene.hpp" /* /root/reference/src */

using namespace raytracer;

#include "embree_record.hpp"

/* not reached by the loader */
void rtcIntersect1(RTCScene, RTCRayHit *) {}
void *rtcGetGeometryUserDataFromScene(RTCScene, unsigned int) { return nullptr; }

namespace {
struct Loaded {
    App *app = nullptr;
    Scene *scene = nullptr;
};
thread_local std::string g_err;
} // namespace

extern "C" {
void *scenref_load(const char *path, float gsx, float gsy, float gsz) {
    Loaded *l = new Loaded();
    try {
        l->app = new App();
        l->scene = new Scene(*l->app, path, glm::vec3(gsx, gsy, gsz));
        return l;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
const char *scenref_last_error() { return g_err.c_str(); }
uint32_t scenref_instance_count(void *h) { return (uint32_t)((Scn *)((Loaded *)h)->scene->scene)->geoms.size(); }
/* material out: [type (1 diffuse, 2 metallic, 3 dielectric), albedo is image, image index] + floats [albedo rgb, roughness, ior, emissive rgb] */
void scenref_instance(void *h, uint32_t i, uint32_t *n_verts, uint32_t *n_idx, const float **pos, const float **nrm, const float **uv,
                      const uint32_t **idx, float *xfm16, float *nmat9, int32_t *mat_i3, float *mat_f8) {
    const Geom *inst = ((Scn *)((Loaded *)h)->scene->scene)->geoms[i];
    const Geom *tri = ((Scn *)inst->instanced)->geoms[0];
    const GeometryData *gd = (const GeometryData *)inst->user;
    *n_verts = (uint32_t)tri->n_vertices;
    *n_idx = (uint32_t)tri->n_triangles * 3;
    *pos = (const float *)gd->vertex_buffer;
    *nrm = (const float *)gd->normal_buffer;
    *uv = (const float *)gd->uv_buffer;
    *idx = gd->index_buffer;
    memcpy(xfm16, inst->xfm, 64);
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) nmat9[c * 3 + r] = gd->obj_to_world[c][r];
    const Material &m = gd->material;
    mat_i3[0] = (int32_t)m.type;
    mat_i3[1] = 0;
    mat_i3[2] = -1;
    for (int k = 0; k < 8; k++) mat_f8[k] = 0.0f;
    auto albedo = [&](const Texture &t) {
        if (t.type == TextureType::eImage) {
            mat_i3[1] = 1;
            mat_i3[2] = (int32_t)t.image_ref.index;
        } else {
            mat_f8[0] = t.color.x();
            mat_f8[1] = t.color.y();
            mat_f8[2] = t.color.z();
        }
    };
    auto emissive = [&](const sycl::float3 &e) {
        mat_f8[5] = e.x();
        mat_f8[6] = e.y();
        mat_f8[7] = e.z();
    };
    switch (m.type) {
    case MaterialType::eDiffuse:
        albedo(m.diffuse.albedo);
        emissive(m.diffuse.emissive);
        break;
    case MaterialType::eMetallic:
        albedo(m.metallic.albedo);
        mat_f8[3] = m.metallic.roughness;
        emissive(m.metallic.emissive);
        break;
    case MaterialType::eDielectric: mat_f8[4] = m.dielectric.ior; break;
    default: break;
    }
}
/* out: sky[3], camera position[3], camera direction[3], focal length, camera node index */
void scenref_globals(void *h, float *out) {
    const Scene *s = ((Loaded *)h)->scene;
    out[0] = s->sky_color.x();
    out[1] = s->sky_color.y();
    out[2] = s->sky_color.z();
    for (int k = 0; k < 3; k++) {
        out[3 + k] = s->camera_position[k];
        out[6 + k] = s->camera_direction[k];
    }
    out[9] = s->camera_focal_length;
    out[10] = (float)s->camera_node_index;
}
uint32_t scenref_layer_count(void *h) { return (uint32_t)((Loaded *)h)->scene->image_baker.images.size(); }
const uint8_t *scenref_layer(void *h, uint32_t i) { return ((Loaded *)h)->scene->image_baker.images[i].data.data(); }
}
