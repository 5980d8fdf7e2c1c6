/* glb_capi.cpp — C view of glb_loader.hpp for the Python tests (libglb_loader.so, host only). */
#include "glb_loader.hpp"

using raytracer::glb::LoadedScene;
static thread_local std::string g_err;

extern "C" {
void *glb_load(const char *path) {
    try {
        return new LoadedScene(raytracer::glb::load(path));
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void *glb_load_scaled(const char *path, float sx, float sy, float sz) { /* Scene(app, path, global_scale), src/scene.hpp:91-93 */
    try {
        const float gs[3] = {sx, sy, sz};
        return new LoadedScene(raytracer::glb::load(path, gs));
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
const char *glb_last_error() { return g_err.c_str(); }
void glb_free(void *h) { delete (LoadedScene *)h; }
uint32_t glb_instance_count(void *h) { return (uint32_t)((LoadedScene *)h)->instances.size(); }
uint32_t glb_layer_count(void *h) { return ((LoadedScene *)h)->texture_layer_count; }
const uint8_t *glb_layers(void *h) { return ((LoadedScene *)h)->texture_layers.data(); }
/* out16: sky[3], cam pos[3], cam dir[3], focal, has_camera */
void glb_globals(void *h, float *out) {
    LoadedScene *s = (LoadedScene *)h;
    for (int k = 0; k < 3; k++) { out[k] = s->sky_color[k]; out[3 + k] = s->camera_position[k]; out[6 + k] = s->camera_direction[k]; }
    out[9] = s->camera_focal_length;
    out[10] = s->has_camera ? 1.0f : 0.0f;
}
void glb_instance(void *h, uint32_t i, uint32_t *n_verts, uint32_t *n_idx, const float **pos, const float **nrm,
                  const float **uv, const uint32_t **idx, float *transform16, rt_material *mat, int32_t *node_mesh_prim) {
    auto &in = ((LoadedScene *)h)->instances[i];
    *n_verts = (uint32_t)(in.positions.size() / 3);
    *n_idx = (uint32_t)in.indices.size();
    *pos = in.positions.data(); *nrm = in.normals.data(); *uv = in.uvs.data(); *idx = in.indices.data();
    memcpy(transform16, in.transform.m, 64);
    *mat = in.material;
    node_mesh_prim[0] = in.node; node_mesh_prim[1] = in.mesh; node_mesh_prim[2] = in.primitive;
}
int glb_png_write(const char *path, const uint8_t *rgba, uint32_t w, uint32_t h) { return raytracer::glb::png_write(path, rgba, w, h) ? 1 : 0; }
/* any embedded image (PNG / JPEG) -> RGBA8; comp = channels in the file (stbi's `comp`) */
int glb_image_decode(const uint8_t *data, size_t n, uint8_t *rgba_out, uint32_t cap, uint32_t *w, uint32_t *h, int *comp) {
    try {
        raytracer::img::Image im = raytracer::img::decode_rgba8(data, n);
        *w = im.w; *h = im.h; *comp = im.channels_in_file;
        if (im.rgba.size() > cap) { g_err = "output buffer too small"; return 0; }
        memcpy(rgba_out, im.rgba.data(), im.rgba.size());
        return 1;
    } catch (const std::exception &e) { g_err = e.what(); return 0; }
}
/* the 256-entry sRGB decode table of the bake (bake_resize.hpp), for the test that compares it with the reference's */
void glb_srgb_decode_table(float *out256) { memcpy(out256, raytracer::glb::bake::tables().to_linear, 256 * sizeof(float)); }
/* the bake's resize to one 512x512 RGBA8 layer (src/image_manager.hpp:52-62) */
void glb_resize_to_layer(const uint8_t *rgba, uint32_t w, uint32_t h, uint8_t *out) {
    std::vector<uint8_t> src(rgba, rgba + (size_t)w * h * 4);
    std::vector<uint8_t> layer = raytracer::glb::resize_to_layer(src, w, h);
    memcpy(out, layer.data(), layer.size());
}
int glb_png_read(const uint8_t *data, size_t n, uint8_t *rgba_out, uint32_t cap, uint32_t *w, uint32_t *h) {
    try {
        auto px = raytracer::glb::png_decode(data, n, *w, *h);
        if (px.size() > cap) return 0;
        memcpy(rgba_out, px.data(), px.size());
        return 1;
    } catch (const std::exception &e) { g_err = e.what(); return 0; }
}
}
