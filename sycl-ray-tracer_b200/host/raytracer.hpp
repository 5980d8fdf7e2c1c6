/*
 * raytracer.hpp — C++ host mirror of the reference's public interface over the C ABI
 * (include/rt_api.h). Same namespace, class names, constructor shapes and call sequence as
 * felipeagc/sycl-ray-tracer, so src/main.cpp:30-70 reads the same with this header:
 *
 *     raytracer::App app;                                              // src/app.hpp:43-55
 *     raytracer::Scene scene(app, desc);                               // src/scene.cpp:54-129
 *     raytracer::Camera camera(img_size, scene.camera_position,
 *                              scene.camera_direction, scene.camera_focal_length);
 *     std::unique_ptr<raytracer::IRenderer> renderer(
 *         new raytracer::MegakernelRenderer(app, img_size, image, max_depth, sample_count));
 *     renderer->render_frame(camera, scene);                           // src/main.cpp:70
 *
 * Differences, all forced by the replacement of SYCL by CUDA streams behind the C ABI:
 *   - App owns an rt_context (device + cudaStream_t) instead of sycl::queue + RTCDevice;
 *   - img_size is raytracer::range2 {width, height} instead of sycl::range<2>;
 *   - the output image is an Image (host RGBA8 buffer) instead of sycl::image<2>&;
 *   - Scene is built from an rt_scene_desc (what the GLB loader produces) — the GLB loader itself
 *     is a "next" row of the scope table;
 *   - errors become std::runtime_error (the reference terminates, src/main.cpp:71-74).
 * render_frame prints the same three metric lines benchmark.py parses
 * (src/render_megakernel.cpp:181-183) unless `quiet` is set.
 */
#pragma once

#include <array>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rt_api.h"

namespace raytracer {

struct range2 {
    size_t v[2];
    range2(size_t w, size_t h) : v{w, h} {}
    size_t operator[](int i) const { return v[i]; }
    size_t size() const { return v[0] * v[1]; }
};

/* replaces sycl::image<2> (rgba, unorm_int8) of src/main.cpp:39-46 */
struct Image {
    range2 size;
    std::vector<uint8_t> rgba8;
    explicit Image(range2 s) : size(s), rgba8(s.size() * 4) {}
};

/* raytracer::App, src/app.hpp:31-58 */
struct App {
    rt_context *ctx = nullptr;

    App(const App &) = delete;
    App &operator=(const App &) = delete;

    explicit App(int device = 0) {
        if (rt_context_create(device, &ctx) != RT_OK)
            throw std::runtime_error(std::string("rt_context_create: ") + rt_last_error(nullptr));
        std::printf("Running on device: %s\n", rt_context_device_name(ctx)); /* src/app.hpp:51-54 */
    }
    ~App() { rt_context_destroy(ctx); }

    void check(rt_status st, const char *what) const {
        if (st != RT_OK) throw std::runtime_error(std::string(what) + ": " + rt_last_error(ctx));
    }
};

/* raytracer::Scene, src/scene.hpp:64-104: owns the committed acceleration structure, the sky colour
 * and the camera extracted from the file */
struct Scene {
    App &app;
    rt_scene *scene = nullptr;
    std::array<float, 3> camera_position{0, 0, 0};
    std::array<float, 3> camera_direction{0, 0, -1};
    float camera_focal_length = 1.0f;
    std::array<float, 3> sky_color{0.5f, 0.7f, 1.0f}; /* src/scene.hpp:76 */

    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;

    Scene(App &app_, const rt_scene_desc &desc) : app(app_) {
        sky_color = {desc.sky_color[0], desc.sky_color[1], desc.sky_color[2]};
        app.check(rt_scene_create(app.ctx, &desc, &scene), "rt_scene_create");
        app.check(rt_scene_commit(scene), "rt_scene_commit"); /* rtcCommitScene, src/scene.cpp:107 */
    }
    ~Scene() { rt_scene_destroy(scene); }

    rt_scene_stats stats() const {
        rt_scene_stats s{};
        app.check(rt_scene_get_stats(scene, &s), "rt_scene_get_stats");
        return s;
    }
};

/* raytracer::Camera, src/camera.hpp:65-131 */
struct Camera {
    rt_camera c{};
    Camera(range2 img_size, const std::array<float, 3> &cam_center, const std::array<float, 3> &cam_dir,
           float focal_length) {
        rt_camera_init(&c, (int32_t)img_size[0], (int32_t)img_size[1], cam_center.data(), cam_dir.data(),
                       focal_length);
    }
};

/* raytracer::IRenderer, src/render.hpp:11-18 */
struct IRenderer {
    virtual void render_frame(const Camera &camera, const Scene &scene) = 0;
    virtual ~IRenderer() {}
};

namespace detail {
struct RendererBase : public IRenderer {
    App &app;
    range2 img_size;
    Image &image;
    const uint32_t max_depth;
    const uint32_t sample_count;
    rt_renderer *handle = nullptr;
    bool quiet = false;
    rt_frame last{}; /* ray_count, device_ms, kernel_launches of the last frame */

    RendererBase(rt_renderer_kind kind, App &app_, range2 size, Image &image_, uint32_t depth, uint32_t spp)
        : app(app_), img_size(size), image(image_), max_depth(depth), sample_count(spp) {
        app.check(rt_renderer_create(app.ctx, kind, (int32_t)size[0], (int32_t)size[1], &handle),
                  "rt_renderer_create");
    }
    ~RendererBase() override { rt_renderer_destroy(handle); }

    void render_frame(const Camera &camera, const Scene &scene) override {
        rt_render_params p{};
        p.max_depth = max_depth;
        p.sample_count = sample_count;
        last = rt_frame{};
        last.rgba8 = image.rgba8.data();
        app.check(rt_render_frame(handle, scene.scene, &camera.c, &p, &last), "rt_render_frame");
        if (!quiet) { /* src/render_megakernel.cpp:181-183, src/render_wavefront.cpp:425-427 */
            const double secs = last.device_ms * 1e-3;
            std::printf("Time measured: %.6f seconds\n", secs);
            std::printf("Total rays: %llu\n", (unsigned long long)last.ray_count);
            std::printf("Rays/sec: %.2fM\n", (double)last.ray_count / secs / 1000000.0);
        }
    }
};
} // namespace detail

/* ---- several GPUs of one node (no reference equivalent: App picks one device, src/app.hpp:43-55) --------------
 * Same call sequence, N devices: GroupApp replaces App, GroupScene replaces Scene, GroupRenderer(kind, ...) replaces
 * the two renderer classes. Mode RT_GROUP_TILES renders exactly the single-device image. */
struct GroupApp {
    rt_group *group = nullptr;
    GroupApp(const GroupApp &) = delete;
    GroupApp &operator=(const GroupApp &) = delete;
    explicit GroupApp(uint32_t n_devices) {
        if (rt_group_create(nullptr, n_devices, &group) != RT_OK)
            throw std::runtime_error(std::string("rt_group_create: ") + rt_last_error(nullptr));
        for (uint32_t i = 0; i < n_devices; i++)
            std::printf("Running on device: %s\n", rt_context_device_name(rt_group_context(group, i)));
    }
    ~GroupApp() { rt_group_destroy(group); }
    uint32_t size() const { return rt_group_size(group); }
    void check(rt_status st, const char *what) const {
        if (st != RT_OK) throw std::runtime_error(std::string(what) + ": " + rt_group_last_error(group));
    }
};

struct GroupScene {
    GroupApp &app;
    rt_group_scene *scene = nullptr;
    GroupScene(const GroupScene &) = delete;
    GroupScene &operator=(const GroupScene &) = delete;
    GroupScene(GroupApp &app_, const rt_scene_desc &desc) : app(app_) {
        app.check(rt_group_scene_create(app.group, &desc, &scene), "rt_group_scene_create");
    }
    ~GroupScene() { rt_group_scene_destroy(scene); }
};

struct GroupRenderer {
    GroupApp &app;
    range2 img_size;
    Image &image;
    const uint32_t max_depth, sample_count;
    rt_group_mode mode;
    rt_group_renderer *handle = nullptr;
    bool quiet = false;
    rt_frame last{};
    GroupRenderer(rt_renderer_kind kind, GroupApp &app_, range2 size, Image &image_, uint32_t depth, uint32_t spp, rt_group_mode mode_)
        : app(app_), img_size(size), image(image_), max_depth(depth), sample_count(spp), mode(mode_) {
        app.check(rt_group_renderer_create(app.group, kind, (int32_t)size[0], (int32_t)size[1], &handle), "rt_group_renderer_create");
    }
    ~GroupRenderer() { rt_group_renderer_destroy(handle); }
    void render_frame(const Camera &camera, const GroupScene &scene) {
        rt_group_params p{};
        p.max_depth = max_depth;
        p.sample_count = sample_count;
        p.mode = (uint32_t)mode;
        last = rt_frame{};
        last.rgba8 = image.rgba8.data();
        app.check(rt_group_render_frame(handle, scene.scene, &camera.c, &p, &last), "rt_group_render_frame");
        if (!quiet) {
            const double secs = last.device_ms * 1e-3;
            std::printf("Time measured: %.6f seconds\n", secs);
            std::printf("Total rays: %llu\n", (unsigned long long)last.ray_count);
            std::printf("Rays/sec: %.2fM\n", (double)last.ray_count / secs / 1000000.0);
        }
    }
};

/* src/render_megakernel.hpp:6-22 */
struct MegakernelRenderer : public detail::RendererBase {
    MegakernelRenderer(App &app, range2 img_size, Image &image, uint32_t max_depth, uint32_t sample_count)
        : RendererBase(RT_MEGAKERNEL, app, img_size, image, max_depth, sample_count) {}
};

/* src/render_wavefront.hpp:40-63 */
struct WavefrontRenderer : public detail::RendererBase {
    WavefrontRenderer(App &app, range2 img_size, Image &image, uint32_t max_depth, uint32_t sample_count)
        : RendererBase(RT_WAVEFRONT, app, img_size, image, max_depth, sample_count) {}
};

} // namespace raytracer
