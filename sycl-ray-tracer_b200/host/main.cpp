/*
 * main.cpp — the reference's CLI (src/main.cpp:8-77) over the C++ mirror in raytracer.hpp:
 *
 *     raytracer [-d N] [-s N] [-m|-w] [--size WxH] [--ppm out.ppm] [--gpus N [--shard tiles|spp]] [scene]
 *
 * Same flags and defaults as the reference: -d/--max-depth 10, -s/--sample-count 32,
 * -w/--wavefront (default), -m/--megakernel (wins if both are given), fixed 1920x1080 unless --size.
 * `scene` is a .glb path (glb_loader.hpp, the reference loader's rules) or the name of the built-in
 * procedural scene "cube" (the shape of assets/cube.glb). Writes out.png like the reference
 * (src/util.hpp:27) unless --no-png, and prints the three lines benchmark.py parses.
 * --gpus N renders the frame on N GPUs of this node (rt_group_*): --shard tiles (default) writes exactly the image one
 * GPU would, --shard spp splits the samples (device i uses seed salt i * 0x9E3779B9).
 */
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "glb_loader.hpp"
#include "raytracer.hpp"

namespace {

struct CubeScene {
    std::vector<float> pos, nrm, uv;
    std::vector<uint32_t> idx;
    rt_instance inst{};
    rt_scene_desc desc{};
    CubeScene() {
        const float f[6][3][3] = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}},  {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}},
                                  {{0, 1, 0}, {0, 0, 1}, {1, 0, 0}},  {{0, -1, 0}, {1, 0, 0}, {0, 0, 1}},
                                  {{0, 0, 1}, {1, 0, 0}, {0, 1, 0}},  {{0, 0, -1}, {0, 1, 0}, {1, 0, 0}}};
        const int s[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        for (uint32_t k = 0; k < 6; k++) {
            for (int c = 0; c < 4; c++) {
                for (int a = 0; a < 3; a++) {
                    pos.push_back(f[k][0][a] + s[c][0] * f[k][1][a] + s[c][1] * f[k][2][a]);
                    nrm.push_back(f[k][0][a]);
                }
                uv.push_back((s[c][0] + 1) * 0.5f);
                uv.push_back((s[c][1] + 1) * 0.5f);
            }
            const uint32_t q[6] = {0, 1, 2, 0, 2, 3};
            for (uint32_t v : q) idx.push_back(4 * k + v);
        }
        inst.positions = pos.data();
        inst.normals = nrm.data();
        inst.uvs = uv.data();
        inst.indices = idx.data();
        inst.vertex_count = 24;
        inst.index_count = 36;
        const float t[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.05813104659318924f, 0.1505535989999771f, -2.920884370803833f, 1};
        std::memcpy(inst.transform, t, sizeof(t));
        inst.material.type = RT_MAT_DIFFUSE;
        inst.material.albedo_image = -1;
        inst.material.albedo_color[0] = inst.material.albedo_color[1] = inst.material.albedo_color[2] = 0.8f;
        inst.material.ior = 1.5f;
        desc.instances = &inst;
        desc.instance_count = 1;
        desc.sky_color[0] = 0.5f;
        desc.sky_color[1] = 0.7f;
        desc.sky_color[2] = 1.0f;
    }
};

} // namespace

int main(int argc, const char *argv[]) {
    uint32_t max_depth = 10, sample_count = 32;
    std::string scene_path = "./assets/sponza.glb", ppm, png = "out.png"; /* src/main.cpp:16 */
    bool use_wavefront = false, use_megakernel = false;
    size_t w = 1920, h = 1080;
    uint32_t gpus = 1;
    rt_group_mode shard = RT_GROUP_TILES;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "-d" || a == "--max-depth") max_depth = (uint32_t)std::atoi(next());
        else if (a == "-s" || a == "--sample-count") sample_count = (uint32_t)std::atoi(next());
        else if (a == "-w" || a == "--wavefront") use_wavefront = true;
        else if (a == "-m" || a == "--megakernel") use_megakernel = true;
        else if (a == "--size") std::sscanf(next(), "%zux%zu", &w, &h);
        else if (a == "--gpus") gpus = (uint32_t)std::atoi(next());
        else if (a == "--shard") shard = std::string(next()) == "spp" ? RT_GROUP_SPP : RT_GROUP_TILES;
        else if (a == "--ppm") ppm = next();
        else if (a == "--png") png = next();
        else if (a == "--no-png") png.clear();
        else scene_path = a;
    }
    if (!use_wavefront && !use_megakernel) use_wavefront = true; /* src/main.cpp:26-28 */
    std::printf("Loading scene: %s\n", scene_path.c_str());
    try {
        /* the device(s) first, like the reference (src/main.cpp:30): without a GPU this is where the program ends */
        std::unique_ptr<raytracer::App> one;
        std::unique_ptr<raytracer::GroupApp> many;
        if (gpus > 1) many.reset(new raytracer::GroupApp(gpus));
        else one.reset(new raytracer::App());
        raytracer::range2 img_size(w, h);
        raytracer::Image image(img_size);
        CubeScene cube;
        raytracer::glb::LoadedScene loaded;
        rt_scene_desc desc = cube.desc;
        std::array<float, 3> cam_pos{0, 0, 0}, cam_dir{0, 0, -1};
        float cam_focal = 1.0f;
        if (scene_path != "cube") {
            loaded = raytracer::glb::load(scene_path); /* throws "Failed to load .glTF : ..." (src/scene.cpp:68-70) */
            desc = loaded.desc();
            cam_pos = {loaded.camera_position[0], loaded.camera_position[1], loaded.camera_position[2]};
            cam_dir = {loaded.camera_direction[0], loaded.camera_direction[1], loaded.camera_direction[2]};
            cam_focal = loaded.camera_focal_length;
        }
        raytracer::Camera camera(img_size, cam_pos, cam_dir, cam_focal);
        if (gpus > 1) { /* the same sequence on a group of devices */
            raytracer::GroupApp &app = *many;
            raytracer::GroupScene scene(app, desc);
            raytracer::GroupRenderer renderer(use_megakernel ? RT_MEGAKERNEL : RT_WAVEFRONT, app, img_size, image, max_depth, sample_count, shard);
            renderer.render_frame(camera, scene);
        } else {
            raytracer::App &app = *one;
            raytracer::Scene scene(app, desc);
            scene.camera_position = cam_pos;
            scene.camera_direction = cam_dir;
            scene.camera_focal_length = cam_focal;
            std::unique_ptr<raytracer::IRenderer> renderer;
            if (use_megakernel) renderer.reset(new raytracer::MegakernelRenderer(app, img_size, image, max_depth, sample_count));
            else renderer.reset(new raytracer::WavefrontRenderer(app, img_size, image, max_depth, sample_count));
            renderer->render_frame(camera, scene);
        }
        if (!png.empty()) {
            std::printf("Writing image to disk\n"); /* src/render_megakernel.cpp:185 */
            if (!raytracer::glb::png_write(png, image.rgba8.data(), (uint32_t)w, (uint32_t)h)) {
                std::printf("Failed to write image to disk.\n"); /* src/util.hpp:28-30 */
                return 1;
            }
        }
        if (!ppm.empty()) {
            std::ofstream o(ppm, std::ios::binary);
            o << "P6\n" << w << " " << h << "\n255\n";
            for (size_t p = 0; p < w * h; p++) o.write((const char *)&image.rgba8[p * 4], 3);
        }
    } catch (const std::exception &e) {
        std::printf("error: %s\n", e.what());
        return 1;
    }
    return 0;
}
