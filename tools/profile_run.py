"""Short run of one renderer for ncu (launch list / --set full): scene build + a few frames."""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sycl-ray-tracer_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_sponza_scale")
ap.add_argument("--renderer", default="megakernel")
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--width", type=int, default=0)
ap.add_argument("--height", type=int, default=0)
a = ap.parse_args()
data, w, h, spp, depth = bench.build_scene_data(a.workload)
w, h = a.width or w, a.height or h
app = pkg.App(0)
scene = pkg.Scene(app, data)
cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
cls = pkg.MegakernelRenderer if a.renderer == "megakernel" else pkg.WavefrontRenderer
r = cls(app, (w, h), None, depth, a.spp)
for i in range(a.frames):
    f = r.render_frame(cam, scene, want=())
    print(f"{a.renderer} {a.workload} {w}x{h} spp={a.spp}: {f.ray_count} rays, {f.device_ms:.3f} ms, "
          f"{f.ray_count / f.device_ms / 1e3:.1f} Mrays/s, launches {f.kernel_launches}", flush=True)
print("stats", scene.stats)
