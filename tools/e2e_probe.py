import importlib, os, sys, time
sys.path.insert(0, os.getcwd())
import bench
pkg = importlib.import_module("sycl-ray-tracer_b200")
import torch
data, w, h, spp, depth = bench.build_scene_data("c3_sponza_scale")
spp = 32
app = pkg.App(0)
cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
r = pkg.MegakernelRenderer(app, (w, h), None, depth, spp)
host_img = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = T(); sc = pkg.Scene(app, data, commit=False); t1 = T(); sc.commit(); t2 = T()
    f = r.render_frame(cam, sc, want=("rgba8",), outputs={"rgba8": host_img}); t3 = T(); sc.close(); t4 = T()
    print(f"create {1e3*(t1-t0):.1f} ms  commit {1e3*(t2-t1):.1f} ms (device {sc.stats['build_ms'] if False else 0})  render+d2h {1e3*(t3-t2):.1f} ms (device {f.device_ms:.1f})  close {1e3*(t4-t3):.1f} ms")
