"""Why is scene create slow inside bench.py's e2e loop? Time pkg.Scene under the bench's conditions, one at a time."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("sycl-ray-tracer_b200")
import bench
data, w, h, spp, depth = bench.build_scene_data("c4_heightfield_10m")
app = pkg.App(0)

def rep(tag, n=3, hold=None):
    for i in range(n):
        t = time.perf_counter(); sc = pkg.Scene(app, data); dt = time.perf_counter() - t
        print(f"{tag} rep {i}: {dt * 1e3:.1f} ms build {sc.stats['build_ms']:.1f}", flush=True)
        sc.close()

rep("own stream")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); app.set_stream(stream.cuda_stream)
rep("torch stream")
main = pkg.Scene(app, data)
rep("torch stream + live main scene")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rep("+ flush buffer")
cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
r = pkg.MegakernelRenderer(app, (w, h), None, depth, spp)
host_img = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
for i in range(3):
    t = time.perf_counter(); sc = pkg.Scene(app, data); t1 = time.perf_counter()
    f = r.render_frame(cam, sc, want=("rgba8",), outputs={"rgba8": host_img}); t2 = time.perf_counter()
    sc.close(); t3 = time.perf_counter()
    print(f"e2e rep {i}: create {(t1 - t) * 1e3:.1f} render {(t2 - t1) * 1e3:.1f} close {(t3 - t2) * 1e3:.1f}", flush=True)
torch.cuda.synchronize()
for i in range(3):
    t = time.perf_counter(); sc = pkg.Scene(app, data); t1 = time.perf_counter()
    f = r.render_frame(cam, sc, want=("rgba8",), outputs={"rgba8": host_img}); t2 = time.perf_counter()
    sc.close(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"e2e+sync rep {i}: create {(t1 - t) * 1e3:.1f} render {(t2 - t1) * 1e3:.1f} close {(t3 - t2) * 1e3:.1f}", flush=True)
uuid = str(torch.cuda.get_device_properties(0).uuid)
s = bench.ClockSampler(0, uuid); s.start()
for i in range(8):
    t = time.perf_counter(); sc = pkg.Scene(app, data); t1 = time.perf_counter()
    f = r.render_frame(cam, sc, want=("rgba8",), outputs={"rgba8": host_img}); t2 = time.perf_counter()
    sc.close(); t3 = time.perf_counter()
    print(f"e2e+sampler rep {i}: create {(t1 - t) * 1e3:.1f} render {(t2 - t1) * 1e3:.1f} close {(t3 - t2) * 1e3:.1f}", flush=True)
print(s.stop(), 'slowest NVML query ms', s.slowest_query_ms)
for i in range(2):
    t = time.perf_counter(); sc = pkg.Scene(app, data); t1 = time.perf_counter()
    print(f"after sampler rep {i}: create {(t1 - t) * 1e3:.1f}", flush=True); sc.close()
