"""The reference's own benchmark protocol (benchmark.py:6-17: depth 10..50 at 128 spp, then 32..512 spp at depth 10, both
renderers, first iteration discarded) on the synthetic stand-ins of its scenes, through the same render_frame call.
Writes the reference's CSV columns.    python tools/benchmark_protocol.py [out.csv] [iterations]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sycl-ray-tracer_b200")
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/benchmark_protocol.csv"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
DEPTH_SAMPLES = [(10, 128), (20, 128), (30, 128), (40, 128), (50, 128), (10, 32), (10, 128), (10, 256), (10, 512)]
app = pkg.App(0)
rows = ["renderer,depth,samples,scene,time,rays_per_sec,ray_count"]
for wl in ("c3_sponza_scale", "c2_cornell"):
    data, w, h, _, _ = bench.build_scene_data(wl)
    scene = pkg.Scene(app, data)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    for depth, spp in DEPTH_SAMPLES:
        for flag, cls in (("-m", pkg.MegakernelRenderer), ("-w", pkg.WavefrontRenderer)):
            r = cls(app, (w, h), None, depth, spp)
            t = rays = 0.0
            for i in range(iters + 1):
                f = r.render_frame(cam, scene, want=())
                if i == 0:
                    continue
                t, rays = t + f.device_ms * 1e-3, rays + f.ray_count
            r.close()
            rows.append(f"{flag},{depth},{spp},{wl},{t / iters:.6f},{rays / t / 1e6:.3f},{rays / iters:.1f}")
            print(rows[-1], flush=True)
    scene.close()
open(out, "w").write("\n".join(rows) + "\n")
