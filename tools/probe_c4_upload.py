import importlib, os, sys, time
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("sycl-ray-tracer_b200")
import bench
data, w, h, spp, depth = bench.build_scene_data("c4_heightfield_10m")
app = pkg.App(0)
for i in range(4):
    t = time.perf_counter(); sc = pkg.Scene(app, data, commit=False); t1 = time.perf_counter(); sc.commit(); t2 = time.perf_counter()
    print(f"rep {i}: create {(t1 - t) * 1e3:.1f} ms commit {(t2 - t1) * 1e3:.1f} ms build {sc.stats['build_ms']:.1f}", flush=True)
    sc.close()
