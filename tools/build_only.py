"""Scene upload + GPU BVH build only (for launch lists of the build): python tools/build_only.py [workload] [repeats]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sycl-ray-tracer_b200")
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "c4_heightfield_10m"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
data, w, h, spp, depth = bench.build_scene_data(wl)
app = pkg.App(0)
for i in range(reps):
    t = time.perf_counter()
    sc = pkg.Scene(app, data)
    dt = time.perf_counter() - t
    print(f"{wl} rep {i}: create+commit {dt * 1e3:.1f} ms  stats {sc.stats}", flush=True)
    sc.close()
