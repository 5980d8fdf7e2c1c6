"""What bounds the scene upload of config 4 (280 MB from pageable host memory): the link, the host copy into pinned memory, or us.
python tools/probe_h2d.py"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n = 256 << 20
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
page = np.ones(n, np.uint8)
for name, fn in (("pinned -> device (cudaMemcpyAsync)", lambda: dev.copy_(pin, non_blocking=True)),
                 ("pageable -> device (torch)", lambda: dev.copy_(torch.from_numpy(page))),
                 ("pageable -> pinned, one host thread (numpy copy)", lambda: np.copyto(pin.numpy(), page))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name:55s} {n / dt / 1e9:6.1f} GB/s  ({dt * 1e3:.1f} ms per 256 MiB)", flush=True)
print("host cores:", len(os.sched_getaffinity(0)))
import bench
pkg = importlib.import_module("sycl-ray-tracer_b200")
data, w, h, spp, depth = bench.build_scene_data("c4_heightfield_10m")
app = pkg.App(0)
os.environ["RT_TRACE"] = "1"
for _ in range(3):
    t0 = time.perf_counter()
    sc = pkg.Scene(app, data)
    print(f"Scene(): {(time.perf_counter() - t0) * 1e3:.1f} ms, build {sc.stats['build_ms']:.2f} ms", flush=True)
    sc.close()
