"""One-process sweep of the scheduling knobs of the persistent kernels (run on the GPU box).

    python tools/tune.py --workload c3_sponza_scale --spp 64 --frames 3 \
        --configs "RT_MEGA_CTX=0;RT_MEGA_CTX=2,RT_TUNE_REFILL=4,RT_TUNE_SHADE=24"

The scene is uploaded once; every configuration creates its own renderer (the knobs are read from the
environment by rt_renderer_create) and renders `frames` frames after one warm-up frame. Prints device-timed
Mrays/s (median and best) per configuration; the image of every configuration is compared with the first one
(scheduling knobs must not change a single bit)."""
import argparse, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sycl-ray-tracer_b200")
KNOBS = ("RT_MEGA_CTX", "RT_TUNE_REFILL", "RT_TUNE_SHADE", "RT_TUNE_IDLE", "RT_BLOCK_ORDER", "RT_WF_PERSIST", "RT_SAMPLE_PARTS", "RT_TUNE_INFLIGHT", "RT_BLOCK_ORDER_MIN_SPP", "RT_ORDER_REGION", "RT_ORDER_PROBES", "RT_TUNE_CARRY")
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_sponza_scale")
ap.add_argument("--renderer", default="megakernel")
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--depth", type=int, default=0)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--width", type=int, default=0)
ap.add_argument("--height", type=int, default=0)
ap.add_argument("--configs", default="")
a = ap.parse_args()
data, w, h, spp, depth = bench.build_scene_data(a.workload)
w, h, spp, depth = a.width or w, a.height or h, a.spp or spp, a.depth or depth
app = pkg.App(0)
scene = pkg.Scene(app, data)
cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
cls = pkg.MegakernelRenderer if a.renderer == "megakernel" else pkg.WavefrontRenderer
ref = None
for cfg in (a.configs.split(";") if a.configs else [""]):
    for k in KNOBS:
        os.environ.pop(k, None)
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        os.environ[k.strip()] = v.strip()
    r = cls(app, (w, h), None, depth, spp)
    f = r.render_frame(cam, scene, want=("rgba8",))
    img = f.rgba8.copy()
    same = "ref" if ref is None else ("same image" if np.array_equal(img, ref[0]) and f.ray_count == ref[1] else "IMAGE DIFFERS")
    if ref is None:
        ref = (img, f.ray_count)
    rates = []
    for i in range(a.frames):
        f = r.render_frame(cam, scene, want=())
        rates.append(f.ray_count / f.device_ms / 1e3)
    rates.sort()
    print(f"{a.workload} {a.renderer} {w}x{h} spp={spp} depth={depth} [{cfg or 'defaults'}]: median {rates[len(rates) // 2]:8.1f} best {rates[-1]:8.1f} Mrays/s  "
          f"({f.device_ms:.2f} ms, {f.kernel_launches} launches, {same})", flush=True)
    r.close()
