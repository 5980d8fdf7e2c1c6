"""Tile-shard balance on ONE GPU: render every rank's shard of a world-N split in turn and compare with 1/N of the
unsharded frame: python tools/tile_probe.py [workload] [world] [tile_size] [renderer]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sycl-ray-tracer_b200")
wl = sys.argv[1] if len(sys.argv) > 1 else "c4_heightfield_10m"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ts = int(sys.argv[3]) if len(sys.argv) > 3 else 64
kind = sys.argv[4] if len(sys.argv) > 4 else "megakernel"
data, w, h, spp, depth = bench.build_scene_data(wl)
app = pkg.App(0)
scene = pkg.Scene(app, data)
cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
r = (pkg.MegakernelRenderer if kind == "megakernel" else pkg.WavefrontRenderer)(app, (w, h), None, depth, spp)
for _ in range(2):
    f = r.render_frame(cam, scene, want=())
full_ms, full_rays = f.device_ms, f.ray_count
print(f"{wl} {kind} unsharded: {full_ms:.3f} ms, {full_rays} rays; ideal per rank at world {world}: {full_ms / world:.3f} ms", flush=True)
ms, rays = [], []
for rank in range(world):
    sh = {"rank": rank, "world": world, "tile_size": ts}
    r.render_frame(cam, scene, want=(), shard=sh)
    f = r.render_frame(cam, scene, want=(), shard=sh)
    ms.append(f.device_ms); rays.append(f.ray_count)
print("tile", ts, "per-rank ms:", " ".join(f"{m:.2f}" for m in ms))
print("per-rank Mrays:", " ".join(f"{x / 1e6:.1f}" for x in rays))
print(f"max {max(ms):.3f} mean {sum(ms) / world:.3f}  -> speed-up at world {world}: {full_ms / max(ms):.2f}x; ray imbalance max/mean {max(rays) * world / sum(rays):.3f}")
