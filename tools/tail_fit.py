"""Fixed cost of one render launch: per-rank time of world-way tile shards on ONE GPU for several spp, fitted as
T = a + b * pixels (a = ramp + tail + launch): python tools/tail_fit.py [workload] [renderer]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sycl-ray-tracer_b200")
wl = sys.argv[1] if len(sys.argv) > 1 else "c4_heightfield_10m"
kind = sys.argv[2] if len(sys.argv) > 2 else "megakernel"
data, w, h, spp0, depth = bench.build_scene_data(wl)
app = pkg.App(0)
scene = pkg.Scene(app, data)
cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
cls = pkg.MegakernelRenderer if kind == "megakernel" else pkg.WavefrontRenderer
for spp in (1, 4, 16, 64):
    r = cls(app, (w, h), None, depth, spp)
    row = []
    for world in (1, 2, 8, 32):
        ms = []
        for rank in range(min(world, 4)):
            sh = {"rank": rank, "world": world, "tile_size": 64} if world > 1 else None
            r.render_frame(cam, scene, want=(), shard=sh)
            ms.append(min(r.render_frame(cam, scene, want=(), shard=sh).device_ms for _ in range(2)))
        row.append((world, sum(ms) / len(ms)))
    t1, t32 = row[0][1], row[-1][1]
    b = (t1 - t32) / (1 - 1 / 32)
    a = t1 - b
    print(f"{wl} {kind} spp={spp}: " + "  ".join(f"1/{wd}: {t:.3f} ms" for wd, t in row) + f"   fit: fixed {a:.3f} ms + {b:.3f} ms per full frame", flush=True)
    r.close()
