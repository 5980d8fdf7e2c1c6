"""Tree quality without a GPU: node visits and triangle tests per ray of the production tree (the host emulation compiles
the build and traversal headers of csrc/, same tree as rt_scene_commit) against a SAH-binned binary tree collapsed the
same way (EMU_SAH=1, experiment switch of tests/hostemu), on a lattice of pixels of the workload's own camera.

    python tools/tree_quality.py stadium_scene c3_sponza_scale ...        # scene factory names or bench workloads
"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module("sycl-ray-tracer_b200")
scenes = importlib.import_module("sycl-ray-tracer_b200.scenes")
import _hostemu


def measure(data, w, h, spp, depth, grid=(24, 14)):
    L = _hostemu.lib()
    L.emu_pixel_costs.restype = C.c_uint32
    L.emu_pixel_costs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    emu = _hostemu.Scene(data)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    p = pkg._capi.rt_render_params()
    p.max_depth, p.sample_count = depth, spp
    cap = spp * depth
    nn, nt, fs = np.zeros(cap, np.uint16), np.zeros(cap, np.uint16), np.zeros(cap, np.uint8)
    rays = visits = tests = 0
    for j in range(grid[1]):
        for i in range(grid[0]):
            x, y = int((i + 0.5) * w / grid[0]), int((j + 0.5) * h / grid[1])
            n = L.emu_pixel_costs(emu.h, 0, C.addressof(cam.c), C.addressof(p), x, y, nn.ctypes.data, nt.ctypes.data, fs.ctypes.data, cap)
            rays += n; visits += int(nn[:n].sum()); tests += int(nt[:n].sum())
    out = dict(nodes=emu.node_count, depth=emu.depth, rays=rays, visits=visits / rays, tests=tests / rays)
    emu.close()
    return out


for name in sys.argv[1:] or ["stadium_scene"]:
    if name in bench.WORKLOADS:
        data, w, h, spp, depth = bench.build_scene_data(name)
    else:
        data, w, h, spp, depth = getattr(scenes, name)(), 1920, 1080, 0, 10
    res = {}
    for tag, env in (("production (Morton LBVH + SAH-optimal collapse)", {}), ("SAH-binned binary tree, same collapse", {"EMU_SAH": "1"})):
        for k in ("EMU_SAH", "EMU_SPLIT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        res[tag] = measure(data, w, h, 4, depth)
        r = res[tag]
        print(f"{name:22s} {data.triangle_count:9d} tris  {tag:48s} nodes {r['nodes']:8d} depth {r['depth']:2d}  "
              f"visits/ray {r['visits']:6.2f}  tests/ray {r['tests']:5.2f}  ({r['rays']} rays)", flush=True)
    a, b = list(res.values())
    print(f"{name:22s} ratio production / SAH: visits {a['visits'] / b['visits']:.2f}x, tests {a['tests'] / b['tests']:.2f}x")
