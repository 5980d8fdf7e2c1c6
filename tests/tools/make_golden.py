"""Generate tests/golden/*.npz from the REFERENCE ITSELF: felipeagc/sycl-ray-tracer's own unmodified
sources (xorshift / camera / material / trace_ray / MegakernelRenderer / WavefrontRenderer), compiled in
place from /root/reference against the API shims in oracle/refshim (make -C oracle ref) and run on the
CPU with the oracle's brute-force closest-hit search standing in for Embree's rtcIntersect1.

Run in the build container only (needs /root/reference):   python tests/tools/make_golden.py
The fixtures it writes are committed; tests/test_golden_reference.py and the GPU suite consume them
without the reference being present."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle  # noqa: E402
import _refshim  # noqa: E402
import golden_cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def fa(v):
    return (C.c_float * len(v))(*[float(x) for x in v])


def primitives():
    L = _refshim.lib()
    scenes = importlib.import_module("sycl-ray-tracer_b200.scenes")
    rs = _refshim.Scene(scenes.cube_scene())
    out = {}
    # xorshift streams (src/xorshift.hpp:11-20)
    seeds = np.array([1, 2463534242, 1081, 1921, 0, 0xFFFFFF80, 12345, 0x9E3779B9], np.uint32)
    st, fl = np.zeros((len(seeds), 16), np.uint32), np.zeros((len(seeds), 16), np.float32)
    for i, s in enumerate(seeds):
        c = C.c_uint32(int(s))
        for k in range(16):
            fl[i, k] = L.ref_xorshift_next(C.byref(c))
            st[i, k] = c.value
    out.update(xs_seeds=seeds, xs_states=st, xs_floats=fl)
    # random_unit_vector (draw order x, y, z; no rejection)
    ruv = np.zeros((len(seeds), 4), np.float32)
    for i, s in enumerate(seeds[:4]):
        c, v = C.c_uint32(int(s)), (C.c_float * 3)()
        L.ref_random_unit_vector(C.byref(c), v)
        ruv[i, :3] = v[:]
        ruv[i, 3] = np.float32(c.value).view(np.float32) if False else 0
        st_after = c.value
        ruv[i, 3] = np.array([st_after], np.uint32).view(np.float32)[0]
    out["ruv"] = ruv
    # camera + primary rays (src/camera.hpp:74-131)
    cams = golden_cases.CAMERAS
    cam_out, rays = np.zeros((len(cams), 12), np.float32), []
    for i, (w, h, pos, d, focal) in enumerate(cams):
        o = (C.c_float * 12)()
        L.ref_camera(w, h, fa(pos), fa(d), focal, o)
        cam_out[i] = o[:]
        for (x, y, seed) in golden_cases.CAMERA_PIXELS:
            c, org, dr = C.c_uint32(seed), (C.c_float * 3)(), (C.c_float * 3)()
            L.ref_camera_get_ray(w, h, fa(pos), fa(d), focal, x % w, y % h, C.byref(c), org, dr)
            rays.append(list(org) + list(dr) + [np.array([c.value], np.uint32).view(np.float32)[0]])
    out.update(cam=cam_out, cam_rays=np.array(rays, np.float32).reshape(len(cams), -1, 7))
    # Material::scatter (src/material.hpp:68-238) on seeded random inputs
    recs = []
    for (mtype, rough, ior, seed, dr, n, uv) in golden_cases.scatter_inputs():
        m = _oracle.orc_material(mtype, -1, fa([0.5, 0.6, 0.7]), rough, ior, fa([0, 0, 0]))
        c, od, oa = C.c_uint32(seed), (C.c_float * 3)(), (C.c_float * 3)()
        ok = L.ref_material_scatter(rs.h, C.byref(m), C.byref(c), fa(dr), fa(n), fa(uv), od, oa)
        recs.append([float(ok)] + list(od) + list(oa) + [np.array([c.value], np.uint32).view(np.float32)[0]])
    out["scatter"] = np.array(recs, np.float32)
    rs.close()
    return out


def renders():
    out = {}
    for name, (factory, w, h, depth, spp) in golden_cases.RENDER_CASES.items():
        data = factory()
        rs = _refshim.Scene(data)
        for mode in (0, 1):
            img, n = rs.render(mode, w, h, depth, spp)
            out[f"{name}_m{mode}_rgba8"] = img
            out[f"{name}_m{mode}_rays"] = np.array([n], np.uint64)
            print(f"{name} mode {mode}: {w}x{h} depth {depth} spp {spp}: {n} rtcIntersect1 calls")
        rs.close()
    return out


if __name__ == "__main__":
    assert _refshim.available(), "build oracle/_ref first: make -C oracle ref (needs /root/reference)"
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "reference_primitives.npz"), **primitives())
    np.savez_compressed(os.path.join(OUT, "reference_renders.npz"), **renders())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
