"""Golden vectors produced by the REFERENCE'S OWN CODE (tests/tools/make_golden.py: the unmodified
sources under /root/reference/src compiled against the API shims in oracle/refshim, Embree replaced by
the oracle's brute-force intersector). They pin the oracle's restatement of xorshift, camera, materials,
trace_ray and both renderers (seed mappings, fp16 quantisation, clamp, queue compaction, output bytes),
and the CUDA path is compared against the same stored outputs."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(f):
    return np.array([f], np.float32).view(np.uint32)[0]


def fa(v):
    return (C.c_float * len(v))(*[float(x) for x in v])


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLD, "reference_primitives.npz"))


@pytest.fixture(scope="module")
def rend():
    return np.load(os.path.join(GOLD, "reference_renders.npz"))


def test_oracle_xorshift_matches_reference(oracle, prim):
    L = oracle.lib()
    for i, s in enumerate(prim["xs_seeds"]):
        c = C.c_uint32(int(s))
        for k in range(16):
            f = L.orc_xorshift_next(C.byref(c))
            assert c.value == prim["xs_states"][i, k] and _bits(f) == _bits(prim["xs_floats"][i, k])
    for i, s in enumerate(prim["xs_seeds"][:4]):
        c, v = C.c_uint32(int(s)), (C.c_float * 3)()
        L.orc_random_unit_vector(C.byref(c), v)
        assert np.array_equal(np.array(v[:], np.float32).view(np.uint32), prim["ruv"][i, :3].view(np.uint32))
        assert c.value == prim["ruv"][i, 3:].view(np.uint32)[0]


def test_oracle_camera_matches_reference(oracle, prim):
    L = oracle.lib()
    for i, (w, h, pos, d, focal) in enumerate(golden_cases.CAMERAS):
        cam = oracle.camera(w, h, pos, d, focal)
        got = np.array(list(cam.center) + list(cam.pixel00_loc) + list(cam.pixel_delta_u) + list(cam.pixel_delta_v), np.float32)
        assert np.array_equal(got.view(np.uint32), prim["cam"][i].view(np.uint32))
        for j, (x, y, seed) in enumerate(golden_cases.CAMERA_PIXELS):
            c, org, dr = C.c_uint32(seed), (C.c_float * 3)(), (C.c_float * 3)()
            L.orc_camera_get_ray(C.byref(cam), x % w, y % h, C.byref(c), org, dr)
            want = prim["cam_rays"][i, j]
            assert np.array_equal(np.array(list(org) + list(dr), np.float32).view(np.uint32), want[:6].view(np.uint32))
            assert c.value == want[6:].view(np.uint32)[0]


def test_oracle_scatter_matches_reference(oracle, prim):
    L = oracle.lib()
    for k, (mtype, rough, ior, seed, dr, n, uv) in enumerate(golden_cases.scatter_inputs()):
        m = oracle.orc_material(mtype, -1, fa([0.5, 0.6, 0.7]), rough, ior, fa([0, 0, 0]))
        c, od, oa = C.c_uint32(seed), (C.c_float * 3)(), (C.c_float * 3)()
        ok = L.orc_material_scatter(C.byref(m), None, 0, C.byref(c), fa(dr), fa(n), fa(uv), od, oa)
        want = prim["scatter"][k]
        assert ok == int(want[0]), k
        assert c.value == want[7:].view(np.uint32)[0], k          # number of draws (0 / 1 / 3)
        if mtype:
            assert np.array_equal(np.array(list(od) + list(oa), np.float32).view(np.uint32), want[1:7].view(np.uint32)), k


@pytest.mark.parametrize("name", sorted(golden_cases.RENDER_CASES))
def test_oracle_renders_match_reference(oracle, rend, name):
    factory, w, h, depth, spp = golden_cases.RENDER_CASES[name]
    data = factory()
    osc, ocam = oracle.Scene(data), oracle.camera_for(data, w, h)
    for mode in (0, 1):
        o = osc.render(ocam, mode, depth, spp)
        assert o["ray_count"] == int(rend[f"{name}_m{mode}_rays"][0])
        assert np.array_equal(o["rgba8"], rend[f"{name}_m{mode}_rgba8"])


@pytest.mark.parametrize("name", sorted(golden_cases.RENDER_CASES))
def test_hostemu_of_kernel_source_matches_reference(pkg, hostemu, rend, name):
    factory, w, h, depth, spp = golden_cases.RENDER_CASES[name]
    data = factory()
    emu = hostemu.Scene(data)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    for kind in (0, 1):
        e = emu.render(cam, kind, depth, spp)
        assert e["ray_count"] == int(rend[f"{name}_m{kind}_rays"][0])
        assert np.array_equal(e["rgba8"], rend[f"{name}_m{kind}_rgba8"])


def test_fixtures_are_current_when_the_reference_is_present(rend):
    """in the build container: re-run the reference's code and compare with the committed fixtures"""
    import _refshim
    if not (_refshim.available() and os.path.isdir("/root/reference/src")):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for name in ("cube", "cornell"):
        factory, w, h, depth, spp = golden_cases.RENDER_CASES[name]
        rs = _refshim.Scene(factory())
        for mode in (0, 1):
            img, n = rs.render(mode, w, h, depth, spp)
            assert n == int(rend[f"{name}_m{mode}_rays"][0]) and np.array_equal(img, rend[f"{name}_m{mode}_rgba8"])
        rs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(golden_cases.RENDER_CASES))
def test_cuda_path_matches_reference_outputs(pkg, app, rend, name):
    """the CUDA path through the C ABI against what the reference's own renderers wrote to out.png"""
    factory, w, h, depth, spp = golden_cases.RENDER_CASES[name]
    data = factory()
    scene = pkg.Scene(app, data)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    for cls, mode in ((pkg.MegakernelRenderer, 0), (pkg.WavefrontRenderer, 1)):
        r = cls(app, (w, h), None, depth, spp)
        f = r.render_frame(cam, scene)
        assert f.ray_count == int(rend[f"{name}_m{mode}_rays"][0])
        assert np.array_equal(f.rgba8, rend[f"{name}_m{mode}_rgba8"])
        r.close()
    scene.close()
