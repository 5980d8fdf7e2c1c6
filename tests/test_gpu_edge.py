"""Edge cases and error behaviour of the C ABI on the GPU."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tri_scene(pkg, n, rs):
    pos = rs.rand(n * 3, 3).astype(np.float32)
    if n >= 2:
        pos[3:6] = pos[0:3]   # duplicate triangle: tie -> lowest id
    if n >= 3:
        pos[8] = pos[7]       # zero-area triangle
    inst = pkg.InstanceData(pos, np.tile(np.float32([0, 0, 1]), (n * 3, 1)), np.zeros((n * 3, 2), np.float32),
                            np.arange(n * 3, dtype=np.uint32))
    return pkg.SceneData([inst], None, (0.5, 0.7, 1.0), (0.5, 0.5, 3), (0, 0, -1), 1.0)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 9, 25])
def test_tiny_and_degenerate_scenes(pkg, oracle, app, n):
    rs = np.random.RandomState(5 + n)
    data = _tri_scene(pkg, n, rs)
    scene = pkg.Scene(app, data)
    org = np.tile(np.float32([0.5, 0.5, 3]), (5000, 1)) + (rs.rand(5000, 3).astype(np.float32) - 0.5)
    d = np.float32([0, 0, -1]) + (rs.rand(5000, 3).astype(np.float32) - 0.5) * 0.6
    g, o = pkg.intersect(app, scene, org, d), oracle.Scene(data).intersect(org, d)
    for k in ("inst", "prim"):
        assert np.array_equal(g[k], o[k])
    assert np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
    if n == 0:
        assert (g["inst"] == -1).all() and np.isinf(g["t"]).all()
    cam = pkg.Camera((32, 24), data.camera_position, data.camera_direction, 1.0)
    for cls, mode in ((pkg.MegakernelRenderer, 0), (pkg.WavefrontRenderer, 1)):
        f = cls(app, (32, 24), None, 4, 2).render_frame(cam, scene)
        oo = oracle.Scene(data).render(oracle.camera_for(data, 32, 24), mode, 4, 2)
        assert np.array_equal(f.rgba8, oo["rgba8"]) and f.ray_count == oo["ray_count"]
    scene.close()


def test_tfar_and_tnear_limits(pkg, oracle, app, scenes):
    data = scenes.cube_scene()
    scene = pkg.Scene(app, data)
    rs = np.random.RandomState(2)
    org = np.zeros((4000, 3), np.float32)
    d = np.float32([0, 0, -1]) + (rs.rand(4000, 3).astype(np.float32) - 0.5) * 0.8
    osc = oracle.Scene(data)
    for tnear, tfar in ((1e-4, 2.5), (2.2, 1e9), (0.0, 1.95)):
        g, o = pkg.intersect(app, scene, org, d, tnear, tfar), osc.intersect(org, d, tnear, tfar)
        assert np.array_equal(g["prim"], o["prim"]) and np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
    assert pkg.intersect(app, scene, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))["t"].size == 0
    scene.close()


def test_depth_zero_and_odd_sizes(pkg, oracle, app, scenes):
    data = scenes.cube_scene()
    scene = pkg.Scene(app, data)
    osc = oracle.Scene(data)
    for (w, h) in ((1, 1), (7, 5), (33, 17), (250, 3)):   # ragged tiles / padding in the seed mapping
        cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, 1.0)
        for cls, mode in ((pkg.MegakernelRenderer, 0), (pkg.WavefrontRenderer, 1)):
            for depth in (0, 1, 8):
                f = cls(app, (w, h), None, depth, 2).render_frame(cam, scene)
                o = osc.render(oracle.camera_for(data, w, h), mode, depth, 2)
                assert f.ray_count == o["ray_count"]
                assert np.array_equal(f.rng_state, o["rng_state"]) and np.array_equal(f.rgba8, o["rgba8"])
    scene.close()


def test_error_behaviour(pkg, app, scenes):
    """the reference terminates on errors; the C ABI returns rt_status + message"""
    lib, cap = app._lib, pkg._capi
    data = scenes.cube_scene()
    s = pkg.Scene(app, data, commit=False)
    cam = pkg.Camera((16, 16), (0, 0, 0), (0, 0, -1), 1.0)
    r = pkg.MegakernelRenderer(app, (16, 16), None, 4, 1)
    with pytest.raises(pkg.RtError, match="not committed"):
        r.render_frame(cam, s)
    s.commit()
    with pytest.raises(pkg.RtError, match="already committed"):
        s.commit()
    r.render_frame(cam, s)
    with pytest.raises(pkg.RtError, match="image size"):
        r.render_frame(pkg.Camera((32, 16), (0, 0, 0), (0, 0, -1), 1.0), s)
    with pytest.raises(pkg.RtError, match="rank"):
        r.render_frame(cam, s, shard={"rank": 5, "world": 2, "tile_size": 32})
    with pytest.raises(pkg.RtError):
        pkg.MegakernelRenderer(app, (0, 16), None, 4, 1)
    bad = scenes.cube_scene()
    bad.instances[0].indices[5] = 999
    with pytest.raises(pkg.RtError, match="index out of range"):
        pkg.Scene(app, bad)
    bad = scenes.cube_scene()
    bad.instances[0].material.albedo_image = 3
    with pytest.raises(pkg.RtError, match="albedo_image"):
        pkg.Scene(app, bad)
    h = C.c_void_p()
    assert lib.rt_context_create(9999, C.byref(h)) != cap.RT_OK
    assert b"rt_context_create" in lib.rt_last_error(None)
    s.close()


@pytest.mark.parametrize("which", ["awkward_transforms", "coincident_centroids"])
def test_awkward_scenes(pkg, oracle, app, which):
    """mirrored / huge / tiny / far-away instance transforms; all-equal Morton keys in a flat scene"""
    import edge_scenes
    data, org, d = getattr(edge_scenes, which)()
    scene = pkg.Scene(app, data)
    g, o = pkg.intersect(app, scene, org, d), oracle.Scene(data).intersect(org, d, use_bvh=False)
    assert (o["inst"] >= 0).sum() > 500
    for k in ("inst", "prim"):
        assert np.array_equal(g[k], o[k])
    for k in ("t", "u", "v"):
        assert np.array_equal(g[k].view(np.uint32), o[k].view(np.uint32))
    cam = pkg.Camera((64, 40), data.camera_position, data.camera_direction, data.camera_focal_length)
    for cls, mode in ((pkg.MegakernelRenderer, 0), (pkg.WavefrontRenderer, 1)):
        f = cls(app, (64, 40), None, 6, 2).render_frame(cam, scene)
        oo = oracle.Scene(data).render(oracle.camera_for(data, 64, 40), mode, 6, 2, use_bvh=True)
        assert f.ray_count == oo["ray_count"] and np.array_equal(f.accum.view(np.uint32), oo["accum"].view(np.uint32))
    scene.close()
