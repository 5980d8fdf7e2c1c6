"""The oracle's CPU BVH (the stand-in for Embree in timed runs and big scenes) must return
exactly what its brute-force intersector returns."""
import numpy as np
import pytest


def _rays(n, seed, extent):
    rs = np.random.RandomState(seed)
    org = ((rs.rand(n, 3) - 0.5) * 2 * extent).astype(np.float32)
    d = (rs.rand(n, 3) - 0.5).astype(np.float32)
    d[::97, 0] = 0.0          # axis-parallel directions
    d[::101, 1:] = 0.0
    return org, d


@pytest.mark.parametrize("which", ["cube", "soup", "cornell", "field"])
def test_bvh_equals_brute_force(oracle, scenes, which):
    data = {"cube": lambda: scenes.cube_scene(), "soup": lambda: scenes.random_soup(700, 5, 1.0, 3),
            "cornell": lambda: scenes.cornell_scene(3), "field": lambda: scenes.big_mesh_scene(48)}[which]()
    ext = {"cube": 4.0, "soup": 2.0, "cornell": 1.4, "field": 55.0}[which]
    sc = oracle.Scene(data)
    org, d = _rays(20000, 11, ext)
    a, b = sc.intersect(org, d, use_bvh=False), sc.intersect(org, d, use_bvh=True)
    assert (a["inst"] >= 0).sum() > 100
    for k in ("inst", "prim"):
        assert np.array_equal(a[k], b[k])
    for k in ("t", "u", "v"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32))


def test_render_bvh_equals_brute_force(oracle, scenes):
    data = scenes.cornell_scene(2)
    sc = oracle.Scene(data)
    cam = oracle.camera_for(data, 40, 30)
    for mode in (oracle.MODE_MEGAKERNEL, oracle.MODE_WAVEFRONT):
        a, b = sc.render(cam, mode, 6, 3, use_bvh=False), sc.render(cam, mode, 6, 3, use_bvh=True)
        assert a["ray_count"] == b["ray_count"]
        assert np.array_equal(a["rng_state"], b["rng_state"])
        assert np.array_equal(a["accum"].view(np.uint32), b["accum"].view(np.uint32))


def test_crop_equals_full(oracle, scenes):
    data = scenes.cube_scene()
    sc = oracle.Scene(data)
    cam = oracle.camera_for(data, 64, 48)
    full = sc.render(cam, oracle.MODE_MEGAKERNEL, 8, 2)
    crop = sc.render(cam, oracle.MODE_MEGAKERNEL, 8, 2, crop=(10, 8, 42, 40))
    assert np.array_equal(full["rgba8"][8:40, 10:42], crop["rgba8"])
    assert np.array_equal(full["rng_state"][8:40, 10:42], crop["rng_state"])


def test_renderer_modes_differ_only_by_seed_and_clamp(oracle, scenes):
    """F3/F9: a bright sky makes samples exceed 1; wavefront clamps each sample, megakernel does not."""
    data = scenes.cube_scene()
    data.sky_color = (3.0, 3.0, 3.0)
    sc = oracle.Scene(data)
    cam = oracle.camera_for(data, 32, 32)
    m = sc.render(cam, oracle.MODE_MEGAKERNEL, 8, 2)
    w = sc.render(cam, oracle.MODE_WAVEFRONT, 8, 2)
    assert m["accum"][..., :3].max() > 2.0 + 1e-3   # unclamped sum of two samples
    assert w["accum"][..., :3].max() <= 2.0          # two samples, each clamped to 1
    assert m["rgba8"][0, 0, 0] == 255 and w["rgba8"][0, 0, 0] == 255
