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


def test_closest_hit_against_exact_rational_arithmetic(oracle, scenes):
    """Embree's contract is "the closest hit". Independent check of the oracle's brute-force intersector (the
    definition everything else is compared with): ray/triangle intersection in EXACT rational arithmetic on the
    same float inputs (single instance, identity transform, so the world-space vertices are the inputs). The
    oracle must name the exactly-closest triangle whenever the exact runner-up is not within 1e-6 relative, report
    t within 1e-5 relative and barycentrics within 1e-4, and agree on hit / miss away from edges."""
    from fractions import Fraction as F
    data = scenes.random_soup(60, 21, 1.0, 1)
    inst = data.instances[0]
    assert np.allclose(np.asarray(inst.transform, np.float64).reshape(4, 4), np.eye(4))
    P = np.asarray(inst.positions, np.float32).reshape(-1, 3)[np.asarray(inst.indices).reshape(-1, 3)]   # (tris, 3, 3)
    rs = np.random.RandomState(3)
    org = ((rs.rand(300, 3) - 0.5) * 4).astype(np.float32)
    d = (rs.rand(300, 3) - 0.5).astype(np.float32)
    w = rs.dirichlet((1, 1, 1), 200).astype(np.float32)                # 200 of the rays aim at a point inside a triangle
    aim = (P[rs.randint(0, len(P), 200)] * w[:, :, None]).sum(1)
    d[:200] = (aim - org[:200]) * rs.uniform(0.3, 3.0, (200, 1)).astype(np.float32)
    got = oracle.Scene(data).intersect(org, d, use_bvh=False)

    def fr(v):
        return [F(float(x)) for x in v]

    def sub(a, b):
        return [a[i] - b[i] for i in range(3)]

    def cross(a, b):
        return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]

    def dot(a, b):
        return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]

    tris = [[fr(v) for v in t] for t in P]
    tnear = F(1, 10000)
    checked = 0
    for r in range(300):
        o, dd = fr(org[r]), fr(d[r])
        hits = []
        for k, (v0, v1, v2) in enumerate(tris):
            e1, e2 = sub(v1, v0), sub(v2, v0)
            pv = cross(dd, e2)
            det = dot(e1, pv)
            if det == 0:
                continue
            tv = sub(o, v0)
            u = dot(tv, pv) / det
            qv = cross(tv, e1)
            v = dot(dd, qv) / det
            t = dot(e2, qv) / det
            margin = min(u, v, 1 - u - v)
            if t > tnear and margin >= 0:
                hits.append((t, k, u, v, margin))
        hits.sort()
        if not hits:
            continue                                     # exact misses are checked below
        t0, k0, u0, v0_, margin0 = hits[0]
        if margin0 < F(1, 100000) or (len(hits) > 1 and (hits[1][0] - t0) <= t0 * F(1, 1000000)):
            continue                                     # on an edge, or an exact near-tie: either answer is legitimate
        checked += 1
        assert got["inst"][r] == 0 and got["prim"][r] == k0, (r, got["prim"][r], k0)
        assert abs(F(float(got["t"][r])) - t0) <= t0 * F(1, 100000)
        assert abs(F(float(got["u"][r])) - u0) <= F(1, 10000) and abs(F(float(got["v"][r])) - v0_) <= F(1, 10000)
    assert checked > 150
    # exact misses: no triangle within the ray at all -> the oracle must miss too
    misses = 0
    for r in range(300):
        o, dd = fr(org[r]), fr(d[r])
        any_near = False
        for (v0, v1, v2) in tris:
            e1, e2 = sub(v1, v0), sub(v2, v0)
            pv = cross(dd, e2)
            det = dot(e1, pv)
            if det == 0:
                any_near = True
                break
            tv = sub(o, v0)
            u = dot(tv, pv) / det
            v = dot(dd, cross(tv, e1)) / det
            t = dot(e2, cross(tv, e1)) / det
            if t > 0 and min(u, v, 1 - u - v) > F(-1, 100000):
                any_near = True
                break
        if not any_near:
            misses += 1
            assert got["inst"][r] == -1, r
    assert misses > 20
