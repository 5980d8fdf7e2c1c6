"""CPU-side check of the CUDA kernels' own source: tests/hostemu compiles the per-item headers
the kernels are made of (csrc/rt_build.h, rt_traverse.h, rt_shade.h, rt_wavefront.h) with g++ and
must agree bit for bit with the oracle. (Host emulation is test-only; the product has no CPU path.)"""
import numpy as np
import pytest


def _rays(n, seed, extent):
    rs = np.random.RandomState(seed)
    org = ((rs.rand(n, 3) - 0.5) * 2 * extent).astype(np.float32)
    d = (rs.rand(n, 3) - 0.5).astype(np.float32)
    d[::89, 2] = 0.0
    d[::103, :2] = 0.0
    return org, d


def _scene(scenes, which):
    return {"cube": lambda: scenes.cube_scene(), "soup": lambda: scenes.random_soup(900, 7, 1.0, 4),
            "cornell": lambda: scenes.cornell_scene(3), "field": lambda: scenes.big_mesh_scene(40),
            "sponza_small": lambda: scenes.sponza_scale_scene(32, 2, 9)}[which]()


EXT = {"cube": 4.0, "soup": 2.0, "cornell": 1.4, "field": 55.0, "sponza_small": 8.0}


@pytest.mark.parametrize("which", ["cube", "soup", "cornell", "field", "sponza_small"])
def test_tree_is_valid_and_traversal_matches_brute_force(oracle, hostemu, scenes, which):
    data = _scene(scenes, which)
    emu, orc = hostemu.Scene(data), oracle.Scene(data)
    assert emu.validate() == 0
    org, d = _rays(15000, 21, EXT[which])
    a, b = orc.intersect(org, d, use_bvh=False), emu.intersect(org, d)
    assert (a["inst"] >= 0).sum() > 50
    for k in ("inst", "prim"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("t", "u", "v"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k


@pytest.mark.parametrize("which", ["cube", "soup", "cornell", "sponza_small"])
@pytest.mark.parametrize("kind", [0, 1])
def test_render_matches_oracle(pkg, oracle, hostemu, scenes, which, kind):
    data = _scene(scenes, which)
    emu, orc = hostemu.Scene(data), oracle.Scene(data)
    w, h = 40, 24
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    ocam = oracle.camera_for(data, w, h)
    for f in ("center", "pixel00_loc", "pixel_delta_u", "pixel_delta_v"):  # rt_camera_init == oracle
        assert list(getattr(cam.c, f)) == list(getattr(ocam, f))
    e = emu.render(cam, kind, 7, 3)
    o = orc.render(ocam, kind, 7, 3, use_bvh=True)
    assert e["ray_count"] == o["ray_count"]
    assert np.array_equal(e["rng_state"], o["rng_state"])
    assert np.array_equal(e["accum"].view(np.uint32), o["accum"].view(np.uint32))
    assert np.array_equal(e["rgba8"], o["rgba8"])


def test_edge_scenes(pkg, oracle, hostemu, scenes):
    """empty scene, 1..4 triangles, degenerate and duplicated triangles"""
    rs = np.random.RandomState(5)
    for n in (0, 1, 2, 3, 4, 9):
        pos = rs.rand(n * 3, 3).astype(np.float32)
        if n >= 2:
            pos[3:6] = pos[0:3]            # exact duplicate of triangle 0: tie -> lowest id
        if n >= 3:
            pos[8] = pos[7]                # degenerate (zero area)
        inst = pkg.InstanceData(pos, np.tile(np.float32([0, 0, 1]), (n * 3, 1)), np.zeros((n * 3, 2), np.float32),
                                np.arange(n * 3, dtype=np.uint32))
        data = pkg.SceneData([inst], None, (0.5, 0.7, 1.0), (0.5, 0.5, 3), (0, 0, -1), 1.0)
        emu, orc = hostemu.Scene(data), oracle.Scene(data)
        assert emu.validate() == 0
        org = np.tile(np.float32([0.5, 0.5, 3]), (4000, 1)) + (rs.rand(4000, 3).astype(np.float32) - 0.5)
        d = np.float32([0, 0, -1]) + (rs.rand(4000, 3).astype(np.float32) - 0.5) * 0.6
        a, b = orc.intersect(org, d), emu.intersect(org, d)
        for k in ("inst", "prim"):
            assert np.array_equal(a[k], b[k]), (n, k)
        assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32)), n
        if n == 0:
            assert (b["inst"] == -1).all()


def test_depth_zero_and_salt(pkg, oracle, hostemu, scenes):
    data = scenes.cube_scene()
    emu, orc = hostemu.Scene(data), oracle.Scene(data)
    cam = pkg.Camera((24, 16), data.camera_position, data.camera_direction, data.camera_focal_length)
    ocam = oracle.camera_for(data, 24, 16)
    for kind in (0, 1):
        e, o = emu.render(cam, kind, 0, 3), orc.render(ocam, kind, 0, 3)   # max_depth 0: black, 2 draws per sample
        assert e["ray_count"] == 0 == o["ray_count"]
        assert np.array_equal(e["rng_state"], o["rng_state"]) and not e["rgba8"][..., :3].any()
        e = emu.render(cam, kind, 5, 2, shard={"rank": 0, "world": 1, "seed_salt": 0xABCDEF01})
        o = orc.render(ocam, kind, 5, 2, seed_salt=0xABCDEF01)
        assert np.array_equal(e["rng_state"], o["rng_state"])
        assert np.array_equal(e["accum"].view(np.uint32), o["accum"].view(np.uint32))


def test_tile_shards_sum_to_the_full_image(pkg, hostemu, scenes):
    data = scenes.cornell_scene(2)
    emu = hostemu.Scene(data)
    cam = pkg.Camera((48, 40), data.camera_position, data.camera_direction, data.camera_focal_length)
    for kind in (0, 1):
        full = emu.render(cam, kind, 6, 2)
        parts = [emu.render(cam, kind, 6, 2, shard={"rank": r, "world": 3, "tile_size": 16}) for r in range(3)]
        acc = sum(p["accum"] for p in parts)
        assert np.array_equal(acc.view(np.uint32), full["accum"].view(np.uint32))   # x + 0 is exact
        assert np.array_equal(sum(p["rgba8"].astype(np.uint32) for p in parts), full["rgba8"].astype(np.uint32))
        assert sum(p["ray_count"] for p in parts) == full["ray_count"]
        owned = [(p["accum"][..., 3] > 0) for p in parts]
        assert (sum(o.astype(int) for o in owned) == 1).all()                      # a partition of the pixels


def test_rays_leaving_axis_aligned_walls(oracle, hostemu, scenes):
    """Regression: a bounce ray that starts one ulp off an axis-aligned wall and grazes back into it
    (t ~ 1.3e-4 > tnear). The quantised slab arithmetic cancels there (|q*id|, |o| >> t), so the
    node test must be conservative by its own rounding error or the hit is culled."""
    data = scenes.cornell_scene(2)
    emu, orc = hostemu.Scene(data), oracle.Scene(data)
    rs = np.random.RandomState(9)
    n = 30000
    org = (rs.rand(n, 3).astype(np.float32) - 0.5) * 2
    d = (rs.rand(n, 3).astype(np.float32) - 0.5)
    axis, side = rs.randint(0, 3, n), rs.randint(0, 2, n) * 2 - 1
    for i in range(n):   # put the origin 0..2 ulp outside / inside a wall, direction barely toward it
        w = np.float32(side[i])
        org[i, axis[i]] = np.nextafter(w, np.float32(w * 2)) if i % 3 == 0 else (w if i % 3 == 1 else np.nextafter(w, np.float32(0)))
        d[i, axis[i]] = np.float32(-side[i] * rs.rand() * 2e-3)
    a, b = orc.intersect(org, d, use_bvh=False), emu.intersect(org, d)
    assert ((a["t"] < 1e-3) & (a["inst"] >= 0)).sum() > 100   # the case is actually exercised
    for k in ("inst", "prim"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))


def test_cornell_paths_regression(pkg, oracle, hostemu, scenes):
    data = scenes.cornell_scene(4)
    emu, orc = hostemu.Scene(data), oracle.Scene(data)
    cam = pkg.Camera((160, 90), data.camera_position, data.camera_direction, data.camera_focal_length)
    e = emu.render(cam, 0, 10, 8)
    o = orc.render(oracle.camera_for(data, 160, 90), 0, 10, 8, use_bvh=True)
    assert e["ray_count"] == o["ray_count"] and np.array_equal(e["rng_state"], o["rng_state"])


def test_awkward_transforms_and_scales(pkg, oracle, hostemu):
    import edge_scenes
    data, org, d = edge_scenes.awkward_transforms()
    emu, orc = hostemu.Scene(data), oracle.Scene(data)
    assert emu.validate() == 0
    a, b = orc.intersect(org, d, use_bvh=False), emu.intersect(org, d)
    assert (a["inst"] >= 0).sum() > 500
    for k in ("inst", "prim"):
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    cam = pkg.Camera((40, 30), data.camera_position, data.camera_direction, data.camera_focal_length)
    e = emu.render(cam, 0, 6, 2)
    o = orc.render(oracle.camera_for(data, 40, 30), 0, 6, 2, use_bvh=True)
    assert e["ray_count"] == o["ray_count"] and np.array_equal(e["accum"].view(np.uint32), o["accum"].view(np.uint32))


def test_coincident_centroids_and_flat_scene(oracle, hostemu):
    import edge_scenes
    data, org, d = edge_scenes.coincident_centroids()
    emu, orc = hostemu.Scene(data), oracle.Scene(data)
    assert emu.validate() == 0
    a, b = orc.intersect(org, d), emu.intersect(org, d)
    assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    assert (a["prim"] >= 0).mean() > 0.5


@pytest.mark.parametrize("size", [(200, 120), (256, 256), (33, 7), (8, 4), (1920, 1080)])
def test_megakernel_block_enumeration_covers_owned_pixels_once(hostemu, size):
    """rt_blocks.h: the pixel slots the megakernel hands out cover every pixel of the rank exactly once —
    unsharded and for image tiles (only the rank's own tiles are enumerated), partial edge tiles included"""
    import ctypes as C
    L = hostemu.lib()
    L.emu_enumerate_blocks.restype = C.c_uint32
    L.emu_enumerate_blocks.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    w, h = size
    for world, ts in [(1, 0), (2, 16), (4, 32), (8, 64), (3, 8), (8, 0), (5, 128)]:
        total = np.zeros((h, w), np.uint32)
        for rank in range(world if ts else 1):
            counts, foreign = np.zeros((h, w), np.uint32), C.c_uint32()
            n = L.emu_enumerate_blocks(w, h, rank, world, ts, counts.ctypes.data, C.byref(foreign))
            assert foreign.value == 0 and counts.max() <= 1
            if ts and world > 1:
                tiles = (np.arange(h)[:, None] // ts) * ((w + ts - 1) // ts) + np.arange(w)[None, :] // ts
                assert np.array_equal(counts == 1, tiles % world == rank)
                assert n % ((ts // 8) * (ts // 4)) == 0
            total += counts
        assert (total == 1).all()


@pytest.mark.parametrize("kind", [0, 1])
def test_progressive_resume_on_the_cpu(pkg, oracle, hostemu, scenes, kind):
    """RT_RENDER_RESUME in the kernels' own per-pixel code (rt_megakernel_pixel, rt_wf_generate_pixel):
    2 + 3 + 1 samples continue the streams and sums of the previous frame and equal one 6-sample frame,
    which equals the oracle"""
    data = scenes.cornell_scene(2)
    emu = hostemu.Scene(data)
    w, h = 40, 24
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    one = emu.render(cam, kind, 7, 6)
    o = oracle.Scene(data).render(oracle.camera_for(data, w, h), kind, 7, 6, use_bvh=True)
    assert np.array_equal(one["accum"].view(np.uint32), o["accum"].view(np.uint32)) and one["ray_count"] == o["ray_count"]
    for shard in (None, {"rank": 1, "world": 2, "tile_size": 8}):
        full = emu.render(cam, kind, 7, 6, shard=shard)
        f, rays = None, 0
        for n in (2, 3, 1):
            f = emu.render(cam, kind, 7, n, shard=shard, resume=f)
            rays += f["ray_count"]
        assert rays == full["ray_count"]
        assert np.array_equal(f["accum"].view(np.uint32), full["accum"].view(np.uint32))
        assert np.array_equal(f["rgba8"], full["rgba8"]) and np.array_equal(f["rng_state"], full["rng_state"])


@pytest.mark.parametrize("split", [False, True])
def test_large_triangles_next_to_dense_detail(monkeypatch, oracle, hostemu, scenes, split):
    """the BVH-hostile input (stadium: 100 m quads and 60 m diagonal strips next to finely tessellated props): the size-class
    bit of the Morton key (default) and the optional early split of large triangles into references (RT_SPLIT=1) change the
    TREE only — the tree stays structurally valid and every hit equals the brute-force oracle's, bit for bit"""
    if split:
        monkeypatch.setenv("RT_SPLIT", "1")
    else:
        monkeypatch.delenv("RT_SPLIT", raising=False)
    data = scenes.stadium_scene(n_props=6, prop_subdiv=2, n_beams=24)
    emu, orc = hostemu.Scene(data), oracle.Scene(data)
    assert emu.validate() == 0
    rs = np.random.RandomState(3)
    org = np.stack([rs.uniform(-45, 45, 20000), rs.uniform(0.5, 18, 20000), rs.uniform(-45, 45, 20000)], 1).astype(np.float32)
    d = (rs.rand(20000, 3) - 0.5).astype(np.float32)
    a, b = orc.intersect(org, d, use_bvh=False), emu.intersect(org, d)
    assert 0.5 < (a["inst"] >= 0).mean() < 0.9   # the stadium is open to the sky
    for k in ("inst", "prim"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("t", "u", "v"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    if split:  # the references really exist: more leaf records than triangles
        monkeypatch.delenv("RT_SPLIT")
        plain = hostemu.Scene(data)
        assert emu.node_count != plain.node_count


def test_texel_byte_to_unit_is_the_ieee_quotient_for_all_bytes(hostemu):
    """rt_u8_to_unit (csrc/rt_hd.h: v * RN(1/255) + one FMA residual + one FMA correction, what the kernels use instead of the
    IEEE division sequence) equals v / 255.0f, correctly rounded, for every byte"""
    import ctypes as C
    L = hostemu.lib()
    L.emu_u8_to_unit.restype = C.c_float
    L.emu_u8_to_unit.argtypes = [C.c_int]
    got = np.array([L.emu_u8_to_unit(v) for v in range(256)], np.float32)
    want = np.arange(256, dtype=np.float32) / np.float32(255.0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_ray_setup_shares_the_dominant_axis_reciprocal(hostemu):
    """rt_ray_pre (csrc/rt_traverse.h) takes 1 / dir[kz] of the triangle shear from the slab reciprocals instead of dividing
    again: the six values equal the plainly divided ones bit for bit, also for axis-aligned, tiny and zero components"""
    import ctypes as C
    L = hostemu.lib()
    L.emu_ray_pre.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    rs = np.random.RandomState(3)
    dirs = (rs.rand(4000, 3).astype(np.float32) - np.float32(0.5))
    dirs[::7] = dirs[::7].astype(np.float16).astype(np.float32)          # what the kernels really see (F6)
    dirs[::11, 0] = 0.0
    dirs[::13, 1] = -0.0
    dirs[::17, :2] = 0.0
    dirs[::19] *= np.float32(1e-22)
    dirs[5] = (0.0, 0.0, 0.0)
    tiny = np.float32(1e-20)
    one = np.float32(1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        for d in dirs:
            out, code = np.zeros(7, np.float32), np.zeros(1, np.uint32)
            L.emu_ray_pre(d.ctypes.data, out.ctypes.data, code.ctypes.data)
            kx, ky, kz = int(code[0] & 3), int((code[0] >> 2) & 3), int((code[0] >> 4) & 3)
            a = np.abs(d)
            assert a[kz] == a.max() and sorted((kx, ky, kz)) == [0, 1, 2]
            c = np.where(a > tiny, d, np.copysign(tiny, d)).astype(np.float32)
            want = np.array([one / c[0], one / c[1], one / c[2], -(d[kx] / d[kz]), -(d[ky] / d[kz]), one / d[kz]], np.float32)
            assert np.array_equal(out[:6].view(np.uint32), want.view(np.uint32)), (d, out[:6], want)
