"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle.

Tolerances (north_star): primary rays — (instID, primID) and barycentrics agree on >= 99.99 % of
pixels, t within 1e-4 relative; full paths — per-pixel linear mean within RMSE 2e-3 at 1024 spp,
xorshift streams bit-exact. The arithmetic contract actually makes the comparison bit-exact, so the
tests assert equality first and report the tolerance figures alongside."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rays(n, seed, extent):
    rs = np.random.RandomState(seed)
    org = ((rs.rand(n, 3) - 0.5) * 2 * extent).astype(np.float32)
    d = (rs.rand(n, 3) - 0.5).astype(np.float32)
    d[::89, 2] = 0.0
    d[::103, :2] = 0.0
    return org, d


def _check_hits(a, b, min_frac=0.9999):
    ids = (a["inst"] == b["inst"]) & (a["prim"] == b["prim"])
    hit = a["inst"] >= 0
    rel_t = np.abs(a["t"][hit & ids] - b["t"][hit & ids]) <= 1e-4 * np.abs(a["t"][hit & ids])
    bary = (np.abs(a["u"] - b["u"]) <= 1e-4) & (np.abs(a["v"] - b["v"]) <= 1e-4)
    ok = ids & bary
    assert ok.mean() >= min_frac, f"only {ok.mean():.6f} of rays agree"
    assert rel_t.all()
    return ok.mean()


def _render_pair(pkg, oracle, app, data, w, h, depth, spp, kind, crop=None, use_bvh=True):
    cls = pkg.MegakernelRenderer if kind == 0 else pkg.WavefrontRenderer
    scene = pkg.Scene(app, data)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    r = cls(app, (w, h), None, depth, spp)
    f = r.render_frame(cam, scene)
    o = oracle.Scene(data).render(oracle.camera_for(data, w, h), kind, depth, spp, use_bvh=use_bvh, crop=crop)
    r.close()
    scene.close()
    if crop:
        x0, y0, x1, y1 = crop
        sub = lambda a: a[y0:y1, x0:x1]
    else:
        sub = lambda a: a
    return f, o, sub


def test_c1_primary_rays_cube_256(pkg, oracle, app, scenes):
    """config 1: assets/cube.glb stand-in, 256x256: primary-hit AOVs against the brute-force oracle"""
    data = scenes.cube_scene()
    scene = pkg.Scene(app, data)
    ocam = oracle.camera_for(data, 256, 256)
    org, d = oracle.primary_rays(ocam, oracle.MODE_MEGAKERNEL, 256, 256)
    g = pkg.intersect(app, scene, org, d)
    o = oracle.Scene(data).intersect(org, d, use_bvh=False)
    frac = _check_hits(o, g)
    assert (g["inst"] >= 0).mean() > 0.2 and (g["inst"] < 0).mean() > 0.2   # both hits and sky
    assert frac == 1.0
    for k in ("t", "u", "v"):
        assert np.array_equal(o[k].view(np.uint32), g[k].view(np.uint32))
    scene.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_c1_full_frame_cube_256(pkg, oracle, app, scenes, kind):
    """config 1 proper: 256x256, 1 spp, depth 8 — image, accumulation, streams, ray count"""
    f, o, _ = _render_pair(pkg, oracle, app, scenes.cube_scene(), 256, 256, 8, 1, kind, use_bvh=False)
    assert f.ray_count == o["ray_count"]
    assert np.array_equal(f.rng_state, o["rng_state"])
    assert np.array_equal(f.accum.view(np.uint32), o["accum"].view(np.uint32))
    assert np.array_equal(f.rgba8, o["rgba8"])


@pytest.mark.parametrize("kind", [0, 1])
def test_full_paths_1024spp_rmse(pkg, oracle, app, scenes, kind):
    """north_star: per-pixel mean within RMSE tolerance at 1024 spp, bit-exact xorshift streams"""
    data = scenes.cube_scene()
    w = h = 48
    f, o, _ = _render_pair(pkg, oracle, app, data, w, h, 8, 1024, kind, use_bvh=False)
    mean_g, mean_o = f.accum[..., :3] / 1024.0, o["accum"][..., :3] / 1024.0
    rmse = float(np.sqrt(np.mean((mean_g - mean_o) ** 2)))
    assert rmse <= 2e-3, rmse
    assert np.array_equal(f.rng_state, o["rng_state"])       # streams bit-exact
    assert rmse == 0.0 and f.ray_count == o["ray_count"]


@pytest.mark.parametrize("which", ["soup", "cornell", "field", "sponza_small"])
def test_intersect_matches_brute_force(pkg, oracle, app, scenes, which):
    data = {"soup": lambda: scenes.random_soup(900, 7, 1.0, 4), "cornell": lambda: scenes.cornell_scene(3),
            "field": lambda: scenes.big_mesh_scene(40), "sponza_small": lambda: scenes.sponza_scale_scene(32, 2, 9)}[which]()
    ext = {"soup": 2.0, "cornell": 1.4, "field": 55.0, "sponza_small": 8.0}[which]
    scene = pkg.Scene(app, data)
    org, d = _rays(60000, 31, ext)
    g = pkg.intersect(app, scene, org, d)
    o = oracle.Scene(data).intersect(org, d, use_bvh=False)
    assert _check_hits(o, g) == 1.0
    assert np.array_equal(o["t"].view(np.uint32), g["t"].view(np.uint32))
    scene.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_c2_cornell_materials(pkg, oracle, app, scenes, kind):
    """config 2's scene (diffuse / metallic / dielectric / emissive), reduced size so the oracle finishes"""
    f, o, _ = _render_pair(pkg, oracle, app, scenes.cornell_scene(4), 160, 90, 10, 8, kind)
    assert f.ray_count == o["ray_count"]
    assert np.array_equal(f.rng_state, o["rng_state"])
    assert np.array_equal(f.accum.view(np.uint32), o["accum"].view(np.uint32))
    assert np.array_equal(f.rgba8, o["rgba8"])
    assert f.accum[..., :3].max() > 0.5   # the emissive ceiling quad lights the box (not an all-black frame)


@pytest.mark.parametrize("kind", [0, 1])
def test_c2_cornell_full_size_crop(pkg, oracle, app, scenes, kind):
    """config 2 at its full 1920x1080 size on the GPU; the oracle checks a crop window (seeds depend
    on the full image size)"""
    crop = (900, 620, 980, 660)
    f, o, sub = _render_pair(pkg, oracle, app, scenes.cornell_scene(5), 1920, 1080, 10, 4, kind, crop=crop)
    assert np.array_equal(sub(f.rng_state), o["rng_state"])
    assert np.array_equal(sub(f.accum).view(np.uint32), o["accum"].view(np.uint32))
    assert np.array_equal(sub(f.rgba8), o["rgba8"])


@pytest.mark.parametrize("kind", [0, 1])
def test_c3_sponza_scale_full_size_crop(pkg, oracle, app, scenes, kind):
    """config 3's ~261 k-triangle textured scene at 1080p; crop checked by the oracle"""
    data = scenes.sponza_scale_scene()
    assert 255000 < data.triangle_count < 265000
    crop = (940, 560, 1004, 592)
    f, o, sub = _render_pair(pkg, oracle, app, data, 1920, 1080, 10, 2, kind, crop=crop)
    assert np.array_equal(sub(f.rng_state), o["rng_state"])
    assert np.array_equal(sub(f.accum).view(np.uint32), o["accum"].view(np.uint32))
    assert np.array_equal(sub(f.rgba8), o["rgba8"])


def test_textures_match_oracle(pkg, oracle, app, scenes):
    """layered texture object (point, repeat computed in code) vs the oracle's sampler restatement"""
    tex = scenes.procedural_textures(3, 1234)
    p, n, uv, i = scenes.grid_mesh(4, (-2, -2, -3), (4, 0, 0), (0, 4, 0), (0, 0, 1), 3.7)
    uv = uv - 1.3   # negative and > 1 coordinates exercise the repeat rule
    insts = [pkg.InstanceData(p, n, uv, i, None, pkg.Material.diffuse(image=2)),
             pkg.InstanceData(p, n, uv * 0.5, i, scenes.trs((0.5, 0.2, 0.5), (0.3, 0.3, 0.3)), pkg.Material.metallic(roughness=0.3, image=1))]
    data = pkg.SceneData(insts, tex, (0.5, 0.7, 1.0), (0, 0, 0), (0, 0, -1), 1.0)
    for kind in (0, 1):
        f, o, _ = _render_pair(pkg, oracle, app, data, 96, 64, 6, 3, kind, use_bvh=False)
        assert np.array_equal(f.accum.view(np.uint32), o["accum"].view(np.uint32))
        assert np.array_equal(f.rng_state, o["rng_state"])


def test_determinism_and_ray_count(pkg, app, scenes):
    """the reference's ray counts are identical run to run (benchmark_raw.csv:2-6)"""
    data = scenes.sponza_scale_scene(64, 3, 12)
    scene = pkg.Scene(app, data)
    cam = pkg.Camera((640, 360), data.camera_position, data.camera_direction, data.camera_focal_length)
    for cls in (pkg.MegakernelRenderer, pkg.WavefrontRenderer):
        r = cls(app, (640, 360), None, 10, 4)
        a, b = r.render_frame(cam, scene), r.render_frame(cam, scene)
        assert a.ray_count == b.ray_count > 640 * 360 * 4
        assert np.array_equal(a.rgba8, b.rgba8) and np.array_equal(a.rng_state, b.rng_state)
        r.close()
    scene.close()


def test_megakernel_and_wavefront_differ_only_by_seed_and_clamp(pkg, app, scenes):
    """F3 + F9: pixel (0,0) has seed 0 in both mappings, so it is the same path in both renderers;
    with a dark sky no sample exceeds 1 and the two must agree there exactly."""
    data = scenes.cube_scene()
    scene = pkg.Scene(app, data)
    cam = pkg.Camera((64, 64), data.camera_position, data.camera_direction, data.camera_focal_length)
    m = pkg.MegakernelRenderer(app, (64, 64), None, 8, 16).render_frame(cam, scene)
    w = pkg.WavefrontRenderer(app, (64, 64), None, 8, 16).render_frame(cam, scene)
    assert np.array_equal(m.accum[0, 0], w.accum[0, 0]) and m.rng_state[0, 0] == w.rng_state[0, 0] == 0
    # seeds agree on column 0 only when H_pad == ... x*H_pad + y == x + y*W  <=> x == y (W == H_pad == 64)
    diag = np.arange(64)
    assert np.array_equal(m.rng_state[diag, diag], w.rng_state[diag, diag])
    assert np.array_equal(m.accum[diag, diag], w.accum[diag, diag])
    assert not np.array_equal(m.rng_state, w.rng_state)
    scene.close()


def test_tile_shards_sum_to_unsharded(pkg, app, scenes):
    """image-tile sharding (config 4's mode): the sum over ranks is bit-identical to one GPU"""
    data = scenes.cornell_scene(3)
    scene = pkg.Scene(app, data)
    cam = pkg.Camera((200, 120), data.camera_position, data.camera_direction, data.camera_focal_length)
    for cls in (pkg.MegakernelRenderer, pkg.WavefrontRenderer):
        r = cls(app, (200, 120), None, 8, 3)
        full = r.render_frame(cam, scene)
        full_acc, full_img, rays = full.accum.copy(), full.rgba8.copy(), full.ray_count
        acc, img, n = np.zeros_like(full_acc), np.zeros(full_img.shape, np.uint32), 0
        for rank in range(4):
            f = r.render_frame(cam, scene, shard={"rank": rank, "world": 4, "tile_size": 32})
            acc += f.accum
            img += f.rgba8
            n += f.ray_count
        assert np.array_equal(acc.view(np.uint32), full_acc.view(np.uint32))
        assert np.array_equal(img, full_img.astype(np.uint32)) and n == rays
        r.close()
    scene.close()


@pytest.mark.parametrize("kind", ["megakernel", "wavefront"])
def test_drain_and_refill_knobs_change_nothing(pkg, app, scenes, monkeypatch, kind):
    """the drain carry-over (RT_TUNE_CARRY: lanes that may take unfinished triangles into the next drain) and the refill
    threshold (RT_TUNE_REFILL) only decide WHEN a lane's triangles are tested: image, accumulation buffer, final stream states
    and ray count equal the frame rendered with a carry of 0 (every drain runs to the end) at every setting, on a scene with
    textures, all three materials and deep paths"""
    data = scenes.sponza_scale_scene(32, 2, 9)
    scene = pkg.Scene(app, data)
    w, h = 192, 128
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    cls = pkg.MegakernelRenderer if kind == "megakernel" else pkg.WavefrontRenderer
    ref = None
    for carry, refill in ((0, 14), (None, None), (1, 16), (5, 8), (12, 24), (32, 31), (3, 1)):
        if carry is None:
            monkeypatch.delenv("RT_TUNE_CARRY", raising=False)
            monkeypatch.delenv("RT_TUNE_REFILL", raising=False)
        else:
            monkeypatch.setenv("RT_TUNE_CARRY", str(carry))
            monkeypatch.setenv("RT_TUNE_REFILL", str(refill))
        r = cls(app, (w, h), None, 8, 6)
        f = r.render_frame(cam, scene)
        got = (f.rgba8.copy(), f.accum.view(np.uint32).copy(), f.rng_state.copy(), f.ray_count)
        r.close()
        if ref is None:
            ref = got
            continue
        assert got[3] == ref[3], (carry, refill)
        for a, b in zip(got[:3], ref[:3]):
            assert np.array_equal(a, b), (carry, refill)
    scene.close()


def test_cost_ordered_block_handout_changes_nothing(pkg, oracle, app, scenes, monkeypatch):
    """from 32 spp the megakernel hands its 8x4 pixel blocks out by decreasing probed cost (k_block_cost):
    scheduling only — every output equals the enumeration-order frame (RT_BLOCK_ORDER=0) and the oracle,
    also under tile sharding"""
    data = scenes.cornell_scene(3)
    scene = pkg.Scene(app, data)
    w, h = 256, 256                                      # 32 x 64 blocks: the ordering is active, also per rank at world 2
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    ordered = pkg.MegakernelRenderer(app, (w, h), None, 6, 32)
    monkeypatch.setenv("RT_BLOCK_ORDER", "0")
    plain = pkg.MegakernelRenderer(app, (w, h), None, 6, 32)
    monkeypatch.delenv("RT_BLOCK_ORDER")
    a, b = ordered.render_frame(cam, scene), plain.render_frame(cam, scene)
    assert a.kernel_launches == 3 and b.kernel_launches == 1   # probe + key + megakernel vs megakernel alone
    assert a.ray_count == b.ray_count
    for k in ("rgba8", "rng_state"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    assert np.array_equal(a.accum.view(np.uint32), b.accum.view(np.uint32))
    crop = (96, 104, 160, 136)
    o = oracle.Scene(data).render(oracle.camera_for(data, w, h), 0, 6, 32, use_bvh=True, crop=crop)
    assert np.array_equal(a.accum[104:136, 96:160].view(np.uint32), o["accum"].view(np.uint32))   # the oracle returns the crop
    acc, rays = np.zeros_like(a.accum), 0
    for rank in range(2):
        f = ordered.render_frame(cam, scene, shard={"rank": rank, "world": 2, "tile_size": 16})
        assert f.kernel_launches == 3
        acc += f.accum
        rays += f.ray_count
    assert np.array_equal(acc.view(np.uint32), a.accum.view(np.uint32)) and rays == a.ray_count
    ordered.close()
    plain.close()
    scene.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_tile_shards_gather_into_one_image(pkg, app, scenes, kind):
    """tile shards with the peer-memory gather: every rank's kernel stores its finished pixels straight
    into the destination renderer's image (here: a second renderer on the same device; across processes
    the pointer comes from the CUDA IPC handle of rt_renderer_export_image) — the destination image equals
    the unsharded frame without any reduction"""
    data = scenes.cornell_scene(3)
    scene = pkg.Scene(app, data)
    w, h = 200, 120
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    cls = pkg.MegakernelRenderer if kind == 0 else pkg.WavefrontRenderer
    ref, dest, peer = cls(app, (w, h), None, 8, 3), cls(app, (w, h), None, 8, 3), cls(app, (w, h), None, 8, 3)
    full = ref.render_frame(cam, scene).rgba8.copy()
    dest.max_depth = 1
    other = dest.render_frame(cam, scene).rgba8.copy()   # the destination starts out holding a different image
    dest.max_depth = 8
    tiles = (np.arange(h)[:, None] // 32) * ((w + 31) // 32) + np.arange(w)[None, :] // 32
    foreign = tiles % 4 != 0
    assert (other != full).any(-1)[foreign].any()
    assert len(dest.export_image()) == 64                # from now on dest leaves foreign pixels alone
    peer.set_gather(device_ptr=dest.device_rgba8_ptr)
    for rank in (1, 2, 3):
        peer.render_frame(cam, scene, want=(), shard={"rank": rank, "world": 4, "tile_size": 32})
    f = dest.render_frame(cam, scene, want=("rgba8",), shard={"rank": 0, "world": 4, "tile_size": 32})
    assert np.array_equal(f.rgba8, full)
    peer.set_gather()                                    # detached: the next peer frame no longer touches dest
    peer.max_depth = 1
    peer.render_frame(cam, scene, want=(), shard={"rank": 1, "world": 4, "tile_size": 32})
    f = dest.render_frame(cam, scene, want=("rgba8",), shard={"rank": 0, "world": 4, "tile_size": 32})
    assert np.array_equal(f.rgba8, full)
    with pytest.raises(pkg.RtError):
        peer.set_gather(handle=b"\0" * 64, device_ptr=dest.device_rgba8_ptr)
    for r in (ref, dest, peer):
        r.close()
    scene.close()


def test_spp_shards_match_salted_oracle_and_resolve(pkg, oracle, app, scenes):
    """spp sharding (config 5's mode): each shard is its own stream (seed ^ salt); the reduced
    buffer resolved by rt_resolve equals the oracle's per-shard sum"""
    data = scenes.cube_scene()
    scene = pkg.Scene(app, data)
    w, h = 64, 40
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    osc, ocam = oracle.Scene(data), oracle.camera_for(data, w, h)
    r = pkg.WavefrontRenderer(app, (w, h), None, 8, 4)
    acc, oacc = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    for rank in range(2):
        salt = (rank * 0x9E3779B9) & 0xFFFFFFFF
        f = r.render_frame(cam, scene, shard={"rank": rank, "world": 2, "tile_size": 0, "seed_salt": salt})
        o = osc.render(ocam, oracle.MODE_WAVEFRONT, 8, 4, seed_salt=salt)
        assert np.array_equal(f.accum.view(np.uint32), o["accum"].view(np.uint32))
        acc += f.accum
        oacc += o["accum"]
    img = pkg.resolve(app, acc, 8, w, h)
    L = oracle.lib()
    mean = oacc[..., :3] / np.float32(8)
    want = np.vectorize(lambda v: L.orc_output_byte(float(v)))(np.sqrt(mean)).astype(np.uint8)
    assert np.array_equal(img[..., :3], want) and (img[..., 3] == 255).all()
    r.close()
    scene.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_progressive_resume_is_bit_identical_to_one_frame(pkg, oracle, app, scenes, kind):
    """config 5 renders 4096 spp progressively: frames chained with RT_RENDER_RESUME continue the
    per-pixel xorshift streams and the accumulation, so 2 + 3 + 1 samples equal one 6-sample frame
    (itself equal to the oracle) in every output — also under tile and spp sharding"""
    data = scenes.cornell_scene(3)
    scene = pkg.Scene(app, data)
    w, h = 160, 96
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    cls = pkg.MegakernelRenderer if kind == 0 else pkg.WavefrontRenderer
    o = oracle.Scene(data).render(oracle.camera_for(data, w, h), kind, 8, 6, use_bvh=True)
    r = cls(app, (w, h), None, 8, 6)
    with pytest.raises(pkg.RtError):
        r.render_frame(cam, scene, resume=True)          # nothing to continue yet
    for shard in (None, {"rank": 1, "world": 3, "tile_size": 32}, {"rank": 1, "world": 2, "tile_size": 0, "seed_salt": 0x9E3779B9}):
        r.sample_count = 6
        one = r.render_frame(cam, scene, shard=shard)
        one = (one.accum.copy(), one.rgba8.copy(), one.rng_state.copy(), one.ray_count)
        rays = 0
        for i, n in enumerate((2, 3, 1)):
            r.sample_count = n
            f = r.render_frame(cam, scene, shard=shard, resume=i > 0)
            rays += f.ray_count
        assert rays == one[3]
        assert np.array_equal(f.accum.view(np.uint32), one[0].view(np.uint32))
        assert np.array_equal(f.rgba8, one[1]) and np.array_equal(f.rng_state, one[2])
        if shard is None:
            assert np.array_equal(f.accum.view(np.uint32), o["accum"].view(np.uint32))
            assert np.array_equal(f.rgba8, o["rgba8"]) and rays == o["ray_count"]
    r.close()
    scene.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_c4_style_half_million_triangles(pkg, oracle, app, scenes, kind):
    """config 4's kind of scene (one displaced height field, grazing camera) at 500 k triangles so
    the oracle's BVH finishes in seconds: random rays + a crop of a 4K frame"""
    data = scenes.big_mesh_scene(500)
    assert data.triangle_count == 500000
    scene = pkg.Scene(app, data)
    st = scene.stats
    assert st["triangle_count"] == 500000 and st["wide_depth"] <= 16
    osc = oracle.Scene(data)
    if kind == 0:
        org, d = _rays(40000, 77, 50.0)
        org[:, 1] = np.abs(org[:, 1]) * 0.2 + 1.0
        g, o = pkg.intersect(app, scene, org, d), osc.intersect(org, d, use_bvh=True)
        assert _check_hits(o, g) == 1.0 and np.array_equal(o["t"].view(np.uint32), g["t"].view(np.uint32))
    w, h, crop = 3840, 2160, (1900, 1300, 1964, 1332)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    cls = pkg.MegakernelRenderer if kind == 0 else pkg.WavefrontRenderer
    r = cls(app, (w, h), None, 10, 2)
    f = r.render_frame(cam, scene)
    o = osc.render(oracle.camera_for(data, w, h), kind, 10, 2, use_bvh=True, crop=crop)
    x0, y0, x1, y1 = crop
    assert np.array_equal(f.rng_state[y0:y1, x0:x1], o["rng_state"])
    assert np.array_equal(f.accum[y0:y1, x0:x1].view(np.uint32), o["accum"].view(np.uint32))
    r.close()
    scene.close()


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("which", ["cornell", "sponza_scale"])
def test_full_paths_1024spp_rmse_on_material_scenes(pkg, oracle, app, scenes, kind, which):
    """north_star's full-path acceptance where the stream-consuming branches live (`reflectance > rng()`, `dot > 0`, textures,
    emissive terminals): configs 2 and 3 at their full 1920x1080 size and 1024 spp on the GPU, a crop window through the
    oracle. Tolerance: per-pixel linear mean RMSE <= 2e-3; it comes out 0 with bit-exact streams and ray-exact accumulation."""
    if which == "cornell":
        data, crop = scenes.cornell_scene(5), (1200, 760, 1232, 772)      # rim of the glass sphere + the wall behind it
    else:
        data, crop = scenes.sponza_scale_scene(), (948, 600, 972, 612)    # textured floor, spheres
    spp = 1024
    f, o, sub = _render_pair(pkg, oracle, app, data, 1920, 1080, 10, spp, kind, crop=crop)
    mean_g, mean_o = sub(f.accum)[..., :3] / float(spp), o["accum"][..., :3] / float(spp)
    rmse = float(np.sqrt(np.mean((mean_g - mean_o) ** 2)))
    assert rmse <= 2e-3, rmse
    assert np.array_equal(sub(f.rng_state), o["rng_state"])               # streams bit-exact after 1024 samples
    assert np.array_equal(sub(f.accum).view(np.uint32), o["accum"].view(np.uint32)) and rmse == 0.0
    assert np.array_equal(sub(f.rgba8), o["rgba8"])
    assert (sub(f.accum)[..., 3] == spp).all()


def test_c4_true_size_ten_million_triangles(pkg, oracle, app, scenes):
    """config 4 at its OWN size: the 9,999,392-triangle height field (11-level wide tree, > 1 M nodes). 40 k random rays
    through rt_intersect against the oracle's SAH BVH, then a crop of the 3840x2160 frame for both renderers."""
    data = scenes.big_mesh_scene()
    assert data.triangle_count == 9999392
    scene = pkg.Scene(app, data)
    st = scene.stats
    assert st["triangle_count"] == 9999392 and st["node_count"] > 1000000
    osc = oracle.Scene(data)
    org, d = _rays(40000, 77, 50.0)
    org[:, 1] = np.abs(org[:, 1]) * 0.2 + 1.0
    g, o = pkg.intersect(app, scene, org, d), osc.intersect(org, d, use_bvh=True)
    assert _check_hits(o, g) == 1.0 and np.array_equal(o["t"].view(np.uint32), g["t"].view(np.uint32))
    assert 0.2 < (g["inst"] >= 0).mean() < 1.0
    w, h, crop = 3840, 2160, (1900, 1300, 1964, 1332)
    x0, y0, x1, y1 = crop
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    for kind, cls in ((0, pkg.MegakernelRenderer), (1, pkg.WavefrontRenderer)):
        r = cls(app, (w, h), None, 10, 2)
        f = r.render_frame(cam, scene)
        ref = osc.render(oracle.camera_for(data, w, h), kind, 10, 2, use_bvh=True, crop=crop)
        assert np.array_equal(f.rng_state[y0:y1, x0:x1], ref["rng_state"])
        assert np.array_equal(f.accum[y0:y1, x0:x1].view(np.uint32), ref["accum"].view(np.uint32))
        assert np.array_equal(f.rgba8[y0:y1, x0:x1], ref["rgba8"])
        r.close()
    scene.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_c5_4k_progressive_crop(pkg, oracle, app, scenes, kind):
    """config 5's shape: the ~261 k-triangle scene at 3840x2160, rendered progressively as four resumed batches (3 + 2 + 2 + 1
    samples); a crop of the final frame equals the oracle's one 8-sample frame in streams, accumulation and image"""
    data = scenes.sponza_scale_scene()
    w, h, crop = 3840, 2160, (1890, 1210, 1938, 1234)
    x0, y0, x1, y1 = crop
    scene = pkg.Scene(app, data)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    cls = pkg.MegakernelRenderer if kind == 0 else pkg.WavefrontRenderer
    r = cls(app, (w, h), None, 10, 8)
    rays = 0
    for i, n in enumerate((3, 2, 2, 1)):
        r.sample_count = n
        f = r.render_frame(cam, scene, resume=i > 0)
        rays += f.ray_count
    r.sample_count = 8
    one = r.render_frame(cam, scene)
    assert rays == one.ray_count and np.array_equal(f.rgba8, one.rgba8) and np.array_equal(f.rng_state, one.rng_state)
    o = oracle.Scene(data).render(oracle.camera_for(data, w, h), kind, 10, 8, use_bvh=True, crop=crop)
    assert np.array_equal(f.rng_state[y0:y1, x0:x1], o["rng_state"])
    assert np.array_equal(f.accum[y0:y1, x0:x1].view(np.uint32), o["accum"].view(np.uint32))
    assert np.array_equal(f.rgba8[y0:y1, x0:x1], o["rgba8"])
    assert (f.accum[..., 3] == 8).all()
    r.close()
    scene.close()


@pytest.mark.parametrize("split", [False, True])
def test_stadium_large_triangles_next_to_dense_detail(monkeypatch, pkg, oracle, app, scenes, split):
    """the BVH-hostile scene on the GPU: size-class key bit (default build) and the optional reference split of large
    triangles (RT_SPLIT=1 at rt_scene_commit) leave every hit and every pixel equal to the oracle's"""
    if split:
        monkeypatch.setenv("RT_SPLIT", "1")
    else:
        monkeypatch.delenv("RT_SPLIT", raising=False)
    data = scenes.stadium_scene(n_props=8, prop_subdiv=3, n_beams=32)
    scene = pkg.Scene(app, data)
    osc = oracle.Scene(data)
    rs = np.random.RandomState(3)
    org = np.stack([rs.uniform(-45, 45, 40000), rs.uniform(0.5, 18, 40000), rs.uniform(-45, 45, 40000)], 1).astype(np.float32)
    d = (rs.rand(40000, 3) - 0.5).astype(np.float32)
    g, o = pkg.intersect(app, scene, org, d), osc.intersect(org, d, use_bvh=True)
    assert _check_hits(o, g) == 1.0 and np.array_equal(o["t"].view(np.uint32), g["t"].view(np.uint32))
    w, h, crop = 960, 540, (440, 250, 520, 290)
    x0, y0, x1, y1 = crop
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    for kind, cls in ((0, pkg.MegakernelRenderer), (1, pkg.WavefrontRenderer)):
        r = cls(app, (w, h), None, 10, 4)
        f = r.render_frame(cam, scene)
        ref = osc.render(oracle.camera_for(data, w, h), kind, 10, 4, use_bvh=True, crop=crop)
        if split:
            # KNOWN LIMIT of the opt-in split: a bounce ray that starts a few 1e-6 behind a 100 m wall re-hits
            # it at an exact distance below tnear, where the triangle test's own rounding reports t = 1.02e-4 > tnear. The
            # brute-force loop (the definition) accepts that hit; the exact, flat box of a wall PIECE culls it (the
            # unsplit wall sits in a coarsely quantised slot and is found). A handful of paths per million differ.
            same = (f.accum[y0:y1, x0:x1].view(np.uint32) == ref["accum"].view(np.uint32)).all(-1)
            assert same.mean() > 0.995
        else:
            assert np.array_equal(f.rng_state[y0:y1, x0:x1], ref["rng_state"])
            assert np.array_equal(f.accum[y0:y1, x0:x1].view(np.uint32), ref["accum"].view(np.uint32))
            assert np.array_equal(f.rgba8[y0:y1, x0:x1], ref["rgba8"])
        r.close()
    if split:
        monkeypatch.delenv("RT_SPLIT")
        plain = pkg.Scene(app, data)
        assert scene.stats["bvh_bytes"] > plain.stats["bvh_bytes"]   # the references are there
        plain.close()
    scene.close()
