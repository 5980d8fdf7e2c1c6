"""The two roadmap items the reference left open (PLAN.md:23-27), implemented as OPTIONS of the path (off by default):
Russian roulette (RT_RENDER_ROULETTE) and independent sample chains per pixel (rt_render_params.sample_chains, the
"splats" item: several samples of a pixel in flight, accumulated into per-chain planes and summed in chain order).
Both change the sampling order, so the comparison is against the oracle run with the same option. CPU part: the kernels'
per-item source (host emulation) against the oracle; GPU part: the CUDA path through the C ABI."""
import numpy as np
import pytest


def _eq(a, b):
    assert a["ray_count"] == b["ray_count"]
    assert np.array_equal(a["rng_state"], b["rng_state"])
    assert np.array_equal(a["accum"].view(np.uint32), b["accum"].view(np.uint32))
    assert np.array_equal(a["rgba8"], b["rgba8"])


@pytest.mark.parametrize("kind", [0, 1])
def test_roulette_in_the_kernel_source_matches_the_oracle(pkg, oracle, hostemu, scenes, kind):
    data = scenes.cornell_scene(2)
    w, h, depth, spp = 40, 28, 12, 3
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    emu = hostemu.Scene(data)
    osc, ocam = oracle.Scene(data), oracle.camera_for(data, w, h)
    on = emu.render(cam, kind, depth, spp, roulette=True)
    _eq(on, osc.render(ocam, kind, depth, spp, roulette=True))
    off = emu.render(cam, kind, depth, spp)
    _eq(off, osc.render(ocam, kind, depth, spp))
    assert on["ray_count"] < off["ray_count"]                      # paths are cut earlier in a closed box ...
    assert not np.array_equal(on["rng_state"], off["rng_state"])   # ... and the streams differ: an option, not the reference


def _frame(f):
    return dict(ray_count=f.ray_count, rng_state=f.rng_state, accum=f.accum, rgba8=f.rgba8)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [0, 1])
def test_roulette_on_the_gpu_matches_the_oracle(pkg, oracle, app, scenes, kind):
    data = scenes.cornell_scene(3)
    w, h, depth, spp = 96, 64, 16, 4
    scene = pkg.Scene(app, data)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    r = (pkg.MegakernelRenderer if kind == 0 else pkg.WavefrontRenderer)(app, (w, h), None, depth, spp)
    osc, ocam = oracle.Scene(data), oracle.camera_for(data, w, h)
    _eq(_frame(r.render_frame(cam, scene, roulette=True)), osc.render(ocam, kind, depth, spp, use_bvh=True, roulette=True))
    _eq(_frame(r.render_frame(cam, scene)), osc.render(ocam, kind, depth, spp, use_bvh=True))   # the default is untouched
    r.close()
    scene.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("chains,spp", [(2, 6), (3, 7), (4, 2)])
def test_sample_chains_equal_the_sum_of_salted_renders(pkg, oracle, app, scenes, kind, chains, spp):
    """chain c = spp_c samples on the stream seed ^ c * 0x9E3779B9; the planes are summed in chain order"""
    data = scenes.cornell_scene(2)
    w, h, depth = 72, 48, 8
    scene = pkg.Scene(app, data)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    r = (pkg.MegakernelRenderer if kind == 0 else pkg.WavefrontRenderer)(app, (w, h), None, depth, spp)
    f = r.render_frame(cam, scene, chains=chains)
    osc, ocam = oracle.Scene(data), oracle.camera_for(data, w, h)
    total, rays, first = None, 0, None
    for c in range(chains):
        s_c = spp // chains + (1 if c < spp % chains else 0)
        if s_c == 0:
            continue
        o = osc.render(ocam, kind, depth, s_c, use_bvh=True, seed_salt=(c * 0x9E3779B9) & 0xFFFFFFFF)
        total = o["accum"].copy() if total is None else total + o["accum"]
        rays += o["ray_count"]
        first = o if first is None else first
    assert f.ray_count == rays
    assert np.array_equal(f.accum.view(np.uint32), total.view(np.uint32))
    assert (f.accum[..., 3] == spp).all()
    assert np.array_equal(f.rng_state, first["rng_state"])          # chain 0 is the reference stream
    L = oracle.lib()
    want = np.vectorize(lambda v: L.orc_output_byte(float(v)))(np.sqrt(total[..., :3] / np.float32(spp))).astype(np.uint8)
    assert np.array_equal(f.rgba8[..., :3], want)
    # progressive frames continue every chain; tile shards of a chained frame sum to the whole
    r.sample_count = spp
    g = r.render_frame(cam, scene, chains=chains, resume=True)
    twice = r.render_frame(cam, scene, chains=chains)  # fresh frame again
    assert np.array_equal(twice.accum.view(np.uint32), f.accum.view(np.uint32))
    assert (g.accum[..., 3] == 2 * spp).all()
    acc = np.zeros_like(f.accum)
    for rank in range(2):
        acc += r.render_frame(cam, scene, chains=chains, shard={"rank": rank, "world": 2, "tile_size": 16}).accum
    assert np.array_equal(acc.view(np.uint32), f.accum.view(np.uint32))
    with pytest.raises(pkg.RtError):
        r.render_frame(cam, scene, chains=17)
    r.close()
    scene.close()
