"""Pins of the oracle itself: the known-answer vectors SURVEY.md 8(c) derives from the reference's
xorshift (src/xorshift.hpp:13-19) and closed-form checks of every restated primitive."""
import ctypes as C

import numpy as np


def _seq(L, seed, n):
    st = C.c_uint32(seed)
    out = []
    for _ in range(n):
        f = L.orc_xorshift_next(C.byref(st))
        out.append((st.value, f))
    return out


def test_xorshift_known_answers(oracle):
    L = oracle.lib()
    s = _seq(L, 1, 4)
    assert [a for a, _ in s] == [0x00042021, 0x04080601, 0x9DCCA8C5, 0x1255994F]
    assert [f for _, f in s] == [6.295018829405308e-05, 0.015747427940368652, 0.616404116153717, 0.07161863148212433]
    s = _seq(L, 2463534242, 4)  # the default state of XorShift32State
    assert [a for a, _ in s] == [0x2B1F4D63, 0x94DACB7A, 0x7B0859A0, 0x77B0567E]
    assert s[0][1] == 0.1684463918209076 and s[1][1] == 0.5814635157585144
    assert [a for a, _ in _seq(L, 1081, 2)] == [0x1063AB3A, 0xC59FEEB8]   # megakernel pixel (1,1) @1080p
    assert [a for a, _ in _seq(L, 1921, 2)] == [0x1EF4D8D9, 0x3818FFEE]   # wavefront pixel (1,1) @1080p


def test_xorshift_seed_zero_and_inclusive_one(oracle):
    L = oracle.lib()
    assert all(a == 0 and f == 0.0 for a, f in _seq(L, 0, 8))  # F2: pixel (0,0) draws 0 forever
    # F1: float = u32 -> f32 (RNE) * 2^-32: a state >= 0xFFFFFF80 rounds to exactly 1.0f.
    for target, want in ((0xFFFFFF80, 1.0), (0xFFFFFF7F, 0.99999994), (0x80000000, 0.5)):
        assert np.float32(np.float32(target) * np.float32(2.0 ** -32)) == np.float32(want)


def test_pixel_seed_mapping(oracle):
    L = oracle.lib()
    # F3: megakernel x * H_pad + y (H_pad = ceil(H/8)*8), wavefront x + y*W
    assert L.orc_pixel_seed(0, 1, 1, 1920, 1080) == 1081
    assert L.orc_pixel_seed(1, 1, 1, 1920, 1080) == 1921
    assert L.orc_pixel_seed(0, 3, 5, 256, 250) == 3 * 256 + 5
    assert L.orc_pixel_seed(0, 0, 0, 64, 64) == 0 and L.orc_pixel_seed(1, 0, 0, 64, 64) == 0


def test_half_rounding(oracle):
    L = oracle.lib()
    for v in (0.1, -0.3333, 1.0, 15.0, 65504.0, 1e-5, 2049.0, 2051.0, 0.0):
        assert L.orc_round_half(v) == float(np.float32(np.float16(np.float32(v))))
    assert L.orc_round_half(1e6) == float("inf")


def test_output_byte_rule(oracle):
    L = oracle.lib()
    # F10: unorm8 store (sat, rte) then (b/255.0f)*255.0f truncation
    for g in np.linspace(-0.2, 1.2, 701, dtype=np.float32):
        q = np.float32(np.rint(np.clip(np.float32(g) * np.float32(255.0), 0, 255)))
        want = int(np.float32(np.float32(q / np.float32(255.0)) * np.float32(255.0)))
        assert L.orc_output_byte(float(g)) == want
    assert L.orc_output_byte(float("nan")) == 0
    assert L.orc_output_byte(1.0) == 255


def test_camera_matches_closed_form(oracle):
    # src/camera.hpp:74-106 for an axis-aligned camera
    c = oracle.camera(1920, 1080, (0, 0, 0), (0, 0, -1), 1.0)
    aspect = np.float32(1920) / np.float32(1080)
    assert np.allclose(list(c.pixel00_loc), [-aspect, 1.0, -1.0], atol=1e-6)
    assert np.allclose(list(c.pixel_delta_u), [2 * aspect / 1920, 0, 0], atol=1e-9)
    assert np.allclose(list(c.pixel_delta_v), [0, -2.0 / 1080, 0], atol=1e-9)
    assert list(c.img_size) == [1920, 1080]


def test_camera_ray_draw_order_and_half_quantisation(oracle):
    L = oracle.lib()
    c = oracle.camera(64, 32, (0.5, 1, 2), (0.2, -0.1, -1), 1.5)
    st = C.c_uint32(77)
    org, d = (C.c_float * 3)(), (C.c_float * 3)()
    L.orc_camera_get_ray(C.byref(c), 10, 7, C.byref(st), org, d)
    s2 = _seq(L, 77, 2)
    assert st.value == s2[1][0]  # exactly two draws: px then py
    px, py = np.float32(-0.5) + np.float32(s2[0][1]), np.float32(-0.5) + np.float32(s2[1][1])
    p00, du, dv = (np.array(list(v), np.float32) for v in (c.pixel00_loc, c.pixel_delta_u, c.pixel_delta_v))
    center = np.array(list(c.center), np.float32)
    want = ((p00 + np.float32(10) * du) + np.float32(7) * dv) + ((px * du) + (py * dv)) - center
    assert list(org) == list(center)
    assert np.array_equal(np.array(list(d), np.float32), want.astype(np.float16).astype(np.float32))


def test_random_unit_vector(oracle):
    L = oracle.lib()
    st = C.c_uint32(12345)
    out = (C.c_float * 3)()
    L.orc_random_unit_vector(C.byref(st), out)
    draws = _seq(L, 12345, 3)
    assert st.value == draws[2][0]
    v = np.array([np.float32(-1) + np.float32(2) * np.float32(f) for _, f in draws], np.float32)  # x, y, z order
    inv = np.float32(1) / np.sqrt(np.float32((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]))
    assert np.array_equal(np.array(list(out), np.float32), v * inv)


def test_texture_sampling_rule(oracle):
    L = oracle.lib()
    tex = np.zeros((2, 512, 512, 4), np.uint8)
    tex[1, 3, 5] = (10, 20, 30, 40)
    tex[1, 511, 0] = (255, 0, 128, 7)
    out = (C.c_float * 3)()

    def samp(u, v):
        uv = (C.c_float * 2)(u, v)
        L.orc_texture_sample(tex.ctypes.data_as(oracle.u8p), 2, 1, uv, out)
        return list(out)
    want = [np.float32(10) / np.float32(255), np.float32(20) / np.float32(255), np.float32(30) / np.float32(255)]
    assert samp(5.5 / 512, 3.5 / 512) == want
    assert samp(5.5 / 512 + 3.0, 3.5 / 512 - 2.0) == want           # repeat addressing
    assert samp(0.0, 511.9 / 512) == [1.0, 0.0, float(np.float32(128) / np.float32(255))]
    assert samp(0.2, 0.2) == [0.0, 0.0, 0.0]


def test_normal_matrix(oracle):
    L = oracle.lib()
    rs = np.random.RandomState(3)
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = rs.rand(3, 3).astype(np.float32) + np.eye(3, dtype=np.float32)
    M[:3, 3] = (1, 2, 3)
    T = np.ascontiguousarray(M.T.reshape(16))  # column-major
    out = (C.c_float * 9)()
    L.orc_normal_matrix(T.ctypes.data_as(oracle.f32p), out)
    got = np.array(list(out), np.float64).reshape(3, 3).T  # back to math layout
    want = np.linalg.inv(M[:3, :3].astype(np.float64)).T
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_scatter_rules(oracle):
    L = oracle.lib()

    def scatter(mat, seed, d, n, uv=(0.3, 0.6)):
        st = C.c_uint32(seed)
        od, oa = (C.c_float * 3)(), (C.c_float * 3)()
        ok = L.orc_material_scatter(C.byref(mat), None, 0, C.byref(st), (C.c_float * 3)(*d), (C.c_float * 3)(*n),
                                    (C.c_float * 2)(*uv), od, oa)
        return ok, st.value, list(od), list(oa)
    n = (0.0, 1.0, 0.0)
    dn = tuple(np.array([0.6, -0.8, 0.0], np.float32))
    dif = oracle.orc_material(1, -1, (C.c_float * 3)(0.5, 0.25, 0.125), 0, 1.5, (C.c_float * 3)(0, 0, 0))
    ok, st, od, oa = scatter(dif, 99, dn, n)
    assert ok == 1 and st == _seq(L, 99, 3)[2][0] and oa == [0.5, 0.25, 0.125]       # 3 draws (F5)
    met = oracle.orc_material(2, -1, (C.c_float * 3)(0.9, 0.9, 0.9), 0.0, 1.5, (C.c_float * 3)(0, 0, 0))
    ok, st, od, oa = scatter(met, 99, dn, n)
    assert ok == 1 and st == _seq(L, 99, 3)[2][0]                                    # 3 draws even at roughness 0
    assert np.allclose(od, [0.6, 0.8, 0.0], atol=1e-6)
    ok, _, _, _ = scatter(met, 99, dn, (0.0, -1.0, 0.0))                             # reflected below the surface
    assert ok == 0                                                                  # F8: path ends, albedo not applied
    die = oracle.orc_material(3, -1, (C.c_float * 3)(1, 1, 1), 0.0, 1.5, (C.c_float * 3)(0, 0, 0))
    # total internal reflection from inside at a grazing angle: ZERO draws (short-circuit ||)
    inside = tuple(np.array([0.95, 0.3122499, 0.0], np.float32))
    ok, st, od, oa = scatter(die, 99, inside, n)
    assert ok == 1 and st == 99 and oa == [1.0, 1.0, 1.0]
    ok, st, od, oa = scatter(die, 99, dn, n)                                         # front face: exactly 1 draw
    assert ok == 1 and st == _seq(L, 99, 1)[0][0]
    none = oracle.orc_material(0, -1, (C.c_float * 3)(1, 1, 1), 0.0, 1.5, (C.c_float * 3)(0, 0, 0))
    assert scatter(none, 5, dn, n)[0] == 0
