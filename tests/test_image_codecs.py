"""Host image codecs (sycl-ray-tracer_b200/host/image_codecs.hpp) against the reference's own stb_image:
every PNG / JPEG variant decodes to exactly the bytes stbi_load_from_memory(..., 4) gives the reference
(fixtures from tests/tools/make_golden_images.py; re-checked live against oracle/_ref/libstbref.so when
/root/reference is present)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "sycl-ray-tracer_b200", "host")
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))
NAMES = sorted(k[3:] for k in GOLD.files if k.startswith("in_"))


@pytest.fixture(scope="module")
def codec():
    subprocess.run(["make", "-s", "-C", HOST, os.path.join(HOST, "libglb_loader.so")], check=True)
    L = C.CDLL(os.path.join(HOST, "libglb_loader.so"))
    L.glb_image_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    L.glb_last_error.restype = C.c_char_p

    def decode(data):
        w, h, comp = C.c_uint32(), C.c_uint32(), C.c_int()
        out = np.zeros(1 << 20, np.uint8)
        if not L.glb_image_decode(bytes(data), len(data), out.ctypes.data, out.size, C.byref(w), C.byref(h), C.byref(comp)):
            raise RuntimeError(L.glb_last_error().decode())
        return out[: w.value * h.value * 4].reshape(h.value, w.value, 4).copy(), comp.value
    return decode


def test_fixture_covers_the_variants():
    assert len(NAMES) >= 55
    for must in ("png_h_rgba16_adam7", "png_palette_trns", "png_h_grey2_colorkey", "jpg_progressive_420", "jpg_restart",
                 "jpg_rgb_adobe", "jpg_grey_progressive", "jpg_odd_9x17_420", "jpg_h_440", "jpg_h_411", "jpg_h_mixed", "jpg_h_4x4",
                 "jpg_h_1px_wide"):
        assert must in NAMES


@pytest.mark.parametrize("name", NAMES)
def test_decodes_to_the_reference_bytes(codec, name):
    got, comp = codec(GOLD["in_" + name].tobytes())
    want = GOLD["out_" + name]
    assert got.shape == want.shape
    diff = np.argwhere(got != want)
    assert diff.size == 0, f"{name}: {len(diff)} differing bytes, first at {diff[0]}: {got[tuple(diff[0])]} != {want[tuple(diff[0])]}"
    assert comp == int(GOLD["comp_" + name])


def test_rejects_what_it_cannot_decode(codec):
    for junk in (b"", b"GIF89a" + b"\0" * 20, b"\xff\xd8\xff", b"\x89PNG\r\n\x1a\n" + b"\0" * 30):
        with pytest.raises(RuntimeError):
            codec(junk)
    data = GOLD["in_jpg_baseline_420"].tobytes()
    for cut in (len(data) // 2, 200):                      # truncated streams must fail or decode, never crash
        try:
            codec(data[:cut])
        except RuntimeError:
            pass
    sof = data.index(b"\xff\xc0")
    with pytest.raises(RuntimeError):                      # 12-bit precision
        codec(data[:sof + 4] + b"\x0c" + data[sof + 5:])


@pytest.mark.skipif(not os.path.isdir("/root/reference/deps/include"), reason="needs the reference's stb headers")
def test_fixtures_are_what_the_reference_stb_returns_today():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mgi", os.path.join(ROOT, "tests", "tools", "make_golden_images.py"))
    mgi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgi)
    L = mgi.stbref()
    for name in NAMES:
        rgba, comp = mgi.stb_decode(L, GOLD["in_" + name].tobytes())
        assert np.array_equal(rgba, GOLD["out_" + name]) and comp == int(GOLD["comp_" + name]), name


def _mgi():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mgi", os.path.join(ROOT, "tests", "tools", "make_golden_images.py"))
    mgi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgi)
    return mgi


def test_bake_resize_is_within_one_code_value_of_the_reference():
    """the 512x512 bake (host/glb_loader.hpp resize_to_layer) against stbir_resize_uint8_srgb(..., STBIR_RGBA)
    of the reference (src/image_manager.hpp:52-62): enlarging (Catmull-Rom), reducing (Mitchell), mixed, with
    alpha weighting. This part of the bake is NOT PINNED bit for bit (stb sums in SIMD single precision in a cost-chosen
    pass order and encodes sRGB through a table; restating that order was judged not worth ~10 k lines of reading):
    the stated tolerance is at most one code value, in under 2 % of the texels. SURVEY 8(f2) therefore stays "partial";
    512x512 inputs (what the synthetic scenes and most real assets use) bypass the resize and are exact."""
    subprocess.run(["make", "-s", "-C", HOST, os.path.join(HOST, "libglb_loader.so")], check=True)
    L = C.CDLL(os.path.join(HOST, "libglb_loader.so"))
    L.glb_resize_to_layer.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    mgi = _mgi()
    inputs = [("resize_" + n, GOLD["out_" + n]) for n in ("png_rgba8", "jpg_baseline_420", "jpg_noisy_128")]
    inputs += [(f"resize_proc_{w}x{h}_{seed}", mgi.test_image(w, h, seed)) for w, h, seed in mgi.RESIZE_PROCEDURAL]
    for key, src in inputs:
        src = np.ascontiguousarray(src)
        h, w, _ = src.shape
        out = np.zeros((512, 512, 4), np.uint8)
        L.glb_resize_to_layer(src.ctypes.data, w, h, out.ctypes.data)
        d = np.abs(out[::8, ::8].astype(int) - GOLD[key].astype(int))
        assert d.max() <= 1, (key, d.max())
        assert (d > 0).mean() < 0.02, (key, (d > 0).mean())
    same = np.ascontiguousarray(mgi.test_image(512, 512, 30))          # already 512x512: untouched
    out = np.zeros_like(same)
    L.glb_resize_to_layer(same.ctypes.data, 512, 512, out.ctypes.data)
    assert np.array_equal(out, same)


def test_mutated_files_never_crash_the_decoders(codec):
    """2000 random corruptions (byte flips, truncation, insertion, deletion) of the fixture files: each one
    decodes or raises, and a decoded image has the size it claims (the same driver ran 100 k iterations under
    ASan + UBSan while the decoders were written)"""
    rs = np.random.RandomState(5)
    decoded = 0
    for _ in range(2000):
        d = bytearray(GOLD["in_" + NAMES[rs.randint(len(NAMES))]].tobytes())
        mode = rs.randint(4)
        if mode == 0:
            for _ in range(rs.randint(1, 6)):
                d[rs.randint(len(d))] = rs.randint(256)
        elif mode == 1:
            d = d[: rs.randint(1, len(d))]
        elif mode == 2:
            i = rs.randint(len(d))
            d[i:i] = bytes(rs.randint(0, 256, rs.randint(1, 8)).astype(np.uint8))
        else:
            i = rs.randint(len(d))
            del d[i:i + rs.randint(1, 8)]
        try:
            img, _ = codec(bytes(d))
            decoded += img.ndim == 3 and img.shape[2] == 4
        except RuntimeError:
            pass
    assert decoded > 100
