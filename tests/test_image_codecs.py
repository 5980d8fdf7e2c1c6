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


def _resize_lib():
    subprocess.run(["make", "-s", "-C", HOST, os.path.join(HOST, "libglb_loader.so")], check=True)
    L = C.CDLL(os.path.join(HOST, "libglb_loader.so"))
    L.glb_resize_to_layer.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.glb_srgb_decode_table.argtypes = [C.c_void_p]

    def resize(src):
        src = np.ascontiguousarray(src)
        h, w, _ = src.shape
        out = np.zeros((512, 512, 4), np.uint8)
        L.glb_resize_to_layer(src.ctypes.data, w, h, out.ctypes.data)
        return out
    return L, resize


def _mgr():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mgr", os.path.join(ROOT, "tests", "tools", "make_golden_resize.py"))
    mgr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgr)
    return mgr


def test_bake_resize_equals_the_reference_bit_for_bit():
    """the 512x512 bake (host/bake_resize.hpp) against stbir_resize_uint8_srgb(..., 512, 512, 0, STBIR_RGBA) of the
    reference's vendored stb_image_resize2 (src/image_manager.hpp:52-62): EQUAL bytes. tests/golden/resize.npz holds the
    SHA-256 of the reference's whole 512x512x4 result (plus one row and one column, so a failure says where) for 24
    procedural inputs — enlarged (Catmull-Rom), reduced (Mitchell), mixed, one axis untouched, 1x1 up to 2048x2048, single
    rows / columns, reductions beyond 8x (the scattering vertical pass), prime sizes, transparent regions — and images.npz
    every 8th texel of seven more."""
    import hashlib
    _, resize = _resize_lib()
    mgr, mgi = _mgr(), _mgi()
    G = np.load(os.path.join(ROOT, "tests", "golden", "resize.npz"))
    for i, (w, h, seed, kind) in enumerate(G["cases"].tolist()):
        out = resize(mgr.resize_input(w, h, seed, kind))
        assert np.array_equal(out[257], G[f"row_{i}"]), (w, h, seed, kind, "row 257")
        assert np.array_equal(out[:, 130], G[f"col_{i}"]), (w, h, seed, kind, "column 130")
        assert hashlib.sha256(out.tobytes()).digest() == G[f"sha_{i}"].tobytes(), (w, h, seed, kind)
    inputs = [("resize_" + n, GOLD["out_" + n]) for n in ("png_rgba8", "jpg_baseline_420", "jpg_noisy_128")]
    inputs += [(f"resize_proc_{w}x{h}_{seed}", mgi.test_image(w, h, seed)) for w, h, seed in mgi.RESIZE_PROCEDURAL]
    for key, src in inputs:
        assert np.array_equal(resize(src)[::8, ::8], GOLD[key]), key
    same = np.ascontiguousarray(mgi.test_image(512, 512, 30))          # already 512x512: untouched (so is the reference's)
    assert np.array_equal(resize(same), same)


@pytest.mark.skipif(not os.path.isdir("/root/reference/deps/include"), reason="needs the reference's stb headers")
def test_bake_resize_against_the_reference_library_live():
    """fresh random sizes and contents through both libraries (oracle/_ref/libstbref.so = the reference's stb compiled in
    place), the sRGB decode table entry by entry against the one in the reference's header, and the resize goldens re-made"""
    import hashlib
    import re
    L, resize = _resize_lib()
    mgr, mgi = _mgr(), _mgi()
    R = mgi.stbref()
    rs = np.random.RandomState(11)
    sizes = [(512, 512), (1, 1), (3, 1), (1, 3), (513, 512), (512, 511), (4200, 5), (5, 4200)]
    sizes += [(int(rs.randint(1, 48)), int(rs.randint(1, 48))) for _ in range(12)]
    sizes += [(int(rs.randint(1, 1500)), int(rs.randint(1, 1500))) for _ in range(14)]
    for i, (w, h) in enumerate(sizes):
        src = mgr.resize_input(w, h, 1000 + i, i % 5)
        want = mgi.stb_resize(R, src)
        got = resize(src)
        assert np.array_equal(got, want), (w, h, i % 5, int(np.abs(got.astype(int) - want.astype(int)).max()))
    G = np.load(os.path.join(ROOT, "tests", "golden", "resize.npz"))
    for i, (w, h, seed, kind) in enumerate(G["cases"].tolist()[:8]):
        assert hashlib.sha256(mgi.stb_resize(R, mgr.resize_input(w, h, seed, kind)).tobytes()).digest() == G[f"sha_{i}"].tobytes()
    header = open("/root/reference/deps/include/stb_image_resize2.h").read()
    body = re.search(r"stbir__srgb_uchar_to_linear_float\[256\] = \{(.*?)\};", header, re.S).group(1)
    want = np.array([np.float32(x.rstrip("f")) for x in re.findall(r"[0-9.]+f", body)], np.float32)
    got = np.zeros(256, np.float32)
    L.glb_srgb_decode_table(got.ctypes.data)
    assert want.shape == (256,) and np.array_equal(got.view(np.uint32), want.view(np.uint32))


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
