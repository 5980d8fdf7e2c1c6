"""GLB loader + PNG writer of the C++ host (host/glb_loader.hpp), i.e. the "next" rows of the scope
table. A synthetic .glb is written here with every feature the reference loader
(src/scene.cpp) interprets; the loader's output is compared with what that file must produce."""
import ctypes as C
import json
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "sycl-ray-tracer_b200", "host")


@pytest.fixture(scope="module")
def glb(pkg):
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    L = C.CDLL(os.path.join(HOST, "libglb_loader.so"))
    L.glb_load.restype = C.c_void_p
    L.glb_load.argtypes = [C.c_char_p]
    L.glb_last_error.restype = C.c_char_p
    L.glb_free.argtypes = [C.c_void_p]
    L.glb_instance_count.argtypes = [C.c_void_p]
    L.glb_layer_count.argtypes = [C.c_void_p]
    L.glb_layers.restype = C.POINTER(C.c_uint8)
    L.glb_layers.argtypes = [C.c_void_p]
    L.glb_globals.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.glb_png_write.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.glb_png_read.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    L.glb_instance.argtypes = [C.c_void_p, C.c_uint32, up, up, C.POINTER(fp), C.POINTER(fp), C.POINTER(fp), C.POINTER(up),
                               fp, C.POINTER(pkg._capi.rt_material), C.POINTER(C.c_int32)]
    return L


def _png(rgba):
    h, w, _ = rgba.shape
    raw = b"".join(b"\x00" + rgba[y].tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")


def _write_glb(path, tex, extra_images=(), uri_images=(), f15=True):
    """two meshes under a parent/child hierarchy, u8/u16/u32 indices, an interleaved (strided) vertex
    buffer, diffuse-textured / metallic / dielectric / emissive materials, sky extras, a camera node"""
    rs = np.random.RandomState(4)
    blobs, views, accessors = [], [], []

    def add_view(data, stride=None):
        off = sum(len(b) for b in blobs)
        pad = (-len(data)) % 4
        blobs.append(data + b"\x00" * pad)
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(data)}
        if stride:
            v["byteStride"] = stride
        views.append(v)
        return len(views) - 1

    def add_acc(view, ctype, count, typ, offset=0):
        accessors.append({"bufferView": view, "componentType": ctype, "count": count, "type": typ, "byteOffset": offset})
        return len(accessors) - 1
    prims, expect = [], []
    for k, (nv, ntri, ict, itype) in enumerate([(12, 7, 5121, np.uint8), (40, 30, 5123, np.uint16), (9, 3, 5125, np.uint32)]):
        pos, nrm, uv = rs.randn(nv, 3).astype(np.float32), rs.randn(nv, 3).astype(np.float32), rs.rand(nv, 2).astype(np.float32)
        idx = rs.randint(0, nv, ntri * 3).astype(itype)
        if k == 1:   # interleaved pos|nrm|uv, stride 32
            inter = np.concatenate([pos, nrm, uv], 1).astype(np.float32)
            v = add_view(inter.tobytes(), 32)
            a = (add_acc(v, 5126, nv, "VEC3", 0), add_acc(v, 5126, nv, "VEC3", 12), add_acc(v, 5126, nv, "VEC2", 24))
        else:
            a = (add_acc(add_view(pos.tobytes()), 5126, nv, "VEC3"), add_acc(add_view(nrm.tobytes()), 5126, nv, "VEC3"),
                 add_acc(add_view(uv.tobytes()), 5126, nv, "VEC2"))
        ia = add_acc(add_view(idx.tobytes()), ict, ntri * 3, "SCALAR")
        prims.append({"attributes": {"POSITION": a[0], "NORMAL": a[1], "TEXCOORD_0": a[2]}, "indices": ia, "material": k})
        expect.append((pos, nrm, uv, idx.astype(np.uint32)))
    img_view = add_view(_png(tex))
    extra_views = [(add_view(bytes(data)), mime) for data, mime in extra_images]
    prims.append({"attributes": prims[0]["attributes"], "indices": prims[0]["indices"]})   # no material: F15 fallback
    if not f15:                      # ... which is undefined behaviour in the reference loader itself (materials[-1])
        prims[-1]["material"] = 1
    expect.append(expect[0])
    q = np.array([0.0, np.sin(0.35), 0.0, np.cos(0.35)])          # rotation about Y
    qc = np.array([np.sin(0.2), 0.0, 0.0, np.cos(0.2)])           # camera pitch
    child_m = np.eye(4)
    child_m[:3, 3] = (0.5, -1.0, 2.0)
    child_m[0, 0] = 2.0
    j = {"asset": {"version": "2.0"}, "scene": 0,
         "scenes": [{"nodes": [0, 3], "extras": {"sky_color": [0.2, 0.4, 0.8], "sky_strength": 2.0}}],
         "nodes": [{"translation": [1.0, 2.0, -3.0], "rotation": q.tolist(), "scale": [1.5, 1.0, 0.5], "children": [1, 2], "mesh": 0},
                   {"matrix": child_m.T.reshape(-1).tolist(), "mesh": 1},
                   {"translation": [0.0, 1.0, 0.0], "mesh": 0},
                   {"translation": [0.0, 1.0, 5.0], "rotation": qc.tolist(), "camera": 0},
                   {"mesh": 1}],                                   # node 4 is not in the scene: ignored
         "cameras": [{"type": "perspective", "perspective": {"yfov": 0.8, "aspectRatio": 1.5, "znear": 0.1}}],
         "meshes": [{"primitives": prims[:2] + [prims[3]]}, {"primitives": [prims[2]]}],
         "materials": [{"pbrMetallicRoughness": {"baseColorFactor": [0.9, 0.8, 0.7, 1], "metallicFactor": 0.0, "baseColorTexture": {"index": 0}},
                        "emissiveFactor": [1.0, 0.5, 0.25], "extensions": {"KHR_materials_emissive_strength": {"emissiveStrength": 4.0}}},
                       {"pbrMetallicRoughness": {"baseColorFactor": [0.5, 0.5, 0.5, 1], "metallicFactor": 0.7, "roughnessFactor": 0.25},
                        "emissiveFactor": [1.0, 1.0, 1.0]},       # no emissive_strength extension -> emissive 0
                       {"pbrMetallicRoughness": {"metallicFactor": 0.9},
                        "extensions": {"KHR_materials_ior": {"ior": 1.33}, "KHR_materials_transmission": {"transmissionFactor": 1}}}],
         "textures": [{"source": 0}], "images": [{"bufferView": img_view, "mimeType": "image/png"}] + [{"bufferView": v, "mimeType": m} for v, m in extra_views] + [{"uri": u} for u in uri_images],
         "accessors": accessors, "bufferViews": views, "buffers": [{"byteLength": sum(len(b) for b in blobs)}]}
    js = json.dumps(j).encode()
    js += b" " * ((-len(js)) % 4)
    binc = b"".join(blobs)
    total = 12 + 8 + len(js) + 8 + len(binc)
    with open(path, "wb") as f:
        f.write(struct.pack("<4sII", b"glTF", 2, total) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(binc), 0x004E4942) + binc)

    def trs(t, qq, s):
        x, y, z, w = qq
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        M = np.eye(4)
        M[:3, :3] = R @ np.diag(s)
        M[:3, 3] = t
        return M
    M0 = trs((1, 2, -3), q, (1.5, 1.0, 0.5))
    M2 = trs((0, 1, 0), (0, 0, 0, 1), (1, 1, 1))
    globals_ = {0: M0, 1: M0 @ child_m, 2: M0 @ M2}
    cam = trs((0, 1, 5), qc, (1, 1, 1))
    return expect, globals_, cam


def test_loader_follows_the_reference_rules(glb, pkg, tmp_path):
    tex = (np.random.RandomState(1).rand(512, 512, 4) * 255).astype(np.uint8)   # already 512x512: baked verbatim
    path = str(tmp_path / "synthetic.glb")
    expect, M, cam = _write_glb(path, tex)
    h = glb.glb_load(path.encode())
    assert h, glb.glb_last_error()
    # node order 0,1,2 (node 3 = camera, node 4 unreachable); node 0 and 2 carry mesh 0 (3 primitives), node 1 mesh 1
    want = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (1, 1, 0), (2, 0, 0), (2, 0, 1), (2, 0, 2)]
    assert glb.glb_instance_count(h) == len(want)
    prim_data = {(0, 0): expect[0], (0, 1): expect[1], (0, 2): expect[3], (1, 0): expect[2]}
    mats = []
    for i, (node, mesh, prim) in enumerate(want):
        nv, ni = C.c_uint32(), C.c_uint32()
        fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        pos, nrm, uv, idx = fp(), fp(), fp(), up()
        T = (C.c_float * 16)()
        mat = pkg._capi.rt_material()
        nmp = (C.c_int32 * 3)()
        glb.glb_instance(h, i, C.byref(nv), C.byref(ni), C.byref(pos), C.byref(nrm), C.byref(uv), C.byref(idx), T, C.byref(mat), nmp)
        assert list(nmp) == [node, mesh, prim]
        e = prim_data[(mesh, prim)]
        assert np.array_equal(np.ctypeslib.as_array(pos, (nv.value, 3)), e[0])
        assert np.array_equal(np.ctypeslib.as_array(nrm, (nv.value, 3)), e[1])
        assert np.array_equal(np.ctypeslib.as_array(uv, (nv.value, 2)), e[2])
        assert np.array_equal(np.ctypeslib.as_array(idx, (ni.value,)), e[3])     # u8 / u16 / u32 widened
        assert np.allclose(np.array(list(T)).reshape(4, 4).T, M[node], atol=1e-5)  # T*R*S*matrix, parents applied
        mats.append((mat.type, mat.albedo_image, list(mat.albedo_color), mat.roughness, mat.ior, list(mat.emissive)))
    D, Mt, Di = pkg._capi.RT_MAT_DIFFUSE, pkg._capi.RT_MAT_METALLIC, pkg._capi.RT_MAT_DIELECTRIC
    t, img, col, rough, ior, em = mats[0]
    assert (t, img) == (D, 0) and np.allclose(col, [0.9, 0.8, 0.7]) and np.allclose(em, [4.0, 2.0, 1.0])   # factor * strength
    t, img, col, rough, ior, em = mats[1]
    assert (t, img) == (Mt, -1) and np.isclose(rough, 0.25) and em == [0, 0, 0]      # strength extension absent -> 0
    assert mats[2][0] == D and np.allclose(mats[2][2], [0.8, 0.8, 0.8])              # F15 fallback
    assert mats[3][0] == Di and np.isclose(mats[3][4], 1.33)                         # ior + transmission beat metallic
    g = (C.c_float * 16)()
    glb.glb_globals(h, g)
    assert np.allclose(list(g)[:3], [0.4, 0.8, 1.6])                                # sky_color * sky_strength
    assert np.allclose(list(g)[3:6], cam[:3, 3]) and np.allclose(list(g)[6:9], -cam[:3, 2], atol=1e-6)
    assert np.isclose(g[9], 1.0 / np.tan(0.4)) and g[10] == 1.0
    assert glb.glb_layer_count(h) == 1
    assert np.array_equal(np.ctypeslib.as_array(glb.glb_layers(h), (512, 512, 4)), tex)
    glb.glb_free(h)


def test_errors_and_png_round_trip(glb, tmp_path):
    assert not glb.glb_load(str(tmp_path / "missing.glb").encode())
    assert b"Failed to load .glTF" in glb.glb_last_error()
    bad = tmp_path / "bad.glb"
    bad.write_bytes(b"not a gltf file at all.....")
    assert not glb.glb_load(str(bad).encode())
    img = (np.random.RandomState(2).rand(37, 53, 4) * 255).astype(np.uint8)
    out = str(tmp_path / "o.png")
    assert glb.glb_png_write(out.encode(), img.ctypes.data, 53, 37) == 1
    data = open(out, "rb").read()
    from PIL import Image
    assert np.array_equal(np.array(Image.open(out).convert("RGBA")), img)           # an independent decoder agrees
    back = np.zeros_like(img)
    w, h = C.c_uint32(), C.c_uint32()
    assert glb.glb_png_read(data, len(data), back.ctypes.data, back.nbytes, C.byref(w), C.byref(h)) == 1
    assert (w.value, h.value) == (53, 37) and np.array_equal(back, img)


def test_reference_assets_when_present(glb):
    """assets/cube.glb (config 1) and assets/triangle.glb load with the documented fallbacks"""
    cube = "/root/reference/assets/cube.glb"
    if not os.path.exists(cube):
        pytest.skip("reference assets not present on this machine")
    h = glb.glb_load(cube.encode())
    assert h and glb.glb_instance_count(h) == 1
    g = (C.c_float * 16)()
    glb.glb_globals(h, g)
    assert list(g)[:3] == [0.5, np.float32(0.7), 1.0] and g[10] == 0.0 and list(g)[6:9] == [0, 0, -1] and g[9] == 1.0
    glb.glb_free(h)
    h = glb.glb_load(b"/root/reference/assets/triangle.glb")
    assert h and glb.glb_instance_count(h) == 1
    glb.glb_free(h)


@pytest.mark.gpu
def test_cli_renders_a_glb_to_png(glb, oracle, scenes, tmp_path):
    """raytracer scene.glb -> out.png: the whole reference CLI flow on the GPU"""
    from PIL import Image
    tex = (np.random.RandomState(1).rand(512, 512, 4) * 255).astype(np.uint8)
    path = str(tmp_path / "synthetic.glb")
    _write_glb(path, tex)
    png = str(tmp_path / "out.png")
    r = subprocess.run([os.path.join(HOST, "raytracer"), "-m", "-d", "6", "-s", "2", "--size", "96x64", "--png", png, path],
                       capture_output=True, text=True, check=True)
    assert "Total rays:" in r.stdout and "Writing image to disk" in r.stdout
    img = np.array(Image.open(png))
    assert img.shape == (64, 96, 4) and (img[..., 3] == 255).all() and img[..., :3].std() > 1.0


def test_embedded_jpeg_and_small_png_are_baked_like_the_reference(glb, tmp_path):
    """images other than a 512x512 RGBA PNG: a JPEG (4:2:0) and a palette PNG with tRNS, decoded to stb_image's
    bytes (tests/golden/images.npz) and resized by the bake's filter"""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))
    tex = (np.random.RandomState(1).rand(512, 512, 4) * 255).astype(np.uint8)
    path = str(tmp_path / "images.glb")
    _write_glb(path, tex, [(gold["in_jpg_baseline_420"], "image/jpeg"), (gold["in_png_palette_trns"], "image/png")])
    h = glb.glb_load(path.encode())
    assert h, glb.glb_last_error()
    assert glb.glb_layer_count(h) == 3
    layers = np.ctypeslib.as_array(glb.glb_layers(h), (3, 512, 512, 4))
    assert np.array_equal(layers[0], tex)
    glb.glb_resize_to_layer.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    for k, name in ((1, "jpg_baseline_420"), (2, "png_palette_trns")):
        src = np.ascontiguousarray(gold["out_" + name])
        want = np.zeros((512, 512, 4), np.uint8)
        glb.glb_resize_to_layer(src.ctypes.data, src.shape[1], src.shape[0], want.ctypes.data)
        assert np.array_equal(layers[k], want)
    glb.glb_free(h)


def test_images_by_uri(glb, tmp_path):
    """tinygltf also resolves image "uri"s inside a .glb: base64 data URIs and files next to the .glb"""
    import base64
    gold = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))
    tex = (np.random.RandomState(1).rand(512, 512, 4) * 255).astype(np.uint8)
    (tmp_path / "side car.jpg").write_bytes(gold["in_jpg_progressive_422"].tobytes())
    data_uri = "data:image/png;base64," + base64.b64encode(gold["in_png_h_rgb8_adam7"].tobytes()).decode()
    path = str(tmp_path / "uris.glb")
    _write_glb(path, tex, uri_images=[data_uri, "side%20car.jpg"])
    h = glb.glb_load(path.encode())
    assert h, glb.glb_last_error()
    assert glb.glb_layer_count(h) == 3
    layers = np.ctypeslib.as_array(glb.glb_layers(h), (3, 512, 512, 4))
    glb.glb_resize_to_layer.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    for k, name in ((1, "png_h_rgb8_adam7"), (2, "jpg_progressive_422")):
        src = np.ascontiguousarray(gold["out_" + name])
        want = np.zeros((512, 512, 4), np.uint8)
        glb.glb_resize_to_layer(src.ctypes.data, src.shape[1], src.shape[0], want.ctypes.data)
        assert np.array_equal(layers[k], want)
    glb.glb_free(h)
    _write_glb(path, tex, uri_images=["missing.png"])
    assert not glb.glb_load(path.encode()) and b"cannot open" in glb.glb_last_error()


def _mine_scaled(glb, pkg, path, scale):
    """host/glb_loader.hpp's view of a .glb in the same shape as tests/_scenref.load()"""
    h = glb.glb_load_scaled(path.encode(), *[float(v) for v in scale])
    assert h, glb.glb_last_error()
    out = {"instances": []}
    fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    for i in range(glb.glb_instance_count(h)):
        nv, ni = C.c_uint32(), C.c_uint32()
        pos, nrm, uv, idx = fp(), fp(), fp(), up()
        T, mat, nmp = (C.c_float * 16)(), pkg._capi.rt_material(), (C.c_int32 * 3)()
        glb.glb_instance(h, i, C.byref(nv), C.byref(ni), C.byref(pos), C.byref(nrm), C.byref(uv), C.byref(idx), T, C.byref(mat), nmp)
        out["instances"].append(dict(
            positions=np.ctypeslib.as_array(pos, (nv.value, 3)).copy(), normals=np.ctypeslib.as_array(nrm, (nv.value, 3)).copy(),
            uvs=np.ctypeslib.as_array(uv, (nv.value, 2)).copy(), indices=np.ctypeslib.as_array(idx, (ni.value,)).copy(),
            transform=np.array(T, np.float32), type=mat.type, albedo_image=mat.albedo_image, albedo=np.array(mat.albedo_color, np.float32),
            roughness=mat.roughness, ior=mat.ior, emissive=np.array(mat.emissive, np.float32)))
    g = (C.c_float * 16)()
    glb.glb_globals(h, g)
    out.update(sky=np.array(g[0:3], np.float32), camera_position=np.array(g[3:6], np.float32), camera_direction=np.array(g[6:9], np.float32),
               focal=float(g[9]))
    n = glb.glb_layer_count(h)
    out["layers"] = list(np.ctypeslib.as_array(glb.glb_layers(h), (n, 512, 512, 4)).copy()) if n else []
    glb.glb_free(h)
    return out


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference sources (build container only)")
def test_loader_equals_the_references_own_loader(glb, pkg, oracle, tmp_path):
    """the reference's OWN src/scene.cpp (+ its vendored tinygltf / stb), compiled in place through the API shims
    (oracle/refshim/scenref.cpp -> oracle/_ref/libscenref.so), loads the same .glb files: instance order,
    buffers, transforms, normal matrices, material records, sky, camera and baked image layers must agree"""
    import _scenref
    gold = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))
    tex = (np.random.RandomState(1).rand(512, 512, 4) * 255).astype(np.uint8)
    path = str(tmp_path / "both.glb")
    _write_glb(path, tex, [(gold["in_jpg_baseline_420"], "image/jpeg"), (gold["in_png_palette_trns"], "image/png"),
                           (gold["in_png_h_rgba16_adam7"], "image/png"), (gold["in_png_grey16"], "image/png"),   # 16-bit: misread by the reference
                           (gold["in_png_h_rgb16_colorkey"], "image/png")], f15=False)
    for scale in ((1.0, 1.0, 1.0), (0.5, 2.0, 3.0)):
        ref = _scenref.load(path, scale)
        subprocess.run(["make", "-s", "-C", HOST], check=True)
        glb.glb_load_scaled.restype = C.c_void_p
        glb.glb_load_scaled.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float]
        mine = _mine_scaled(glb, pkg, path, scale)
        assert len(mine["instances"]) == len(ref["instances"]) == 7
        TYPE = {1: pkg._capi.RT_MAT_DIFFUSE, 2: pkg._capi.RT_MAT_METALLIC, 3: pkg._capi.RT_MAT_DIELECTRIC}
        for a, b in zip(mine["instances"], ref["instances"]):
            for k in ("positions", "normals", "uvs", "indices"):
                assert np.array_equal(a[k], b[k]), k
            assert np.array_equal(a["transform"].view(np.uint32), b["transform"].view(np.uint32))     # same arithmetic order as glm
            assert a["type"] == TYPE[b["type"]]
            if b["type"] != 3:
                assert a["albedo_image"] == (b["albedo_image"] if b["albedo_is_image"] else -1)
                if not b["albedo_is_image"]:
                    assert np.array_equal(a["albedo"], b["albedo"])
                assert np.array_equal(a["emissive"], b["emissive"])
            if b["type"] == 2:
                assert a["roughness"] == b["roughness"]
            if b["type"] == 3:
                assert a["ior"] == b["ior"]
            # the normal matrix of the reference's GeometryData against the oracle's restatement, same matrix in
            nm = np.zeros(9, np.float32)
            oracle.lib().orc_normal_matrix(np.ascontiguousarray(b["transform"]).ctypes.data_as(oracle.f32p), nm.ctypes.data_as(oracle.f32p))
            assert np.array_equal(nm.view(np.uint32), b["normal_matrix"].view(np.uint32))
        assert np.array_equal(mine["sky"], ref["sky"])
        assert np.array_equal(mine["camera_position"], ref["camera_position"])
        assert np.array_equal(mine["camera_direction"], ref["camera_direction"])
        assert mine["focal"] == ref["focal"]
        assert len(mine["layers"]) == len(ref["layers"]) == 6
        assert np.array_equal(mine["layers"][0], ref["layers"][0])            # 512x512: verbatim in both
        for k in (1, 2, 3, 4, 5):                                              # resized: the reference's bytes
            assert np.array_equal(mine["layers"][k], ref["layers"][k]), k


def test_malformed_files_are_rejected_not_trusted(glb, tmp_path):
    """indices, offsets and counts in the JSON are untrusted input: out-of-range accessors / buffer views /
    node and texture indices and cyclic hierarchies raise (the loader ran 20 k mutated files under ASan + UBSan)"""
    tex = np.zeros((8, 8, 4), np.uint8)
    path = str(tmp_path / "base.glb")
    _write_glb(path, tex)
    raw = open(path, "rb").read()
    jlen = struct.unpack("<I", raw[12:16])[0]
    js, rest = raw[20:20 + jlen].decode(), raw[20 + jlen:]

    def load(mutated_json):
        b = mutated_json.encode()
        b += b" " * ((-len(b)) % 4)
        p = str(tmp_path / "m.glb")
        with open(p, "wb") as f:
            f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(b) + len(rest)) + struct.pack("<II", len(b), 0x4E4F534A) + b + rest)
        h = glb.glb_load(p.encode())
        if h:
            glb.glb_free(h)
        return bool(h), glb.glb_last_error().decode()

    assert load(js)[0]
    j = json.loads(js)
    cases = []
    a = json.loads(js); a["accessors"][0]["count"] = 10 ** 9; cases.append(a)
    a = json.loads(js); a["accessors"][0]["byteOffset"] = 10 ** 8; cases.append(a)
    a = json.loads(js); a["accessors"][0]["bufferView"] = 999; cases.append(a)
    a = json.loads(js); a["bufferViews"][0]["byteOffset"] = 2 ** 31 - 8; cases.append(a)
    a = json.loads(js); a["bufferViews"][0]["byteLength"] = 2 ** 31 - 1; cases.append(a)
    a = json.loads(js); a["meshes"][0]["primitives"][0]["indices"] = -7; cases.append(a)
    a = json.loads(js); a["meshes"][0]["primitives"][0]["attributes"]["NORMAL"] = 12345; cases.append(a)
    a = json.loads(js); a["nodes"][0]["children"] = [1, 77]; cases.append(a)
    a = json.loads(js); a["scenes"][0]["nodes"] = [0, -2]; cases.append(a)
    a = json.loads(js); a["textures"][0]["source"] = 5; cases.append(a)
    a = json.loads(js); a["images"][0]["bufferView"] = 4000; cases.append(a)
    for c in cases:
        ok, err = load(json.dumps(c))
        assert not ok and "Failed to load .glTF" in err, err
    a = json.loads(js); a["nodes"][1]["children"] = [0]; a["nodes"][2]["children"] = [2]     # cycles: must terminate
    assert load(json.dumps(a))[0]
    assert not load(js[: len(js) // 2])[0] and not load("[1, 2")[0] and not load('{"nodes": 3')[0]


def _write_random_glb(path, seed, n_nodes=60, n_meshes=9, n_materials=7, tex512=False):
    """a random forest of nodes (TRS and/or matrix, some unreachable), meshes with 1-3 primitives, every
    material class, PNG images of odd sizes; the camera is never node 0 and every primitive has a material,
    so the reference loader itself stays clear of its F15 undefined behaviour"""
    rs = np.random.RandomState(seed)
    blobs, views, accessors = [], [], []

    def add_view(data, stride=None):
        off = sum(len(b) for b in blobs)
        blobs.append(data + b"\x00" * ((-len(data)) % 4))
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(data)}
        if stride:
            v["byteStride"] = stride
        views.append(v)
        return len(views) - 1

    def add_acc(view, ctype, count, typ, offset=0):
        accessors.append({"bufferView": view, "componentType": ctype, "count": count, "type": typ, "byteOffset": offset})
        return len(accessors) - 1

    images = []
    for k in range(3):
        w, h = (512, 512) if tex512 else (rs.randint(3, 40), rs.randint(3, 40))   # 512x512 is baked verbatim by both loaders
        images.append({"bufferView": add_view(_png(rs.randint(0, 256, (h, w, 4)).astype(np.uint8))), "mimeType": "image/png"})
    materials = []
    for k in range(n_materials):
        pbr = {"baseColorFactor": [float(v) for v in rs.rand(3)] + [1.0], "metallicFactor": float(rs.choice([0.0, 0.005, 0.02, 0.8])),
               "roughnessFactor": float(rs.rand())}
        if rs.rand() < 0.5:
            pbr["baseColorTexture"] = {"index": int(rs.randint(3))}
        m = {"pbrMetallicRoughness": pbr, "emissiveFactor": [float(v) for v in rs.rand(3)]}
        ext = {}
        if rs.rand() < 0.5:
            ext["KHR_materials_emissive_strength"] = {"emissiveStrength": float(rs.rand() * 10)} if rs.rand() < 0.8 else {}
        if rs.rand() < 0.3:
            ext["KHR_materials_ior"] = {"ior": float(1 + rs.rand())} if rs.rand() < 0.8 else {}
            if rs.rand() < 0.7:
                ext["KHR_materials_transmission"] = {"transmissionFactor": 1.0}
        if ext:
            m["extensions"] = ext
        materials.append(m)
    meshes = []
    for k in range(n_meshes):
        prims = []
        for _ in range(rs.randint(1, 4)):
            nv, ntri = rs.randint(3, 30), rs.randint(1, 20)
            pos, nrm, uv = rs.randn(nv, 3).astype(np.float32), rs.randn(nv, 3).astype(np.float32), rs.rand(nv, 2).astype(np.float32)
            ict, dt = [(5121, np.uint8), (5123, np.uint16), (5125, np.uint32)][rs.randint(3)]
            idx = rs.randint(0, nv, ntri * 3).astype(dt)
            if rs.rand() < 0.5:                                  # interleaved vertex buffer
                inter = np.concatenate([pos, nrm, uv], axis=1).astype(np.float32)
                v = add_view(inter.tobytes(), 32)
                a = (add_acc(v, 5126, nv, "VEC3", 0), add_acc(v, 5126, nv, "VEC3", 12), add_acc(v, 5126, nv, "VEC2", 24))
            else:
                a = (add_acc(add_view(pos.tobytes()), 5126, nv, "VEC3"), add_acc(add_view(nrm.tobytes()), 5126, nv, "VEC3"),
                     add_acc(add_view(uv.tobytes()), 5126, nv, "VEC2"))
            prims.append({"attributes": {"POSITION": a[0], "NORMAL": a[1], "TEXCOORD_0": a[2]},
                          "indices": add_acc(add_view(idx.tobytes()), ict, ntri * 3, "SCALAR"), "material": int(rs.randint(n_materials))})
        meshes.append({"primitives": prims})
    nodes = []
    for n in range(n_nodes):
        nd = {}
        if rs.rand() < 0.7:
            nd["translation"] = [float(v) for v in rs.randn(3)]
        if rs.rand() < 0.6:
            q = rs.randn(4)
            nd["rotation"] = [float(v) for v in q / np.linalg.norm(q)]
        if rs.rand() < 0.5:
            nd["scale"] = [float(v) for v in 0.5 + rs.rand(3)]
        if rs.rand() < 0.2:
            nd["matrix"] = [float(v) for v in (np.eye(4) + 0.3 * rs.randn(4, 4) * np.array([1, 1, 1, 0])[None, :]).T.reshape(-1)]
        if rs.rand() < 0.6:
            nd["mesh"] = int(rs.randint(n_meshes))
        nodes.append(nd)
    roots = []
    for n in range(n_nodes):                                     # parent has a lower index or none: a forest
        r = rs.rand()
        if n == 0 or r < 0.15:
            roots.append(n)
        elif r < 0.9:
            nodes[rs.randint(0, n)].setdefault("children", []).append(n)
        # else: unreachable node, ignored by both loaders
    cam = int(rs.randint(1, n_nodes))                            # a root, so that it is reachable for sure
    nodes[cam]["camera"] = 0
    for nd in nodes:
        if cam in nd.get("children", []):
            nd["children"].remove(cam)
    if cam not in roots:
        roots.append(cam)
    j = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": roots, "extras": {"sky_color": [0.1, 0.2, 0.3]}}], "nodes": nodes,
         "cameras": [{"type": "perspective", "perspective": {"yfov": 0.9, "aspectRatio": 1.7, "znear": 0.1}}], "meshes": meshes,
         "materials": materials, "textures": [{"source": 2}, {"source": 0}, {"source": 1}], "images": images,
         "accessors": accessors, "bufferViews": views, "buffers": [{"byteLength": sum(len(b) for b in blobs)}]}
    js = json.dumps(j).encode()
    js += b" " * ((-len(js)) % 4)
    binc = b"".join(blobs)
    with open(path, "wb") as f:
        f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(binc)) + struct.pack("<II", len(js), 0x4E4F534A) + js +
                struct.pack("<II", len(binc), 0x004E4942) + binc)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference sources (build container only)")
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8])
def test_random_scenes_load_like_the_reference(glb, pkg, tmp_path, seed):
    """random node forests / meshes / materials through the reference's own loader and through ours"""
    import _scenref
    path = str(tmp_path / f"random{seed}.glb")
    _write_random_glb(path, seed)
    ref = _scenref.load(path)
    glb.glb_load_scaled.restype = C.c_void_p
    glb.glb_load_scaled.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float]
    mine = _mine_scaled(glb, pkg, path, (1.0, 1.0, 1.0))
    assert len(mine["instances"]) == len(ref["instances"]) > 20
    TYPE = {1: pkg._capi.RT_MAT_DIFFUSE, 2: pkg._capi.RT_MAT_METALLIC, 3: pkg._capi.RT_MAT_DIELECTRIC}
    for a, b in zip(mine["instances"], ref["instances"]):
        for k in ("positions", "normals", "uvs", "indices"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a["transform"].view(np.uint32), b["transform"].view(np.uint32))
        assert a["type"] == TYPE[b["type"]]
        if b["type"] != 3:
            assert a["albedo_image"] == (b["albedo_image"] if b["albedo_is_image"] else -1)
            if not b["albedo_is_image"]:
                assert np.array_equal(a["albedo"], b["albedo"])
            assert np.array_equal(a["emissive"], b["emissive"])
        if b["type"] == 2:
            assert a["roughness"] == b["roughness"]
        if b["type"] == 3:
            assert a["ior"] == b["ior"]
    assert np.array_equal(mine["sky"], ref["sky"])
    assert np.array_equal(mine["camera_position"], ref["camera_position"])
    assert np.array_equal(mine["camera_direction"], ref["camera_direction"]) and mine["focal"] == ref["focal"]
    assert len(mine["layers"]) == len(ref["layers"]) == 3
    for k in range(3):
        assert np.array_equal(mine["layers"][k], ref["layers"][k]), k


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference sources (build container only)")
def test_uri_images_load_like_the_reference(glb, pkg, tmp_path):
    """image "uri"s (base64 data URI, percent-encoded file name next to the .glb) through tinygltf and through ours"""
    import base64
    import _scenref
    gold = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))
    tex = (np.random.RandomState(1).rand(512, 512, 4) * 255).astype(np.uint8)
    (tmp_path / "side car.jpg").write_bytes(gold["in_jpg_progressive_422"].tobytes())
    data_uri = "data:image/png;base64," + base64.b64encode(gold["in_png_h_rgb8_adam7"].tobytes()).decode()
    path = str(tmp_path / "uris.glb")
    _write_glb(path, tex, uri_images=[data_uri, "side%20car.jpg"], f15=False)
    ref = _scenref.load(path)
    glb.glb_load_scaled.restype = C.c_void_p
    glb.glb_load_scaled.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float]
    mine = _mine_scaled(glb, pkg, path, (1.0, 1.0, 1.0))
    assert len(mine["layers"]) == len(ref["layers"]) == 3
    for k in range(3):
        assert np.array_equal(mine["layers"][k], ref["layers"][k]), k


def _scene_data(pkg, mine):
    """our loader's output as the SceneData the oracle (and the CUDA path) render"""
    insts = [pkg.InstanceData(i["positions"], i["normals"], i["uvs"], i["indices"], i["transform"],
                              pkg.Material(i["type"], tuple(i["albedo"]), i["albedo_image"], i["roughness"], i["ior"], tuple(i["emissive"])))
             for i in mine["instances"]]
    tex = np.stack(mine["layers"]) if mine["layers"] else None
    return pkg.SceneData(insts, tex, tuple(mine["sky"]), tuple(mine["camera_position"]), tuple(mine["camera_direction"]), mine["focal"], "glb")


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference sources (build container only)")
@pytest.mark.parametrize("kind", [0, 1])
def test_whole_reference_program_equals_loader_plus_oracle(glb, pkg, oracle, tmp_path, kind):
    """END TO END on the CPU: the reference's whole program — src/scene.cpp loader, Camera, Megakernel /
    WavefrontRenderer::render_frame, write_image, i.e. src/main.cpp:30-70 compiled in place
    (oracle/refshim/fullref.cpp -> oracle/_ref/libfullref.so) — renders a .glb; our loader + the oracle's renderer
    (the thing every CUDA result is compared with) must produce the same out.png bytes and the same ray count."""
    so = os.path.join(ROOT, "oracle", "_ref", "libfullref.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    F = C.CDLL(so)
    F.fullref_main.restype = C.c_uint64
    F.fullref_main.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    F.fullref_last_error.restype = C.c_char_p
    tex = (np.random.RandomState(1).rand(512, 512, 4) * 255).astype(np.uint8)      # 512x512: baked verbatim by both
    path = str(tmp_path / "e2e.glb")
    _write_glb(path, tex, f15=False)
    w, h, depth, spp = 72, 48, 6, 3
    ref_img = np.zeros((h, w, 4), np.uint8)
    rays = F.fullref_main(path.encode(), 1 if kind == 0 else 0, w, h, depth, spp, ref_img.ctypes.data)
    assert rays != 2 ** 64 - 1, F.fullref_last_error()
    glb.glb_load_scaled.restype = C.c_void_p
    glb.glb_load_scaled.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float]
    data = _scene_data(pkg, _mine_scaled(glb, pkg, path, (1.0, 1.0, 1.0)))
    o = oracle.Scene(data).render(oracle.camera_for(data, w, h), kind, depth, spp, use_bvh=False)
    assert o["ray_count"] == rays > w * h * spp
    assert np.array_equal(o["rgba8"], ref_img)
    assert ref_img[..., :3].std() > 5 and (ref_img[..., 3] == 255).all()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference sources (build container only)")
@pytest.mark.parametrize("seed", [11, 12])
def test_whole_reference_program_on_random_scenes(glb, pkg, oracle, tmp_path, seed):
    """the end-to-end comparison on random scenes: ~60 instances under a random node forest, textured diffuse /
    metallic / dielectric / emissive materials, both renderers"""
    F = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfullref.so"))
    F.fullref_main.restype = C.c_uint64
    F.fullref_main.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    path = str(tmp_path / "e2e_random.glb")
    _write_random_glb(path, seed, n_nodes=40, tex512=True)
    glb.glb_load_scaled.restype = C.c_void_p
    glb.glb_load_scaled.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float]
    data = _scene_data(pkg, _mine_scaled(glb, pkg, path, (1.0, 1.0, 1.0)))
    w, h, depth, spp = 48, 32, 5, 2
    for kind in (0, 1):
        ref_img = np.zeros((h, w, 4), np.uint8)
        rays = F.fullref_main(path.encode(), 1 if kind == 0 else 0, w, h, depth, spp, ref_img.ctypes.data)
        o = oracle.Scene(data).render(oracle.camera_for(data, w, h), kind, depth, spp, use_bvh=False)
        assert o["ray_count"] == rays > w * h * spp * 1.02     # some paths do bounce
        assert np.array_equal(o["rgba8"], ref_img)


E2E_GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_program_e2e.npz")
# name -> (writer, width, height, max_depth, samples)
E2E_CASES = {
    "synthetic": (lambda p: _write_glb(p, (np.random.RandomState(1).rand(512, 512, 4) * 255).astype(np.uint8), f15=False), 72, 48, 6, 3),
    "random11": (lambda p: _write_random_glb(p, 11, n_nodes=40, tex512=True), 48, 32, 5, 2),   # ~50 textured instances
}


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference sources (build container only)")
def test_e2e_golden_is_what_the_reference_program_renders_today(tmp_path):
    """tests/golden/reference_program_e2e.npz = out.png bytes + ray counts of the reference's whole program
    (libfullref.so) on generated .glb files; regenerate with RT_WRITE_GOLDEN=1"""
    F = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfullref.so"))
    F.fullref_main.restype = C.c_uint64
    F.fullref_main.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    out = {}
    for case, (writer, w, h, depth, spp) in E2E_CASES.items():
        path = str(tmp_path / f"{case}.glb")
        writer(path)
        out[f"{case}_glb_crc"] = np.uint32(zlib.crc32(open(path, "rb").read()))
        for kind, name in ((0, "megakernel"), (1, "wavefront")):
            img = np.zeros((h, w, 4), np.uint8)
            rays = F.fullref_main(path.encode(), 1 if kind == 0 else 0, w, h, depth, spp, img.ctypes.data)
            out[f"{case}_img_{name}"], out[f"{case}_rays_{name}"] = img, np.uint64(rays)
    if os.environ.get("RT_WRITE_GOLDEN"):
        np.savez_compressed(E2E_GOLDEN, **out)
    gold = np.load(E2E_GOLDEN)
    for k in out:
        assert np.array_equal(gold[k], out[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(E2E_CASES))
@pytest.mark.parametrize("flag,name", [("-m", "megakernel"), ("-w", "wavefront")])
def test_cli_output_equals_the_reference_programs(glb, tmp_path, flag, name, case):
    """`raytracer [-m|-w] scene.glb` on the B200 writes the out.png the reference's whole program writes
    (fixture rendered by the reference's own code, see above) and prints the same ray count"""
    from PIL import Image
    gold = np.load(E2E_GOLDEN)
    writer, w, h, depth, spp = E2E_CASES[case]
    path = str(tmp_path / "e2e.glb")
    writer(path)
    assert zlib.crc32(open(path, "rb").read()) == int(gold[f"{case}_glb_crc"])   # the very file the fixture was rendered from
    png = str(tmp_path / "out.png")
    r = subprocess.run([os.path.join(HOST, "raytracer"), flag, "-d", str(depth), "-s", str(spp), "--size", f"{w}x{h}", "--png", png, path],
                       capture_output=True, text=True, check=True)
    assert f"Total rays: {int(gold[f'{case}_rays_{name}'])}" in r.stdout
    assert np.array_equal(np.array(Image.open(png)), gold[f"{case}_img_{name}"])


def _patch_glb_json(path, fn):
    """rewrite the JSON chunk of a .glb in place (the BIN chunk is kept)"""
    raw = open(path, "rb").read()
    jlen = struct.unpack_from("<I", raw, 12)[0]
    j = json.loads(raw[20:20 + jlen].decode())
    fn(j)
    js = json.dumps(j).encode()
    js += b" " * ((-len(js)) % 4)
    rest = raw[20 + jlen:]
    with open(path, "wb") as f:
        f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + len(rest)) + struct.pack("<II", len(js), 0x4E4F534A) + js + rest)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference sources (build container only)")
@pytest.mark.parametrize("roots", [[0, 3, 5], [5, 0, 3], [3, 5, 0], [5, 3]])
def test_several_cameras_pick_the_references_camera(glb, pkg, tmp_path, roots):
    """Scene::load_node recurses in pre-order and the LAST camera node it visits wins (src/scene.cpp:444-457): sibling
    camera roots, a camera below a camera, and every order of the scene roots must frame the image like the reference"""
    import _scenref
    tex = (np.random.RandomState(1).rand(512, 512, 4) * 255).astype(np.uint8)
    path = str(tmp_path / "cams.glb")
    _write_glb(path, tex, f15=False)

    def patch(j):
        j["nodes"].append({"translation": [3.0, 2.0, 1.0], "rotation": [0.0, np.sin(0.4), 0.0, np.cos(0.4)], "camera": 0, "children": [6, 7]})  # 5
        j["nodes"].append({"translation": [0.0, 0.5, 1.0], "camera": 0})                                                                        # 6
        j["nodes"].append({"translation": [1.0, 0.0, 0.0], "mesh": 1})                                                                          # 7: no camera
        j["scenes"][0]["nodes"] = roots
    _patch_glb_json(path, patch)
    ref = _scenref.load(path)
    glb.glb_load_scaled.restype = C.c_void_p
    glb.glb_load_scaled.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float]
    mine = _mine_scaled(glb, pkg, path, (1.0, 1.0, 1.0))
    assert ref["camera_node"] == {(0, 3, 5): 6, (5, 0, 3): 3, (3, 5, 0): 6, (5, 3): 3}[tuple(roots)]
    assert np.array_equal(mine["camera_position"], ref["camera_position"])
    assert np.array_equal(mine["camera_direction"], ref["camera_direction"])
    assert mine["focal"] == ref["focal"]
    assert len(mine["instances"]) == len(ref["instances"])
    for a, b in zip(mine["instances"], ref["instances"]):
        assert np.array_equal(a["transform"].view(np.uint32), b["transform"].view(np.uint32))
