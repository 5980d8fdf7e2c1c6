"""Deterministic synthetic scenes for BASELINE.json's configs.

Pure numpy, no file or network access; the same arrays feed the CUDA path, the oracle and the
host emulation. All randomness comes from a splitmix64 hash seeded 0x5EED0001 + scene id, so the
scenes are identical across numpy versions and machines.

    cube_scene()          C1  assets/cube.glb stand-in: 12 triangles, diffuse 0.8 grey, camera at the
                              origin looking down -Z, focal 1 (explicit fallbacks for F15)
    cornell_scene()       C2  Cornell box + dielectric and metallic icospheres (~41 k triangles)
    sponza_scale_scene()  C3/C5  height-field floor + walls + 25 icospheres, textured (~261 k)
    big_mesh_scene(n)     C4  displaced height field, n x n cells (2236 -> 9,999,392 triangles)
    stadium_scene()       a BVH-HOSTILE input (not one of BASELINE's configs): "teapots in a stadium" — 100 m architectural
                              quads and 60 m long, 0.2 m wide diagonal beams next to 5,120-triangle props (~154 k triangles);
                              what real glTF scenes hand to Embree's SAH builder (src/scene.cpp:406-439)
"""
import numpy as np

from . import InstanceData, Material, SceneData

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _hash01(seed, *coords):
    """uniform [0,1) from integer coordinates (vectorised)."""
    with np.errstate(over="ignore"):
        h = np.uint64(seed)
        for c in coords:
            h = _splitmix64(h ^ (np.asarray(c).astype(np.int64).astype(np.uint64) * np.uint64(0xD6E8FEB86659FD93)))
        return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class _Stream:
    def __init__(self, seed):
        self.seed, self.k = int(seed), 0

    def uniform(self, lo=0.0, hi=1.0):
        self.k += 1
        return float(lo + (hi - lo) * _hash01(self.seed, np.array([self.k]))[0])


def _value_noise(seed, x, y):
    """bilinear value noise on the integer lattice, smoothstep weights."""
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = x - x0, y - y0
    fx, fy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    a, b = _hash01(seed, x0, y0), _hash01(seed, x0 + 1, y0)
    c, d = _hash01(seed, x0, y0 + 1), _hash01(seed, x0 + 1, y0 + 1)
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def _fbm(seed, x, y, octaves=4):
    out, amp, f = 0.0, 0.5, 1.0
    for o in range(octaves):
        out = out + amp * _value_noise(seed + 101 * o, x * f, y * f)
        amp *= 0.5
        f *= 2.0
    return out


def quad(p0, p1, p2, p3, normal, uv_scale=1.0):
    """two triangles p0 p1 p2 / p0 p2 p3 with a constant normal."""
    pos = np.array([p0, p1, p2, p3], np.float32)
    nrm = np.tile(np.asarray(normal, np.float32), (4, 1))
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32) * uv_scale
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    return pos, nrm, uv, idx


def grid_mesh(n, origin, eu, ev, normal, uv_tiles=1.0):
    """flat (n x n)-cell grid spanned by eu, ev from origin."""
    s = np.linspace(0.0, 1.0, n + 1)
    u, v = np.meshgrid(s, s, indexing="xy")
    pos = (np.asarray(origin)[None, :] + u.reshape(-1, 1) * np.asarray(eu)[None, :]
           + v.reshape(-1, 1) * np.asarray(ev)[None, :]).astype(np.float32)
    nrm = np.tile(np.asarray(normal, np.float32), (pos.shape[0], 1))
    uv = np.stack([u.reshape(-1) * uv_tiles, v.reshape(-1) * uv_tiles], 1).astype(np.float32)
    return pos, nrm, uv, _grid_indices(n)


def _grid_indices(n):
    i, j = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64), indexing="xy")
    a = (j * (n + 1) + i).reshape(-1)
    b, c, d = a + 1, a + (n + 1) + 1, a + (n + 1)
    return np.stack([a, b, c, a, c, d], 1).reshape(-1).astype(np.uint32)


def heightfield(n, half_extent, height, seed, freq, uv_tiles=8.0):
    """(n x n)-cell height field on the XZ plane, smooth analytic-difference normals."""
    s = np.linspace(-1.0, 1.0, n + 1)
    x, z = np.meshgrid(s, s, indexing="xy")

    def h(xx, zz):
        return height * _fbm(seed, (xx + 1.0) * freq, (zz + 1.0) * freq)

    y = h(x, z)
    e = 1.0 / n
    dydx = (h(x + e, z) - h(x - e, z)) / (2 * e * half_extent)
    dydz = (h(x, z + e) - h(x, z - e)) / (2 * e * half_extent)
    nrm = np.stack([-dydx, np.ones_like(y), -dydz], -1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    pos = np.stack([x * half_extent, y, z * half_extent], -1).reshape(-1, 3).astype(np.float32)
    uv = np.stack([(x + 1) * 0.5 * uv_tiles, (z + 1) * 0.5 * uv_tiles], -1).reshape(-1, 2).astype(np.float32)
    # winding so that the geometric normal points +Y: (i,j) (i,j+1) (i+1,j+1)
    idx = _grid_indices(n).reshape(-1, 3)[:, [0, 2, 1]].reshape(-1)
    return pos, nrm.reshape(-1, 3).astype(np.float32), uv, idx.astype(np.uint32)


def icosphere(subdivisions):
    """unit icosphere: 20 * 4^s triangles, shared vertices, smooth normals = positions."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t],
                  [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdivisions):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        es = np.sort(e, 1)
        key = es[:, 0] * (len(v) + 1) + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // (len(v) + 1), uniq % (len(v) + 1)
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        m = len(v) + inv.reshape(3, -1)  # midpoint ids of edges 01, 12, 20 per face
        v = np.concatenate([v, mid], 0)
        f = np.concatenate([np.stack([f[:, 0], m[0], m[2]], 1), np.stack([f[:, 1], m[1], m[0]], 1),
                            np.stack([f[:, 2], m[2], m[1]], 1), np.stack([m[0], m[1], m[2]], 1)], 0)
    uv = np.stack([0.5 + np.arctan2(v[:, 2], v[:, 0]) / (2 * np.pi), 0.5 - np.arcsin(np.clip(v[:, 1], -1, 1)) / np.pi], 1)
    return v.astype(np.float32), v.astype(np.float32).copy(), uv.astype(np.float32), f.reshape(-1).astype(np.uint32)


def trs(translate=(0, 0, 0), scale=(1, 1, 1), rot_y=0.0):
    c, s = np.cos(rot_y), np.sin(rot_y)
    r = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], np.float64)
    sc = np.diag([scale[0], scale[1], scale[2], 1.0])
    t = np.eye(4)
    t[:3, 3] = translate
    return (t @ r @ sc).astype(np.float32)


def cube_mesh():
    """24 vertices in {+-1}^3, flat per-face normals, 12 triangles (the shape of assets/cube.glb)."""
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
             ((0, -1, 0), (1, 0, 0), (0, 0, 1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    pos, nrm, uv, idx = [], [], [], []
    for k, (n, a, b) in enumerate(faces):
        n, a, b = np.array(n, np.float32), np.array(a, np.float32), np.array(b, np.float32)
        for (sa, sb) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            pos.append(n + sa * a + sb * b)
            nrm.append(n)
            uv.append(((sa + 1) / 2, (sb + 1) / 2))
        idx += [4 * k, 4 * k + 1, 4 * k + 2, 4 * k, 4 * k + 2, 4 * k + 3]
    return np.array(pos, np.float32), np.array(nrm, np.float32), np.array(uv, np.float32), np.array(idx, np.uint32)


def cube_scene():
    """C1. Node translation is the one in assets/cube.glb; material / camera are the
    explicit fallbacks for what the reference leaves to undefined behaviour (F15)."""
    p, n, uv, i = cube_mesh()
    inst = InstanceData(p, n, uv, i, trs((0.05813104659318924, 0.1505535989999771, -2.920884370803833)),
                        Material.diffuse((0.8, 0.8, 0.8)))
    return SceneData([inst], None, (0.5, 0.7, 1.0), (0, 0, 0), (0, 0, -1), 1.0, "cube")


def procedural_textures(layers=8, seed=0x5EED0101):
    """seeded checker / noise RGBA8 layers, 512 x 512 (the baked array of src/image_manager.hpp)."""
    ys, xs = np.meshgrid(np.arange(512), np.arange(512), indexing="ij")
    out = np.zeros((layers, 512, 512, 4), np.uint8)
    for l in range(layers):
        st = _Stream(seed + l)
        c0 = np.array([st.uniform(0.55, 0.95) for _ in range(3)])
        c1 = np.array([st.uniform(0.15, 0.6) for _ in range(3)])
        cell = 2 ** (3 + l % 4)
        chk = (((xs // cell) + (ys // cell)) & 1).astype(np.float64)
        noise = _fbm(seed + 31 * l, xs / 37.0, ys / 37.0, 3)
        w = np.clip(0.75 * chk + 0.5 * (noise - 0.45), 0, 1)[..., None]
        rgb = c0[None, None, :] * (1 - w) + c1[None, None, :] * w
        out[l, :, :, :3] = np.clip(np.rint(rgb * 255.0), 0, 255).astype(np.uint8)
        out[l, :, :, 3] = 255
    return out


def cornell_scene(sphere_subdiv=5):
    """C2: box [-1,1]^3 open toward +Z, emissive ceiling quad, dielectric and metallic
    icospheres; sky black; camera (0,0,3.4) -> -Z, focal 2."""
    white, red, green = (0.73, 0.73, 0.73), (0.65, 0.05, 0.05), (0.12, 0.45, 0.15)
    I = []
    I.append(InstanceData(*quad((-1, -1, 1), (1, -1, 1), (1, -1, -1), (-1, -1, -1), (0, 1, 0)), None, Material.diffuse(white)))
    I.append(InstanceData(*quad((-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1), (0, -1, 0)), None, Material.diffuse(white)))
    I.append(InstanceData(*quad((-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (0, 0, 1)), None, Material.diffuse(white)))
    I.append(InstanceData(*quad((-1, -1, 1), (-1, -1, -1), (-1, 1, -1), (-1, 1, 1), (1, 0, 0)), None, Material.diffuse(red)))
    I.append(InstanceData(*quad((1, -1, -1), (1, -1, 1), (1, 1, 1), (1, 1, -1), (-1, 0, 0)), None, Material.diffuse(green)))
    I.append(InstanceData(*quad((-0.25, 0.998, -0.25), (0.25, 0.998, -0.25), (0.25, 0.998, 0.25), (-0.25, 0.998, 0.25),
                                (0, -1, 0)), None, Material.diffuse((0.78, 0.78, 0.78), emissive=(15, 15, 15))))
    # (the light keeps a non-black albedo on purpose: under the reference's radiance rule (F7) the
    # emitted term is scaled by the WHOLE path attenuation, emitter included, so a black-albedo
    # emitter would contribute nothing and the box would render black)
    sp = icosphere(sphere_subdiv)
    I.append(InstanceData(*sp, trs((0.4, -0.65, 0.3), (0.35,) * 3), Material.dielectric(1.5)))
    I.append(InstanceData(*sp, trs((-0.4, -0.65, -0.3), (0.35,) * 3), Material.metallic((0.9, 0.9, 0.9), 0.1)))
    return SceneData(I, None, (0, 0, 0), (0, 0, 3.4), (0, 0, -1), 2.0, "cornell")


def sponza_scale_scene(field_cells=256, sphere_subdiv=4, n_spheres=25, seed=0x5EED0003):
    """C3 / C5: 256x256-cell value-noise floor (131,072 tris) inside four 16x16-quad
    walls open to the sky + 25 icospheres of 5,120 tris => 261,120 triangles; 8 textures;
    70 % textured diffuse / 20 % metallic / 10 % dielectric by seeded draw; default sky."""
    st = _Stream(seed)
    tex = procedural_textures(8, seed + 0x100)
    E, wall_h = 8.0, 6.0
    I = [InstanceData(*heightfield(field_cells, E, 0.9, seed + 1, 3.0, 8.0), None, Material.diffuse(image=0))]
    walls = [((-E, -0.2, -E), (2 * E, 0, 0), (0, wall_h, 0), (0, 0, 1)), ((E, -0.2, E), (-2 * E, 0, 0), (0, wall_h, 0), (0, 0, -1)),
             ((-E, -0.2, E), (0, 0, -2 * E), (0, wall_h, 0), (1, 0, 0)), ((E, -0.2, -E), (0, 0, 2 * E), (0, wall_h, 0), (-1, 0, 0))]
    for k, (o, eu, ev, n) in enumerate(walls):
        I.append(InstanceData(*grid_mesh(16, o, eu, ev, n, 4.0), None, Material.diffuse(image=1 + k % 3)))
    sp = icosphere(sphere_subdiv)
    for k in range(n_spheres):
        r = st.uniform(0.35, 0.95)
        x, z = st.uniform(-6.5, 6.5), st.uniform(-6.5, 5.0)
        y = 0.9 * float(_fbm(seed + 1, np.array([(x / E + 1.0) * 3.0]), np.array([(z / E + 1.0) * 3.0]))[0]) + r * st.uniform(0.8, 1.6)
        kind = st.uniform()
        if kind < 0.7:
            m = Material.diffuse(image=int(st.uniform(0, 8)) % 8)
        elif kind < 0.9:
            m = Material.metallic((st.uniform(0.6, 0.95), st.uniform(0.6, 0.95), st.uniform(0.6, 0.95)), st.uniform(0.0, 0.5))
        else:
            m = Material.dielectric(1.5)
        I.append(InstanceData(*sp, trs((x, y, z), (r, r, r), st.uniform(0, 6.28)), m))
    return SceneData(I, tex, (0.5, 0.7, 1.0), (0.0, 2.2, 7.6), (0.0, -0.22, -1.0), 1.5, "sponza_scale")


def _cached(tag, make):
    """the 10 M-triangle field takes ~45 s of numpy per process: keep the arrays of big meshes in a per-machine
    scratch file (RT_SCENE_CACHE overrides the directory, empty disables). Deterministic content, so a cache hit
    returns the same bytes the generator would."""
    import os
    import tempfile
    d = os.environ.get("RT_SCENE_CACHE", tempfile.gettempdir())
    path = os.path.join(d, f"rt_b200_scene_{tag}.npz") if d else None
    if path and os.path.exists(path):
        try:
            z = np.load(path)
            return z["pos"], z["nrm"], z["uv"], z["idx"]
        except Exception:
            pass
    arrs = make()
    if path:
        try:
            tmp = path + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, pos=arrs[0], nrm=arrs[1], uv=arrs[2], idx=arrs[3])
            os.replace(tmp, path)
        except Exception:
            pass
    return arrs


def big_mesh_scene(cells=2236, seed=0x5EED0004):
    """C4: one displaced height field, 2 * cells^2 triangles (2236 -> 9,999,392), diffuse 0.7 grey,
    grazing camera so rays span the whole BVH."""
    mesh = (lambda: heightfield(cells, 50.0, 4.0, seed + 1, 12.0, 32.0))
    arrs = _cached(f"heightfield_{cells}_{seed:x}", mesh) if cells >= 1024 else mesh()
    I = [InstanceData(*arrs, None, Material.diffuse((0.7, 0.7, 0.7)))]
    return SceneData(I, None, (0.5, 0.7, 1.0), (0.0, 6.0, 49.0), (0.0, -0.08, -1.0), 1.5, f"heightfield_{cells}")


def _beam(a, b, width, up=(0.0, 1.0, 0.0)):
    """a flat strip from a to b, `width` wide: two long thin triangles"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = b - a
    side = np.cross(d, np.asarray(up, np.float64))
    if np.linalg.norm(side) < 1e-9:
        side = np.cross(d, (1.0, 0.0, 0.0))
    side *= 0.5 * width / np.linalg.norm(side)
    n = np.cross(d, side)
    n /= np.linalg.norm(n)
    return quad(tuple(a - side), tuple(b - side), tuple(b + side), tuple(a + side), tuple(n), 8.0)


def stadium_scene(n_props=30, prop_subdiv=4, n_beams=48, seed=0x5EED0006):
    """The input that is hard for a Morton-ordered tree ("Tree quality vs Embree"): a 100 m x 100 m stadium whose
    floor, walls and stands are a handful of huge triangles, crossed by `n_beams` long thin DIAGONAL strips (their boxes span
    most of the scene), with `n_props` finely tessellated props (5,120 triangles, 0.3 - 0.9 m) standing on the pitch in front
    of the camera. 30 props + 48 beams => 153,726 triangles, 70 % diffuse / 20 % metallic / 10 % dielectric props."""
    st = _Stream(seed)
    E, H = 50.0, 20.0
    grey, green = (0.6, 0.6, 0.62), (0.25, 0.5, 0.22)
    I = [InstanceData(*quad((-E, 0, E), (E, 0, E), (E, 0, -E), (-E, 0, -E), (0, 1, 0), 16.0), None, Material.diffuse(green))]
    walls = [((-E, 0, -E), (E, 0, -E), (E, H, -E), (-E, H, -E), (0, 0, 1)), ((E, 0, E), (-E, 0, E), (-E, H, E), (E, H, E), (0, 0, -1)),
             ((-E, 0, E), (-E, 0, -E), (-E, H, -E), (-E, H, E), (1, 0, 0)), ((E, 0, -E), (E, 0, E), (E, H, E), (E, H, -E), (-1, 0, 0))]
    for p0, p1, p2, p3, n in walls:
        I.append(InstanceData(*quad(p0, p1, p2, p3, n, 8.0), None, Material.diffuse(grey)))
    # sloped stands behind the far goal and along one side: again two triangles each
    I.append(InstanceData(*quad((-E, 0, -30.0), (E, 0, -30.0), (E, 12.0, -E), (-E, 12.0, -E), (0, 0.8, 0.6), 8.0), None, Material.diffuse((0.5, 0.45, 0.4))))
    I.append(InstanceData(*quad((30.0, 0, E), (30.0, 0, -E), (E, 12.0, -E), (E, 12.0, E), (-0.6, 0.8, 0), 8.0), None, Material.diffuse((0.5, 0.45, 0.4))))
    for k in range(n_beams):  # roof trusses and cables: corner to corner through the volume
        a = (st.uniform(-E, E), st.uniform(8.0, H), st.uniform(-E, E))
        ang, ln = st.uniform(0.0, 6.28318), st.uniform(40.0, 70.0)
        b = (a[0] + ln * np.cos(ang), st.uniform(8.0, H), a[2] + ln * np.sin(ang))
        b = (float(np.clip(b[0], -E, E)), b[1], float(np.clip(b[2], -E, E)))
        I.append(InstanceData(*_beam(a, b, 0.2), None, Material.metallic((0.8, 0.8, 0.85), 0.3) if k % 3 == 0 else Material.diffuse((0.3, 0.3, 0.35))))
    sp = icosphere(prop_subdiv)
    for k in range(n_props):
        r = st.uniform(0.3, 0.9)
        x, z = st.uniform(-9.0, 9.0), st.uniform(-14.0, 4.0)
        kind = st.uniform()
        if kind < 0.7:
            m = Material.diffuse((st.uniform(0.2, 0.9), st.uniform(0.2, 0.9), st.uniform(0.2, 0.9)))
        elif kind < 0.9:
            m = Material.metallic((st.uniform(0.6, 0.95), st.uniform(0.6, 0.95), st.uniform(0.6, 0.95)), st.uniform(0.0, 0.5))
        else:
            m = Material.dielectric(1.5)
        I.append(InstanceData(*sp, trs((x, r * st.uniform(1.0, 1.8), z), (r, r, r), st.uniform(0, 6.28)), m))
    return SceneData(I, None, (0.5, 0.7, 1.0), (0.0, 3.0, 14.0), (0.0, -0.12, -1.0), 1.5, "stadium")


def random_soup(n_tris, seed=1, extent=1.0, instances=1):
    """small random triangle soup for intersection tests (ragged / degenerate cases included)."""
    out = []
    for k in range(instances):
        ids = np.arange(n_tris * 9).reshape(n_tris, 3, 3)
        c = (_hash01(seed + 17 * k, ids[:, :1, :] // 9 * 9 + np.arange(3)[None, None, :]) - 0.5) * 2 * extent
        p = c + (_hash01(seed + 17 * k + 5, ids) - 0.5) * 0.6 * extent
        pos = p.reshape(-1, 3).astype(np.float32)
        nrm = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
        nrm = np.repeat(nrm, 3, 0).astype(np.float32)
        uv = _hash01(seed + 17 * k + 9, np.arange(n_tris * 6).reshape(-1, 2)).astype(np.float32)
        idx = np.arange(n_tris * 3, dtype=np.uint32)
        mats = [Material.diffuse((0.7, 0.6, 0.5)), Material.metallic((0.9, 0.8, 0.7), 0.2), Material.dielectric(1.5)]
        out.append(InstanceData(pos, nrm, uv, idx, trs((0.3 * k, 0.1 * k, -0.2 * k), (1, 1, 1), 0.4 * k), mats[k % 3]))
    return SceneData(out, None, (0.5, 0.7, 1.0), (0, 0, 3.0 * extent), (0, 0, -1), 1.0, f"soup_{n_tris}x{instances}")
