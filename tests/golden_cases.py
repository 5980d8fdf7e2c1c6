"""Inputs of the golden fixtures (shared by tests/tools/make_golden.py, which runs the reference's own
code on them, and by the tests that check the oracle and the CUDA path against the stored outputs)."""
import importlib

import numpy as np

CAMERAS = [(1920, 1080, (0, 0, 0), (0, 0, -1), 1.0), (64, 48, (0.5, 1.0, 2.0), (0.2, -0.1, -1.0), 1.5),
           (250, 3, (0, 2.2, 7.6), (0, -0.22, -1.0), 1.5)]
CAMERA_PIXELS = [(0, 0, 0), (1, 1, 1081), (1, 1, 1921), (37, 2, 777), (63, 47, 0xDEADBEEF), (249, 0, 5)]


def scatter_inputs(n_per=200):
    rs = np.random.RandomState(2024)
    out = []
    for mtype in (1, 2, 3, 0):
        for k in range(n_per if mtype else 4):
            d = rs.randn(3).astype(np.float32)
            d /= np.linalg.norm(d)
            n = rs.randn(3).astype(np.float32)
            n /= np.linalg.norm(n)
            if mtype == 3 and k % 3 == 0:   # grazing from inside: total internal reflection (0 draws)
                d = (n * np.float32(0.3) + np.cross(n, rs.randn(3)).astype(np.float32)).astype(np.float32)
                d /= np.linalg.norm(d)
            out.append((mtype, float(rs.rand() * 0.5), 1.5 if k % 2 else 1.33, int(rs.randint(1, 2 ** 31 - 1)),
                        [float(v) for v in d], [float(v) for v in n], [float(rs.rand() * 3 - 1), float(rs.rand() * 3 - 1)]))
    return out


def _scenes():
    return importlib.import_module("sycl-ray-tracer_b200.scenes")


def _textured():
    pkg, sc = importlib.import_module("sycl-ray-tracer_b200"), _scenes()
    tex = sc.procedural_textures(3, 1234)
    p, n, uv, i = sc.grid_mesh(4, (-2, -2, -3), (4, 0, 0), (0, 4, 0), (0, 0, 1), 3.7)
    uv = uv - 1.3
    insts = [pkg.InstanceData(p, n, uv, i, None, pkg.Material.diffuse(image=2)),
             pkg.InstanceData(p, n, uv * 0.5, i, sc.trs((0.5, 0.2, 0.5), (0.3, 0.3, 0.3)), pkg.Material.metallic(roughness=0.3, image=1)),
             pkg.InstanceData(*sc.icosphere(1), sc.trs((-0.6, -0.4, -1.5), (0.4, 0.4, 0.4), 0.7), pkg.Material.dielectric(1.5))]
    return pkg.SceneData(insts, tex, (0.5, 0.7, 1.0), (0, 0, 0), (0, 0, -1), 1.0, "textured")


# name -> (scene factory, width, height, max_depth, spp)
RENDER_CASES = {
    "cube": (lambda: _scenes().cube_scene(), 64, 48, 8, 3),           # config 1's scene
    "cornell": (lambda: _scenes().cornell_scene(2), 48, 32, 10, 4),   # diffuse / metallic / dielectric / emissive
    "textured": (_textured, 40, 30, 6, 3),                             # Texture::sample through the image array
    "soup": (lambda: _scenes().random_soup(60, 11, 1.0, 3), 36, 24, 5, 2),  # transforms + all materials
    "odd": (lambda: _scenes().cube_scene(), 33, 17, 3, 2),            # ragged size: padding in the seed mapping
}
