"""Awkward scenes shared by the host-emulation and GPU parity tests."""
import importlib

import numpy as np


def awkward_transforms():
    """mirrored (negative determinant), strongly non-uniform and huge/tiny instance transforms; many
    instances sharing one mesh; coordinates far from the origin"""
    pkg = importlib.import_module("sycl-ray-tracer_b200")
    scenes = importlib.import_module("sycl-ray-tracer_b200.scenes")
    sp = scenes.icosphere(1)
    rs = np.random.RandomState(8)
    insts = []
    for k in range(40):
        s = rs.rand(3) * 2 + 0.05
        if k % 4 == 0:
            s[0] = -s[0]                      # mirror: det < 0
        if k % 7 == 0:
            s *= 40.0                         # huge next to tiny
        if k % 9 == 0:
            s *= 0.02
        t = rs.randn(3) * 6 + (1000.0 if k % 11 == 0 else 0.0)
        mats = [pkg.Material.diffuse((0.6, 0.7, 0.8)), pkg.Material.metallic((0.9, 0.9, 0.9), 0.15), pkg.Material.dielectric(1.4)]
        insts.append(pkg.InstanceData(*sp, scenes.trs(tuple(t), tuple(s), rs.rand() * 6.28), mats[k % 3]))
    data = pkg.SceneData(insts, None, (0.5, 0.7, 1.0), (0, 0, 30), (0, 0, -1), 1.2)
    org = (rs.randn(20000, 3) * 8).astype(np.float32)
    d = rs.randn(20000, 3).astype(np.float32)
    return data, org, d


def coincident_centroids():
    """all Morton keys equal (every triangle has the same centroid) in a perfectly flat scene: the
    radix tree falls back to index bits, the quantisation exponent of a zero extent is clamped"""
    pkg = importlib.import_module("sycl-ray-tracer_b200")
    rs = np.random.RandomState(3)
    n = 300
    ang = rs.rand(n, 3) * 6.28
    r = rs.rand(n, 1) + 0.2
    p = np.zeros((n, 3, 3), np.float32)
    for k in range(3):
        a = ang[:, 0] + k * 2.0943951
        p[:, k, 0], p[:, k, 1] = (r[:, 0] * np.cos(a)), (r[:, 0] * np.sin(a))
    inst = pkg.InstanceData(p.reshape(-1, 3), np.tile(np.float32([0, 0, 1]), (n * 3, 1)), np.zeros((n * 3, 2), np.float32),
                            np.arange(n * 3, dtype=np.uint32))
    data = pkg.SceneData([inst], None, (0.5, 0.7, 1.0), (0, 0, 3), (0, 0, -1), 1.0)
    org = np.tile(np.float32([0, 0, 3]), (6000, 1)) + (rs.rand(6000, 3).astype(np.float32) - 0.5)
    d = np.float32([0, 0, -1]) + (rs.rand(6000, 3).astype(np.float32) - 0.5) * 0.9
    return data, org, d
