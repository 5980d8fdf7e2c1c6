"""The C++ host mirror (host/raytracer.hpp + host/main.cpp, the reference's CLI shape)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "sycl-ray-tracer_b200", "host")
BIN = os.path.join(HOST, "raytracer")


def _build():
    subprocess.run(["make", "-s", "-C", HOST], check=True)


def test_cli_builds_and_fails_loudly_without_gpu():
    import torch
    _build()
    assert os.path.exists(BIN)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([BIN, "-m", "-s", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "rt_context_create" in r.stdout   # no CPU path


@pytest.mark.gpu
def test_cli_prints_the_benchmark_lines_and_matches_oracle(tmp_path, oracle, scenes):
    """stdout protocol of src/render_megakernel.cpp:181-183 (what benchmark.py:49-55 parses)"""
    _build()
    for flag, mode in (("-m", 0), ("-w", 1)):
        ppm = str(tmp_path / f"out{mode}.ppm")
        r = subprocess.run([BIN, flag, "-d", "8", "-s", "2", "--size", "128x96", "--ppm", ppm, "cube"],
                           capture_output=True, text=True, check=True)
        t = float(re.search(r"Time measured: ([\d\.]+) seconds", r.stdout).group(1))
        rays = int(re.search(r"Total rays: (\d+)", r.stdout).group(1))
        rate = float(re.search(r"Rays/sec: ([\d\.]+)M", r.stdout).group(1))
        assert t > 0 and rate > 0
        data = scenes.cube_scene()
        o = oracle.Scene(data).render(oracle.camera_for(data, 128, 96), mode, 8, 2)
        assert rays == o["ray_count"]
        raw = open(ppm, "rb").read()
        img = np.frombuffer(raw[raw.index(b"255\n") + 4:], np.uint8).reshape(96, 128, 3)
        assert np.array_equal(img, o["rgba8"][..., :3])
