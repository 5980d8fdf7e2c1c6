"""The C-ABI library loads on a machine without a GPU, exports every symbol include/rt_api.h
declares, and refuses to create a context (no CPU path) instead of silently falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "rt_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._capi.load()
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rt_api.h but not exported"
        assert n in pkg._capi.SYMBOLS, f"{n} has no ctypes binding"
    assert lib.rt_api_version() == 2


def test_struct_layouts_match_the_header(pkg):
    cap = pkg._capi
    assert C.sizeof(cap.rt_material) == 40
    assert C.sizeof(cap.rt_instance) == 4 * 8 + 8 + 64 + 40
    assert C.sizeof(cap.rt_camera) == 56
    assert C.sizeof(cap.rt_render_params) == 32
    assert C.sizeof(cap.rt_group_params) == 24
    assert C.sizeof(cap.rt_ipc_handle) == 64
    assert C.sizeof(cap.rt_frame) == 40
    assert C.sizeof(cap.rt_scene_stats) == 48


def test_camera_init_is_host_only(pkg, oracle):
    cam = pkg.Camera((1920, 1080), (1, 2, 3), (0.3, -0.2, -1), 1.7)
    o = oracle.camera(1920, 1080, (1, 2, 3), (0.3, -0.2, -1), 1.7)
    for f in ("center", "pixel00_loc", "pixel_delta_u", "pixel_delta_v", "img_size"):
        assert list(getattr(cam.c, f)) == list(getattr(o, f))


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.RtError) as e:
        pkg.App(0)
    assert "rt_context_create failed" in str(e.value)
