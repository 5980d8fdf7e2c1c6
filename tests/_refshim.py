"""ctypes binding of oracle/_ref/librefshim.so: the reference's OWN unmodified sources compiled against
API shims (oracle/refshim). Only exists where /root/reference does (the build container); used to pin
the oracle and to generate tests/golden/."""
import ctypes as C
import os

import numpy as np

import _oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "librefshim.so")
f32p, u32p, u8p = _oracle.f32p, _oracle.u32p, _oracle.u8p
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.POINTER(_oracle.orc_instance), C.c_uint32, u8p, C.c_uint32, f32p, C.c_int]
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_xorshift_next.restype = C.c_float
        L.ref_xorshift_next.argtypes = [u32p]
        L.ref_random_unit_vector.argtypes = [u32p, f32p]
        L.ref_camera.argtypes = [C.c_int, C.c_int, f32p, f32p, C.c_float, f32p]
        L.ref_camera_get_ray.argtypes = [C.c_int, C.c_int, f32p, f32p, C.c_float, C.c_int, C.c_int, u32p, f32p, f32p]
        L.ref_material_scatter.restype = C.c_int
        L.ref_material_scatter.argtypes = [C.c_void_p, C.POINTER(_oracle.orc_material), u32p, f32p, f32p, f32p, f32p, f32p]
        L.ref_trace_ray.restype = C.c_int
        L.ref_trace_ray.argtypes = [C.c_void_p, u32p, f32p, f32p, f32p, f32p, f32p]
        L.ref_render.restype = C.c_ulonglong
        L.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f32p, f32p, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


class Scene:
    def __init__(self, data, use_bvh=False):
        import importlib
        pkg = importlib.import_module("sycl-ray-tracer_b200")
        self.data = data
        self._insts = pkg.fill_instances(_oracle.orc_instance, _oracle.orc_material, data)
        tex = data.textures.ctypes.data_as(u8p) if data.textures is not None else None
        self.h = lib().ref_scene_create(self._insts, len(data.instances), tex, data.texture_layer_count,
                                        _oracle._fa(data.sky_color), int(use_bvh))

    def render(self, wavefront, w, h, depth, spp):
        """runs the reference's MegakernelRenderer / WavefrontRenderer::render_frame; returns the bytes
        it hands to stbi_write_png and the number of rtcIntersect1 calls"""
        d = self.data
        img = np.zeros((h, w, 4), np.uint8)
        devnull = os.open(os.devnull, os.O_WRONLY)   # the reference prints per-sample chatter
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            n = lib().ref_render(self.h, int(wavefront), w, h, _oracle._fa(d.camera_position), _oracle._fa(d.camera_direction),
                                 d.camera_focal_length, depth, spp, img.ctypes.data)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
        return img, int(n)

    def close(self):
        if self.h:
            lib().ref_scene_destroy(self.h)
            self.h = None
