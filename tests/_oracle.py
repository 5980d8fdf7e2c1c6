"""ctypes binding of oracle/rt_oracle.h — the CHECKER. Only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

f32p, u32p, u8p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)


class orc_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("albedo_image", C.c_int32), ("albedo_color", C.c_float * 3),
                ("roughness", C.c_float), ("ior", C.c_float), ("emissive", C.c_float * 3)]


class orc_instance(C.Structure):
    _fields_ = [("positions", f32p), ("normals", f32p), ("uvs", f32p), ("indices", u32p),
                ("vertex_count", C.c_uint32), ("index_count", C.c_uint32), ("transform", C.c_float * 16),
                ("material", orc_material)]


class orc_camera(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("pixel00_loc", C.c_float * 3), ("pixel_delta_u", C.c_float * 3),
                ("pixel_delta_v", C.c_float * 3), ("img_size", C.c_int32 * 2)]


class orc_render_params(C.Structure):
    _fields_ = [("mode", C.c_int32), ("max_depth", C.c_uint32), ("sample_count", C.c_uint32),
                ("seed_salt", C.c_uint32), ("use_bvh", C.c_int32), ("threads", C.c_int32),
                ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32), ("roulette", C.c_int32)]


MODE_MEGAKERNEL, MODE_WAVEFRONT = 0, 1
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "rt_oracle.cpp")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            build()
        L = C.CDLL(LIB)
        L.orc_xorshift_next.restype = C.c_float
        L.orc_xorshift_next.argtypes = [u32p]
        L.orc_random_unit_vector.argtypes = [u32p, f32p]
        L.orc_round_half.restype = C.c_float
        L.orc_round_half.argtypes = [C.c_float]
        L.orc_output_byte.restype = C.c_uint8
        L.orc_output_byte.argtypes = [C.c_float]
        L.orc_pixel_seed.restype = C.c_uint32
        L.orc_pixel_seed.argtypes = [C.c_int32] * 5
        L.orc_camera_init.argtypes = [C.POINTER(orc_camera), C.c_int32, C.c_int32, f32p, f32p, C.c_float]
        L.orc_camera_get_ray.argtypes = [C.POINTER(orc_camera), C.c_int32, C.c_int32, u32p, f32p, f32p]
        L.orc_material_scatter.restype = C.c_int
        L.orc_material_scatter.argtypes = [C.POINTER(orc_material), u8p, C.c_uint32, u32p, f32p, f32p, f32p, f32p, f32p]
        L.orc_texture_sample.argtypes = [u8p, C.c_uint32, C.c_int32, f32p, f32p]
        L.orc_normal_matrix.argtypes = [f32p, f32p]
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.POINTER(orc_instance), C.c_uint32, u8p, C.c_uint32, f32p]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_triangle_count.restype = C.c_uint64
        L.orc_scene_triangle_count.argtypes = [C.c_void_p]
        L.orc_scene_world_triangles.restype = f32p
        L.orc_scene_world_triangles.argtypes = [C.c_void_p]
        L.orc_intersect.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_float,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_render.restype = C.c_uint64
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(orc_camera), C.POINTER(orc_render_params), C.c_void_p,
                                 C.c_void_p, C.c_void_p]
        L.orc_last_render_seconds.restype = C.c_double
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _fa(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


class Scene:
    def __init__(self, data):
        import importlib
        pkg = importlib.import_module("sycl-ray-tracer_b200")
        self.data = data
        self._insts = pkg.fill_instances(orc_instance, orc_material, data)  # same field layout, own struct
        tex = data.textures.ctypes.data_as(u8p) if data.textures is not None else None
        self.h = lib().orc_scene_create(self._insts, len(data.instances), tex, data.texture_layer_count,
                                        _fa(data.sky_color))

    @property
    def triangle_count(self):
        return int(lib().orc_scene_triangle_count(self.h))

    def world_triangles(self):
        n = self.triangle_count
        p = lib().orc_scene_world_triangles(self.h)
        return np.ctypeslib.as_array(p, shape=(n, 3, 3)).copy()

    def intersect(self, org, dir, tnear=1e-4, tfar=float("inf"), use_bvh=False, threads=0):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dir = np.ascontiguousarray(dir, np.float32).reshape(-1, 3)
        n = org.shape[0]
        inst, prim = np.empty(n, np.int32), np.empty(n, np.int32)
        u, v, t = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
        lib().orc_intersect(self.h, int(use_bvh), threads, n, org.ctypes.data, dir.ctypes.data, tnear, tfar,
                            inst.ctypes.data, prim.ctypes.data, u.ctypes.data, v.ctypes.data, t.ctypes.data)
        return dict(inst=inst, prim=prim, u=u, v=v, t=t)

    def render(self, cam, mode, max_depth, spp, use_bvh=False, seed_salt=0, crop=None, threads=0, roulette=False):
        W, H = cam.img_size[0], cam.img_size[1]
        x0, y0, x1, y1 = crop if crop else (0, 0, W, H)
        p = orc_render_params(mode, max_depth, spp, seed_salt, int(use_bvh), threads, x0, y0, x1, y1, int(roulette))
        ch, cw = y1 - y0, x1 - x0
        accum, rgba8 = np.empty((ch, cw, 4), np.float32), np.empty((ch, cw, 4), np.uint8)
        rng = np.empty((ch, cw), np.uint32)
        rays = lib().orc_render(self.h, C.byref(cam), C.byref(p), accum.ctypes.data, rgba8.ctypes.data, rng.ctypes.data)
        return dict(accum=accum, rgba8=rgba8, rng_state=rng, ray_count=int(rays),
                    seconds=float(lib().orc_last_render_seconds()))

    def close(self):
        if self.h:
            lib().orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def camera(width, height, pos, direction, focal):
    c = orc_camera()
    lib().orc_camera_init(C.byref(c), width, height, _fa(pos), _fa(direction), float(focal))
    return c


def camera_for(data, width, height):
    return camera(width, height, data.camera_position, data.camera_direction, data.camera_focal_length)


def primary_rays(cam, mode, width, height, salt=0):
    """sample-0 camera rays for every pixel (row-major), with the mode's seed mapping"""
    L = lib()
    org, d = np.empty((height, width, 3), np.float32), np.empty((height, width, 3), np.float32)
    o3, d3 = (C.c_float * 3)(), (C.c_float * 3)()
    for y in range(height):
        for x in range(width):
            st = C.c_uint32(L.orc_pixel_seed(mode, x, y, width, height) ^ salt)
            L.orc_camera_get_ray(C.byref(cam), x, y, C.byref(st), o3, d3)
            org[y, x] = o3[:]
            d[y, x] = d3[:]
    return org.reshape(-1, 3), d.reshape(-1, 3)
