"""ctypes binding of tests/hostemu (TEST-ONLY host emulation of the kernels' per-item source)."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu")
LIB = os.path.join(HERE, "libhostemu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "hostemu.cpp")
        csrc = os.path.join(os.path.dirname(HERE), "..", "sycl-ray-tracer_b200", "csrc")
        newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc) if f.endswith(".h")])
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
            subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-std=c++17",
                            "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-o", LIB, src] + os.environ.get("EMU_CXXFLAGS", "").split(), check=True)
        L = C.CDLL(LIB)
        L.emu_scene_create.restype = C.c_void_p
        L.emu_scene_destroy.argtypes = [C.c_void_p]
        L.emu_scene_node_count.restype = C.c_uint32
        L.emu_scene_node_count.argtypes = [C.c_void_p]
        L.emu_scene_depth.restype = C.c_uint32
        L.emu_scene_depth.argtypes = [C.c_void_p]
        L.emu_scene_validate.argtypes = [C.c_void_p]
        L.emu_intersect.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_float, C.c_float] + [C.c_void_p] * 5
        L.emu_render.restype = C.c_uint64
        L.emu_render.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        _lib = L
    return _lib


class Scene:
    def __init__(self, data):
        pkg = importlib.import_module("sycl-ray-tracer_b200")
        cap = pkg._capi
        self._insts = pkg.fill_instances(cap.rt_instance, cap.rt_material, data)
        desc = cap.rt_scene_desc()
        desc.instances = self._insts
        desc.instance_count = len(data.instances)
        if data.textures is not None:
            desc.texture_layers = data.textures.ctypes.data_as(cap.u8p)
            desc.texture_layer_count = data.textures.shape[0]
        desc.sky_color = (C.c_float * 3)(*data.sky_color)
        self.data, self.cap = data, cap
        lib().emu_scene_create.argtypes = [C.POINTER(cap.rt_scene_desc)]
        self.h = lib().emu_scene_create(C.byref(desc))

    node_count = property(lambda s: int(lib().emu_scene_node_count(s.h)))
    depth = property(lambda s: int(lib().emu_scene_depth(s.h)))

    def validate(self):
        return int(lib().emu_scene_validate(self.h))

    def intersect(self, org, dir, tnear=1e-4, tfar=float("inf")):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dir = np.ascontiguousarray(dir, np.float32).reshape(-1, 3)
        n = org.shape[0]
        inst, prim = np.empty(n, np.int32), np.empty(n, np.int32)
        u, v, t = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
        lib().emu_intersect(self.h, n, org.ctypes.data, dir.ctypes.data, tnear, tfar, inst.ctypes.data,
                            prim.ctypes.data, u.ctypes.data, v.ctypes.data, t.ctypes.data)
        return dict(inst=inst, prim=prim, u=u, v=v, t=t)

    def render(self, camera, kind, max_depth, spp, shard=None, resume=None, roulette=False):
        """resume = the dict a previous render() returned: continue it (RT_RENDER_RESUME)"""
        w, h = camera.img_size
        p = self.cap.rt_render_params()
        p.max_depth, p.sample_count = max_depth, spp
        p.flags = (self.cap.RT_RENDER_RESUME if resume else 0) | (self.cap.RT_RENDER_ROULETTE if roulette else 0)
        if shard:
            p.shard.rank, p.shard.world = shard.get("rank", 0), shard.get("world", 1)
            p.shard.tile_size, p.shard.seed_salt = shard.get("tile_size", 0), shard.get("seed_salt", 0)
        accum, rgba8 = np.empty((h, w, 4), np.float32), np.empty((h, w, 4), np.uint8)
        rng = np.empty((h, w), np.uint32)
        if resume:
            accum[...] = resume["accum"]
            rng[...] = resume["rng_state"]
        rays = lib().emu_render(self.h, kind, C.addressof(camera.c), C.addressof(p), accum.ctypes.data,
                                rgba8.ctypes.data, rng.ctypes.data)
        return dict(accum=accum, rgba8=rgba8, rng_state=rng, ray_count=int(rays))

    def close(self):
        if self.h:
            lib().emu_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
