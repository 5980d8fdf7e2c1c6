"""sycl-ray-tracer_b200 — B200-native (sm_100a) path-tracing hot path behind the reference's
App / Scene / Camera / IRenderer API.

Python mirror of the reference's host interface (the reference itself is C++; the C++ mirror is
host/raytracer.hpp). Every class is a thin handle over the C ABI in include/rt_api.h; all
computation happens in CUDA kernels inside librt_b200.so. There is no CPU path.

    App                      src/app.hpp:31-58            -> rt_context (device + stream)
    Scene                    src/scene.hpp:64-104         -> rt_scene  (commit = GPU BVH build)
    Camera                   src/camera.hpp:65-131        -> rt_camera
    IRenderer.render_frame   src/render.hpp:11-18         -> rt_render_frame
    MegakernelRenderer       src/render_megakernel.hpp    -> RT_MEGAKERNEL
    WavefrontRenderer        src/render_wavefront.hpp     -> RT_WAVEFRONT
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import (RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE, RT_MAT_METALLIC, RT_MAT_NONE, RT_MEGAKERNEL,
                    RT_WAVEFRONT)

__all__ = ["App", "Scene", "Camera", "IRenderer", "MegakernelRenderer", "WavefrontRenderer", "Frame",
           "SceneData", "InstanceData", "Material", "RtError", "intersect", "resolve", "Group", "GroupScene", "GroupRenderer"]


class RtError(RuntimeError):
    """The reference throws / terminates (src/main.cpp:71-74); the C ABI returns rt_status."""


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):  # torch tensor (host or device)
        return x.data_ptr()
    return int(x)


class Material:
    """Flattened Texture + Material (src/material.hpp:17-238)."""

    def __init__(self, type=RT_MAT_DIFFUSE, albedo=(0.8, 0.8, 0.8), albedo_image=-1, roughness=0.0,
                 ior=1.5, emissive=(0.0, 0.0, 0.0)):
        self.type, self.albedo, self.albedo_image = int(type), tuple(albedo), int(albedo_image)
        self.roughness, self.ior, self.emissive = float(roughness), float(ior), tuple(emissive)

    @staticmethod
    def diffuse(albedo=(0.8, 0.8, 0.8), image=-1, emissive=(0, 0, 0)):
        return Material(RT_MAT_DIFFUSE, albedo, image, 0.0, 1.5, emissive)

    @staticmethod
    def metallic(albedo=(0.8, 0.8, 0.8), roughness=0.0, image=-1, emissive=(0, 0, 0)):
        return Material(RT_MAT_METALLIC, albedo, image, roughness, 1.5, emissive)

    @staticmethod
    def dielectric(ior=1.5):
        return Material(RT_MAT_DIELECTRIC, (1, 1, 1), -1, 0.0, ior, (0, 0, 0))


class InstanceData:
    """One Embree instance = glTF node x primitive + GeometryData (src/scene.cpp:483-507)."""

    def __init__(self, positions, normals, uvs, indices, transform=None, material=None):
        self.positions = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        self.normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        self.uvs = np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
        self.indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        t = np.eye(4, dtype=np.float32) if transform is None else np.asarray(transform, dtype=np.float32)
        # stored column-major like glm::mat4 (src/scene.cpp:491-494); a (4,4) array is taken as the
        # mathematical matrix M (M @ [x,y,z,1]), a flat (16,) array as already column-major
        self.transform = np.ascontiguousarray(t.T.reshape(16) if t.shape == (4, 4) else t.reshape(16))
        self.material = material or Material()

    @property
    def triangle_count(self):
        return self.indices.size // 3


class SceneData:
    """Host-side description of what Scene hands to the kernels (src/scene.hpp:64-91)."""

    def __init__(self, instances, textures=None, sky_color=(0.5, 0.7, 1.0), camera_position=(0, 0, 0),
                 camera_direction=(0, 0, -1), camera_focal_length=1.0, name="scene"):
        self.instances = list(instances)
        self.textures = None if textures is None else np.ascontiguousarray(textures, dtype=np.uint8)
        self.sky_color = tuple(float(c) for c in sky_color)
        self.camera_position = tuple(float(c) for c in camera_position)
        self.camera_direction = tuple(float(c) for c in camera_direction)
        self.camera_focal_length = float(camera_focal_length)
        self.name = name

    @property
    def triangle_count(self):
        return sum(i.triangle_count for i in self.instances)

    @property
    def texture_layer_count(self):
        return 0 if self.textures is None else int(self.textures.shape[0])


def fill_instances(struct_type, material_type, data):
    """Marshal SceneData into a ctypes array of rt_instance-shaped structs (also used by the
    tests to fill the oracle's identically laid out struct)."""
    arr = (struct_type * max(1, len(data.instances)))()
    for k, inst in enumerate(data.instances):
        s = arr[k]
        s.positions = inst.positions.ctypes.data_as(_capi.f32p)
        s.normals = inst.normals.ctypes.data_as(_capi.f32p)
        s.uvs = inst.uvs.ctypes.data_as(_capi.f32p)
        s.indices = inst.indices.ctypes.data_as(_capi.u32p)
        s.vertex_count = inst.positions.shape[0]
        s.index_count = inst.indices.size
        s.transform = (C.c_float * 16)(*[float(v) for v in inst.transform])
        m = material_type()
        m.type, m.albedo_image = inst.material.type, inst.material.albedo_image
        m.albedo_color = (C.c_float * 3)(*inst.material.albedo)
        m.roughness, m.ior = inst.material.roughness, inst.material.ior
        m.emissive = (C.c_float * 3)(*inst.material.emissive)
        s.material = m
    return arr


class App:
    """raytracer::App (src/app.hpp:31-58): owns the device and the in-order stream."""

    def __init__(self, device=0):
        self._lib = _capi.load()
        h = C.c_void_p()
        st = self._lib.rt_context_create(int(device), C.byref(h))
        if st != _capi.RT_OK:
            raise RtError(f"rt_context_create failed ({st}): {self._lib.rt_last_error(None).decode()}")
        self.handle = h
        self.device = int(device)

    @property
    def device_name(self):
        return self._lib.rt_context_device_name(self.handle).decode()

    @property
    def stream(self):
        return self._lib.rt_context_stream(self.handle)

    def set_stream(self, cuda_stream):
        """borrow a caller-owned cudaStream_t (int), e.g. torch.cuda.current_stream().cuda_stream"""
        self.check(self._lib.rt_context_set_stream(self.handle, C.c_void_p(int(cuda_stream))), "rt_context_set_stream")

    def check(self, st, what):
        if st != _capi.RT_OK:
            raise RtError(f"{what} failed ({st}): {self._lib.rt_last_error(self.handle).decode()}")

    def close(self):
        if getattr(self, "handle", None):
            self._lib.rt_context_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """raytracer::Scene (src/scene.hpp:64-104). The reference builds from a .glb path; here the
    loader output (SceneData) is passed in and the constructor uploads + commits (GPU BVH build,
    the rtcCommitScene of src/scene.cpp:107)."""

    def __init__(self, app, data, commit=True):
        self.app, self.data, self._lib = app, data, app._lib
        self.sky_color = data.sky_color
        self.camera_position, self.camera_direction = data.camera_position, data.camera_direction
        self.camera_focal_length = data.camera_focal_length
        insts = fill_instances(_capi.rt_instance, _capi.rt_material, data)
        desc = _capi.rt_scene_desc()
        desc.instances = insts
        desc.instance_count = len(data.instances)
        if data.textures is not None:
            assert data.textures.shape[1:] == (_capi.RT_TEX_SIZE, _capi.RT_TEX_SIZE, 4)
            desc.texture_layers = data.textures.ctypes.data_as(_capi.u8p)
            desc.texture_layer_count = data.textures.shape[0]
        desc.sky_color = (C.c_float * 3)(*data.sky_color)
        h = C.c_void_p()
        app.check(self._lib.rt_scene_create(app.handle, C.byref(desc), C.byref(h)), "rt_scene_create")
        self.handle = h
        if commit:
            self.commit()

    def commit(self):
        self.app.check(self._lib.rt_scene_commit(self.handle), "rt_scene_commit")

    @property
    def stats(self):
        s = _capi.rt_scene_stats()
        self.app.check(self._lib.rt_scene_get_stats(self.handle, C.byref(s)), "rt_scene_get_stats")
        return {k: getattr(s, k) for k, _ in s._fields_}

    def close(self):
        if getattr(self, "handle", None) and getattr(self.app, "handle", None):
            self._lib.rt_scene_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Camera:
    """raytracer::Camera (src/camera.hpp:65-106): Camera(img_size, position, direction, focal)."""

    def __init__(self, img_size, cam_center, cam_dir, focal_length):
        lib = _capi.load()
        self.c = _capi.rt_camera()
        pos = (C.c_float * 3)(*[float(v) for v in cam_center])
        d = (C.c_float * 3)(*[float(v) for v in cam_dir])
        lib.rt_camera_init(C.byref(self.c), int(img_size[0]), int(img_size[1]), pos, d, float(focal_length))
        self.img_size = (int(img_size[0]), int(img_size[1]))


class Frame:
    """What one render_frame produced (the reference writes out.png and prints three lines)."""

    def __init__(self, rgba8, accum, rng_state, ray_count, device_ms, kernel_launches, sample_count):
        self.rgba8, self.accum, self.rng_state = rgba8, accum, rng_state
        self.ray_count, self.device_ms, self.kernel_launches = ray_count, device_ms, kernel_launches
        self.sample_count = sample_count

    def metric_lines(self):
        """The three stdout lines benchmark.py parses (src/render_megakernel.cpp:181-183)."""
        secs = self.device_ms * 1e-3
        return [f"Time measured: {secs:.6f} seconds", f"Total rays: {self.ray_count}",
                f"Rays/sec: {self.ray_count / max(secs, 1e-12) / 1e6:.2f}M"]


class IRenderer:
    """raytracer::IRenderer (src/render.hpp:11-18)."""
    kind = None

    def __init__(self, app, img_size, image=None, max_depth=10, sample_count=32):
        # ctor shape of src/render_megakernel.hpp:13-19 / src/render_wavefront.hpp:55-61;
        # `image` is the RGBA8 output (numpy (H,W,4) uint8, or None to allocate one)
        self.app, self._lib = app, app._lib
        self.img_size = (int(img_size[0]), int(img_size[1]))
        self.max_depth, self.sample_count = int(max_depth), int(sample_count)
        self.image = image
        h = C.c_void_p()
        app.check(self._lib.rt_renderer_create(app.handle, self.kind, self.img_size[0], self.img_size[1],
                                               C.byref(h)), "rt_renderer_create")
        self.handle = h

    def render_frame(self, camera, scene, want=("rgba8", "accum", "rng_state"), shard=None, outputs=None, resume=False,
                     roulette=False, chains=0):
        """IRenderer::render_frame. `want` selects which host copies to make; `outputs` may map
        names to caller buffers (numpy arrays or torch tensors, host or device)."""
        w, h = self.img_size
        outputs = dict(outputs or {})
        if "rgba8" in want and "rgba8" not in outputs:
            outputs["rgba8"] = self.image if self.image is not None else np.empty((h, w, 4), np.uint8)
        if "accum" in want and "accum" not in outputs:
            outputs["accum"] = np.empty((h, w, 4), np.float32)
        if "rng_state" in want and "rng_state" not in outputs:
            outputs["rng_state"] = np.empty((h, w), np.uint32)
        p = _capi.rt_render_params()
        p.max_depth, p.sample_count = self.max_depth, self.sample_count
        p.flags = (_capi.RT_RENDER_RESUME if resume else 0) | (_capi.RT_RENDER_ROULETTE if roulette else 0)
        p.sample_chains = int(chains)  # > 1: independent sample chains per pixel (not the reference's order)
        if shard:
            p.shard.rank, p.shard.world = int(shard.get("rank", 0)), int(shard.get("world", 1))
            p.shard.tile_size, p.shard.seed_salt = int(shard.get("tile_size", 0)), int(shard.get("seed_salt", 0))
        f = _capi.rt_frame()
        f.rgba8, f.accum, f.rng_state = _ptr(outputs.get("rgba8")), _ptr(outputs.get("accum")), _ptr(outputs.get("rng_state"))
        self.app.check(self._lib.rt_render_frame(self.handle, scene.handle, C.byref(camera.c), C.byref(p),
                                                 C.byref(f)), "rt_render_frame")
        return Frame(outputs.get("rgba8"), outputs.get("accum"), outputs.get("rng_state"), int(f.ray_count),
                     float(f.device_ms), int(f.kernel_launches), self.sample_count)

    @property
    def device_accum_ptr(self):
        return self._lib.rt_renderer_device_accum(self.handle)

    @property
    def device_rgba8_ptr(self):
        return self._lib.rt_renderer_device_rgba8(self.handle)

    def export_image(self):
        """Make this renderer's device RGBA8 image the gather destination of tile-sharded peers:
        returns the 64 opaque bytes to send them (rt_renderer_export_image)."""
        h = _capi.rt_ipc_handle()
        self.app.check(self._lib.rt_renderer_export_image(self.handle, C.byref(h)), "rt_renderer_export_image")
        return bytes(h.bytes)

    def export_accum(self):
        """64 opaque bytes naming this renderer's device accumulation buffer for the other ranks (CUDA IPC)"""
        h = _capi.rt_ipc_handle()
        self.app.check(self._lib.rt_renderer_export_accum(self.handle, C.byref(h)), "rt_renderer_export_accum")
        return bytes(h.bytes)

    def set_peers(self, handles=None, rank=0):
        """spp slices across processes: attach every rank's accumulation buffer (list of export_accum() bytes, own entry
        ignored); None detaches"""
        if not handles:
            self.app.check(self._lib.rt_renderer_set_peers(self.handle, None, 0, 0), "rt_renderer_set_peers")
            return
        arr = (_capi.rt_ipc_handle * len(handles))()
        for k, hb in enumerate(handles):
            C.memmove(arr[k].bytes, bytes(hb), 64)
        self.app.check(self._lib.rt_renderer_set_peers(self.handle, arr, len(handles), int(rank)), "rt_renderer_set_peers")

    def reduce_resolve(self):
        """enqueue the fused reduce-scatter + resolve + gather kernel for this rank's slice (asynchronous on the stream)"""
        self.app.check(self._lib.rt_renderer_reduce_resolve(self.handle), "rt_renderer_reduce_resolve")

    def set_gather(self, handle=None, device_ptr=None):
        """Tile shards: also store owned pixels' RGBA8 into a peer's image — `handle` = bytes from the
        destination's export_image() (another process) or `device_ptr` (same process); neither = detach."""
        hp = None
        if handle is not None:
            hs = _capi.rt_ipc_handle()
            C.memmove(hs.bytes, bytes(handle), 64)
            hp = C.byref(hs)
        self.app.check(self._lib.rt_renderer_set_gather(self.handle, hp, C.c_void_p(int(device_ptr)) if device_ptr else None),
                       "rt_renderer_set_gather")

    def close(self):
        if getattr(self, "handle", None) and getattr(self.app, "handle", None):
            self._lib.rt_renderer_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MegakernelRenderer(IRenderer):
    """src/render_megakernel.hpp:6-22"""
    kind = RT_MEGAKERNEL


class WavefrontRenderer(IRenderer):
    """src/render_wavefront.hpp:40-63"""
    kind = RT_WAVEFRONT


def intersect(app, scene, org, dir, tnear=1e-4, tfar=float("inf")):
    """Batch rtcIntersect1 (src/trace_ray.hpp:18-22). org, dir: (n,3) float32 numpy arrays.
    Returns dict(inst, prim, u, v, t, device_ms)."""
    org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
    dir = np.ascontiguousarray(dir, np.float32).reshape(-1, 3)
    n = org.shape[0]
    inst, prim = np.empty(n, np.int32), np.empty(n, np.int32)
    u, v, t = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    ms = C.c_float(0)
    app.check(app._lib.rt_intersect(app.handle, scene.handle, n, _ptr(org), _ptr(dir), float(tnear), float(tfar),
                                    _ptr(inst), _ptr(prim), _ptr(u), _ptr(v), _ptr(t), C.byref(ms)), "rt_intersect")
    return dict(inst=inst, prim=prim, u=u, v=v, t=t, device_ms=ms.value)


def resolve(app, accum, sample_count, width, height, rgba8=None):
    """mean / sqrt gamma / F10 bytes from an accumulation buffer (after a cross-GPU reduction)."""
    if rgba8 is None:
        rgba8 = np.empty((height, width, 4), np.uint8)
    app.check(app._lib.rt_resolve(app.handle, _ptr(accum), int(sample_count), int(width), int(height), _ptr(rgba8)),
              "rt_resolve")
    return rgba8


def device_count():
    """CUDA devices visible to the library (0 without a GPU)."""
    try:
        import ctypes.util
        rt = C.CDLL(ctypes.util.find_library("cudart") or "libcudart.so")
    except OSError:
        try:
            import torch
            return int(torch.cuda.device_count())
        except Exception:
            return 0
    n = C.c_int(0)
    return int(n.value) if rt.cudaGetDeviceCount(C.byref(n)) == 0 else 0


class Group:
    """Several GPUs of one node behind one handle (rt_group_*, include/rt_api.h): what lets the reference's
    main() sequence (src/main.cpp:57-70) run on N B200s from one process."""

    def __init__(self, devices):
        self._lib = _capi.load()
        devices = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        st = self._lib.rt_group_create(arr, len(devices), C.byref(h))
        if st != _capi.RT_OK:
            raise RtError(f"rt_group_create failed ({st}): {self._lib.rt_last_error(None).decode()}")
        self.handle, self.devices = h, devices

    def __len__(self):
        return int(self._lib.rt_group_size(self.handle))

    def check(self, st, what):
        if st != _capi.RT_OK:
            raise RtError(f"{what} failed ({st}): {self._lib.rt_group_last_error(self.handle).decode()}")

    def close(self):
        if getattr(self, "handle", None):
            self._lib.rt_group_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupScene:
    """Scene replicated on every device of a Group (upload + GPU BVH build run concurrently)."""

    def __init__(self, group, data):
        self.group, self.data, self._lib = group, data, group._lib
        self._insts = fill_instances(_capi.rt_instance, _capi.rt_material, data)
        desc = _capi.rt_scene_desc()
        desc.instances = self._insts
        desc.instance_count = len(data.instances)
        if data.textures is not None:
            desc.texture_layers = data.textures.ctypes.data_as(_capi.u8p)
            desc.texture_layer_count = data.textures.shape[0]
        desc.sky_color = (C.c_float * 3)(*data.sky_color)
        h = C.c_void_p()
        group.check(self._lib.rt_group_scene_create(group.handle, C.byref(desc), C.byref(h)), "rt_group_scene_create")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None) and getattr(self.group, "handle", None):
            self._lib.rt_group_scene_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupRenderer:
    """IRenderer over a Group: mode "tiles" (bit-identical to one device) or "spp" (salted sample slices)."""

    def __init__(self, group, kind, img_size, max_depth=10, sample_count=32, mode="tiles", tile_size=64):
        self.group, self._lib = group, group._lib
        self.img_size = (int(img_size[0]), int(img_size[1]))
        self.max_depth, self.sample_count, self.mode, self.tile_size = int(max_depth), int(sample_count), mode, int(tile_size)
        h = C.c_void_p()
        group.check(self._lib.rt_group_renderer_create(group.handle, int(kind), self.img_size[0], self.img_size[1], C.byref(h)),
                    "rt_group_renderer_create")
        self.handle = h

    def render_frame(self, camera, scene, want=("rgba8", "accum", "rng_state"), resume=False, roulette=False, chains=0):
        w, h = self.img_size
        out = {"rgba8": np.empty((h, w, 4), np.uint8) if "rgba8" in want else None,
               "accum": np.empty((h, w, 4), np.float32) if "accum" in want else None,
               "rng_state": np.empty((h, w), np.uint32) if "rng_state" in want else None}
        p = _capi.rt_group_params()
        p.max_depth, p.sample_count = self.max_depth, self.sample_count
        p.mode = _capi.RT_GROUP_SPP if self.mode == "spp" else _capi.RT_GROUP_TILES
        p.tile_size = self.tile_size
        p.flags = (_capi.RT_RENDER_RESUME if resume else 0) | (_capi.RT_RENDER_ROULETTE if roulette else 0)
        p.sample_chains = int(chains)
        f = _capi.rt_frame()
        f.rgba8, f.accum, f.rng_state = _ptr(out["rgba8"]), _ptr(out["accum"]), _ptr(out["rng_state"])
        self.group.check(self._lib.rt_group_render_frame(self.handle, scene.handle, C.byref(camera.c), C.byref(p), C.byref(f)),
                         "rt_group_render_frame")
        return Frame(out["rgba8"], out["accum"], out["rng_state"], int(f.ray_count), float(f.device_ms), int(f.kernel_launches),
                     self.sample_count)

    def close(self):
        if getattr(self, "handle", None) and getattr(self.group, "handle", None):
            self._lib.rt_group_renderer_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

