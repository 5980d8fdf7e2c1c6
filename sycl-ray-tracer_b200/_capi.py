"""ctypes binding of include/rt_api.h (librt_b200.so).

This is the only way the Python layer reaches the hot path: there is no Python or CPU
implementation behind it. Importing the package without the built library raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_LIB_PATH") or os.path.join(_HERE, "librt_b200.so")  # RT_LIB_PATH: development builds

RT_OK = 0
RT_MEGAKERNEL, RT_WAVEFRONT = 0, 1
RT_MAT_NONE, RT_MAT_DIFFUSE, RT_MAT_METALLIC, RT_MAT_DIELECTRIC = 0, 1, 2, 3
RT_TEX_SIZE, RT_MAX_IMAGES = 512, 128
RT_RENDER_RESUME = 1
RT_RENDER_ROULETTE = 2

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)


class rt_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("albedo_image", C.c_int32), ("albedo_color", C.c_float * 3),
                ("roughness", C.c_float), ("ior", C.c_float), ("emissive", C.c_float * 3)]


class rt_instance(C.Structure):
    _fields_ = [("positions", f32p), ("normals", f32p), ("uvs", f32p), ("indices", u32p),
                ("vertex_count", C.c_uint32), ("index_count", C.c_uint32),
                ("transform", C.c_float * 16), ("material", rt_material)]


class rt_scene_desc(C.Structure):
    _fields_ = [("instances", C.POINTER(rt_instance)), ("instance_count", C.c_uint32),
                ("texture_layers", u8p), ("texture_layer_count", C.c_uint32),
                ("sky_color", C.c_float * 3)]


class rt_camera(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("pixel00_loc", C.c_float * 3),
                ("pixel_delta_u", C.c_float * 3), ("pixel_delta_v", C.c_float * 3),
                ("img_size", C.c_int32 * 2)]


class rt_shard(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32), ("tile_size", C.c_uint32),
                ("seed_salt", C.c_uint32)]


class rt_ipc_handle(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 64)]


class rt_render_params(C.Structure):
    _fields_ = [("max_depth", C.c_uint32), ("sample_count", C.c_uint32), ("shard", rt_shard), ("flags", C.c_uint32),
                ("sample_chains", C.c_uint32)]


class rt_frame(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("accum", C.c_void_p), ("rng_state", C.c_void_p),
                ("ray_count", C.c_uint64), ("device_ms", C.c_float), ("kernel_launches", C.c_uint32)]


class rt_group_params(C.Structure):
    _fields_ = [("max_depth", C.c_uint32), ("sample_count", C.c_uint32), ("mode", C.c_uint32), ("tile_size", C.c_uint32),
                ("flags", C.c_uint32), ("sample_chains", C.c_uint32)]


RT_GROUP_TILES, RT_GROUP_SPP = 0, 1


class rt_scene_stats(C.Structure):
    _fields_ = [("triangle_count", C.c_uint64), ("node_count", C.c_uint64), ("bvh_bytes", C.c_uint64),
                ("shading_bytes", C.c_uint64), ("build_ms", C.c_float), ("max_leaf_tris", C.c_uint32),
                ("wide_depth", C.c_uint32)]


# every symbol include/rt_api.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "rt_api_version": (C.c_uint32, []),
    "rt_last_error": (C.c_char_p, [C.c_void_p]),
    "rt_context_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "rt_context_destroy": (None, [C.c_void_p]),
    "rt_context_stream": (C.c_void_p, [C.c_void_p]),
    "rt_context_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_context_device_name": (C.c_char_p, [C.c_void_p]),
    "rt_scene_create": (C.c_int, [C.c_void_p, C.POINTER(rt_scene_desc), C.POINTER(C.c_void_p)]),
    "rt_scene_commit": (C.c_int, [C.c_void_p]),
    "rt_scene_get_stats": (C.c_int, [C.c_void_p, C.POINTER(rt_scene_stats)]),
    "rt_scene_destroy": (None, [C.c_void_p]),
    "rt_camera_init": (None, [C.POINTER(rt_camera), C.c_int32, C.c_int32, f32p, f32p, C.c_float]),
    "rt_intersect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_float,
                               C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               f32p]),
    "rt_renderer_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "rt_renderer_destroy": (None, [C.c_void_p]),
    "rt_render_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(rt_camera),
                                  C.POINTER(rt_render_params), C.POINTER(rt_frame)]),
    "rt_renderer_device_accum": (C.c_void_p, [C.c_void_p]),
    "rt_renderer_device_rgba8": (C.c_void_p, [C.c_void_p]),
    "rt_renderer_export_image": (C.c_int, [C.c_void_p, C.POINTER(rt_ipc_handle)]),
    "rt_renderer_set_gather": (C.c_int, [C.c_void_p, C.POINTER(rt_ipc_handle), C.c_void_p]),
    "rt_resolve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p]),
    "rt_renderer_device_rng": (C.c_void_p, [C.c_void_p]),
    "rt_renderer_export_accum": (C.c_int, [C.c_void_p, C.POINTER(rt_ipc_handle)]),
    "rt_renderer_set_peers": (C.c_int, [C.c_void_p, C.POINTER(rt_ipc_handle), C.c_uint32, C.c_uint32]),
    "rt_renderer_reduce_resolve": (C.c_int, [C.c_void_p]),
    "rt_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_uint32, C.POINTER(C.c_void_p)]),
    "rt_group_destroy": (None, [C.c_void_p]),
    "rt_group_size": (C.c_uint32, [C.c_void_p]),
    "rt_group_context": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "rt_group_last_error": (C.c_char_p, [C.c_void_p]),
    "rt_group_scene_create": (C.c_int, [C.c_void_p, C.POINTER(rt_scene_desc), C.POINTER(C.c_void_p)]),
    "rt_group_scene_destroy": (None, [C.c_void_p]),
    "rt_group_scene_get": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "rt_group_renderer_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "rt_group_renderer_destroy": (None, [C.c_void_p]),
    "rt_group_renderer_get": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "rt_group_render_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(rt_camera), C.POINTER(rt_group_params), C.POINTER(rt_frame)]),
}

_lib = None


def load():
    """Load librt_b200.so and bind every declared symbol. Raises if the CUDA library is missing:
    there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The path has no CPU or pure-Python implementation.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
