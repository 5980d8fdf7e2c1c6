"""ctypes view of oracle/_ref/libscenref.so: the reference's own glTF loader (src/scene.cpp + its vendored
tinygltf / stb), compiled in place through the API shims of oracle/refshim. TEST INFRASTRUCTURE; exists only
where /root/reference does."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libscenref.so")
_lib = None


def available():
    return os.path.isdir("/root/reference/src")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
        L = C.CDLL(LIB)
        L.scenref_load.restype = C.c_void_p
        L.scenref_load.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float]
        L.scenref_last_error.restype = C.c_char_p
        L.scenref_instance_count.argtypes = [C.c_void_p]
        fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L.scenref_instance.argtypes = [C.c_void_p, C.c_uint32, up, up, C.POINTER(fp), C.POINTER(fp), C.POINTER(fp), C.POINTER(up),
                                       fp, fp, C.POINTER(C.c_int32), fp]
        L.scenref_globals.argtypes = [C.c_void_p, fp]
        L.scenref_layer_count.argtypes = [C.c_void_p]
        L.scenref_layer.restype = C.POINTER(C.c_uint8)
        L.scenref_layer.argtypes = [C.c_void_p, C.c_uint32]
        _lib = L
    return _lib


def load(path, global_scale=(1.0, 1.0, 1.0)):
    """-> dict(instances=[...], sky, camera_position, camera_direction, focal, camera_node, layers)"""
    L = lib()
    h = L.scenref_load(path.encode(), *[float(v) for v in global_scale])
    if not h:
        raise RuntimeError(L.scenref_last_error().decode())
    out = {"instances": []}
    fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    for i in range(L.scenref_instance_count(h)):
        nv, ni = C.c_uint32(), C.c_uint32()
        pos, nrm, uv, idx = fp(), fp(), fp(), up()
        xf, nm, mi, mf = (C.c_float * 16)(), (C.c_float * 9)(), (C.c_int32 * 3)(), (C.c_float * 8)()
        L.scenref_instance(h, i, C.byref(nv), C.byref(ni), C.byref(pos), C.byref(nrm), C.byref(uv), C.byref(idx), xf, nm, mi, mf)
        out["instances"].append(dict(
            positions=np.ctypeslib.as_array(pos, (nv.value, 3)).copy(), normals=np.ctypeslib.as_array(nrm, (nv.value, 3)).copy(),
            uvs=np.ctypeslib.as_array(uv, (nv.value, 2)).copy(), indices=np.ctypeslib.as_array(idx, (ni.value,)).copy(),
            transform=np.array(xf, np.float32).copy(), normal_matrix=np.array(nm, np.float32).copy(),
            type=int(mi[0]), albedo_is_image=bool(mi[1]), albedo_image=int(mi[2]), albedo=np.array(mf[0:3], np.float32),
            roughness=float(mf[3]), ior=float(mf[4]), emissive=np.array(mf[5:8], np.float32)))
    g = (C.c_float * 11)()
    L.scenref_globals(h, g)
    out.update(sky=np.array(g[0:3], np.float32), camera_position=np.array(g[3:6], np.float32), camera_direction=np.array(g[6:9], np.float32),
               focal=float(g[9]), camera_node=int(g[10]))
    n = L.scenref_layer_count(h)
    out["layers"] = [np.ctypeslib.as_array(L.scenref_layer(h, i), (512, 512, 4)).copy() for i in range(n)]
    return out
