"""ctypes binding of libbdpt_host.so (include/bdpt_host.h): the C++ .dae loader of the host side."""
from __future__ import annotations

import ctypes as C
import os

from . import flatscene as fs

LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "host", "libbdpt_host.so")
_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C bidirectional-pathtracing_b200/host`")
        L = C.CDLL(LIB_PATH)
        L.bdpt_host_load_dae.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_ubyte)), C.POINTER(C.c_size_t)]
        L.bdpt_host_load_dae.restype = C.c_int
        L.bdpt_host_apply_camera_settings.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.bdpt_host_apply_camera_settings.restype = C.c_int
        L.bdpt_host_load_envmap.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.bdpt_host_load_envmap.restype = C.c_int
        L.bdpt_host_free.argtypes = [C.c_void_p]
        L.bdpt_host_free.restype = None
        L.bdpt_host_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def load_dae(path: str, width: int = 0, height: int = 0) -> fs.FlatScene:
    """COLLADA file -> flat scene, by the C++ loader (reference semantics: camera placed at 800x600; `-r` optional)."""
    L = load_library()
    blob, size = C.POINTER(C.c_ubyte)(), C.c_size_t()
    rc = L.bdpt_host_load_dae(os.fsencode(path), width, height, C.byref(blob), C.byref(size))
    if rc != 0:
        raise RuntimeError(f"bdpt_host_load_dae({path}) failed ({rc}): {L.bdpt_host_last_error().decode()}")
    try:
        return fs.frombytes(C.string_at(blob, size.value))
    finally:
        L.bdpt_host_free(blob)


def apply_camera_settings(scene: fs.FlatScene, path: str, width: int, height: int) -> fs.FlatScene:
    """The reference's `-c <file>` (Camera::load_settings) on a copy of the scene's camera."""
    import numpy as np
    L = load_library()
    cam = np.ascontiguousarray(scene.camera.copy())
    rc = L.bdpt_host_apply_camera_settings(cam.ctypes.data, os.fsencode(path), width, height)
    if rc != 0:
        raise RuntimeError(f"bdpt_host_apply_camera_settings({path}) failed ({rc}): {L.bdpt_host_last_error().decode()}")
    return fs.FlatScene(cam, scene.materials, scene.lights, scene.prims)


def load_envmap(spec: str):
    """The reference's `-e <path>`: (h, w, 3) float32 lat-long map, row 0 = up (.exr, .pfm or "r,g,b")."""
    import numpy as np
    L = load_library()
    p, w, h = C.POINTER(C.c_float)(), C.c_int(), C.c_int()
    rc = L.bdpt_host_load_envmap(os.fsencode(spec), C.byref(p), C.byref(w), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"bdpt_host_load_envmap({spec}) failed ({rc}): {L.bdpt_host_last_error().decode()}")
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()
    finally:
        L.bdpt_host_free(p)
