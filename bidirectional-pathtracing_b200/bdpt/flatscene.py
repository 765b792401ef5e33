"""Flat POD scene (mirror of include/bdpt_flatscene.h) as numpy structured arrays.

The flat scene is what crosses the C-ABI: the primitive list in the reference's
BVHAccel order (reference src/pathtracer/raytraced_renderer.cpp:355-360), the
BSDF table (src/pathtracer/bsdf.h:132-300), the light table
(src/scene/light.h:63-139) and the placed camera (src/pathtracer/camera.h:104-128).
"""
from __future__ import annotations

import gzip
import io
from dataclasses import dataclass

import numpy as np

MAGIC = 0x3130534654504442  # "BDPTFS01"

MAT_DIFFUSE, MAT_EMISSION, MAT_MIRROR, MAT_GLASS, MAT_REFRACTION, MAT_MICROFACET = range(6)
LIGHT_AREA, LIGHT_POINT, LIGHT_ENV = 0, 1, 2
PRIM_TRIANGLE, PRIM_SPHERE = 0, 1

HEADER_DT = np.dtype([("magic", "<u8"), ("n_prims", "<u4"), ("n_materials", "<u4"),
                      ("n_lights", "<u4"), ("reserved", "<u4")])
CAMERA_DT = np.dtype([("pos", "<f8", 3), ("c2w", "<f8", 9), ("hfov_deg", "<f8"), ("vfov_deg", "<f8"),
                      ("nclip", "<f8"), ("fclip", "<f8"), ("screen_dist", "<f8"),
                      ("screen_w", "<f8"), ("screen_h", "<f8"), ("target", "<f8", 3)])
MATERIAL_DT = np.dtype([("type", "<i4"), ("pad", "<i4"), ("reflectance", "<f8", 3),
                        ("transmittance", "<f8", 3), ("radiance", "<f8", 3), ("eta", "<f8", 3),
                        ("k", "<f8", 3), ("alpha", "<f8"), ("ior", "<f8"), ("roughness", "<f8")])
LIGHT_DT = np.dtype([("type", "<i4"), ("pad", "<i4"), ("radiance", "<f8", 3), ("position", "<f8", 3),
                     ("direction", "<f8", 3), ("dim_x", "<f8", 3), ("dim_y", "<f8", 3)])
PRIM_DT = np.dtype([("type", "<i4"), ("material", "<i4"), ("data", "<f8", 18)])

assert HEADER_DT.itemsize == 24 and CAMERA_DT.itemsize == 176
assert MATERIAL_DT.itemsize == 152 and LIGHT_DT.itemsize == 128 and PRIM_DT.itemsize == 152


@dataclass
class FlatScene:
    camera: np.ndarray      # shape (1,) CAMERA_DT
    materials: np.ndarray   # (n_materials,) MATERIAL_DT
    lights: np.ndarray      # (n_lights,) LIGHT_DT
    prims: np.ndarray       # (n_prims,) PRIM_DT

    @property
    def n_prims(self) -> int:
        return int(self.prims.shape[0])

    def tobytes(self) -> bytes:
        hdr = np.zeros(1, HEADER_DT)
        hdr["magic"] = MAGIC
        hdr["n_prims"] = self.prims.shape[0]
        hdr["n_materials"] = self.materials.shape[0]
        hdr["n_lights"] = self.lights.shape[0]
        return b"".join(a.tobytes() for a in (hdr, self.camera, self.materials, self.lights, self.prims))

    def save(self, path: str) -> None:
        data = self.tobytes()
        if str(path).endswith(".gz"):
            # mtime=0 keeps regenerated fixtures byte-identical
            with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
                f.write(data)
        else:
            with open(path, "wb") as f:
                f.write(data)

    def with_frame(self, w: int, h: int) -> "FlatScene":
        """Camera::set_screen_size (reference camera.cpp:83-89): screen_dist is kept,
        both fields of view are recomputed from the new pixel size."""
        cam = self.camera.copy()
        sd = float(cam["screen_dist"][0])
        cam["screen_w"] = w
        cam["screen_h"] = h
        cam["hfov_deg"] = 2.0 * np.degrees(np.arctan(w / (2.0 * sd)))
        cam["vfov_deg"] = 2.0 * np.degrees(np.arctan(h / (2.0 * sd)))
        return FlatScene(cam, self.materials, self.lights, self.prims)


def frombytes(data: bytes) -> FlatScene:
    buf = io.BytesIO(data)

    def take(dt, n):
        raw = buf.read(dt.itemsize * n)
        if len(raw) != dt.itemsize * n:
            raise ValueError("truncated flat scene")
        return np.frombuffer(raw, dt, n).copy()

    hdr = take(HEADER_DT, 1)[0]
    if int(hdr["magic"]) != MAGIC:
        raise ValueError("not a BDPTFS01 flat scene")
    cam = take(CAMERA_DT, 1)
    mats = take(MATERIAL_DT, int(hdr["n_materials"]))
    lights = take(LIGHT_DT, int(hdr["n_lights"]))
    prims = take(PRIM_DT, int(hdr["n_prims"]))
    return FlatScene(cam, mats, lights, prims)


def load(path: str) -> FlatScene:
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        return frombytes(f.read())
