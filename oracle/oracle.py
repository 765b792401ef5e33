"""TEST INFRASTRUCTURE (oracle) — ctypes wrapper over oracle/libbdpt_oracle.so (the fp64
CPU restatement, oracle/bdpt_oracle.cpp) and a runner for oracle/_ref/ref_harness (the
reference itself, built from /root/reference by oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (bidirectional-pathtracing_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbdpt_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_HARNESS = os.path.join(REF_DIR, "ref_harness")
REF_CLI = os.path.join(REF_DIR, "pathtracer_ref")
REF_SCENES = os.path.join(REF_DIR, "scenes")

_lib = None


def build() -> None:
    subprocess.run(["make", "-C", HERE, "port"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_frame.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_set_cell.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_set_envmap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_set_lens.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.oracle_lens_ray.argtypes = [C.c_void_p, dp, C.c_long, dp]
        for name in ("oracle_intersect", "oracle_intersect_brute", "oracle_bsdf_eval", "oracle_bsdf_sample",
                     "oracle_light_eval", "oracle_camera_eval"):
            getattr(L, name).argtypes = [C.c_void_p, dp, C.c_long, dp]
            getattr(L, name).restype = None
        L.oracle_bsdf_sample_ref.argtypes = [C.c_void_p, dp, C.c_long, dp, dp]
        L.oracle_bsdf_sample_ref.restype = None
        L.oracle_render.restype = C.c_ulonglong
        L.oracle_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, dp, dp, dp]
        L.oracle_trace_sample.restype = C.c_long
        L.oracle_trace_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_int, C.c_int,
                                          C.c_int, dp, C.c_long]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """fp64 CPU restatement bound to one flat scene."""

    def __init__(self, scene, width: int | None = None, height: int | None = None):
        L = lib()
        self._keep = (np.ascontiguousarray(scene.camera), np.ascontiguousarray(scene.materials),
                      np.ascontiguousarray(scene.lights), np.ascontiguousarray(scene.prims))
        cam, mats, lights, prims = self._keep
        self.h = L.oracle_create(cam.ctypes.data, mats.ctypes.data, len(mats), lights.ctypes.data, len(lights),
                                 prims.ctypes.data, len(prims))
        self.W = int(cam["screen_w"][0])
        self.H = int(cam["screen_h"][0])
        if width and height:
            self.set_frame(width, height)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    def set_frame(self, w: int, h: int) -> None:
        lib().oracle_set_frame(self.h, w, h)
        self.W, self.H = w, h

    def set_lens(self, radius: float, focal_distance: float) -> None:
        """Thin lens (-b radius, -d focal distance); radius 0 = the pinhole the reference's BDPT always uses."""
        lib().oracle_set_lens(self.h, radius, focal_distance)

    def lens_ray(self, q):
        return self._batch("oracle_lens_ray", q, 4, 6)

    def set_envmap(self, rgb) -> None:
        """Lat-long radiance map (h, w, 3) float32 of the environment light, row 0 = up; None removes it."""
        if rgb is None:
            lib().oracle_set_envmap(self.h, None, 0, 0)
            return
        self._env = np.ascontiguousarray(rgb, np.float32)
        lib().oracle_set_envmap(self.h, self._env.ctypes.data, self._env.shape[1], self._env.shape[0])

    def envmap_lookup(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        out = np.empty_like(dirs)
        lib().oracle_envmap_lookup.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_long, C.POINTER(C.c_double)]
        lib().oracle_envmap_lookup(self.h, _dp(dirs), dirs.shape[0], _dp(out))
        return out

    def set_cell(self, x: int = 0, y: int = 0, dx: int = 0, dy: int = 0) -> None:
        """Cell render (reference `-p x y dx dy`): only these pixels are sampled; dx = 0 restores the whole frame."""
        lib().oracle_set_cell(self.h, x, y, dx, dy)

    def _batch(self, fn, arr, n_in, n_out):
        arr = np.ascontiguousarray(arr, np.float64).reshape(-1, n_in)
        out = np.empty((arr.shape[0], n_out), np.float64)
        getattr(lib(), fn)(self.h, _dp(arr), arr.shape[0], _dp(out))
        return out

    def intersect(self, rays):
        return self._batch("oracle_intersect", rays, 8, 6)

    def intersect_brute(self, rays):
        return self._batch("oracle_intersect_brute", rays, 8, 5)

    def bsdf_eval(self, q):
        return self._batch("oracle_bsdf_eval", q, 7, 8)

    def bsdf_sample(self, q):
        return self._batch("oracle_bsdf_sample", q, 7, 7)

    def bsdf_sample_ref(self, q):
        """sample_f in the reference's own draw order (fresh default-seeded streams).  q: n x {material, wo[3]}.
        Returns (n x {wi[3], pdf, f[3]}, n x {u1, u2, coin}): the outputs and the uniforms each call drew."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 4)
        out = np.empty((q.shape[0], 7), np.float64)
        uni = np.empty((q.shape[0], 3), np.float64)
        lib().oracle_bsdf_sample_ref(self.h, _dp(q), q.shape[0], _dp(out), _dp(uni))
        return out, uni

    def light_eval(self, q):
        return self._batch("oracle_light_eval", q, 7, 6)

    def camera_eval(self, q):
        return self._batch("oracle_camera_eval", q, 5, 11)

    def render(self, ns: int, m: int, *, roulette: int = 0, no_mis: int = 0, balance: int = 0, rng_mode: int = 1, seed: int = 0, first_sample: int = 0,
               n_samples: int | None = None, threads: int = 1, strategy: tuple[int, int] | None = None):
        """Returns (sum, eye, light) as (H, W, 3) fp64 arrays (row 0 = bottom) and the ray count."""
        n_samples = ns if n_samples is None else n_samples
        bufs = [np.zeros((self.H, self.W, 3), np.float64) for _ in range(3)]
        t, s = strategy if strategy else (-1, -1)
        rays = lib().oracle_render(self.h, ns, m, (1 if roulette else 0) | (2 if no_mis else 0) | (4 if balance else 0), rng_mode, seed, first_sample, n_samples, threads, t, s,
                                   _dp(bufs[0]), _dp(bufs[1]), _dp(bufs[2]))
        return bufs[0], bufs[1], bufs[2], int(rays)

    def trace_sample(self, ns: int, m: int, x: int, y: int, sample: int, *, seed: int = 0, roulette: int = 0):
        cap = 1 << 16
        out = np.zeros(cap, np.float64)
        n = lib().oracle_trace_sample(self.h, ns, m, roulette, seed, x, y, sample, _dp(out), cap)
        out = out[:n]
        ne, nl = int(out[0]), int(out[1])
        verts = out[2:2 + 13 * (ne + nl)].reshape(-1, 13)
        strat = out[2 + 13 * (ne + nl):].reshape(-1, 8)
        return {"eye": verts[:ne], "light": verts[ne:], "strategies": strat}


# ----------------------------------------------------------------------------- the reference itself
def have_ref() -> bool:
    return os.path.exists(REF_HARNESS)


def ref_scene(name: str) -> str:
    return os.path.join(REF_SCENES, name if name.endswith(".dae") else name + ".dae")


def _run(args, **kw):
    return subprocess.run(args, check=True, capture_output=True, text=True, **kw)


def ref_batch(cmd: str, scene_dae: str, arr: np.ndarray, n_in: int, n_out: int, extra=()):
    arr = np.ascontiguousarray(arr, np.float64).reshape(-1, n_in)
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        arr.tofile(fi)
        _run([REF_HARNESS, cmd, scene_dae, *map(str, extra), fi, fo])
        return np.fromfile(fo, np.float64).reshape(-1, n_out)


def ref_dump(scene_dae: str, out_path: str) -> None:
    _run([REF_HARNESS, "dump", scene_dae, out_path])


def ref_render(scene_dae: str, w: int, h: int, ns: int, m: int, threads: int = 1, strategy=None):
    """Runs the reference's own renderer; returns (sum, eye, light, info)."""
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "r")
        args = [REF_HARNESS, "render", scene_dae, w, h, ns, m, threads, prefix]
        if strategy:
            args += list(strategy)
        r = _run(list(map(str, args)))
        info = {}
        for line in r.stdout.splitlines():
            if line.startswith("HARNESS render"):
                for kv in line.split()[2:]:
                    k, v = kv.split("=")
                    info[k] = float(v)
        bufs = [np.fromfile(prefix + ext, np.float64).reshape(h, w, 3) for ext in (".sum.f64", ".eye.f64", ".light.f64")]
        return bufs[0], bufs[1], bufs[2], info
