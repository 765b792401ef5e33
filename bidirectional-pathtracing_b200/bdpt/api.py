"""ctypes binding of libbdpt_b200.so (include/bdpt_b200.h) and a host-side mirror of the reference's
renderer interface for the BDPT path.

`Renderer` keeps the names of the reference's OfflineRenderer / RaytracedRenderer surface
(reference src/application/renderer.h:14-99, src/pathtracer/raytraced_renderer.h): set_scene,
set_camera, set_frame_size, start_raytracing (here: render), clear, plus the fields the caller
sets on the integrator (ns_aa, max_ray_depth: src/pathtracer/pathtracer.h:73-82).

There is no fallback of any kind: if the CUDA library is missing or no device is present this
module raises.  It never imports anything from oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import flatscene as fs

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "libbdpt_b200.so")

BUFFER_SUM, BUFFER_EYE, BUFFER_LIGHT = 0, 1, 2
FLAG_ROULETTE = 1
FLAG_NO_MIS = 2
FLAG_BALANCE = 4
COMM_ID_BYTES = 128


class BdptError(RuntimeError):
    pass


class RenderParams(C.Structure):
    _fields_ = [("ns_total", C.c_int32), ("max_ray_depth", C.c_int32), ("first_sample", C.c_int32),
                ("n_samples", C.c_int32), ("flags", C.c_uint32), ("only_t", C.c_int32), ("only_s", C.c_int32),
                ("reserved", C.c_int32), ("seed", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("connections", C.c_uint64), ("kernel_launches", C.c_uint64), ("ms_total", C.c_double),
                ("ms_walk", C.c_double), ("ms_connect", C.c_double), ("ms_shadow", C.c_double),
                ("ms_bvh_build", C.c_double), ("bvh_nodes", C.c_uint32), ("batch_samples", C.c_uint32),
                ("n_batches", C.c_uint32), ("timed_batches", C.c_uint32)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["bdpt_create", "bdpt_destroy", "bdpt_last_error", "bdpt_set_stream", "bdpt_upload_scene",
           "bdpt_set_camera", "bdpt_set_environment_map", "bdpt_set_frame", "bdpt_set_lens", "bdpt_set_cell", "bdpt_clear", "bdpt_render", "bdpt_read_framebuffer",
           "bdpt_framebuffer_device", "bdpt_tonemap_rgba8", "bdpt_get_stats", "bdpt_intersect_batch",
           "bdpt_occluded_batch", "bdpt_bsdf_eval_batch", "bdpt_bsdf_sample_batch", "bdpt_light_eval_batch",
           "bdpt_camera_eval_batch", "bdpt_trace_sample", "bdpt_host_build_bvh", "bdpt_device_bvh",
           "bdpt_comm_unique_id", "bdpt_comm_init_rank", "bdpt_comm_init_all", "bdpt_reduce", "bdpt_reduce_all",
           "bdpt_comm_destroy"]

_lib = None


def load_library():
    """Loads libbdpt_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BdptError(f"{LIB_PATH} is missing: build it with `make -C bidirectional-pathtracing_b200` "
                        "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    L.bdpt_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.bdpt_destroy.argtypes = [vp]; L.bdpt_destroy.restype = None
    L.bdpt_last_error.argtypes = [vp]; L.bdpt_last_error.restype = C.c_char_p
    L.bdpt_set_stream.argtypes = [vp, vp]
    L.bdpt_upload_scene.argtypes = [vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32]
    L.bdpt_set_camera.argtypes = [vp, vp]
    L.bdpt_set_environment_map.argtypes = [vp, vp, C.c_int, C.c_int]
    L.bdpt_set_frame.argtypes = [vp, C.c_int, C.c_int]
    L.bdpt_set_lens.argtypes = [vp, C.c_double, C.c_double]
    L.bdpt_set_cell.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.bdpt_clear.argtypes = [vp]
    L.bdpt_render.argtypes = [vp, C.POINTER(RenderParams)]
    L.bdpt_read_framebuffer.argtypes = [vp, C.c_int, vp]
    L.bdpt_framebuffer_device.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.bdpt_tonemap_rgba8.argtypes = [vp, C.c_int, vp]
    L.bdpt_get_stats.argtypes = [vp, C.POINTER(Stats)]
    for name in ("bdpt_intersect_batch", "bdpt_bsdf_eval_batch", "bdpt_bsdf_sample_batch", "bdpt_light_eval_batch",
                 "bdpt_camera_eval_batch"):
        getattr(L, name).argtypes = [vp, dp, C.c_size_t, dp]
    L.bdpt_occluded_batch.argtypes = [vp, dp, C.c_size_t, C.POINTER(C.c_int32)]
    L.bdpt_trace_sample.argtypes = [vp, C.POINTER(RenderParams), C.c_int, C.c_int, C.c_int, dp, C.c_size_t,
                                    C.POINTER(C.c_size_t)]
    L.bdpt_device_bvh.argtypes = [vp, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.bdpt_comm_unique_id.argtypes = [vp]
    L.bdpt_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, vp]
    L.bdpt_comm_init_all.argtypes = [C.POINTER(vp), C.c_int]
    L.bdpt_reduce.argtypes = [vp, C.c_int]
    L.bdpt_reduce_all.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    L.bdpt_comm_destroy.argtypes = [vp]
    u32p = C.POINTER(C.c_uint32)
    L.bdpt_host_build_bvh.argtypes = [vp, C.c_uint32, u32p, C.c_uint32, u32p, u32p, u32p]
    for name in EXPORTS:
        if name not in ("bdpt_destroy", "bdpt_last_error"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Renderer:
    """One GPU, one scene, one frame.  Mirrors the reference's RaytracedRenderer for the BDPT path."""

    def __init__(self, device: int = 0, *, ns_aa: int = 1, max_ray_depth: int = 1, seed: int = 0,
                 roulette: bool = False, no_mis: bool = False, balance: bool = False):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.bdpt_create(device, C.byref(self.h))
        if rc != 0:
            msg = self.L.bdpt_last_error(None).decode()
            self.h = None
            raise BdptError(f"bdpt_create failed ({rc}): {msg}")
        # the fields RaytracedRenderer's constructor writes onto the integrator (raytraced_renderer.cpp:55-63)
        self.ns_aa = ns_aa
        self.max_ray_depth = max_ray_depth
        self.seed = seed
        self.roulette = roulette
        self.no_mis = no_mis
        self.balance = balance      # Veach's balance heuristic instead of the reference's power heuristic (beta = 2)
        self.device = device
        self.W = self.H = 0
        self._keep = None

    # -- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.L.bdpt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise BdptError(f"libbdpt_b200 error {rc}: {self.L.bdpt_last_error(self.h).decode()}")

    def set_stream(self, cuda_stream: int):
        self._ck(self.L.bdpt_set_stream(self.h, C.c_void_p(cuda_stream)))

    # -- RaytracedRenderer::set_scene / set_camera / set_frame_size
    def set_scene(self, scene: fs.FlatScene):
        cam, mats, lights, prims = (np.ascontiguousarray(a) for a in (scene.camera, scene.materials, scene.lights, scene.prims))
        self._keep = (cam, mats, lights, prims)
        self._ck(self.L.bdpt_upload_scene(self.h, cam.ctypes.data, mats.ctypes.data, len(mats), lights.ctypes.data,
                                          len(lights), prims.ctypes.data, len(prims)))

    def set_camera(self, camera: np.ndarray):
        cam = np.ascontiguousarray(camera)
        self._ck(self.L.bdpt_set_camera(self.h, cam.ctypes.data))

    def set_environment_map(self, rgb):
        """Lat-long radiance map (h, w, 3) float32 of the scene's environment light, row 0 = up; None: constant."""
        if rgb is None:
            self._ck(self.L.bdpt_set_environment_map(self.h, None, 0, 0))
            return
        a = np.ascontiguousarray(rgb, np.float32)
        self._ck(self.L.bdpt_set_environment_map(self.h, a.ctypes.data, a.shape[1], a.shape[0]))

    def set_frame_size(self, width: int, height: int):
        self._ck(self.L.bdpt_set_frame(self.h, width, height))
        self.W, self.H = width, height

    def set_lens(self, lens_radius: float, focal_distance: float):
        """Thin lens (extension; the reference's BDPT is a pinhole whatever -b / -d say); radius 0 = pinhole."""
        self._ck(self.L.bdpt_set_lens(self.h, lens_radius, focal_distance))

    def set_cell(self, x: int = 0, y: int = 0, dx: int = 0, dy: int = 0):
        """The reference's cell render (`-p x y dx dy`); dx = dy = 0: the whole frame again."""
        self._ck(self.L.bdpt_set_cell(self.h, x, y, dx, dy))

    def clear(self):
        self._ck(self.L.bdpt_clear(self.h))

    # -- start_raytracing (+ wait): renders sample indices [first_sample, first_sample + n_samples) of every pixel
    def _params(self, first_sample=0, n_samples=None, strategy=None):
        p = RenderParams()
        p.ns_total = self.ns_aa
        p.max_ray_depth = self.max_ray_depth
        p.first_sample = first_sample
        p.n_samples = self.ns_aa if n_samples is None else n_samples
        p.flags = ((FLAG_ROULETTE if self.roulette else 0) | (FLAG_NO_MIS if self.no_mis else 0) |
                   (FLAG_BALANCE if self.balance else 0))
        p.only_t, p.only_s = strategy if strategy else (-1, -1)
        p.seed = self.seed
        return p

    def render(self, first_sample: int = 0, n_samples: int | None = None, strategy=None):
        p = self._params(first_sample, n_samples, strategy)
        self._ck(self.L.bdpt_render(self.h, C.byref(p)))

    def framebuffer(self, which: int = BUFFER_SUM) -> np.ndarray:
        """(H, W, 3) float32, row 0 = bottom, linear radiance divided by ns_aa."""
        out = np.empty((self.H, self.W, 3), np.float32)
        self._ck(self.L.bdpt_read_framebuffer(self.h, which, out.ctypes.data))
        return out

    def framebuffer_device(self, which: int):
        ptr, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.bdpt_framebuffer_device(self.h, which, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    # -- multi-GPU: one NCCL sum-reduce of the accumulators (eye + light, one allocation) to the root
    def comm_init_rank(self, nranks: int, rank: int, id_bytes: bytes):
        buf = C.create_string_buffer(bytes(id_bytes), COMM_ID_BYTES)
        self._ck(self.L.bdpt_comm_init_rank(self.h, nranks, rank, C.cast(buf, C.c_void_p)))

    def reduce(self, root: int = 0):
        """Enqueues the reduce on the renderer's stream (after the renders issued so far)."""
        self._ck(self.L.bdpt_reduce(self.h, root))

    def tonemap(self, which: int = BUFFER_SUM) -> np.ndarray:
        """(H, W, 4) uint8, row 0 = top (PNG order): the reference's toColor + save_image flip."""
        out = np.empty((self.H, self.W, 4), np.uint8)
        self._ck(self.L.bdpt_tonemap_rgba8(self.h, which, out.ctypes.data))
        return out

    def stats(self) -> dict:
        s = Stats()
        self._ck(self.L.bdpt_get_stats(self.h, C.byref(s)))
        return s.asdict()

    # -- test hooks
    def device_bvh(self):
        """(nodes (n,20) uint32, prim_order, depth, PLOC rounds) of the tree the device built for the current scene."""
        u32p = C.POINTER(C.c_uint32)
        nn, depth, rounds = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._ck(self.L.bdpt_device_bvh(self.h, None, 0, C.byref(nn), None, C.byref(depth), C.byref(rounds)))
        nodes = np.zeros((nn.value, 20), np.uint32)
        order = np.zeros(len(self._keep[3]), np.uint32)
        self._ck(self.L.bdpt_device_bvh(self.h, nodes.ctypes.data_as(u32p), nn.value, C.byref(nn), order.ctypes.data_as(u32p),
                                        C.byref(depth), C.byref(rounds)))
        return nodes, order, depth.value, rounds.value

    def _batch(self, fn, arr, n_in, n_out):
        arr = np.ascontiguousarray(arr, np.float64).reshape(-1, n_in)
        out = np.empty((arr.shape[0], n_out), np.float64)
        self._ck(getattr(self.L, fn)(self.h, _dp(arr), arr.shape[0], _dp(out)))
        return out

    def intersect(self, rays):
        return self._batch("bdpt_intersect_batch", rays, 8, 6)

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays, np.float64).reshape(-1, 8)
        out = np.empty(rays.shape[0], np.int32)
        self._ck(self.L.bdpt_occluded_batch(self.h, _dp(rays), rays.shape[0], out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def bsdf_eval(self, q):
        return self._batch("bdpt_bsdf_eval_batch", q, 7, 8)

    def bsdf_sample(self, q):
        return self._batch("bdpt_bsdf_sample_batch", q, 7, 7)

    def light_eval(self, q):
        return self._batch("bdpt_light_eval_batch", q, 7, 6)

    def camera_eval(self, q):
        return self._batch("bdpt_camera_eval_batch", q, 5, 11)

    def trace_sample(self, x: int, y: int, sample: int):
        p = self._params()
        cap = 1 << 16
        out = np.zeros(cap, np.float64)
        n = C.c_size_t()
        self._ck(self.L.bdpt_trace_sample(self.h, C.byref(p), x, y, sample, _dp(out), cap, C.byref(n)))
        out = out[: n.value]
        ne, nl = int(out[0]), int(out[1])
        verts = out[2:2 + 13 * (ne + nl)].reshape(-1, 13)
        strat = out[2 + 13 * (ne + nl):].reshape(-1, 8)
        return {"eye": verts[:ne], "light": verts[ne:], "strategies": strat}


def comm_unique_id() -> bytes:
    """A fresh NCCL id (rank 0 makes it, every rank passes it to Renderer.comm_init_rank)."""
    L = load_library()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = L.bdpt_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise BdptError(f"bdpt_comm_unique_id failed ({rc}): {L.bdpt_last_error(None).decode()}")
    return buf.raw


def host_build_bvh(scene: fs.FlatScene):
    """Host-only: (nodes as (n,20) uint32, prim_order, max_depth) of the BVH bdpt_upload_scene would build."""
    L = load_library()
    prims = np.ascontiguousarray(scene.prims)
    n_nodes, depth = C.c_uint32(), C.c_uint32()
    rc = L.bdpt_host_build_bvh(prims.ctypes.data, len(prims), None, 0, C.byref(n_nodes), None, C.byref(depth))
    if rc != 0:
        raise BdptError(f"bdpt_host_build_bvh failed ({rc})")
    nodes = np.zeros((n_nodes.value, 20), np.uint32)
    order = np.zeros(len(prims), np.uint32)
    u32p = C.POINTER(C.c_uint32)
    rc = L.bdpt_host_build_bvh(prims.ctypes.data, len(prims), nodes.ctypes.data_as(u32p), n_nodes.value, C.byref(n_nodes),
                               order.ctypes.data_as(u32p), C.byref(depth))
    if rc != 0:
        raise BdptError(f"bdpt_host_build_bvh failed ({rc})")
    return nodes, order, depth.value
