"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol
include/bdpt_b200.h declares, refuses to run without a device (no CPU fallback), the host BVH
builder produces a sound tree, the flat scene format round-trips, and frames shard across ranks."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ALL_SCENES, GOLDEN, ROOT, load_scene


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from bdpt import api
    hdr = open(os.path.join(ROOT, "include", "bdpt_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(bdpt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = api.load_library()
    for name in declared:
        assert hasattr(lib, name), f"libbdpt_b200.so does not export {name}"
    assert sorted(api.EXPORTS) == declared


def test_host_library_exports_every_declared_symbol():
    from bdpt import host
    hdr = open(os.path.join(ROOT, "include", "bdpt_host.h")).read()
    declared = sorted(set(re.findall(r"\b(bdpt_host_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == ["bdpt_host_apply_camera_settings", "bdpt_host_free", "bdpt_host_last_error", "bdpt_host_load_dae",
                        "bdpt_host_load_envmap"]
    lib = host.load_library()
    for name in declared:
        assert hasattr(lib, name)


def test_struct_layouts_match_header():
    from bdpt import api, flatscene as fs
    assert C.sizeof(api.RenderParams) == 40
    assert C.sizeof(api.Stats) == 96
    assert (fs.HEADER_DT.itemsize, fs.CAMERA_DT.itemsize, fs.MATERIAL_DT.itemsize, fs.LIGHT_DT.itemsize,
            fs.PRIM_DT.itemsize) == (24, 176, 152, 128, 152)
    # compile the header as C and compare sizeof
    src = r'''
    #include <stdio.h>
    #include "bdpt_b200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(BdptRenderParams), sizeof(BdptStats), sizeof(BdptSceneHeader),
      sizeof(BdptCamera), sizeof(BdptMaterial), sizeof(BdptLight), sizeof(BdptPrim)); return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o",
                        os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    assert list(map(int, out)) == [40, 96, 24, 176, 152, 128, 152]


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_device():
    from bdpt import api
    with pytest.raises(api.BdptError, match="no CUDA device|CPU fallback"):
        api.Renderer(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bidirectional-pathtracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libbdpt_oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert not re.search(r"#include\s+\"[^\"]*oracle/", txt), f


def test_flatscene_roundtrip(tmp_path):
    from bdpt import flatscene as fs
    sc = load_scene("CBgems")
    for ext in (".fscene", ".fscene.gz"):
        p = str(tmp_path / ("x" + ext))
        sc.save(p)
        b = fs.load(p)
        assert b.tobytes() == sc.tobytes()
    with pytest.raises(ValueError):
        fs.frombytes(b"\0" * 64)
    w = sc.with_frame(480, 360)
    sd = float(sc.camera["screen_dist"][0])
    assert np.isclose(w.camera["hfov_deg"][0], 2 * np.degrees(np.arctan(480 / (2 * sd))))   # camera.cpp:83-89


# ------------------------------------------------------------------------------ host BVH builder
def decode_nodes(nodes):
    """(n,20) uint32 -> dict of arrays, layout of csrc/bvh8.h."""
    p = nodes[:, 0:3].copy().view(np.float32)
    e = np.stack([(nodes[:, 3] >> s) & 0xff for s in (0, 8, 16)], 1).astype(np.int64)
    imask = (nodes[:, 3] >> 24) & 0xff
    scale = np.ldexp(1.0, e - 127)
    by = nodes[:, 6:20].copy().view(np.uint8).reshape(len(nodes), 7, 8)   # meta, qlo xyz, qhi xyz
    return p.astype(np.float64), scale, imask, nodes[:, 4], nodes[:, 5], by[:, 0], by[:, 1:4].astype(np.float64), by[:, 4:7].astype(np.float64)


def bvh_cost(nodes, prim_cost=0.8):
    """Expected work of a random ray, relative to the root's area: (wide-node visits, primitive tests) — the SAH figure
    the collapse minimises, read back from the quantised child boxes."""
    p, scale, imask, child_base, prim_base, meta, qlo, qhi = decode_nodes(nodes)
    clo = p[:, None, :] + np.transpose(qlo, (0, 2, 1)) * scale[:, None, :]
    chi = p[:, None, :] + np.transpose(qhi, (0, 2, 1)) * scale[:, None, :]
    d = np.maximum(chi - clo, 0)
    A = d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]
    used = meta != 0
    internal = ((imask[:, None] >> np.arange(8)) & 1).astype(bool)
    cnt = np.where(used & ~internal, np.select([meta >> 5 == 1, meta >> 5 == 3, meta >> 5 == 7], [1, 2, 3], 0), 0)
    dd = chi[0][used[0]].max(0) - clo[0][used[0]].min(0)
    A0 = dd[0] * dd[1] + dd[1] * dd[2] + dd[2] * dd[0]
    visits, tests = 1 + (A * internal).sum() / A0, (A * cnt).sum() / A0
    return visits, tests, visits + prim_cost * tests


@pytest.mark.parametrize("algo", ["sah", "ploc"])
@pytest.mark.parametrize("name", ALL_SCENES)
def test_host_bvh_is_sound(name, algo, monkeypatch):
    """Both host builders: the binned-SAH one (BDPT_BVH=host, the checker) and the host restatement of the DEVICE
    build's algorithm (Morton order + PLOC + the shared collapse; csrc/bvh_ploc.h, bvh8_wide.h)."""
    from bdpt import api
    monkeypatch.setenv("BDPT_BVH_HOST_ALGO", algo)
    sc = load_scene(name)
    nodes, order, depth = api.host_build_bvh(sc)
    check_bvh_sound(sc, nodes, order, depth)


def test_ploc_tree_is_as_good_as_the_sah_tree(monkeypatch):
    """Quality of the device build's algorithm against the binned-SAH build, as the SAH cost of the final wide tree."""
    from bdpt import api
    for name in ("CBbunny", "CBcoil", "CBgems"):
        sc = load_scene(name)
        cost = {}
        for algo in ("sah", "ploc"):
            monkeypatch.setenv("BDPT_BVH_HOST_ALGO", algo)
            cost[algo] = bvh_cost(api.host_build_bvh(sc)[0])[2]
        assert cost["ploc"] <= 1.02 * cost["sah"], (name, cost)


def check_bvh_sound(sc, nodes, order, depth):
    n = sc.n_prims
    assert sorted(order.tolist()) == list(range(n)), "every primitive exactly once"
    assert 1 <= depth < 32
    p, scale, imask, child_base, prim_base, meta, qlo, qhi = decode_nodes(nodes)
    # fp32 bounds of the primitives as the kernels see them
    d = sc.prims["data"]
    tri = sc.prims["type"] == 0
    v = d[:, :9].reshape(n, 3, 3).astype(np.float32).astype(np.float64)
    lo = np.where(tri[:, None], v.min(1), d[:, :3] - d[:, 3:4])
    hi = np.where(tri[:, None], v.max(1), d[:, :3] + d[:, 3:4])
    seen_nodes = np.zeros(len(nodes), bool); seen_nodes[0] = True
    seen_prims = np.zeros(n, bool)
    node_lo = {0: None}
    for i in range(len(nodes)):
        clo = p[i][None, :] + qlo[i].T * scale[i][None, :]     # (8,3)
        chi = p[i][None, :] + qhi[i].T * scale[i][None, :]
        rank = 0
        for s in range(8):
            m = int(meta[i, s])
            if m == 0:
                assert not (imask[i] >> s) & 1
                continue
            if (imask[i] >> s) & 1:
                assert m == (0x20 | (24 + s))
                c = int(child_base[i]) + rank
                rank += 1
                assert 0 < c < len(nodes) and not seen_nodes[c]
                seen_nodes[c] = True
                node_lo[c] = (clo[s], chi[s])
            else:
                cnt = {1: 1, 3: 2, 7: 3}[m >> 5]
                off = m & 31
                assert off + cnt <= 24
                for k in range(cnt):
                    slot = int(prim_base[i]) + off + k
                    pid = int(order[slot])
                    assert not seen_prims[slot]; seen_prims[slot] = True
                    assert np.all(clo[s] <= lo[pid] + 1e-12) and np.all(chi[s] >= hi[pid] - 1e-12), "leaf box contains its primitive"
                    # ... with the slack the traversal's folded byte decode relies on (bvh8.cpp "quantisation grid"),
                    # up to what fp32 can represent at these coordinates
                    ulp = np.spacing(np.maximum(np.abs(lo[pid]), np.abs(hi[pid])).astype(np.float32)).astype(np.float64)
                    assert np.all(clo[s] <= lo[pid] - scale[i] / 128 + 4 * ulp) and np.all(chi[s] >= hi[pid] + scale[i] / 128 - 4 * ulp)
        if node_lo.get(i) is not None:
            plo, phi = node_lo[i]
            used = meta[i] != 0
            # a node's own grid starts 1/32 of ITS quantum below its exact box and rounds outward by less than one
            # quantum, so its children may stick out of the (also outward-rounded) slot box of the parent by that
            assert (np.all(clo[used].min(0) >= plo - scale[i] / 16 - 1e-6)
                    and np.all(chi[used].max(0) <= phi + scale[i] + 1e-6)), "child near parent's slot box"
    assert seen_nodes.all() and seen_prims.all()


def test_host_bvh_degenerate_inputs():
    from bdpt import api, flatscene as fs
    sc = load_scene("CBempty")
    one = fs.FlatScene(sc.camera, sc.materials, sc.lights, sc.prims[:1])
    nodes, order, depth = api.host_build_bvh(one)
    assert len(nodes) == 1 and order.tolist() == [0]
    dup = fs.FlatScene(sc.camera, sc.materials, sc.lights, np.repeat(sc.prims[:1], 40))   # identical centroids
    nodes, order, depth = api.host_build_bvh(dup)
    assert sorted(order.tolist()) == list(range(40)) and depth < 32


# -------------------------------------------------------------------------------------- sharding
def test_shard_samples_partition():
    from bdpt import shard
    for ns in (1, 4, 7, 128, 1024, 4096):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.shard_samples(ns, r, world) for r in range(world)]
            cover = []
            for f, n in blocks:
                cover += list(range(f, f + n))
            assert cover == list(range(ns))
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1
    with pytest.raises(ValueError):
        shard.shard_samples(4, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
    from bdpt import shard
    from oracle import oracle as O
    from conftest import load_scene as ls
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    W, H, ns, m = 24, 18, 6, 3
    orc = O.Oracle(ls("CBspheres_lambertian"), W, H)
    first, n = shard.shard_samples(ns, rank, world)
    # each rank renders its block of sample indices of every pixel (the oracle stands in for the GPU here)
    s, e, l, _ = orc.render(ns, m, rng_mode=1, seed=9, first_sample=first, n_samples=n)
    te, tl = torch.from_numpy(e.copy()), torch.from_numpy(l.copy())
    shard.reduce_to_root([te, tl], root=0)
    if rank == 0:
        q.put((te.numpy() + tl.numpy()))
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_rank_frame():
    """world_size 2 over gloo: the union of the shards, reduced to rank 0, is the single-rank frame."""
    import torch.multiprocessing as mp
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    got = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    ref = O.Oracle(load_scene("CBspheres_lambertian"), 24, 18).render(6, 3, rng_mode=1, seed=9)[0]
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-14)


# ------------------------------------------------------------------------------------- bench arms
def test_bench_reference_arm_prints_contract_line():
    from oracle import oracle as O
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "config0",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, check=True, timeout=600)
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["value"] > 1e3
    assert line["cpu_baseline"]["kind"] == ("reference" if O.have_ref() else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0
