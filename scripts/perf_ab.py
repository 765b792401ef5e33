"""A/B of kernel variants on the bench workload at reduced spp (run on the GPU box).
usage: perf_ab.py scene W H spp m [VARIANT ...]   VARIANT = "K=V,K2=V2" (tuning knobs of DESIGN.md §4, read at
bdpt_create / scene upload); scene = a fixture name or dae:<name> (staged under oracle/_ref/scenes).
Every variant must render the image of the default build (relMSE < 1e-5)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import api
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "CBbunny"
W, H, spp, m = (int(a) for a in sys.argv[2:6]) if len(sys.argv) > 5 else (1024, 768, 64, 5)
variants = [dict(kv.split("=") for kv in v.split(",") if kv) for v in sys.argv[6:]]
sc = bench.load_fixture(name)
KNOBS = [k for k in os.environ if k.startswith("BDPT_")]

def run(env):
    for k in list(os.environ):
        if k.startswith("BDPT_") and k not in KNOBS:
            os.environ.pop(k)
    os.environ.update({k: str(v) for k, v in env.items()})
    R = api.Renderer(0, ns_aa=spp, max_ray_depth=m, seed=1)
    R.set_scene(sc); R.set_frame_size(W, H)
    R.render(); R.clear(); R.render()
    st = R.stats()
    img = R.framebuffer(0)
    R.close()
    tb = max(1, st["timed_batches"])
    print("%-44s %8.1f ms %7.1f Ms/s | walk %.3f connect %.3f shadow %.3f ms/batch | rays/sample %.2f | scene %.1f ms" % (
        env, st["ms_total"], W * H * spp / st["ms_total"] / 1e3, st["ms_walk"] / tb, st["ms_connect"] / tb, st["ms_shadow"] / tb,
        (st["rays_closest"] + st["rays_shadow"]) / st["samples"], st["ms_bvh_build"]), flush=True)
    return img

base = run({})
for v in variants:
    img = run(v)
    rel = float(np.mean((img - base) ** 2 / (base ** 2 + 1e-2)))
    if not (rel < 1e-5 and abs(img.mean() - base.mean()) < 1e-4 * base.mean()):
        print("   !! image changed: relMSE %.3e, means %.6f vs %.6f" % (rel, img.mean(), base.mean()), flush=True)
