"""A/B of kernel variants on the bench workload at reduced spp (run on the GPU box).
usage: perf_ab.py [scene] [W H spp m]  — env knobs are applied per Renderer (read at bdpt_create)."""
import os, sys, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import api, flatscene as fs

name = sys.argv[1] if len(sys.argv) > 1 else "CBbunny"
W, H, spp, m = (int(a) for a in sys.argv[2:6]) if len(sys.argv) > 5 else (1024, 768, 64, 5)
sc = fs.load(os.path.join(ROOT, "tests", "golden", "scenes", name + ".fscene.gz"))

def run(env):
    for k in ("BDPT_SCHED", "BDPT_WALK_CTAS", "BDPT_SHADOW_CTAS", "BDPT_VARIANT", "BDPT_CONNECT_MINB", "BDPT_BATCH_LOG2", "BDPT_BATCH_GIB", "BDPT_BVH_LEAF", "BDPT_BVH_CP", "BDPT_BVH_BINS", "BDPT_TOP_NODES", "BDPT_SPEC", "BDPT_PAD", "BDPT_SKEW"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    R = api.Renderer(0, ns_aa=spp, max_ray_depth=m, seed=1)
    R.set_scene(sc); R.set_frame_size(W, H)
    R.render(); R.clear(); R.render()
    st = R.stats()
    img = R.framebuffer(0)
    R.close()
    tb = max(1, st["timed_batches"])
    print("%-60s %8.1f ms  %7.1f Msamples/s | per batch walk %.3f connect %.3f shadow %.3f ms | rays/sample %.2f" % (
        env, st["ms_total"], W * H * spp / st["ms_total"] / 1e3, st["ms_walk"] / tb, st["ms_connect"] / tb, st["ms_shadow"] / tb,
        (st["rays_closest"] + st["rays_shadow"]) / st["samples"]), flush=True)
    return img

base = run({})
variants = [{"BDPT_SPEC": 0}, {"BDPT_SCHED": 0}, {"BDPT_BATCH_LOG2": 22, "BDPT_BATCH_GIB": 12}]
for v in variants:
    img = run(v)
    d = np.abs(img - base)
    rel = float(np.mean((img - base) ** 2 / (base ** 2 + 1e-2)))
    assert rel < 1e-5 and abs(img.mean() - base.mean()) < 1e-4 * base.mean(), ("image changed", d.max(), rel)
print("all variants produce the same image (relMSE < 1e-5; kernels compiled separately differ in the last fp32 bit)")
