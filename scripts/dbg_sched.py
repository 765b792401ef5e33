import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import api, flatscene as fs
name = sys.argv[1]; W, H, spp, m = map(int, sys.argv[2:6])
sc = fs.load(os.path.join(ROOT, "tests", "golden", "scenes", name + ".fscene.gz"))
out = {}
for sched in (0, 1, 2):
    os.environ["BDPT_SCHED"] = str(sched)
    R = api.Renderer(0, ns_aa=spp, max_ray_depth=m, seed=1); R.set_scene(sc); R.set_frame_size(W, H); R.render()
    out[sched] = (R.framebuffer(1), R.framebuffer(2), R.stats())
    st = out[sched][2]
    print(sched, "rays", st["rays_closest"], st["rays_shadow"], "conn", st["connections"], "mean", out[sched][0].mean(), out[sched][1].mean())
for s in (1, 2):
    for k, nm in ((0, "eye"), (1, "light")):
        d = np.abs(out[s][k] - out[0][k])
        bad = np.argwhere(d.max(-1) > 1e-4 * (np.abs(out[0][k]).max(-1) + 0.1))
        print("sched", s, nm, "max diff", d.max(), "pixels differing", len(bad), bad[:5].tolist())
