import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import api, flatscene as fs
sc = fs.load(os.path.join(ROOT, "tests", "golden", "scenes", "CBbunny.fscene.gz"))
W, H, spp, m = 256, 192, 16, 5
Rs = {}
for sched in (1, 2):
    os.environ["BDPT_SCHED"] = str(sched)
    R = api.Renderer(0, ns_aa=spp, max_ray_depth=m, seed=1); R.set_scene(sc); R.set_frame_size(W, H)
    Rs[sched] = R
# determinism of sched 2
a = []
for _ in range(2):
    Rs[2].clear(); Rs[2].render(); a.append(Rs[2].framebuffer(1))
print("sched2 run-to-run max diff", np.abs(a[0] - a[1]).max())
for (y, x) in [(17, 95), (18, 166), (18, 253), (22, 154), (61, 64)]:
    for k in range(spp):
        t1 = Rs[1].trace_sample(x, y, k); t2 = Rs[2].trace_sample(x, y, k)
        same_shape = all(t1[q].shape == t2[q].shape for q in ("eye", "light", "strategies"))
        if not same_shape:
            print((x, y, k), "shapes differ", [(t1[q].shape, t2[q].shape) for q in t1]); continue
        for q in ("eye", "light", "strategies"):
            d = np.nan_to_num(np.abs(t1[q] - t2[q]))
            if d.max() > 0:
                i, j = np.unravel_index(np.argmax(d), d.shape)
                print((x, y, k), q, "max diff", d.max(), "at row", i, "col", j, t1[q][i], t2[q][i])
