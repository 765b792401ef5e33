"""Does k_connect's time depend on which physical memory the batch slab got?  Several contexts alive at once (so each
slab is different memory), same scene, same render; prints per-batch kernel times of each.  usage: placement_probe.py [n [scene W H spp]]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import api, flatscene as fs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
name = sys.argv[2] if len(sys.argv) > 2 else "CBbunny"
W, H, spp = (int(a) for a in sys.argv[3:6]) if len(sys.argv) > 5 else (1024, 768, 128)
sc = fs.load(os.path.join(ROOT, "tests", "golden", "scenes", name + ".fscene.gz"))
Rs = []
for i in range(n):
    R = api.Renderer(0, ns_aa=spp, max_ray_depth=5, seed=1)
    R.set_scene(sc); R.set_frame_size(W, H)
    R.render(); R.clear(); R.render()
    st = R.stats(); tb = max(1, st["timed_batches"])
    print("context %d: walk %.3f connect %.3f shadow %.3f ms per batch of %d" % (i, st["ms_walk"] / tb, st["ms_connect"] / tb, st["ms_shadow"] / tb, st["batch_samples"]), flush=True)
    Rs.append(R)
for i, R in enumerate(Rs):   # again, all alive
    R.clear(); R.render(); st = R.stats(); tb = max(1, st["timed_batches"])
    print("again   %d: walk %.3f connect %.3f shadow %.3f" % (i, st["ms_walk"] / tb, st["ms_connect"] / tb, st["ms_shadow"] / tb), flush=True)
