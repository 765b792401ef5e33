import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import api, flatscene as fs
sc = fs.load(os.path.join(ROOT, "tests", "golden", "scenes", "CBbunny.fscene.gz"))
R = api.Renderer(0, ns_aa=64, max_ray_depth=5)
for it in range(3):
    t0 = time.perf_counter(); R.set_scene(sc); t1 = time.perf_counter(); R.set_frame_size(1024, 768); t2 = time.perf_counter()
    R.render(); t3 = time.perf_counter(); img = R.framebuffer(0); t4 = time.perf_counter()
    print("set_scene %.1f ms set_frame %.1f ms render %.1f ms (device %.1f) read %.1f ms; bvh build %.1f ms" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, R.stats()["ms_total"], (t4 - t3) * 1e3, R.stats()["ms_bvh_build"]))
