"""One render of a scene (profiling target). usage: one_render.py scene W H spp m   (scene: fixture name or dae:<name>)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import api
import bench
name = sys.argv[1]; W, H, spp, m = map(int, sys.argv[2:6])
sc = bench.load_fixture(name)
R = api.Renderer(0, ns_aa=spp, max_ray_depth=m, seed=1)
R.set_scene(sc); R.set_frame_size(W, H)
R.render()
st = R.stats()
print("%s %dx%d %dspp m=%d: %.1f ms, %.1f Msamples/s" % (name, W, H, spp, m, st["ms_total"], W * H * spp / st["ms_total"] / 1e3))
