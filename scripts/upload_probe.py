"""Scene ingest time (bdpt_upload_scene: H2D + device BVH build + records), first and repeated uploads.
usage: upload_probe.py scene [scene ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import api
import bench
for name in sys.argv[1:]:
    sc = bench.load_fixture(name)
    if os.environ.get("PROBE_PINNED"):
        sc = bench.pinned_scene(sc)
    R = api.Renderer(0)
    ts = []
    for i in range(6):
        t0 = time.perf_counter(); R.set_scene(sc); ts.append((time.perf_counter() - t0) * 1e3)
    nodes, order, depth, rounds = R.device_bvh() if os.environ.get("BDPT_BVH") != "host" else (R.stats()["bvh_nodes"] * [0], None, 0, 0)
    print("%-14s %7d prims: upload ms %s | nodes %d depth %d ploc rounds %d" % (name, sc.n_prims, " ".join("%.2f" % t for t in ts), len(nodes), depth, rounds), flush=True)
    R.close()
