"""Developer smoke script (run on the GPU box): CUDA path vs the fp64 oracle port."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import flatscene as fs, api
from oracle import oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")
def scene(name): return fs.load(os.path.join(GOLD, "scenes", name + ".fscene.gz"))

def relmse(a, b): return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))

names = sys.argv[1:] or ["CBspheres_lambertian", "CBspheres", "CBgems", "CBbunny"]
for name in names:
    sc = scene(name)
    g = np.load(os.path.join(GOLD, name + ".units.npz"))
    R = api.Renderer(0, ns_aa=4, max_ray_depth=5)
    R.set_scene(sc); R.set_frame_size(480, 360)
    print("==", name, "prims", sc.n_prims, "stats", {k: v for k, v in R.stats().items() if k in ("bvh_nodes", "ms_bvh_build")})
    orc = O.Oracle(sc, 480, 360)
    rays = g["rays"]
    got = R.intersect(rays); ref = g["hits"]
    same = got[:, 0] == ref[:, 0]
    hit = (ref[:, 0] >= 0) & same
    print(" hits: id equal %.5f (mismatch %d of %d); max rel dt %.2e; max |dn| %.2e; mat eq %.5f" % (
        same.mean(), (~same).sum(), len(same),
        np.max(np.abs(got[hit, 1] - ref[hit, 1]) / ref[hit, 1]) if hit.any() else 0,
        np.abs(got[hit, 2:5] - ref[hit, 2:5]).max() if hit.any() else 0, (got[:, 5] == ref[:, 5]).mean()))
    bad = np.where(~same)[0][:5]
    if len(bad):
        br = orc.intersect_brute(rays[bad])
        for i, b in zip(bad, br): print("   ray", i, "gpu", got[i, :2], "ref", ref[i, :2], "second t", b[2])
    occ = R.occluded(rays); print(" occluded agrees with closest-hit: %.5f" % np.mean((occ > 0) == (ref[:, 0] >= 0)))
    a = R.bsdf_eval(g["bsdf_in"]); b = g["bsdf_out"]
    print(" bsdf max abs diff %.2e (nan match %s)" % (np.nanmax(np.abs(a - b)), np.array_equal(np.isnan(a), np.isnan(b))))
    a = R.light_eval(g["light_in"]); b = g["light_out"]; print(" light max abs diff %.2e" % np.abs(a - b).max())
    a = R.camera_eval(g["cam_in"]); b = g["cam_out"]
    print(" camera max abs diff ray %.2e We %.2e dirpdf rel %.2e raster eq %.4f" % (np.abs(a[:, :6] - b[:, :6]).max(), np.abs(a[:, 6] - b[:, 6]).max(), np.max(np.abs(a[:, 7] - b[:, 7]) / np.abs(b[:, 7])), np.mean(a[:, 8:10] == b[:, 8:10])))
    # one-sample traces
    R2 = api.Renderer(0, ns_aa=4, max_ray_depth=5, seed=3); R2.set_scene(sc); R2.set_frame_size(64, 48)
    orc.set_frame(64, 48)
    nbad = 0; ncmp = 0
    for (x, y, k) in [(10, 10, 0), (32, 24, 1), (50, 40, 2), (5, 30, 3), (20, 8, 0), (40, 20, 1), (60, 45, 2), (33, 3, 3)]:
        tg = R2.trace_sample(x, y, k); to = orc.trace_sample(4, 5, x, y, k, seed=3)
        ne = min(len(tg["eye"]), len(to["eye"])); nl = min(len(tg["light"]), len(to["light"]))
        pe = np.nanmax(np.abs(tg["eye"][2:ne, :9] - to["eye"][2:ne, :9])) if ne > 2 else 0
        pl = np.nanmax(np.abs(tg["light"][1:nl, :9] - to["light"][1:nl, :9])) if nl > 1 else 0
        so = {(int(r[0]), int(r[1])): r for r in to["strategies"]}
        worst = 0
        for r in tg["strategies"]:
            key = (int(r[0]), int(r[1])); ncmp += 1
            if key not in so: nbad += 1; print("   missing in oracle", key); continue
            q = so[key]
            if np.isnan(q[2]) != np.isnan(r[2]): nbad += 1; print("   occlusion differs", (x, y, k), key, r, q); continue
            if np.isnan(q[2]): continue
            ill_g = r[2:5] * r[5]; ill_o = q[2:5] * q[5]
            err = np.abs(ill_g - ill_o).max() / (np.abs(ill_o).max() + 1e-3)
            worst = max(worst, err)
            if err > 1e-2: nbad += 1; print("   strategy differs", (x, y, k), key, "gpu", r[2:8], "oracle", q[2:8])
        print("  trace", (x, y, k), "verts gpu %d/%d oracle %d/%d; max vertex diff eye %.2e light %.2e; worst strategy rel err %.2e" % (len(tg["eye"]), len(tg["light"]), len(to["eye"]), len(to["light"]), pe, pl, worst))
    print(" strategies compared", ncmp, "bad", nbad)
    # small render with the same random numbers
    for (ns, m) in [(8, 5)]:
        R3 = api.Renderer(0, ns_aa=ns, max_ray_depth=m, seed=11); R3.set_scene(sc); R3.set_frame_size(64, 48)
        R3.render(); gs = R3.framebuffer(0).astype(np.float64); ge = R3.framebuffer(1); gl = R3.framebuffer(2)
        os_, oe, ol, rays_o = orc.render(ns, m, rng_mode=1, seed=11, threads=8)
        d = np.abs(gs - os_)
        print(" render 64x48 ns=%d m=%d: mean gpu %.5f oracle %.5f; relMSE %.3e; frac pixels within 1e-3 rel: %.4f; eye relMSE %.3e light relMSE %.3e; stats %s" % (
            ns, m, gs.mean(), os_.mean(), relmse(gs, os_), np.mean((d <= 1e-3 * (np.abs(os_) + 1e-2)).all(-1)), relmse(ge, oe), relmse(gl, ol),
            {k: v for k, v in R3.stats().items() if k in ("rays_closest", "rays_shadow", "connections")}), "oracle rays", rays_o)
    # quick speed
    R.ns_aa = 64; R.max_ray_depth = 5
    R.render(); R.clear()
    t0 = time.time(); R.render(); dt = time.time() - t0
    st = R.stats()
    print(" speed 480x360 ns=64 m=5: %.1f ms device (%.1f wall) -> %.2f Msamples/s; %.1f Mrays/s; walk/connect/shadow ms %.1f/%.1f/%.1f (first %d of %d batches)" % (
        st["ms_total"], dt * 1e3, 480 * 360 * 64 / st["ms_total"] / 1e3, (st["rays_closest"] + st["rays_shadow"]) / st["ms_total"] / 1e3,
        st["ms_walk"], st["ms_connect"], st["ms_shadow"], st["timed_batches"], st["n_batches"]))
