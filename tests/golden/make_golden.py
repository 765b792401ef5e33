"""Generates the golden fixtures under tests/golden/ by running THE REFERENCE ITSELF
(oracle/_ref/ref_harness, built from /root/reference by `make -C oracle ref`).

Run here (the container that has /root/reference); the fixtures are committed so that the
GPU box, which has no reference tree, can still check the oracle port and the CUDA path.

    python tests/golden/make_golden.py

Writes:
  scenes/<name>.fscene.gz   flat scene dumped by the reference's own loader (Application::load ->
                            get_static_scene -> build_accel primitive order)
  <name>.units.npz          ray batch + BVHAccel::intersect hits; BSDF f / sample_pdf batch;
                            light contain_point / sample_pdf batch; camera generate_ray /
                            sample_ray_pdf batch (480x360)
  <name>.render.npz         reference `-t 1` render (fp64 sum / eye / light buffers, ray count)
  CBspheres_lambertian.strategies.npz   reference per-(t,s) images (debug filter of bidirection.cpp:492-493)
  <name>.converged.npz      reference render at 1024 spp, 64x48, all host threads (statistical pin)
  published_pooled.npz      the author's published 128-spp PNGs (assets/), box-filtered to 60x45
  <name>.units.npz for CBspheres_microfacet_al_ag / CBspheres_refract (scenes the reference cannot RENDER with
                            BDPT, but whose BSDF::f, lights, camera and BVH it evaluates fine) and
  <name>.bsdfsample.npz     BSDF::sample_f of the reference called in order from a fresh process (its own
                            default-seeded streams): inputs {material, wo}, outputs {wi, pdf, f}   [`extra` mode]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
from bdpt import flatscene as fs  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# name -> (render W, H, ns, m)
SCENES = {
    "CBspheres_lambertian": (64, 48, 2, 5),
    "CBspheres": (64, 48, 2, 5),
    "CBgems": (48, 36, 1, 16),
    "CBspheres_lambertian_occluded": (48, 36, 2, 4),
    "CBempty": (40, 30, 2, 3),
    "CBbunny": (40, 30, 1, 5),
    "CBcoil": (40, 30, 1, 5),
}


# scenes the reference cannot render with BDPT (MicrofacetBSDF::sample_pdf aborts, advanced_bsdf.cpp:144-148;
# RefractionBSDF::sample_f leaves its outputs unset on total internal reflection, :166-168): the scene dump, the
# unit batches (f, lights, camera, rays) and the sample_f replay — everything but a render
EXTRA_SCENES = ["CBspheres_microfacet_al_ag", "CBspheres_refract"]
# sample_f replays: every BSDF class the reference has (diffuse, emission, mirror, glass, refraction, microfacet x3)
SAMPLE_SCENES = ["CBspheres", "CBgems", "CBspheres_microfacet_al_ag", "CBspheres_refract", "CBbunny_microfacet_cu"]


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def ray_batch(sc, rng, n):
    """Interior random rays, primary-like rays from the camera, and short shadow-like segments."""
    lo, hi = np.array([-1.0, 0.0, -1.0]), np.array([1.0, 1.5, 1.0])
    k = n // 3
    o1 = rng.uniform(lo, hi, (k, 3)); d1 = unit(rng.normal(size=(k, 3)))
    r1 = np.concatenate([o1, d1, np.full((k, 1), float(np.float32(1e-5))), np.full((k, 1), np.inf)], 1)
    cam = sc.camera[0]
    tgt = rng.uniform(lo, hi, (k, 3)); o2 = np.broadcast_to(cam["pos"], (k, 3)); d2 = unit(tgt - o2)
    r2 = np.concatenate([o2, d2, np.full((k, 1), cam["nclip"]), np.full((k, 1), cam["fclip"])], 1)
    a = rng.uniform(lo, hi, (n - 2 * k, 3)); b = rng.uniform(lo, hi, (n - 2 * k, 3))
    dist = np.linalg.norm(b - a, axis=1, keepdims=True)
    r3 = np.concatenate([a, (b - a) / dist, np.full_like(dist, float(np.float32(1e-5))), dist - float(np.float32(1e-5))], 1)
    # every number exactly representable in fp32, so the fp64 reference and the fp32 CUDA path are given the SAME ray
    return np.concatenate([r1, r2, r3], 0).astype(np.float32).astype(np.float64)


def upper_biased(rng, n):
    """Unit vectors, two thirds of them in the upper hemisphere (where MicrofacetBSDF is non-zero)."""
    v = unit(rng.normal(size=(n, 3)))
    k = 2 * n // 3
    v[:k, 2] = np.abs(v[:k, 2])
    return v


def extra_units(name, with_scene_units=True):
    """Unit batches of a scene the reference cannot render, and the sample_f replay."""
    dae = O.ref_scene(name)
    tmp = os.path.join(OUT, "scenes", name + ".fscene")
    O.ref_dump(dae, tmp)
    sc = fs.load(tmp)
    os.remove(tmp)
    rng = np.random.default_rng(sum(map(ord, name)) + 1)
    M = 900
    if with_scene_units:
        sc.save(os.path.join(OUT, "scenes", name + ".fscene.gz"))
        rays = ray_batch(sc, rng, 3000)
        hits = O.ref_batch("intersect", dae, rays, 8, 6)
        bq = np.zeros((M, 7)); bq[:, 0] = rng.integers(0, len(sc.materials), M)
        bq[:, 1:4] = upper_biased(rng, M); bq[:, 4:7] = upper_biased(rng, M)
        bq = bq.astype(np.float32).astype(np.float64)
        bout = O.ref_batch("bsdf", dae, bq, 7, 8)
        lq = np.zeros((M, 7)); lq[:, 0] = rng.integers(0, len(sc.lights), M)
        lq[:, 1:4] = rng.uniform([-1, 0, -1], [1, 1.5, 1], (M, 3))
        on = sc.lights["position"][lq[: M // 2, 0].astype(int)]
        lq[: M // 2, 2] = on[:, 1]
        lq[:, 4:7] = unit(rng.normal(size=(M, 3)))
        lq = lq.astype(np.float32).astype(np.float64)
        lout = O.ref_batch("light", dae, lq, 7, 6)
        cq = np.zeros((M, 5)); cq[:, 0:2] = rng.uniform(0, 1, (M, 2)); cq[:, 2:5] = rng.uniform([-1, 0, -1], [1, 1.5, 1], (M, 3))
        cq = cq.astype(np.float32).astype(np.float64)
        cout = O.ref_batch("camera", dae, cq, 5, 11, extra=(480, 360))
        np.savez_compressed(os.path.join(OUT, name + ".units.npz"), rays=rays, hits=hits, bsdf_in=bq, bsdf_out=bout,
                            light_in=lq, light_out=lout, cam_in=cq, cam_out=cout, cam_wh=np.array([480, 360]))
    return sc


def sample_replay(name):
    dae = O.ref_scene(name)
    tmp = os.path.join(OUT, "scenes", name + ".fscene.tmp")
    O.ref_dump(dae, tmp)
    sc = fs.load(tmp)
    os.remove(tmp)
    rng = np.random.default_rng(sum(map(ord, name)) + 2)
    M = 1500
    q = np.zeros((M, 4)); q[:, 0] = rng.integers(0, len(sc.materials), M)
    q[:, 1:4] = upper_biased(rng, M)
    q = q.astype(np.float32).astype(np.float64)
    out = O.ref_batch("bsdf_sample", dae, q, 4, 7)
    # for the copper bunny only material 0 differs from CBbunny's: f of that one material as well
    bq = np.zeros((M, 7)); bq[:, 0] = rng.integers(0, len(sc.materials), M)
    bq[:, 1:4] = upper_biased(rng, M); bq[:, 4:7] = upper_biased(rng, M)
    bq = bq.astype(np.float32).astype(np.float64)
    bout = O.ref_batch("bsdf", dae, bq, 7, 8)
    np.savez_compressed(os.path.join(OUT, name + ".bsdfsample.npz"), samp_in=q, samp_out=out, bsdf_in=bq, bsdf_out=bout,
                        mat_types=sc.materials["type"].astype(np.int32))
    print(name, "sample_f replay:", M, "calls;", int(np.isnan(out[:, 3]).sum()), "left unset by the reference (refraction TIR)")


def envmap_units():
    """EnvironmentLight::sample_dir of the reference on a small synthetic lat-long map (envdir command)."""
    import subprocess, tempfile
    rng = np.random.default_rng(77)
    w, h = 16, 8
    m = rng.uniform(0.1, 3.0, (h, w, 3)).astype(np.float32).astype(np.float64)
    d = unit(rng.normal(size=(2000, 3)))
    d[:6] = [[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1]]     # poles and the seam
    d = d.astype(np.float32).astype(np.float64)
    with tempfile.TemporaryDirectory() as td:
        m.tofile(os.path.join(td, "m.bin")); d.tofile(os.path.join(td, "d.bin"))
        subprocess.run([O.REF_HARNESS, "envdir", str(w), str(h), "m.bin", "d.bin", "o.bin"], check=True, cwd=td, capture_output=True)
        out = np.fromfile(os.path.join(td, "o.bin"), np.float64).reshape(-1, 3)
    np.savez_compressed(os.path.join(OUT, "envmap.units.npz"), map=m.astype(np.float32), dirs=d, out=out)
    print("envmap units:", out.shape, out.mean())


def lens_units():
    """Camera::generate_ray_for_thin_lens (camera_lens.cpp:22-43) of the reference, CBspheres camera at 480x360."""
    rng = np.random.default_rng(78)
    q = rng.uniform(0, 1, (1500, 4)); q[:, 3] *= 2 * np.pi
    q = q.astype(np.float32).astype(np.float64)
    out = O.ref_batch("lensray", O.ref_scene("CBspheres"), q, 4, 6, extra=(480, 360, 0.125, 4.5))
    np.savez_compressed(os.path.join(OUT, "lens.units.npz"), q=q, out=out, params=np.array([480, 360, 0.125, 4.5]))
    print("lens units:", out.shape)


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    if only == "envmap":
        envmap_units()
        return
    if only == "lens":
        lens_units()
        return
    assert O.have_ref(), "build the reference first: make -C oracle ref"
    os.makedirs(os.path.join(OUT, "scenes"), exist_ok=True)
    for name in EXTRA_SCENES:
        extra_units(name)
    for name in SAMPLE_SCENES:
        sample_replay(name)
    envmap_units()
    lens_units()
    if only == "extra":
        return
    for name, (W, H, ns, m) in SCENES.items():
        dae = O.ref_scene(name)
        tmp = os.path.join(OUT, "scenes", name + ".fscene")
        O.ref_dump(dae, tmp)
        sc = fs.load(tmp)
        os.remove(tmp)
        sc.save(os.path.join(OUT, "scenes", name + ".fscene.gz"))
        rng = np.random.default_rng(sum(map(ord, name)))
        big = sc.n_prims > 1000
        rays = ray_batch(sc, rng, 3000 if not big else 6000)
        hits = O.ref_batch("intersect", dae, rays, 8, 6)
        M = 600
        bq = np.zeros((M, 7)); bq[:, 0] = rng.integers(0, len(sc.materials), M)
        bq[:, 1:4] = unit(rng.normal(size=(M, 3))); bq[:, 4:7] = unit(rng.normal(size=(M, 3)))
        bq = bq.astype(np.float32).astype(np.float64)
        bout = O.ref_batch("bsdf", dae, bq, 7, 8)
        lq = np.zeros((M, 7)); lq[:, 0] = rng.integers(0, len(sc.lights), M)
        lq[:, 1:4] = rng.uniform([-1, 0, -1], [1, 1.5, 1], (M, 3))
        on = sc.lights["position"][lq[: M // 2, 0].astype(int)]
        lq[: M // 2, 2] = on[:, 1]  # on the light's plane (scenes are Y-up, lights face -Y)
        lq[:, 4:7] = unit(rng.normal(size=(M, 3)))
        lq = lq.astype(np.float32).astype(np.float64)
        lout = O.ref_batch("light", dae, lq, 7, 6)
        cq = np.zeros((M, 5)); cq[:, 0:2] = rng.uniform(0, 1, (M, 2)); cq[:, 2:5] = rng.uniform([-1, 0, -1], [1, 1.5, 1], (M, 3))
        cq = cq.astype(np.float32).astype(np.float64)
        cout = O.ref_batch("camera", dae, cq, 5, 11, extra=(480, 360))
        np.savez_compressed(os.path.join(OUT, name + ".units.npz"), rays=rays, hits=hits, bsdf_in=bq, bsdf_out=bout,
                            light_in=lq, light_out=lout, cam_in=cq, cam_out=cout, cam_wh=np.array([480, 360]))
        if only == "units":
            continue
        s, e, l, info = O.ref_render(dae, W, H, ns, m, threads=1)
        np.savez_compressed(os.path.join(OUT, name + ".render.npz"), sum=s, eye=e, light=l,
                            params=np.array([W, H, ns, m]), rays=np.array([info["rays"]]))
        print(name, "prims", sc.n_prims, "render rays", info["rays"])

    if only == "units":
        return
    # converged reference renders (the reference's own binary, all host threads): the level-3 pin
    for name, (W, H, ns, m) in {"CBspheres_lambertian": (64, 48, 1024, 5), "CBspheres": (64, 48, 1024, 5),
                                "CBgems": (64, 48, 1024, 5), "CBbunny": (64, 48, 512, 5),
                                "CBspheres_lambertian_occluded": (64, 48, 1024, 5), "CBgems_m16": (48, 36, 512, 16)}.items():
        if only not in ("all", "converged") or (only == "converged" and os.path.exists(os.path.join(OUT, name + ".converged.npz"))):
            continue
        s, e, l, info = O.ref_render(O.ref_scene(name.split("_m")[0]), W, H, ns, m, threads=os.cpu_count())
        np.savez_compressed(os.path.join(OUT, name + ".converged.npz"), sum=s.astype(np.float32), eye=e.astype(np.float32),
                            light=l.astype(np.float32), params=np.array([W, H, ns, m]))
        print(name, "converged", info)

    if only == "converged":
        return
    # the author's PUBLISHED renders (assets/*_bdpt_128.png: 480x360, 128 spp, tone-mapped 8-bit) reduced to 60x45 box-filtered
    # means: a loose pin that is independent of our build of the reference (framing, tone map, overall light transport)
    from PIL import Image
    pub = {}
    for name in ("CBspheres", "CBgems"):
        a = np.asarray(Image.open(f"/root/reference/assets/{name}_bdpt_128.png").convert("RGB")).astype(np.float64)
        pub[name] = a.reshape(45, 8, 60, 8, 3).mean((1, 3)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "published_pooled.npz"), **pub)

    # per-strategy images, small
    name, (W, H, ns, m) = "CBspheres_lambertian", (32, 24, 2, 3)
    dae = O.ref_scene(name)
    out = {}
    for t in range(1, m + 2):
        for s_ in range(0, m + 2):
            s, e, l, info = O.ref_render(dae, W, H, ns, m, threads=1, strategy=(t, s_))
            out[f"sum_t{t}_s{s_}"] = s
    np.savez_compressed(os.path.join(OUT, name + ".strategies.npz"), params=np.array([W, H, ns, m]), **out)
    print("strategies done")


if __name__ == "__main__":
    main()
