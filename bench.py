#!/usr/bin/env python
"""bench.py — BDPT samples/s (and Mrays/s) of the B200 path on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--impl reference]

Workload (config.workload): BASELINE.json configs[2], "CBlucy 1024x768, 1024 spp, m=5 — large-mesh BVH
traversal throughput", the configuration the north_star's target is quoted on.  dae/sky/CBlucy.dae is
not in the reference checkout, so the documented CBlucy-class stand-in is used: CBbunny4,
the Cornell-box bunny refined to 114,304 triangles (scripts/make_standin.py; a .dae both arms load;
SURVEY.md §8d).  One "step" = one whole frame: 1024x768 pixels x 1024 spp = 805,306,368 BDPT samples
(1 sample = 1 eye subpath + 1 light subpath + all (s,t) connections).  With N GPUs the frame's sample
indices are sharded N ways (strong scaling: the frame is fixed) and the accumulators (eye + light image,
one allocation) are sum-reduced to rank 0 with ONE ncclReduce inside the step (bdpt_reduce).
At N = 1 the line also carries `extras`: the other configs of BASELINE.json and of SURVEY §8d (CBbunny,
CBbunny16, the glass bunny, CBspheres, CBgems m=16 with roulette) measured in the same run.

  value      whole-job samples/s, scene resident in HBM, timed with CUDA events on the launching
             stream, max over ranks (includes the NCCL reduce for N > 1)
  e2e        same metric through the C-ABI from HOST buffers: bdpt_upload_scene (host BVH build + H2D)
             + bdpt_set_frame + bdpt_render + reduce + bdpt_read_framebuffer (D2H) every step
  roofline   dominant kernel: algorithmic HBM bytes per launch / its measured duration vs the measured
             HBM peak.  This path is NOT HBM-bound (it is SM-issue / L1-L2 latency bound);
             the fraction says how far from the memory roof it runs, profiles/ holds the issue-side numbers.
  cpu_baseline   the reference's own binary (oracle/_ref/pathtracer_ref, kind "reference") on all host
             cores, on a bounded sample of the same workload; N = 1, rank 0 only.  Two runs (s and 3s spp):
             `value` is the MARGINAL samples/s — the reference spends a fixed ~4.5 s per run re-tonemapping
             the whole frame after every tile (raytraced_renderer.cpp:619), which a 1024-spp frame amortises
             and a 3-spp sample does not
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (flat scene fixture, reference .dae, W, H, spp, m, roulette, description)
    "config0": ("CBspheres_lambertian", "CBspheres_lambertian", 480, 360, 4, 5, 0, "configs[0] CBspheres_lambertian 480x360 4spp m=5"),
    "config1": ("CBspheres", "CBspheres", 480, 360, 128, 5, 0, "configs[1] CBspheres 480x360 128spp m=5"),
    "config2": ("CBbunny", "CBbunny", 1024, 768, 1024, 5, 0,
                "configs[2] CBlucy 1024x768 1024spp m=5 on the CBbunny stand-in (CBlucy.dae absent from the reference checkout)"),
    "config2x4": ("dae:CBbunny4", "CBbunny4", 1024, 768, 1024, 5, 0,
                  "configs[2] CBlucy 1024x768 1024spp m=5 on the CBlucy-class stand-in CBbunny4 (bunny refined to 114,304 triangles by scripts/make_standin.py)"),
    "config2x16": ("dae:CBbunny16", "CBbunny16", 1024, 768, 1024, 5, 0,
                   "configs[2] CBlucy 1024x768 1024spp m=5 on CBbunny16 (bunny refined to 457,216 triangles by scripts/make_standin.py)"),
    "config2glass": ("glass:CBbunny", "CBbunny", 1024, 768, 1024, 5, 0,
                     "configs[2] shape on CBbunny with the bunny's material switched to CBlucy-style glass (ior 1.45): generic kernel builds"),
    "config3": ("CBgems", "CBgems", 1024, 768, 1024, 16, 1, "configs[3] CBgems 1024x768 1024spp m=16 roulette"),
    "config4": ("CBbunny", "CBbunny", 1920, 1080, 4096, 8, 0, "configs[4] CBlucy 1920x1080 4096spp m=8 on the CBbunny stand-in"),
}
DEFAULT_CONFIG = "config2x4"
EXTRA_CONFIGS = ["config2", "config2x16", "config2glass", "config1", "config3", "config0"]
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def load_fixture(fixture):
    """Committed flat-scene fixture, or `dae:<name>`: a staged .dae read by the C++ loader (libbdpt_host.so)."""
    from bdpt import flatscene as fs
    if fixture.startswith("glass:"):
        sc = load_fixture(fixture[6:])
        mats = sc.materials.copy()
        g = mats[0]
        g["type"] = fs.MAT_GLASS
        g["reflectance"] = 1.0; g["transmittance"] = 1.0; g["ior"] = 1.45
        mats[0] = g
        return fs.FlatScene(sc.camera, mats, sc.lights, sc.prims)
    if fixture.startswith("dae:"):
        from bdpt import host
        path = os.path.join(ROOT, "oracle", "_ref", "scenes", fixture[4:] + ".dae")
        if not os.path.exists(path):
            raise SystemExit(f"{path} is missing: run `python scripts/make_standin.py` where /root/reference is mounted")
        return host.load_dae(path)
    return fs.load(os.path.join(ROOT, "tests", "golden", "scenes", fixture + ".fscene.gz"))


def pinned_scene(scene):
    """The same scene with its arrays in page-locked host memory (what the e2e contract asks the step's inputs to be in):
    bdpt_upload_scene's H2D copy of the primitives then runs at PCIe rate instead of through the driver's staging buffer."""
    import numpy as np
    import torch
    from bdpt import flatscene as fs

    def pin(a):
        a = np.ascontiguousarray(a)
        if a.nbytes == 0:
            return a
        t = torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory()
        return t.numpy().view(a.dtype).reshape(a.shape)
    out = fs.FlatScene(pin(scene.camera), pin(scene.materials), pin(scene.lights), pin(scene.prims))
    return out


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap," \
             "clocks.mem,power.draw"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("BENCH_SMI_MS", "200"), "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, mem, pw, reasons = [], [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                mem.append(float(f[9])); pw.append(float(f[10]))
            except (ValueError, IndexError):
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "mem_mhz": float(np.median(mem)) if mem else None,
                "power_w": float(np.median(pw)) if pw else None}


# ------------------------------------------------------------------------------- reference arm
def run_reference_cli(dae_name, W, H, spp, m, threads):
    """One run of the reference's own binary; returns (samples/s from its 'Rendering... 100%! (Xs)' line, seconds)."""
    from oracle import oracle as O
    out = os.path.join("/tmp", f"bench_ref_{os.getpid()}.png")
    r = subprocess.run([O.REF_CLI, "-t", str(threads), "-s", str(spp), "-m", str(m), "-r", str(W), str(H), "-f", out,
                        O.ref_scene(dae_name)], capture_output=True, text=True, check=True)
    txt = (r.stdout + r.stderr).replace("\r", "\n")
    mt = re.search(r"100%! \(([0-9.]+)s\)", txt)
    secs = float(mt.group(1))
    for f in (out, out.replace(".png", "_rate.png")):
        if os.path.exists(f):
            os.remove(f)
    return W * H * spp / secs, secs


def cpu_sample_spp(W, H, budget_s=30.0):
    """Smallest of the two spp values of the bounded CPU sample: about budget_s seconds of CPU work for the PAIR
    (s and 3s spp) at ~2e4 samples/s/core plus the reference's fixed ~4.5 s per run; at least 1 spp."""
    cores = os.cpu_count() or 1
    per_spp = W * H / (2.0e4 * cores)
    s = int((budget_s - 9.0) / (4.0 * per_spp)) if per_spp > 0 else 1
    return max(1, min(8, s))


def reference_baseline(cfg, budget_s=30.0):
    """cpu_baseline object: the reference binary if it was built here, else the oracle port.  Two-point fit
    (BASELINE.md §3): runs at s and 3s spp; value = marginal samples/s, the fixed seconds are reported."""
    from oracle import oracle as O
    fixture, dae, W, H, spp, m, roulette, _ = cfg
    cores = os.cpu_count() or 1
    s = cpu_sample_spp(W, H, budget_s)
    if os.path.exists(O.REF_CLI) and os.path.exists(O.ref_scene(dae)):
        v1, t1 = run_reference_cli(dae, W, H, s, m, cores)
        v3, t3 = run_reference_cli(dae, W, H, 3 * s, m, cores)
        marginal = W * H * 2 * s / max(t3 - t1, 1e-9)
        fixed = max(0.0, t1 - W * H * s / marginal)
        return {"value": marginal, "unit": "samples/s", "cores": cores, "kind": "reference",
                "single_point_value": v3, "fixed_seconds": fixed, "seconds": [t1, t3], "spp": [s, 3 * s],
                "sample": f"pathtracer_ref -t {cores} -m {m} -r {W} {H} {dae}.dae at -s {s} ({t1:.2f} s) and -s {3 * s} ({t3:.2f} s): "
                          f"marginal rate {marginal:.4g} samples/s, fixed cost {fixed:.2f} s per run (per-tile whole-frame tonemap, "
                          f"raytraced_renderer.cpp:619); the single {3 * s}-spp run alone gives {v3:.4g}"}
    sc = load_fixture(fixture)
    orc = O.Oracle(sc, W, H)
    t0 = time.time()
    orc.render(s, m, rng_mode=1, threads=cores, roulette=roulette)
    secs = time.time() - t0
    return {"value": W * H * s / secs, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"oracle port, {s} spp of {W}x{H} m={m} ({secs:.2f} s)"}


def main_reference(args, cfg, cfg_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    vals, last, secs = [], None, 0.0
    n_runs = max(1, args.warmup + args.steps)
    budget = max(12.0, min(30.0, 170.0 / n_runs))     # the whole --steps K --warmup W run ends within a few minutes
    for i in range(n_runs):
        t0 = time.time()
        last = reference_baseline(cfg, budget)
        if i >= args.warmup:
            vals.append(last["value"]); secs += time.time() - t0
    v = float(np.mean(vals))
    fixture, dae, W, H, spp, m, roulette, desc = cfg
    last["value"] = v
    line = {"impl": "reference", "metric": "BDPT samples/s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(vals)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, None, 1), "rays_per_sample": REF_RAYS_PER_SAMPLE.get(dae),
            "cpu_baseline": last, "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# rays the REFERENCE traces per sample (exact, `-t 1` runs of oracle/_ref/ref_harness).  It traces rays
# whose contribution is already known to be zero (e.g. after an EmissionBSDF), this path does not: Mrays/s of the two
# are only comparable through `mrays_per_s_reference_equivalent` = samples/s x this number.
REF_RAYS_PER_SAMPLE = {"CBspheres_lambertian": 28.2, "CBspheres": 28.6, "CBgems": 82.0, "CBbunny": 27.2,
                       "CBbunny4": 27.2, "CBbunny16": 27.2}


def config_dict(cfg, scene, world, extra=None):
    """The same keys on both arms (the reference arm has no scene object: prims come from the table below)."""
    fixture, dae, W, H, spp, m, roulette, desc = cfg
    prims = {"CBbunny": 28588, "CBbunny4": 114316, "CBbunny16": 457228, "CBspheres": 14, "CBspheres_lambertian": 14, "CBgems": 252}
    d = {"workload": desc, "scene": dae, "prims": int(scene.n_prims) if scene is not None else prims.get(dae), "width": W,
         "height": H, "spp": spp, "max_ray_depth": m, "roulette": bool(roulette), "samples_per_step": W * H * spp}
    if extra:
        d.update(extra)
    return d


# --------------------------------------------------------------------------------- our arm
def algorithmic_bytes(st):
    """Algorithmic HBM bytes per sample of each kernel: 64 B per stored vertex record (written once by
    the walk, read once by k_connect), 8 B of subpath lengths, 48 B per queued shadow ray (written by k_connect,
    read by k_shadow), 16 B per frame-buffer reduction.  V and R_s are counted by the kernels themselves."""
    n = max(1, st["samples"])
    verts = st["rays_closest"] / n        # stored vertices <= closest-hit rays; + 1 light origin record
    shadow = st["rays_shadow"] / n
    return {"k_walk": 64.0 * (verts + 1) + 8, "k_connect": 64.0 * (verts + 1) + 8 + 48.0 * shadow,
            "k_shadow": 48.0 * shadow + 16.0 * shadow}, verts, shadow


def measure_config(api, shard, torch, dist, cfg, args, rank, world, local, stream, steps, warmup, e2e_steps):
    """Renders cfg's frame `steps` times resident and `e2e_steps` times end to end; returns the numbers of the line."""
    fixture, dae, W, H, spp, m, roulette, desc = cfg
    scene = pinned_scene(load_fixture(fixture))   # host inputs in pinned memory (the e2e step copies them up every time)
    first, n_mine = shard.shard_samples(spp, rank, world)
    R = api.Renderer(local, ns_aa=spp, max_ray_depth=m, seed=0, roulette=bool(roulette))
    R.set_stream(stream.cuda_stream)
    R.set_scene(scene)
    R.set_frame_size(W, H)
    shard.join_ranks(R, rank, world)
    samples_per_step = W * H * spp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        R.clear()
        R.render(first_sample=first, n_samples=n_mine)
        shard.reduce_frame(R, world)            # ONE ncclReduce of eye + light, on the same stream

    def step_e2e():
        R.set_scene(scene)          # host buffers -> BVH build -> H2D
        R.set_frame_size(W, H)      # clears the accumulators
        R.render(first_sample=first, n_samples=n_mine)
        shard.reduce_frame(R, world)
        return R.framebuffer(api.BUFFER_SUM) if rank == 0 else None   # D2H of the result

    def timed(fn, k):
        """k steps bracketed by barrier + synchronize; device time by CUDA events on the launching stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(warmup):
        step_resident()
    R.clear()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, _ = timed(step_resident, steps)
    st = R.stats()       # per-kernel times of the last step's first batches; counters of the last step (cleared per step)
    clock_info = clocks.stop() if rank == 0 else None
    out = {"scene": scene, "value": samples_per_step * steps / (ms / 1e3), "ms_per_step": ms / steps, "stats": st,
           "clocks": clock_info, "gpu_launches": int(st["kernel_launches"]) * steps * world, "e2e": None}
    if e2e_steps > 0:
        step_e2e()
        _, e2e_wall = timed(step_e2e, e2e_steps)
        # the e2e region has host work (scene ingest, D2H wait) the device events do not see: wall clock, max over ranks
        h2d = int(scene.prims.nbytes + scene.materials.nbytes + scene.lights.nbytes + scene.camera.nbytes)
        out["e2e"] = {"value": samples_per_step * e2e_steps / (e2e_wall / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": W * H * 3 * 4, "ms_per_step": e2e_wall / e2e_steps,
                      "ms_scene_upload": R.stats()["ms_bvh_build"]}
    R.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return main_reference(args, cfg, args.config)

    import torch
    import torch.distributed as dist
    from bdpt import api, shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()

    fixture, dae, W, H, spp, m, roulette, desc = cfg
    r = measure_config(api, shard, torch, dist, cfg, args, rank, world, local, stream, args.steps, args.warmup, args.steps)

    if rank == 0:
        st, value = r["stats"], r["value"]
        rays_per_sample = (st["rays_closest"] + st["rays_shadow"]) / max(1, st["samples"])  # rank 0's shard
        # roofline of the dominant kernel (per-kernel times cover the first timed_batches batches of the last step)
        kern = {"k_walk": st["ms_walk"], "k_connect": st["ms_connect"], "k_shadow": st["ms_shadow"]}
        top = max(kern, key=kern.get)
        tb = max(1, st["timed_batches"])
        B = st["batch_samples"]
        bytes_per_sample, verts, shadow_per_sample = algorithmic_bytes(st)
        peak, peak_kind = hbm_peak()
        dur_ms = kern[top] / tb
        achieved = bytes_per_sample[top] * B / (dur_ms / 1e3) / 1e9 if dur_ms > 0 else 0.0
        traffic, issue = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get(top)
                if traffic is not None and tj.get("batch_samples") and tj["batch_samples"] != B:
                    traffic = traffic * B / tj["batch_samples"]   # captured at another batch size: scale per sample
            except Exception:
                traffic = None
        ip = os.path.join(ROOT, "profiles", "issue.json")     # ncu issue-side numbers of the committed capture, per kernel
        if os.path.exists(ip):
            try:
                issue = json.load(open(ip))
            except Exception:
                issue = None
        ref_rps = REF_RAYS_PER_SAMPLE.get(dae)
        line = {
            "metric": "BDPT samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg, r["scene"], world, {
                "parallelism": f"sample-sharded x{world}, scene replicated, 1 ncclReduce of eye+light accumulators (bdpt_reduce)",
                "l2_policy": "inputs larger than L2: each batch of %d samples streams ~%.1f GB of vertex records and ~%.1f GB of shadow queue (L2 = 126 MB)"
                             % (B, 64.0 * (verts + 1) * B / 1e9, 48.0 * shadow_per_sample * B / 1e9)}),
            "mrays_per_s": value * rays_per_sample / 1e6, "rays_per_sample": rays_per_sample,
            # the reference traces more rays per sample than this path (it keeps tracing zero-throughput paths): the same
            # samples/s expressed in ITS rays
            "mrays_per_s_reference_equivalent": value * ref_rps / 1e6 if ref_rps else None,
            "kernel_ms_first_batches": {**kern, "batches": tb, "batch_samples": B},
            "render_ms_last_step": st["ms_total"],   # CUDA events around the batch loop inside bdpt_render (rank 0)
            "e2e": r["e2e"],
            "gpu_launches": r["gpu_launches"],
            "clocks": r["clocks"],
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                         "algorithmic_bytes_per_sample": bytes_per_sample,
                         "issue": issue,
                         "note": "issue/latency-bound path (SURVEY.md §8d): the HBM fraction is small by nature; `issue` holds the "
                                 "ncu numbers of the binding roof (issue-slot utilisation, active lanes per instruction, "
                                 "long-scoreboard stalls) from the committed capture under profiles/"},
        }
        if world == 1 and not args.no_extras:
            # the other configs, same run, one GPU: the generic / no-microfacet kernel builds, spheres, glass, m = 16 with
            # roulette and a 457 k-triangle BVH get a driver-run number too
            extras = {}
            for name in EXTRA_CONFIGS:
                try:
                    x = measure_config(api, shard, torch, dist, CONFIGS[name], args, 0, 1, local, stream, 1, 1, 1)
                    xs = x["stats"]
                    extras[name] = {"workload": CONFIGS[name][7], "prims": int(x["scene"].n_prims), "value": x["value"],
                                    "ms_per_step": x["ms_per_step"], "e2e": x["e2e"]["value"], "e2e_ms_per_step": x["e2e"]["ms_per_step"],
                                    "rays_per_sample": (xs["rays_closest"] + xs["rays_shadow"]) / max(1, xs["samples"]),
                                    "kernel_ms": {"k_walk": xs["ms_walk"], "k_connect": xs["ms_connect"], "k_shadow": xs["ms_shadow"],
                                                  "batches": xs["timed_batches"]}}
                except (Exception, SystemExit) as e:
                    extras[name] = {"error": str(e)}
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = reference_baseline(cfg)
            except Exception as e:  # the checker failing must not hide the measurement
                line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
