#!/usr/bin/env python
"""bench.py — BDPT samples/s (and Mrays/s) of the B200 path on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--impl reference]

Workload (config.workload): BASELINE.json configs[2], "CBlucy 1024x768, 1024 spp, m=5 — large-mesh BVH
traversal throughput", the configuration the north_star's target is quoted on.  dae/sky/CBlucy.dae is
not in the reference checkout, so the documented stand-in is used: CBbunny (28,588
primitives, the largest Cornell-box mesh present).  One "step" = one whole frame:
1024x768 pixels x 1024 spp = 805,306,368 BDPT samples (1 sample = 1 eye subpath + 1 light subpath +
all (s,t) connections).  With N GPUs the frame's sample indices are sharded N ways (strong scaling: the
frame is fixed) and the two fp32 accumulators are sum-reduced to rank 0 with NCCL inside the step.

  value      whole-job samples/s, scene resident in HBM, timed with CUDA events on the launching
             stream, max over ranks (includes the NCCL reduce for N > 1)
  e2e        same metric through the C-ABI from HOST buffers: bdpt_upload_scene (host BVH build + H2D)
             + bdpt_set_frame + bdpt_render + reduce + bdpt_read_framebuffer (D2H) every step
  roofline   dominant kernel: algorithmic HBM bytes per launch / its measured duration vs the measured
             HBM peak.  This path is NOT HBM-bound (it is SM-issue / L1-L2 latency bound);
             the fraction says how far from the memory roof it runs, profiles/ holds the issue-side numbers.
  cpu_baseline   the reference's own binary (oracle/_ref/pathtracer_ref, kind "reference") on all host
             cores, on a bounded sample of the same workload; N = 1, rank 0 only
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
    "config3": ("CBgems", "CBgems", 1024, 768, 1024, 16, 1, "configs[3] CBgems 1024x768 1024spp m=16 roulette"),
    "config4": ("CBbunny", "CBbunny", 1920, 1080, 4096, 8, 0, "configs[4] CBlucy 1920x1080 4096spp m=8 on the CBbunny stand-in"),
}
DEFAULT_CONFIG = "config2"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def load_fixture(fixture):
    """Committed flat-scene fixture, or `dae:<name>`: a staged .dae read by the C++ loader (libbdpt_host.so)."""
    from bdpt import flatscene as fs
    if fixture.startswith("dae:"):
        from bdpt import host
        path = os.path.join(ROOT, "oracle", "_ref", "scenes", fixture[4:] + ".dae")
        if not os.path.exists(path):
            raise SystemExit(f"{path} is missing: run `python scripts/make_standin.py` where /root/reference is mounted")
        return host.load_dae(path)
    return fs.load(os.path.join(ROOT, "tests", "golden", "scenes", fixture + ".fscene.gz"))


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


def cpu_sample_spp(W, H):
    # bounded sample: about 10-30 s of CPU work at ~1e4 samples/s/core; at least 1 spp
    cores = os.cpu_count() or 1
    budget = 15.0 * 1.0e4 * cores
    return max(1, min(4, int(budget // (W * H))))


def reference_baseline(cfg):
    """cpu_baseline object: the reference binary if it was built here, else the oracle port."""
    from oracle import oracle as O
    fixture, dae, W, H, spp, m, roulette, _ = cfg
    cores = os.cpu_count() or 1
    s = cpu_sample_spp(W, H)
    if os.path.exists(O.REF_CLI) and os.path.exists(O.ref_scene(dae)):
        v, secs = run_reference_cli(dae, W, H, s, m, cores)
        return {"value": v, "unit": "samples/s", "cores": cores, "kind": "reference",
                "sample": f"pathtracer_ref -t {cores} -s {s} -m {m} -r {W} {H} {dae}.dae ({W * H * s} samples, {secs:.2f} s; "
                          f"time is linear in spp, SURVEY.md §6)"}
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
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = reference_baseline(cfg)
        if i >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals))
    fixture, dae, W, H, spp, m, roulette, desc = cfg
    last["value"] = v
    line = {"impl": "reference", "metric": "BDPT samples/s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * W * H * cpu_sample_spp(W, H) / v,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "scene": fixture, "width": W, "height": H, "spp": spp, "max_ray_depth": m},
            "cpu_baseline": last, "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return main_reference(args, cfg, args.config)

    import torch
    import torch.distributed as dist
    from bdpt import api, flatscene as fs, shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    fixture, dae, W, H, spp, m, roulette, desc = cfg
    scene = load_fixture(fixture)
    first, n_mine = shard.shard_samples(spp, rank, world)

    R = api.Renderer(local, ns_aa=spp, max_ray_depth=m, seed=0, roulette=bool(roulette))
    stream = torch.cuda.current_stream()
    R.set_stream(stream.cuda_stream)
    R.set_scene(scene)
    R.set_frame_size(W, H)
    acc = shard.accumulator_tensors(R)
    samples_per_step = W * H * spp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        R.clear()
        R.render(first_sample=first, n_samples=n_mine)
        shard.reduce_to_root(acc)

    def step_e2e():
        R.set_scene(scene)          # host buffers -> BVH build -> H2D
        R.set_frame_size(W, H)      # clears the accumulators
        R.render(first_sample=first, n_samples=n_mine)
        shard.reduce_to_root(shard.accumulator_tensors(R))
        return R.framebuffer(api.BUFFER_SUM) if rank == 0 else None   # D2H of the result

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time by CUDA events on the launching stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(args.warmup):
        step_resident()
    R.clear()
    launches0 = R.stats()["kernel_launches"]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, wall_ms = timed(step_resident, args.steps)
    st = R.stats()       # per-kernel times of the last step's first batches; counters of the last step (cleared per step)
    clock_info = clocks.stop() if rank == 0 else None
    value = samples_per_step * args.steps / (ms / 1e3)
    launches = (st["kernel_launches"] + 0) * 1  # launches of the last step (stats are cleared each step)
    gpu_launches = int(launches) * args.steps * world   # every rank launches the same kernels

    # end to end through the C-ABI from host buffers
    for _ in range(1):
        step_e2e()
    e2e_ms, e2e_wall = timed(step_e2e, args.steps)
    # the e2e region has host work (BVH build, D2H wait) the device events do not see: use the wall clock
    e2e_value = samples_per_step * args.steps / (e2e_wall / 1e3)
    h2d = int(scene.prims.nbytes // 152 * (48 + 48 + 80) + st["bvh_nodes"] * 80 + scene.materials.shape[0] * 80 + scene.lights.shape[0] * 80)
    d2h = W * H * 3 * 4 if True else 0

    if rank == 0:
        rays_per_sample = (st["rays_closest"] + st["rays_shadow"]) / max(1, st["samples"])  # rank 0's shard
        # roofline of the dominant kernel (per-kernel times cover the first timed_batches batches of the last step)
        kern = {"k_walk": st["ms_walk"], "k_connect": st["ms_connect"], "k_shadow": st["ms_shadow"]}
        top = max(kern, key=kern.get)
        tb = max(1, st["timed_batches"])
        B = st["batch_samples"]
        shadow_per_sample = st["rays_shadow"] / max(1, st["samples"])
        closest_per_sample = st["rays_closest"] / max(1, st["samples"])
        # algorithmic HBM bytes per sample and kernel: 64 B per stored vertex, 48 B per queued shadow ray,
        # 12 B per frame-buffer add
        verts = max(0.0, closest_per_sample)  # stored vertices <= closest-hit rays (+1 light origin)
        bytes_per_sample = {
            "k_walk": 64.0 * (verts + 1) + 8,
            "k_connect": 64.0 * (verts + 1) + 8 + 48.0 * shadow_per_sample,
            "k_shadow": 48.0 * shadow_per_sample + 12.0 * shadow_per_sample,
        }
        peak, peak_kind = hbm_peak()
        dur_ms = kern[top] / tb
        achieved = bytes_per_sample[top] * B / (dur_ms / 1e3) / 1e9 if dur_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get(top)
                if traffic is not None and tj.get("batch_samples") and tj["batch_samples"] != B:
                    traffic = traffic * B / tj["batch_samples"]   # captured at another batch size: scale per sample
            except Exception:
                traffic = None
        line = {
            "metric": "BDPT samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "scene": fixture, "prims": int(scene.n_prims), "width": W, "height": H, "spp": spp,
                       "max_ray_depth": m, "roulette": bool(roulette), "samples_per_step": samples_per_step,
                       "parallelism": f"sample-sharded x{world}, scene replicated, 1 NCCL reduce per accumulator",
                       "l2_policy": "inputs larger than L2: each batch of %d samples streams ~%.1f GB of vertex planes and ~%.1f GB of shadow queue (L2 = 126 MB)"
                                    % (B, 64.0 * (verts + 1) * B / 1e9, 48.0 * shadow_per_sample * B / 1e9)},
            "mrays_per_s": value * rays_per_sample / 1e6, "rays_per_sample": rays_per_sample,
            "kernel_ms_first_batches": {**kern, "batches": tb, "batch_samples": B},
            "render_ms_last_step": st["ms_total"],   # CUDA events around the batch loop inside bdpt_render (rank 0)
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_wall / args.steps},
            "gpu_launches": gpu_launches,
            "clocks": clock_info,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                         "note": "latency/issue-bound path (SURVEY.md §8d): the HBM fraction is expected to be small; see profiles/"},
        }
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
