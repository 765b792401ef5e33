"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : the N-way sample-sharded frame, reduced with NCCL,
equals the single-GPU frame (same seed) up to fp32 summation order."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))
import torch, torch.distributed as dist
from bdpt import api, flatscene as fs, shard

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = fs.load(os.path.join(ROOT, "tests", "golden", "scenes", "CBspheres.fscene.gz"))
W, H, ns, m = 480, 360, 16, 5
R = api.Renderer(local, ns_aa=ns, max_ray_depth=m, seed=4)
R.set_stream(torch.cuda.current_stream().cuda_stream)
R.set_scene(sc); R.set_frame_size(W, H)
first, n = shard.shard_samples(ns, rank, world)
R.render(first_sample=first, n_samples=n)
shard.join_ranks(R, rank, world)
shard.reduce_frame(R, world)
torch.cuda.synchronize()
if rank == 0:
    got = [R.framebuffer(i) for i in range(3)]
    S = api.Renderer(local, ns_aa=ns, max_ray_depth=m, seed=4); S.set_scene(sc); S.set_frame_size(W, H); S.render()
    ref = [S.framebuffer(i) for i in range(3)]
    for g, r in zip(got, ref):
        assert np.allclose(g, r, rtol=2e-5, atol=1e-6), np.abs(g - r).max()
    print(f"multi_gpu_check OK: {world} ranks, shard of rank 0 = samples [{first},{first + n}), max |diff| = {max(np.abs(g - r).max() for g, r in zip(got, ref)):.2e}")
dist.destroy_process_group()
