"""Sample-index sharding of one frame across the GPUs of a box, and the single reduce that ends it.

The reference treats pixels x samples as embarrassingly parallel and funnels every write through
one mutex (reference src/pathtracer/bidirection.cpp:544-551).  Here the scene is replicated, rank r
renders a contiguous block of the frame's sample indices for EVERY pixel (light-image splats can
land on any pixel, so pixels do not shard), each rank accumulates privately, and one sum-reduce of
the two accumulators to rank 0 finishes the frame.  The random streams are keyed on
(seed, pixel, global sample index), so the union of the shards is the same set of samples a single
GPU would have drawn.
"""
from __future__ import annotations


def shard_samples(ns_total: int, rank: int, world: int) -> tuple[int, int]:
    """(first_sample, n_samples) of `rank`: contiguous blocks, remainder spread over the low ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(ns_total, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def join_ranks(renderer, rank: int, world: int) -> None:
    """Makes the renderer a member of the job's NCCL communicator (the library's own: include/bdpt_b200.h).
    torch.distributed is only the courier of rank 0's 128-byte id."""
    import torch
    import torch.distributed as dist
    from . import api
    if world == 1:
        return
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", renderer.device)
        t = torch.zeros(api.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        ident = bytes(t.cpu().numpy().tobytes())
    else:
        box = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ident = box[0]
    renderer.comm_init_rank(world, rank, ident)


def reduce_frame(renderer, world: int, root: int = 0) -> None:
    """The single collective of the path: ONE ncclReduce(sum) of the accumulators, on the renderer's stream."""
    if world > 1:
        renderer.reduce(root)


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can view the library's accumulators in place."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def accumulator_tensor(renderer):
    """torch view (no copy) of the un-normalised accumulators — eye image then light image, W*H*4 floats each
    {r, g, b, unused} — on the renderer's device, for callers that reduce with torch.distributed themselves.
    The view lives on the renderer's device whatever torch's current device is.  Ordering is the caller's job: either
    run the renderer on torch's stream (Renderer.set_stream(torch.cuda.current_stream().cuda_stream)) or synchronise
    between rendering and reducing — the renderer's own stream is a non-blocking one."""
    import torch
    from . import api
    ptr, n = renderer.framebuffer_device(api.BUFFER_SUM)
    t = torch.as_tensor(_CudaArray(ptr, n), device=torch.device("cuda", renderer.device))
    if t.data_ptr() != ptr:
        raise RuntimeError("torch copied the accumulators instead of viewing them (wrong device?)")
    return t


def reduce_to_root(tensors, root: int = 0):
    """Sum-reduce of caller-held tensors (gloo in the CPU tests; the GPU path uses reduce_frame)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
