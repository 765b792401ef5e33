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


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can view the library's accumulators in place."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def accumulator_tensors(renderer):
    """torch views (no copy) of the un-normalised eye and light accumulators on the renderer's device."""
    import torch
    from . import api
    out = []
    for which in (api.BUFFER_EYE, api.BUFFER_LIGHT):
        ptr, n = renderer.framebuffer_device(which)
        out.append(torch.as_tensor(_CudaArray(ptr, n), device="cuda"))
    return out


def reduce_to_root(tensors, root: int = 0):
    """One sum-reduce per accumulator (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
