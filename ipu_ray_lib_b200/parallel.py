"""Ray-data-parallel partitioning of a TraceResult stream over replicas (one replica per GPU).

Mirrors the reference's only inter-device strategy: the stream is cut into batches of
`rays_per_batch` rays and batch i is processed by replica i % R (src/IpuScene.cpp:676-684); the scene is
replicated; replicas never talk to each other; results are gathered per replica at the end.
"""
from __future__ import annotations

import numpy as np


def batch_owner_mask(num_rays: int, rays_per_batch: int, world_size: int, rank: int) -> np.ndarray:
    """Boolean mask over the stream: True where the ray's batch belongs to `rank`."""
    if world_size < 1 or not (0 <= rank < world_size) or rays_per_batch < 1:
        raise ValueError("bad partition arguments")
    batch = np.arange(num_rays, dtype=np.int64) // rays_per_batch
    return (batch % world_size) == rank


def scatter_shards(stream: np.ndarray, rays_per_batch: int, world_size: int):
    """Per-rank contiguous copies of the rays each rank owns."""
    return [np.ascontiguousarray(stream[batch_owner_mask(stream.size, rays_per_batch, world_size, r)])
            for r in range(world_size)]


def merge_shards(shards, num_rays: int, rays_per_batch: int) -> np.ndarray:
    """Inverse of :func:`scatter_shards`: place every rank's results back at their stream positions."""
    world = len(shards)
    out = np.empty(num_rays, dtype=shards[0].dtype)
    for r, s in enumerate(shards):
        out[batch_owner_mask(num_rays, rays_per_batch, world, r)] = s
    return out


def gather_stream(local: np.ndarray, num_rays: int, rays_per_batch: int, dist=None, dst: int = 0):
    """Gather every rank's shard to `dst` with torch.distributed (gloo on CPU, NCCL on GPUs) and merge.

    Returns the merged stream on `dst`, None elsewhere. With dist=None (single process) returns `local`.
    """
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [int(batch_owner_mask(num_rays, rays_per_batch, world, r).sum()) for r in range(world)]
    item = local.dtype.itemsize
    # shards differ by up to one ray batch and NCCL's gather wants equal sizes: every rank sends max(shard) bytes (its
    # own rays, zero-padded) and `dst` keeps the first counts[r] rays of rank r's buffer
    size = max(counts) * item
    mine = torch.zeros(size, dtype=torch.uint8)
    mine[:local.size * item] = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
    use_cuda = dist.get_backend() == "nccl"
    if use_cuda:
        mine = mine.cuda()
    bufs = None
    if rank == dst:
        bufs = [torch.empty(size, dtype=torch.uint8, device=mine.device) for _ in counts]
    dist.gather(mine, bufs, dst=dst)
    if rank != dst:
        return None
    shards = [b.cpu().numpy()[:c * item].view(local.dtype) for b, c in zip(bufs, counts)]
    return merge_shards(shards, num_rays, rays_per_batch)
