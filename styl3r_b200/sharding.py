"""Multi-GPU partitioning of the hot path: one process per GPU, units = scenes (each with its target views), no
data-path collective.  The only optional exchange is a gather of finished images to rank 0."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def shard_range(n_units: int, world_size: int, rank: int) -> range:
    """Contiguous, balanced split: the first (n_units % world_size) ranks get one extra unit."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(max(n_units, 0), world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_scene_views(n_scenes: int, views_per_scene: int, world_size: int, rank: int):
    """Units are whole scenes while there are at least `world_size` of them (the Gaussians stay on one GPU);
    with fewer scenes than ranks the (scene, view) pairs of a scene are split instead (the 64 B/Gaussian scene
    buffer is then needed on several ranks — an NVLink broadcast in a real deployment).
    Returns a list of (scene, [views])."""
    if n_scenes >= world_size:
        return [(s, list(range(views_per_scene))) for s in shard_range(n_scenes, world_size, rank)]
    out = {}
    for u in shard_range(n_scenes * views_per_scene, world_size, rank):
        out.setdefault(u // views_per_scene, []).append(u % views_per_scene)
    return sorted(out.items())


def gather_images(local: torch.Tensor, counts: Sequence[int], group=None, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank image stacks [n_i, 3, H, W] (n_i = counts[rank]) to `dst`; returns the concatenation on dst,
    None elsewhere.  Works with NCCL (CUDA tensors) and gloo (CPU tensors, used by the CPU tests)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nmax = max(counts) if len(counts) else 0
    pad = local.new_zeros((nmax, *local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
