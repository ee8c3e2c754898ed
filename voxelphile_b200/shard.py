"""Sharding of a chunk batch across GPUs: one process per GPU, chunks
partitioned by position in the coordinate list, no device-side exchange.

A chunk's mesh needs only its own voxels and a one-voxel halo, so a shard's batch is its
owned chunks followed by read-only copies of the neighbouring seam chunks ("halo chunks",
stored but not meshed).  Terrain is a pure function of (seed, world xyz), so the fused
path needs no halo at all.  The only collective is the final host gather of per-chunk
quad counts.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .mesher import neighbour_table


def partition(n: int, rank: int, world: int) -> tuple[int, int]:
    """Balanced contiguous range [lo, hi) of rank among world ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class Shard:
    batch_index: np.ndarray  # global chunk indices of the shard batch: owned first, halo chunks after
    n_owned: int             # chunks [0, n_owned) are meshed; the rest are read-only neighbours
    nbr6: np.ndarray         # int32 [n_owned, 6], indices into the shard batch (-1 = air)


def shard_with_halo(coords: np.ndarray, rank: int, world: int, nbr6: np.ndarray | None = None) -> Shard:
    coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
    n = coords.shape[0]
    if nbr6 is None:
        nbr6 = neighbour_table(coords)
    lo, hi = partition(n, rank, world)
    owned = np.arange(lo, hi, dtype=np.int64)
    links = nbr6[lo:hi]
    outside = np.unique(links[(links >= 0) & ((links < lo) | (links >= hi))])
    batch = np.concatenate([owned, outside.astype(np.int64)])
    local = np.full(n, -1, dtype=np.int32)
    local[batch] = np.arange(batch.shape[0], dtype=np.int32)
    local_nbr = np.where(links >= 0, local[np.clip(links, 0, n - 1)], -1).astype(np.int32)
    return Shard(batch_index=batch, n_owned=hi - lo, nbr6=np.ascontiguousarray(local_nbr))


def gather_counts(counts, n_total: int, rank: int, world: int):
    """Final host gather of per-chunk quad counts (the one collective of the path).

    counts: 1-D int64 torch tensor (CPU for gloo, CUDA for nccl) of this rank's owned chunks.
    Returns the int64 tensor of all n_total counts in global chunk order on every rank.
    """
    import torch
    import torch.distributed as dist

    if world == 1:
        return counts.clone()
    width = max(partition(n_total, r, world)[1] - partition(n_total, r, world)[0] for r in range(world))
    padded = torch.zeros(width, dtype=torch.int64, device=counts.device)
    padded[: counts.numel()] = counts
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    out = []
    for r in range(world):
        lo, hi = partition(n_total, r, world)
        out.append(parts[r][: hi - lo])
    return torch.cat(out)
