"""Data parallelism of the denoising path: one process per GPU, independent trajectories.

Every batch element / frame / shift is an independent DDIM trajectory (the UNet has no
cross-sample op; GroupNorm is per sample), so the batch is cut into contiguous per-rank slices
with no collective inside the step loop, and the decoded frames are collected by exactly one
all-gather at the end.  ``torch.distributed`` is the plumbing: NCCL over
NVLink / NVSwitch on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, as-even-as-possible slice [lo, hi) of n trajectories for ``rank``."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int = None, world_size: int = None) -> torch.Tensor:
    """This rank's slice of a batch that every rank holds (e.g. seed-0 latents generated identically
    on all ranks, so that the sharded run reproduces the single-GPU run sample for sample)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi]


def gather_frames(local: torch.Tensor, total: int) -> torch.Tensor:
    """The single collective of the path: all-gather of the per-rank results along the batch axis.
    Ranks may hold unequal slices (``shard_bounds``); results come back in global batch order."""
    rank, w = world()
    if w == 1:
        return local
    counts = [hi - lo for lo, hi in (shard_bounds(total, r, w) for r in range(w))]
    m = max(counts)
    buf = local.new_zeros((m,) + tuple(local.shape[1:]))
    buf[:local.shape[0]] = local
    out = local.new_empty((w * m,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, buf.contiguous())
    parts = [out[r * m: r * m + counts[r]] for r in range(w)]
    return torch.cat(parts, dim=0)
