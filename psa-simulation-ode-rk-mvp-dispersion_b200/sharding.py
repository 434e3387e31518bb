"""Multi-GPU plumbing of the sweeps: scan points are independent, so a sweep is partitioned by
contiguous index range across ranks (one process per GPU) and the only communication is the final
gather of the per-point result maps (torch.distributed: NCCL over NVLink on the GPU box, gloo in
the CPU tests).  No collective touches the integration itself.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """[start, stop) of rank's contiguous share of n items; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(n), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_axis(axis, world: int, rank: int) -> np.ndarray:
    """Rank's slice of a sweep axis (e.g. the pump-wavelength rows of the 2-D sweep)."""
    a = np.asarray(axis)
    lo, hi = shard_range(a.shape[0], world, rank)
    return np.ascontiguousarray(a[lo:hi])


def gather_rows(local, n_rows: int, dist, world: int, rank: int):
    """All-gather row blocks of unequal height into the full [n_rows, ...] map (every rank gets
    it).  `local` is a torch tensor on the backend's device; padding rows are dropped."""
    import torch
    if world == 1:
        return local
    tall = max(shard_range(n_rows, world, r)[1] - shard_range(n_rows, world, r)[0] for r in range(world))
    padded = torch.zeros((tall,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * tall,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    pieces = []
    for r in range(world):
        lo, hi = shard_range(n_rows, world, r)
        pieces.append(out[r * tall: r * tall + (hi - lo)])
    return torch.cat(pieces, dim=0)
