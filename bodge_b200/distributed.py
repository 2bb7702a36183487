"""Column sharding across GPUs: one process per GPU, a replica of the matrix on each, the start
vectors (probe sites or stochastic-trace columns) split contiguously over the ranks, and ONE
collective at the end that combines the Chebyshev moments.

The recursion itself needs no communication.  ``torch.distributed`` provides the plumbing:
NCCL over NVLink when the moments live on the GPU, gloo for the CPU-side tests.
"""

from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first ``n_items % world`` ranks get one extra item."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def resolve(group="auto"):
    """``(rank, world, group)`` of the process group to shard over; ``(0, 1, None)`` when
    ``torch.distributed`` is not in use (``group=None`` forces single-process behaviour)."""
    if group is None:
        return 0, 1, None
    try:
        import torch.distributed as dist
    except ModuleNotFoundError:
        return 0, 1, None
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1, None
    pg = None if group == "auto" else group
    return dist.get_rank(pg), dist.get_world_size(pg), pg


def combine(local: np.ndarray, summed: bool, n_total: int, rank: int, world: int, group=None, device: int = 0):
    """Combine per-rank moments: all-reduce(SUM) of ``[n_moments]`` traces, or all-gather of the
    ``[n_moments, n_local]`` column blocks into ``[n_moments, n_total]`` (rank order = column order)."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist

    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", device) if on_gpu else torch.device("cpu")
    if summed:
        t = torch.as_tensor(np.ascontiguousarray(local), dtype=torch.float64).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()
    n_moments = local.shape[0]
    widest = shard_range(n_total, 0, world)[1]
    padded = np.zeros((n_moments, widest))
    padded[:, : local.shape[1]] = local
    mine = torch.as_tensor(padded, dtype=torch.float64).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.empty((n_moments, n_total))
    for r, part in enumerate(parts):
        lo, hi = shard_range(n_total, r, world)
        out[:, lo:hi] = part.cpu().numpy()[:, : hi - lo]
    return out
