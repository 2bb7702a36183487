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


class Replicas:
    """ONE process driving several GPUs through the C ABI's own multi-GPU entry point (``bdg_cheb_moments_multi``):
    a replica of the Hamiltonian per device, the start columns sharded over them, one NCCL collective at the end --
    no torch, no process group.  (With ``torchrun`` -- one process per GPU -- ``Hamiltonian`` shards over the
    initialised ``torch.distributed`` group by itself; this is the same partition for callers that own all GPUs.)

        systems = Replicas(CubicLattice((1000, 1000, 1)), devices=[0, 1, 2, 3])
        systems.fill(*packed)                       # or:  with systems as (H, Δ): ...
        mu = systems.chebyshev_moments(2048, vectors=64, summed=True)
    """

    def __init__(self, lattice, devices):
        from .hamiltonian import Hamiltonian

        if len(set(devices)) != len(devices) or not devices:
            raise ValueError("one replica per device: give a list of distinct device indices")
        self.devices = list(devices)
        self.replicas = [Hamiltonian(lattice, device=d) for d in self.devices]
        self.lattice = lattice

    # the reference's context manager, applied to every replica
    def __enter__(self):
        self._hopp, self._pair = {}, {}
        return self._hopp, self._pair

    def __exit__(self, exc_type, exc_val, exc_tb):
        from .hamiltonian import _pack_entries

        packed = _pack_entries(self.lattice, self._hopp) + _pack_entries(self.lattice, self._pair)
        self.fill(*packed)
        del self._hopp, self._pair

    def fill(self, *packed, **kw):
        return max(r.fill(*packed, **kw) for r in self.replicas)

    def spectral_bound(self):
        return self.replicas[0].spectral_bound()

    def chebyshev_moments(self, moments: int, *, rows=None, vectors=None, seed: int = 1234, scale=None, summed: bool = False):
        from . import _native

        if (rows is None) == (vectors is None):
            raise ValueError("give either rows= (probe columns) or vectors= (random columns)")
        scale = self.spectral_bound() if scale is None else float(scale)
        return _native.cheb_moments_multi([r._sys for r in self.replicas], int(moments), probe_rows=rows,
                                          n_random=0 if vectors is None else int(vectors), seed=seed, scale=scale, summed=summed)
