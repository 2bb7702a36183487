"""World-size-2 test of the multi-GPU path's host logic on CPU (gloo): columns sharded by
``shard_range``, per-rank moments from the oracle (standing in for the per-GPU CUDA engine),
combined by ``distributed.combine`` -- all-reduce for traces, all-gather for per-column moments."""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import cases
from oracle import bdg_oracle as orc
from util import oracle_assemble


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_cols, out_dir):
    import torch.distributed as dist

    from bodge_b200 import distributed

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rec = cases.dwave_rashba(cases.recorder_api(), (6, 5, 1))
        _, (ptr, idx, dat) = oracle_assemble((6, 5, 1), [rec.packed()])
        H = orc.to_scipy(ptr, idx, dat)
        scale = 1.01 * orc.norm_inf(ptr, idx, dat)
        r, w, group = distributed.resolve("auto")
        assert (r, w) == (rank, world)
        lo, hi = distributed.shard_range(n_cols, r, w)
        x0 = orc.rademacher(7, H.shape[0], np.arange(lo, hi))       # global column ids
        local = orc.cheb_moments(H, x0, 24, scale) if hi > lo else np.zeros((24, 0))
        per_col = distributed.combine(local, False, n_cols, r, w, group)
        trace = distributed.combine(local.sum(axis=1), True, n_cols, r, w, group)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), per_col=per_col, trace=trace)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cols", [5, 1])
def test_column_sharding_world2(tmp_path, n_cols):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_cols, str(tmp_path)), nprocs=world, join=True)
    rec = cases.dwave_rashba(cases.recorder_api(), (6, 5, 1))
    _, (ptr, idx, dat) = oracle_assemble((6, 5, 1), [rec.packed()])
    H = orc.to_scipy(ptr, idx, dat)
    scale = 1.01 * orc.norm_inf(ptr, idx, dat)
    want = orc.cheb_moments(H, orc.rademacher(7, H.shape[0], np.arange(n_cols)), 24, scale)
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(got["per_col"], want)                 # gathered columns are bit-identical
        assert np.allclose(got["trace"], want.sum(axis=1), rtol=1e-13, atol=1e-13)
