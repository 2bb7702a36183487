"""Multi-GPU path on real devices: one process per GPU (NCCL), a replica of the matrix on each,
columns sharded over the ranks, one collective for the moments.  Needs >= 2 visible GPUs
(``gpurun --gpus 2``); on a single-GPU box the test is skipped and the same host logic is
covered by ``test_distributed_gloo.py``."""

import os
import socket

import numpy as np
import pytest

import cases
from oracle import bdg_oracle as orc
from util import rel_err

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import types

    import torch
    import torch.distributed as dist

    import bodge_b200 as b

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        api = types.SimpleNamespace(CubicLattice=b.CubicLattice, Hamiltonian=b.Hamiltonian, σ0=b.σ0, σ1=b.σ1,
                                    σ2=b.σ2, σ3=b.σ3, jσ2=b.jσ2, dwave=b.dwave)
        system = cases.dwave_rashba(api, (9, 8, 1))       # every rank assembles its own replica
        assert system.device == rank
        scale = system.spectral_bound()
        per_col = system.chebyshev_moments(40, vectors=11, seed=7, scale=scale)          # all-gather
        trace = system.chebyshev_moments(40, vectors=11, seed=7, scale=scale, summed=True)  # all-reduce
        rho = system.ldos_map([(4, 4, 0), (1, 2, 0), (7, 7, 0)], np.linspace(-0.3, 0.3, 7), moments=400)
        F = system.free_energy(0.2, cuda=True)                                              # exact trace, sharded
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), per_col=per_col, trace=trace, rho=rho, F=F)
    finally:
        dist.destroy_process_group()


def test_column_sharding_two_gpus(gpu_api, tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)

    system = cases.dwave_rashba(gpu_api, (9, 8, 1))   # single-GPU run of the same calls (no process group here)
    scale = system.spectral_bound()
    H = system.matrix("bsr")
    want = orc.cheb_moments(H, orc.rademacher(7, H.shape[0], np.arange(11)), 40, scale)
    single = system.chebyshev_moments(40, vectors=11, seed=7, scale=scale)
    rho1 = system.ldos_map([(4, 4, 0), (1, 2, 0), (7, 7, 0)], np.linspace(-0.3, 0.3, 7), moments=400)
    F1 = system.free_energy(0.2, cuda=True)
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert rel_err(got["per_col"], want) <= 1e-10
        assert rel_err(got["per_col"], single) <= 1e-13           # per-column results do not depend on the sharding
        assert rel_err(got["trace"], want.sum(axis=1)) <= 1e-10
        assert np.allclose(got["rho"], rho1, rtol=1e-10, atol=1e-12)
        assert abs(float(got["F"]) - F1) <= 1e-11 * abs(F1)


def test_c_abi_multi_gpu_entry_point(gpu_api):
    """One process, two GPUs, through the raw C ABI (``bdg_cheb_moments_multi``, NCCL inside): the sharded
    result equals the single-GPU one per column (bit for bit: a column's arithmetic does not depend on which GPU it
    runs on when the panels line up) and the oracle's to 1e-10, for summed moments, per-column moments and probe columns."""
    import ctypes as C

    import bodge_b200 as b
    from bodge_b200 import _native, workloads

    if _native.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    shape = (18, 12, 1)
    packed = workloads.junction(shape)
    replicas = b.Replicas(b.CubicLattice(shape), devices=[0, 1])
    assert replicas.fill(*packed) == 0.0
    single = replicas.replicas[0]
    scale = single.spectral_bound()
    H = single.matrix("bsr")
    want = orc.cheb_moments(H, orc.rademacher(5, H.shape[0], np.arange(16)), 30, scale)
    # raw ABI call, no Python helper in between
    lib = _native.load()
    handles = (C.c_void_p * 2)(*[r._sys._h for r in replicas.replicas])
    mu = np.empty((30, 16))
    rc = lib.bdg_cheb_moments_multi(handles, 2, _native.X0_RADEMACHER, 16, None, C.c_uint64(5), scale, 30, _native.MU_PER_COLUMN,
                                    mu.ctypes.data_as(C.c_void_p))
    assert rc == 0, _native.last_error()
    assert rel_err(mu, want) <= 1e-10
    assert np.array_equal(mu, single.chebyshev_moments(30, vectors=16, seed=5, scale=scale, group=None))  # 8 + 8 columns: whole panels
    # helper: uneven shards (11 = 6 + 5), summed and per column, probe columns
    got = replicas.chebyshev_moments(30, vectors=11, seed=5, scale=scale)
    assert rel_err(got, want[:, :11]) <= 1e-10
    total = replicas.chebyshev_moments(30, vectors=11, seed=5, scale=scale, summed=True)
    assert rel_err(total, want[:, :11].sum(axis=1)) <= 1e-10
    rows = single._probe_rows([(3, 4, 0), (9, 9, 0), (17, 0, 0)])
    probe = replicas.chebyshev_moments(30, rows=rows, scale=scale)
    assert rel_err(probe, orc.cheb_moments(H, orc.probes(H.shape[0], rows), 30, scale)) <= 1e-10
    # the dict API on all replicas at once, then an incremental update that every replica follows
    with replicas as (Hd, Dd):
        Hd[(2, 2, 0), (2, 2, 0)] = 2.5 * gpu_api.σ0
    H2 = single.matrix("bsr")
    scale2 = replicas.spectral_bound()
    again = replicas.chebyshev_moments(30, vectors=16, seed=5, scale=scale2, summed=True)
    assert rel_err(again, orc.cheb_moments(H2, orc.rademacher(5, H2.shape[0], np.arange(16)), 30, scale2).sum(axis=1)) <= 1e-10
    assert lib.bdg_multi_release() == 0
