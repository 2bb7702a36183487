"""CPU baseline of the Chebyshev step -- TEST/BENCH INFRASTRUCTURE (see bdg_oracle.py header).

Times the reference-side arithmetic of the hot path on the host cores: scipy's ``bsr_matvecs``
(``H @ X`` on a 4x4-block BSR matrix, exactly what the reference's ``matrix("bsr")`` would be
multiplied with) driving ``T_{n+1} = 2 H~ T_n - T_{n-1}``.  scipy's kernel is single-threaded, so
to use every host core the block rows are split into contiguous slabs, one worker process per
slab, with the two vector buffers in shared memory (the update is in place, like on the GPU).

Only ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``) and tests import this.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import time
from multiprocessing import shared_memory

import numpy as np
import scipy.sparse as sp

from . import bdg_oracle as orc


def assemble(shape, packed):
    """Oracle assembly of one packed with-block -> (indptr, indices, data) after eliminate_zeros."""
    indptr, indices = orc.cubic_skeleton(shape)
    data = orc.scatter(indptr, indices, orc.zero_data(indices), *packed)
    return orc.eliminate_zeros(indptr, indices, data)


def _row_slab(indptr, indices, data, r0, r1, n_sites):
    """Block rows [r0, r1) of a BSR matrix as their own (r1-r0)*4 x 4N BSR matrix (views, no copy)."""
    lo, hi = int(indptr[r0]), int(indptr[r1])
    ptr = (indptr[r0 : r1 + 1] - indptr[r0]).astype(np.int32)
    return sp.bsr_matrix((data[lo:hi], indices[lo:hi], ptr), shape=(4 * (r1 - r0), 4 * n_sites), blocksize=(4, 4))


def _worker(slab, r0, r1, shm_names, shape, start, done, stop):
    bufs = [shared_memory.SharedMemory(name=n) for n in shm_names]
    vecs = [np.ndarray(shape, dtype=np.complex128, buffer=b.buf) for b in bufs]
    cur = 1  # vecs[1] = T_n, vecs[0] = T_{n-1} (overwritten in place with T_{n+1})
    while True:
        start.wait()
        if stop.value:
            break
        out = vecs[cur ^ 1]
        out[4 * r0 : 4 * r1] = 2 * (slab @ vecs[cur]) - out[4 * r0 : 4 * r1]
        cur ^= 1
        done.wait()
    for b in bufs:
        b.close()


class ParallelStepper:
    """``step()`` performs one Chebyshev step on ``row_fraction`` of every worker's slab."""

    def __init__(self, indptr, indices, data, scale, x0, n_procs=None, row_fraction=1.0):
        n_sites = len(indptr) - 1
        self.n_procs = n_procs or min(os.cpu_count() or 1, 64)
        self.row_fraction = float(row_fraction)
        k = x0.shape[1]
        self.shape = (4 * n_sites, k)
        nbytes = int(np.prod(self.shape)) * 16
        self._shm = [shared_memory.SharedMemory(create=True, size=nbytes) for _ in range(2)]
        self.vecs = [np.ndarray(self.shape, dtype=np.complex128, buffer=s.buf) for s in self._shm]
        scaled = data / scale
        H = sp.bsr_matrix((scaled, indices, indptr), shape=(4 * n_sites, 4 * n_sites), blocksize=(4, 4))
        self.vecs[0][...] = x0
        self.vecs[1][...] = H @ x0  # T_1
        ctx = mp.get_context("fork")
        self._start = ctx.Barrier(self.n_procs + 1)
        self._done = ctx.Barrier(self.n_procs + 1)
        self._stop = ctx.Value("i", 0)
        bounds = np.linspace(0, n_sites, self.n_procs + 1).astype(int)
        self.rows_per_step = 0
        self._procs = []
        for w in range(self.n_procs):
            r0 = int(bounds[w])
            r1 = r0 + max(1, int(round((int(bounds[w + 1]) - r0) * self.row_fraction)))
            r1 = min(r1, int(bounds[w + 1]))
            self.rows_per_step += r1 - r0
            slab = _row_slab(indptr, indices, scaled, r0, r1, n_sites)
            p = ctx.Process(target=_worker, args=(slab, r0, r1, [s.name for s in self._shm], self.shape,
                                                  self._start, self._done, self._stop), daemon=True)
            p.start()
            self._procs.append(p)
        self.n_sites = n_sites
        self.cur = 1

    def step(self) -> float:
        t0 = time.perf_counter()
        self._start.wait()
        self._done.wait()
        self.cur ^= 1
        return time.perf_counter() - t0

    def current(self) -> np.ndarray:
        return self.vecs[self.cur]

    def close(self):
        if self._procs:
            self._stop.value = 1
            self._start.wait()
            for p in self._procs:
                p.join(timeout=10)
            self._procs = []
        for s in self._shm:
            try:
                s.close()
                s.unlink()
            except FileNotFoundError:
                pass
        self._shm = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def time_steps(indptr, indices, data, scale, x0, steps, warmup, budget_s=60.0, n_procs=None):
    """Run ``warmup + steps`` CPU Chebyshev steps within roughly ``budget_s`` seconds.

    If full steps would not fit, every step processes only a fraction of the block rows (a bounded
    sample of the same workload) and the time is scaled to a full step.  Returns a dict with
    ``ms_per_step`` (scaled to the full workload), ``fraction`` and ``cores``.
    """
    probe = ParallelStepper(indptr, indices, data, scale, x0, n_procs=n_procs, row_fraction=1.0)
    try:
        probe.step()
        t_full = min(probe.step(), probe.step())
    finally:
        probe.close()
    need = t_full * (steps + warmup)
    fraction = 1.0 if need <= budget_s else max(budget_s / need, 1e-3)
    stepper = ParallelStepper(indptr, indices, data, scale, x0, n_procs=n_procs, row_fraction=fraction)
    try:
        for _ in range(warmup):
            stepper.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            stepper.step()
        elapsed = time.perf_counter() - t0
        covered = stepper.rows_per_step / stepper.n_sites
        cores = stepper.n_procs
    finally:
        stepper.close()
    return dict(ms_per_step=1e3 * elapsed / steps / covered, fraction=covered, cores=cores,
                full_step_s=t_full, elapsed_s=elapsed)


def moments_parallel(indptr, indices, data, scale, x0, n_moments, n_procs=None):
    """``mu[n, c] = <x0_c| T_n(H/scale) |x0_c>`` for ``n < n_moments`` by the literal three-term recursion
    (``bdg_oracle.cheb_moments``: same arithmetic, scipy ``bsr_matvecs``), with the block rows split over the host
    cores so that a 10^6-site check takes seconds.  The CHECKER of the full-size parity tests and of ``bench.py``'s
    ``parity_check``."""
    x0 = np.ascontiguousarray(x0, dtype=np.complex128)
    mu = np.empty((n_moments, x0.shape[1]))
    stepper = ParallelStepper(indptr, indices, data, scale, x0, n_procs=n_procs, row_fraction=1.0)
    try:
        mu[0] = np.einsum("ic,ic->c", x0.conj(), x0).real
        if n_moments > 1:
            mu[1] = np.einsum("ic,ic->c", x0.conj(), stepper.current()).real
        for n in range(2, n_moments):
            stepper.step()
            mu[n] = np.einsum("ic,ic->c", x0.conj(), stepper.current()).real
    finally:
        stepper.close()
    return mu
