"""Shared test systems, written against the *reference's public API only*.

Every builder takes ``api`` -- any namespace exposing ``CubicLattice``, ``Hamiltonian``,
``σ0..σ3``, ``jσ2``, ``dwave`` -- so the very same code runs against

* the unmodified reference (``tests/golden/make_golden.py``, build container only),
* ``bodge_b200`` (the ``-m gpu`` parity tests), and
* ``Recorder`` below (captures the dict entries as packed arrays for the CPU oracle).

Models follow SURVEY.md §8(d); random systems use a seeded ``default_rng`` so the golden
fixtures are reproducible.
"""

from __future__ import annotations

import types

import numpy as np


# --------------------------------------------------------------------------------------
# Model builders (dict API)
# --------------------------------------------------------------------------------------
def readme_swave(api, shape, mu=-3.0, m=0.05, ds=0.10, t=1.0):
    """README model (reference README.md:73-86): C1 = (40,40,1), C2 = (100,100,1)."""
    lattice = api.CubicLattice(shape)
    system = api.Hamiltonian(lattice)
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = -mu * api.σ0 - m * api.σ3
            D[i, i] = -ds * api.jσ2
        for i, j in lattice.bonds():
            H[i, j] = -t * api.σ0
    return system


def dwave_rashba(api, shape, mu=-0.5, alpha=0.2, dd=0.1, t=1.0):
    """C3: d-wave + Rashba spin-orbit coupling."""
    lattice = api.CubicLattice(shape)
    system = api.Hamiltonian(lattice)
    sd = api.dwave()
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = -mu * api.σ0
        for i, j in lattice.bonds():
            dx, dy = j[0] - i[0], j[1] - i[1]
            H[i, j] = -t * api.σ0 + 1j * alpha * (dy * api.σ1 - dx * api.σ2)
            D[i, j] = -dd * sd(i, j)
    return system


def swave_3d(api, shape, mu=-3.0, ds=0.1, t=1.0):
    """C4 analogue: README on-site terms (m = 0) on a 3-D lattice."""
    return readme_swave(api, shape, mu=mu, m=0.0, ds=ds, t=t)


def junction(api, shape, mu=-3.0, d0=0.2, phi=np.pi / 2, m=0.3, t=1.0):
    """C5 analogue: S / altermagnet / S Josephson junction along x."""
    lattice = api.CubicLattice(shape)
    system = api.Hamiltonian(lattice)
    Lx = shape[0]
    x1, x2 = Lx // 3, Lx - Lx // 3  # 333 / 667 at Lx = 1000
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = -mu * api.σ0
            if i[0] < x1:
                D[i, i] = -d0 * api.jσ2 * np.exp(-0.5j * phi)
            elif i[0] >= x2:
                D[i, i] = -d0 * api.jσ2 * np.exp(+0.5j * phi)
        for i, j in lattice.bonds():
            mid = (x1 <= i[0] < x2) and (x1 <= j[0] < x2)
            if not mid:
                H[i, j] = -t * api.σ0
            elif i[0] != j[0]:
                H[i, j] = -t * api.σ0 - m * api.σ3
            elif i[1] != j[1]:
                H[i, j] = -t * api.σ0 + m * api.σ3
            else:
                H[i, j] = -t * api.σ0
    return system


def random_periodic(api, shape, seed):
    """Dense-ish random Hermitian system incl. periodic edges
    (after the reference's tests/test_hamiltonian.py:17-51, but seeded)."""
    rng = np.random.default_rng(seed)
    r = rng.random
    lattice = api.CubicLattice(shape)
    system = api.Hamiltonian(lattice)

    def spin4():
        return r() * api.σ0 + r() * api.σ1 + r() * api.σ2 + r() * api.σ3

    def triplet():
        return (r() * api.σ1 + r() * api.σ2 + r() * api.σ3) @ api.jσ2

    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = spin4()
            D[i, i] = triplet()
        # A bond/edge pair can coincide when an axis has length 2; the dict then keeps
        # the last assignment, for both directions alike, so the result stays Hermitian.
        for i, j in lattice.bonds():
            hop = spin4()
            H[i, j] = hop
            H[j, i] = hop
            D[i, j] = triplet()
        for i, j in lattice.edges():
            if i == j:
                continue
            hop = spin4()
            H[i, j] = hop
            H[j, i] = hop
            D[i, j] = triplet()
    return system


def kat_export(api, shape=(3, 5, 7)):
    """Known-answer system of the reference's tests/test_hamiltonian.py:68-93."""
    lattice = api.CubicLattice(shape)
    system = api.Hamiltonian(lattice)
    with system as (H, D):
        for i, j in lattice:
            H[i, j] = 3 * api.σ0 - 4 * api.σ2
            D[i, j] = 2 * api.σ3 + 5 * api.σ2
    return system


def snf_free_energy(api, shape=(10, 7, 3)):
    """S/N/F system of the reference's tests/test_hamiltonian.py:431-444."""
    lattice = api.CubicLattice(shape)
    system = api.Hamiltonian(lattice)
    with system as (H, D):
        for i in lattice.sites():
            if i[0] <= 3:
                H[i, i] = -0.5 * api.σ0
                D[i, i] = -1.0 * api.jσ2
            if i[0] >= 7:
                H[i, i] = +0.5 * api.σ0 + 1.5 * api.σ3
        for i, j in lattice.bonds():
            H[i, j] = -1 * api.σ0
    return system


BUILDERS = {
    "readme": readme_swave,
    "dwave_rashba": dwave_rashba,
    "swave_3d": swave_3d,
    "junction": junction,
    "kat": kat_export,
    "snf": snf_free_energy,
}


# --------------------------------------------------------------------------------------
# Recorder: the dict API captured as packed arrays (for the CPU oracle)
# --------------------------------------------------------------------------------------
class _RecordingHamiltonian:
    """Context manager with the reference's ``with system as (H, Δ)`` protocol that only
    records; ``packed()`` returns what the last ``with`` block wrote."""

    def __init__(self, lattice):
        self.lattice = lattice
        self.blocks = []  # one (hopp, pair) dict pair per with-block

    def __enter__(self):
        self._hopp, self._pair = {}, {}
        return self._hopp, self._pair

    def __exit__(self, *exc):
        self.blocks.append((self._hopp, self._pair))

    def packed(self, which=-1):
        hopp, pair = self.blocks[which]
        return pack_dicts(self.lattice, hopp, pair)


def pack_dicts(lattice, hopp, pair):
    def one(d):
        if not d:
            z = np.zeros(0, dtype=np.int64)
            return z, z.copy(), np.zeros((0, 2, 2), dtype=np.complex128)
        i = np.array([lattice[k[0]] for k in d], dtype=np.int64)
        j = np.array([lattice[k[1]] for k in d], dtype=np.int64)
        v = np.stack([np.asarray(v, dtype=np.complex128) for v in d.values()])
        return i, j, v

    return one(hopp) + one(pair)


def recorder_api():
    """Namespace for the builders above that records instead of assembling."""
    from bodge_b200 import common, helpers, lattice

    return types.SimpleNamespace(
        CubicLattice=lattice.CubicLattice,
        Hamiltonian=_RecordingHamiltonian,
        σ0=common.σ0, σ1=common.σ1, σ2=common.σ2, σ3=common.σ3, jσ2=common.jσ2,
        dwave=helpers.dwave, pwave=helpers.pwave,
    )
