"""Physics scenarios of the reference's integration tests (``tests/test_physics.py``), as CALLERS of the hot path.

Each scenario builds its system through the dict API of ``api`` (see ``cases.py``: the unmodified reference, ``bodge_b200``
or the recorder in front of the CPU oracle), re-enters ``with`` where the reference's test does, and returns the numbers
that test looks at -- but through ``observe.ldos(system, site, energies)`` / ``observe.free_energy(system, T)``, so the
same code yields

* the golden values (reference ``spsolve`` LDOS and dense ``eigvalsh`` free energy; ``golden/make_physics_golden.py``),
* the oracle's KPM values (``-m "not gpu"``), and
* the CUDA path's (``free_energy(T, cuda=True)`` and ``ldos`` of ``bodge_b200``; ``-m gpu``).

The reference's tests that only look at ``diagonalize()`` (gap scaling, Josephson minigap; out of the hot path) appear here
through the free energy of the same systems instead.  ``check(values)`` holds the inequalities the reference asserts.
"""

from __future__ import annotations

import numpy as np


def gap_existence(api, observe):
    """Normal metal -> s-wave superconductor: the LDOS leaves the gap (tests/test_physics.py:16-67)."""
    lattice = api.CubicLattice((16, 16, 1))
    system = api.Hamiltonian(lattice)
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = -1.5 * api.σ0
        for i, j in lattice.bonds():
            H[i, j] = -1.0 * api.σ0
    gap = 0.5
    site = (8, 8, 0)
    w = gap * np.array([-1.2, -0.8, 0.8, 1.2])
    out = {"normal": observe.ldos(system, site, w)}
    with system as (H, D):
        for i in lattice.sites():
            D[i, i] = gap * api.jσ2
    out["superconducting"] = observe.ldos(system, site, w)
    return out


def check_gap_existence(v):
    n, s = v["normal"], v["superconducting"]
    assert s[1] < n[1] and s[2] < n[2]   # inside the gap
    assert s[0] > n[0] and s[3] > n[3]   # coherence peaks outside


def gap_sweep(api, observe):
    """One chain, the order parameter stepped up by re-entering ``with`` (tests/test_physics.py:70-112); the free
    energy at a low temperature stands in for the lowest eigenvalue."""
    lattice = api.CubicLattice((32, 1, 1))
    system = api.Hamiltonian(lattice)
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = -1.5 * api.σ0
        for i, j in lattice.bonds():
            H[i, j] = -1.0 * api.σ0
    F = []
    for gap in [0.0, 0.01, 0.03, 0.1, 0.3, 1.0]:
        with system as (H, D):
            for i in lattice.sites():
                D[i, i] = gap * api.jσ2
        F.append(observe.free_energy(system, 0.02))
    return {"F": np.array(F)}


def check_gap_sweep(v):
    assert np.all(np.diff(v["F"]) < 0)   # condensation energy grows with the gap


def spin_valve(api, observe):
    """F / S / F chain, parallel against antiparallel magnets; only the right magnet is rewritten
    (tests/test_physics.py:175-228)."""
    lattice = api.CubicLattice((128, 1, 1))
    system = api.Hamiltonian(lattice)
    t, gap, m, T = 1.0, 0.3, 0.7, 0.001
    with system as (H, D):
        for i, j in lattice.bonds():
            H[i, j] = -t * api.σ0
        for i in lattice.sites():
            if i[0] < 32 or i[0] >= 96:
                H[i, i] = -m * api.σ3
            else:
                D[i, i] = -gap * api.jσ2
    parallel = observe.free_energy(system, T)
    with system as (H, D):
        for i in lattice.sites():
            if i[0] >= 96:
                H[i, i] = +m * api.σ3
    return {"F": np.array([parallel, observe.free_energy(system, T)])}


def check_spin_valve(v):
    assert v["F"][1] < v["F"][0]


def odd_frequency(api, observe):
    """Zero-energy LDOS of an s-wave chain without and with an exchange field (tests/test_physics.py:231-269)."""
    lattice = api.CubicLattice((128, 1, 1))
    system = api.Hamiltonian(lattice)
    gap = 0.3
    site, w = (63, 0, 0), [0.0, 0.05 * gap]
    with system as (H, D):
        for i, j in lattice.bonds():
            H[i, j] = -1.0 * api.σ0
        for i in lattice.sites():
            D[i, i] = -gap * api.jσ2
    out = {"plain": observe.ldos(system, site, w)}
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = -0.5 * gap * api.σ2
    out["magnetic"] = observe.ldos(system, site, w)
    return out


def check_odd_frequency(v):
    assert v["plain"][0] >= 0 and v["magnetic"][0] >= v["plain"][0]


def energy_temperature(api, observe):
    """Free energy of a 2-D metal against temperature (tests/test_physics.py:272-297)."""
    lattice = api.CubicLattice((10, 10, 1))
    system = api.Hamiltonian(lattice)
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = -2.0 * api.σ0
        for i, j in lattice.bonds():
            H[i, j] = -1.0 * api.σ0
    return {"F": np.array([observe.free_energy(system, T) for T in [0.01, 0.1, 0.5, 1.0]])}


def check_energy_temperature(v):
    assert np.all(np.diff(v["F"]) < 0)


def pwave_edges(api, observe):
    """p_x-wave superconductor: the gap closes at edges perpendicular to x (tests/test_physics.py:300-339)."""
    lattice = api.CubicLattice((31, 31, 1))
    system = api.Hamiltonian(lattice)
    gap = 0.1
    sp = api.pwave("e_z * p_x")
    with system as (H, D):
        for i, j in lattice.bonds():
            H[i, j] = -1.0 * api.σ0
            D[i, j] = -gap * sp(i, j)
    w = [0.0, gap / 4]
    sites = [(15, 15, 0), (15, 0, 0), (0, 15, 0), (0, 0, 0)]
    return {"rho0": np.array([observe.ldos(system, s, w)[0] for s in sites])}


def check_pwave_edges(v):
    bulk, y_edge, x_edge, corner = v["rho0"]
    assert x_edge > bulk and x_edge > y_edge and corner > bulk and corner > y_edge


def josephson_phase(api, observe):
    """S / N / S chain against the phase difference (tests/test_physics.py:342-390): free energy instead of the lowest
    eigenvalue; F(φ) = F(2π − φ), and the junction's ground state is φ = 0."""
    lattice = api.CubicLattice((128, 1, 1))
    gap = 3.0
    F = []
    for phi in np.pi * np.array([0.0, 0.5, 1.0, 1.5, 2.0]):
        system = api.Hamiltonian(lattice)
        with system as (H, D):
            for i in lattice.sites():
                if i[0] < 32:
                    D[i, i] = -gap * api.jσ2 * np.exp(-0.5j * phi)
                if i[0] >= 96:
                    D[i, i] = -gap * api.jσ2 * np.exp(+0.5j * phi)
            for i, j in lattice.bonds():
                H[i, j] = -1.0 * api.σ0
        F.append(observe.free_energy(system, 0.05))
    return {"F": np.array(F)}


def check_josephson_phase(v):
    F = v["F"]
    assert F[0] < F[1] < F[2]
    assert np.allclose(F[0], F[4], rtol=1e-9, atol=0) and np.allclose(F[1], F[3], rtol=1e-9, atol=0)


SCENARIOS = {
    "gap_existence": (gap_existence, check_gap_existence),
    "gap_sweep": (gap_sweep, check_gap_sweep),
    "spin_valve": (spin_valve, check_spin_valve),
    "odd_frequency": (odd_frequency, check_odd_frequency),
    "energy_temperature": (energy_temperature, check_energy_temperature),
    "pwave_edges": (pwave_edges, check_pwave_edges),
    "josephson_phase": (josephson_phase, check_josephson_phase),
}
