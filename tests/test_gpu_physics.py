"""The reference's physics integration scenarios (``tests/test_physics.py`` there, ``physics_cases.py`` here) through the
CUDA path: ``with`` blocks re-entered on one handle, ``ldos()`` and ``free_energy(T, cuda=True)`` -- compared with the
numbers the UNMODIFIED reference gives for the same scripts (``golden/physics.npz``: ``spsolve`` LDOS, dense ``eigvalsh``
free energy) to 1e-10, plus the inequalities the reference's tests assert.  CPU twin: ``test_physics_oracle.py``."""

import os
import types
import warnings

import numpy as np
import pytest

import physics_cases

pytestmark = pytest.mark.gpu

TOL = 1e-10
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "physics.npz")

CUDA = types.SimpleNamespace(
    ldos=lambda system, site, energies: system.ldos(site, energies),
    free_energy=lambda system, T: system.free_energy(T, cuda=True),
)


@pytest.fixture(scope="module")
def physics():
    return dict(np.load(GOLDEN))


@pytest.mark.parametrize("name", list(physics_cases.SCENARIOS))
def test_scenario_matches_the_reference(gpu_api, physics, name):
    import bodge_b200 as b

    scenario, check = physics_cases.SCENARIOS[name]
    with warnings.catch_warnings():
        warnings.simplefilter("error", b.hamiltonian.AccuracyWarning)  # default series lengths must reach their tolerance here
        values = scenario(gpu_api, CUDA)
    for key, got in values.items():
        want = physics[f"{name}/{key}"]
        err = np.max(np.abs(np.asarray(got) - want)) / np.max(np.abs(want))
        assert err <= TOL, (name, key, err, got, want)
    # (the spin valve's two configurations differ by 2.3e-11 of F: the exact-trace expansion resolves that too -- the
    # oracle's KPM agrees with the reference to 8e-16 there)
    check(values)


def test_spin_valve_update_is_patched_in_place(gpu_api):
    """The spin valve rewrites the right magnet only (32 of 128 on-site blocks, tests/test_physics.py:219-224): that
    scatter is patched into the kernel-native copies instead of rebuilding them.  (The gap sweep rewrites EVERY on-site
    block, more than a quarter of all blocks, where the streaming rebuild is the cheaper of the two and is taken:
    ``bdg_scatter``)"""
    seen = []

    def free_energy(system, T):
        F = system.free_energy(T, cuda=True)
        seen.append(system._sys.stats())
        return F

    physics_cases.spin_valve(gpu_api, types.SimpleNamespace(ldos=None, free_energy=free_energy))
    assert seen[1]["native_builds"] == seen[0]["native_builds"] == 1, seen
    assert seen[1]["patched_scatters"] == seen[0]["patched_scatters"] + 1, seen


def test_observables_before_anything_is_set(gpu_api):
    """A Hamiltonian with no entries (H = 0): the reference returns F = 0 (no positive eigenvalue) and the LDOS of the bare
    resolvent (ε + iΓ)^-1; so does the CUDA path's front end, without a spectral interval to expand on."""
    system = gpu_api.Hamiltonian(gpu_api.CubicLattice((6, 5, 1)))
    assert system.spectral_bound() == 0.0
    assert system.free_energy(0.1, cuda=True) == 0.0 == system.free_energy(0.1)
    rho = system.ldos((2, 2, 0), [0.0, 0.1])
    assert np.allclose(rho, [2 / (np.pi * 0.1), 2 * 0.1 / (np.pi * 0.02)], rtol=1e-15, atol=0)
    # ... and once something is set the expansion takes over
    lattice = system.lattice
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = -1.5 * gpu_api.σ0
        for i, j in lattice.bonds():
            H[i, j] = -1.0 * gpu_api.σ0
    F = system.free_energy(0.1)
    assert system.spectral_bound() > 0 and abs(system.free_energy(0.1, cuda=True) - F) <= 1e-10 * abs(F)


def test_sites_left_out_of_the_geometry(gpu_api):
    """A geometry cut out of the lattice by leaving sites unset (here the two last columns): their rows are zero, the
    reference's ``ε > 0`` filter drops their exactly-zero eigenvalues, and ``free_energy(cuda=True)`` takes their share
    ``-(T/2) ln 2`` per row out of the trace again.  (Formula checked against the unmodified reference on the CPU:
    oracle KPM + correction == reference to 2e-16 on this system.)"""
    api = gpu_api
    lattice = api.CubicLattice((6, 5, 1))
    system = api.Hamiltonian(lattice)
    with system as (H, D):
        for i in lattice.sites():
            if i[0] < 4:
                H[i, i] = -1.5 * api.σ0
                D[i, i] = 0.3 * api.jσ2
        for i, j in lattice.bonds():
            if i[0] < 4 and j[0] < 4:
                H[i, j] = -1.0 * api.σ0
    assert system._sys.zero_scalar_rows() == 40
    assert abs(-39.09478965835833 - system.free_energy(0.1)) <= 1e-12 * 39.1      # the reference's value for this script
    for T in (0.1, 0.5):
        F = system.free_energy(T)
        assert abs(system.free_energy(T, cuda=True) - F) <= 1e-10 * abs(F)
    Fs = system.free_energy(0.1, cuda=True, vectors=4096, seed=3)   # sigma ~ 0.2 %; the zero rows' share is 3.5 %
    assert abs(Fs - system.free_energy(0.1)) <= 1e-2 * abs(Fs)
