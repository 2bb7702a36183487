"""The CPU oracle on the reference's physics scenarios (``physics_cases.py``) against values computed by the
UNMODIFIED reference (``golden/physics.npz``, made by ``golden/make_physics_golden.py``).

This pins the oracle beyond single systems: several ``with`` blocks on one handle (partial overwrites), bond pairing
(p-wave), complex on-site pairing, and the KPM read-out (resolvent / free-energy series) at the broadenings and
temperatures the reference's own integration tests use.  The GPU twin is ``test_gpu_physics.py``."""

import os
import types

import numpy as np
import pytest

import cases
import physics_cases
from oracle import bdg_oracle as orc
from util import oracle_assemble

TOL = 1e-10
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "physics.npz")


@pytest.fixture(scope="module")
def physics():
    return dict(np.load(GOLDEN))


def assembled(rec):
    """Everything the recorder saw so far, through the oracle's scatter; zeros eliminated like ``matrix("bsr")``."""
    shape = rec.lattice.shape
    _, (ptr, idx, dat) = oracle_assemble(shape, [rec.packed(k) for k in range(len(rec.blocks))])
    return shape, ptr, idx, dat


def kpm_ldos(rec, site, energies):
    shape, ptr, idx, dat = assembled(rec)
    H = orc.to_scipy(ptr, idx, dat)
    scale = 1.01 * orc.norm_inf(ptr, idx, dat)
    eps = np.unique(np.abs(np.asarray(energies, dtype=float)))
    gamma = float(np.min(np.abs(np.gradient(eps))))
    n_mom = int(np.ceil(32 * scale / gamma))
    i = int(orc.cubic_index(shape, [site])[0])
    mu = orc.cheb_moments_doubling(H, orc.probes(H.shape[0], [4 * i + a for a in range(4)]), n_mom + (n_mom & 1), scale)
    return orc.ldos_from_moments(mu, np.asarray(energies, dtype=float), scale)


def kpm_free_energy(rec, T):
    """Exact Chebyshev trace (all unit columns), series long enough for 1e-13 at this temperature."""
    _, ptr, idx, dat = assembled(rec)
    H = orc.to_scipy(ptr, idx, dat)
    scale = 1.01 * orc.norm_inf(ptr, idx, dat)
    n_mom = int(np.ceil(30.0 * scale / (np.pi * T)))
    n_mom += n_mom & 1
    # a temperature sweep on one matrix runs the recursion once (kept on the recorder, dropped by the next with-block)
    seen, trace = getattr(rec, "_trace", (None, ()))
    if seen != len(rec.blocks) or len(trace) < n_mom:
        trace = orc.cheb_moments_doubling(H, np.eye(H.shape[0], dtype=np.complex128), n_mom, scale).sum(axis=1)
        rec._trace = (len(rec.blocks), trace)
    return orc.free_energy_from_moments(trace[:n_mom], T, scale)


def dense_free_energy(rec, T):
    _, ptr, idx, dat = assembled(rec)
    return orc.free_energy_dense(orc.to_scipy(ptr, idx, dat).toarray(), T)


KPM = types.SimpleNamespace(ldos=kpm_ldos, free_energy=kpm_free_energy)
DENSE = types.SimpleNamespace(ldos=kpm_ldos, free_energy=dense_free_energy)


def compare(name, values, physics, tol=TOL):
    for key, got in values.items():
        want = physics[f"{name}/{key}"]
        err = np.max(np.abs(np.asarray(got) - want)) / np.max(np.abs(want))
        assert err <= tol, (name, key, err, got, want)


@pytest.mark.parametrize("name", ["gap_existence", "gap_sweep", "odd_frequency", "energy_temperature", "pwave_edges",
                                  "josephson_phase"])
def test_kpm_oracle_reproduces_the_reference(physics, name):
    scenario, check = physics_cases.SCENARIOS[name]
    values = scenario(cases.recorder_api(), KPM)
    compare(name, values, physics)
    check(values)


@pytest.mark.parametrize("name", ["gap_sweep", "spin_valve", "energy_temperature", "josephson_phase"])
def test_dense_oracle_reproduces_the_reference(physics, name):
    """The dense restatement (eigvalsh) on the oracle-assembled matrix: agreement to rounding, so even the spin valve's
    2e-11 relative difference between its two configurations keeps its sign."""
    scenario, check = physics_cases.SCENARIOS[name]
    values = scenario(cases.recorder_api(), DENSE)
    compare(name, values, physics, tol=1e-13)
    check(values)
