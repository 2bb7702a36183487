"""pytest configuration: the ``gpu`` marker, golden-fixture loaders, shared helpers.

``-m "not gpu"``: oracle vs. the reference-generated golden fixtures, host logic, ABI symbols.
``-m gpu``:       parity of the CUDA path (through the C ABI) against the oracle and the fixtures.
"""

import json
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def structures():
    return dict(np.load(os.path.join(GOLDEN, "structures.npz")))


@pytest.fixture(scope="session")
def observables():
    return dict(np.load(os.path.join(GOLDEN, "observables.npz")))


@pytest.fixture(scope="session")
def digests():
    with open(os.path.join(GOLDEN, "digests.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def gpu_api():
    """``bodge_b200`` as the namespace the shared model builders expect."""
    import bodge_b200 as b

    if b._native.device_count() == 0:
        pytest.fail("no CUDA device visible: the -m gpu tests must run on the GPU box")
    return types.SimpleNamespace(
        CubicLattice=b.CubicLattice, Hamiltonian=b.Hamiltonian,
        σ0=b.σ0, σ1=b.σ1, σ2=b.σ2, σ3=b.σ3, jσ2=b.jσ2, dwave=b.dwave, pwave=b.pwave,
    )
