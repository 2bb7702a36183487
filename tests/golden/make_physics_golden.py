"""Golden values of the physics scenarios (``tests/physics_cases.py``) from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_physics_golden.py        # ~1 min: spsolve LDOS and dense eigvalsh free energies

Output (committed): ``tests/golden/physics.npz`` -- ``<scenario>/<quantity>`` arrays.
"""

from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.environ.get("BODGE_REFERENCE", "/root/reference"))
sys.path.insert(0, os.path.join(REPO, "tests"))

import bodge  # noqa: E402  (the reference)

import physics_cases  # noqa: E402

REF = types.SimpleNamespace(
    CubicLattice=bodge.CubicLattice, Hamiltonian=bodge.Hamiltonian,
    σ0=bodge.σ0, σ1=bodge.σ1, σ2=bodge.σ2, σ3=bodge.σ3, jσ2=bodge.jσ2, dwave=bodge.dwave, pwave=bodge.pwave,
)

OBSERVE = types.SimpleNamespace(
    ldos=lambda system, site, energies: system.ldos(site, energies),          # bodge/hamiltonian.py:323-387 (spsolve)
    free_energy=lambda system, T: system.free_energy(T),                      # bodge/hamiltonian.py:253-321 (eigvalsh)
)


def main():
    out = {}
    for name, (scenario, check) in physics_cases.SCENARIOS.items():
        values = scenario(REF, OBSERVE)
        check(values)  # the reference's own assertions hold for the reference
        for key, val in values.items():
            out[f"{name}/{key}"] = np.asarray(val, dtype=np.float64)
            print(name, key, out[f"{name}/{key}"])
    np.savez_compressed(os.path.join(HERE, "physics.npz"), **out)
    print("physics.npz", os.path.getsize(os.path.join(HERE, "physics.npz")), "bytes")


if __name__ == "__main__":
    main()
