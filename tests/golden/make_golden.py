"""Generate the golden fixtures by running the UNMODIFIED reference (jabirali/bodge v1.3.0).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py            # everything incl. the 40x40 dense eigvalsh (~1 min)

Outputs (committed):
    tests/golden/structures.npz   full indptr/indices/data of small systems built by the reference
    tests/golden/observables.npz  free energies, LDOS, Chebyshev moments (scipy on the reference BSR)
    tests/golden/digests.json     sha256 digests + block counts of the larger configs (C1, C2, C3 ...)
"""

from __future__ import annotations

import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.environ.get("BODGE_REFERENCE", "/root/reference"))
sys.path.insert(0, os.path.join(REPO, "tests"))  # `cases` by path: the reference ships a `tests` package too
sys.path.insert(0, REPO)

import bodge  # noqa: E402  (the reference)

from oracle import bdg_oracle as orc  # noqa: E402  (only for the Chebyshev restatement on the reference BSR)
import cases  # noqa: E402

REF = types.SimpleNamespace(
    CubicLattice=bodge.CubicLattice, Hamiltonian=bodge.Hamiltonian,
    σ0=bodge.σ0, σ1=bodge.σ1, σ2=bodge.σ2, σ3=bodge.σ3, jσ2=bodge.jσ2, dwave=bodge.dwave,
)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def struct_of(system):
    sk = system._matrix
    ex = system.matrix("bsr")
    return dict(
        sk_indptr=sk.indptr.copy(), sk_indices=sk.indices.copy(), sk_data=sk.data.copy(),
        ex_indptr=ex.indptr.copy(), ex_indices=ex.indices.copy(), ex_data=ex.data.copy(),
    )


def format_digests(system):
    """sha256 of what the reference's matrix("csr") / ("csc") / ("dense") return (hamiltonian.py:144-151)."""
    csr, csc = system.matrix("csr"), system.matrix("csc")
    assert csr.indptr.dtype == np.int32 and csc.indices.dtype == np.int32
    return dict(
        csr_structure=digest(csr.indptr, csr.indices), csr_data=digest(csr.data), csr_nnz=int(csr.nnz),
        csc_structure=digest(csc.indptr, csc.indices), csc_data=digest(csc.data), csc_nnz=int(csc.nnz),
        dense=digest(np.asarray(system.matrix("dense"))),
    )


def main():
    structures = {}
    digests = {}
    obs = {}

    # ---- skeletons of degenerate and small shapes (hamiltonian.py:37-64) --------------
    for shape in [(1, 1, 1), (2, 1, 1), (2, 2, 2), (5, 1, 1), (1, 6, 1), (1, 1, 4), (4, 4, 1),
                  (2, 3, 1), (3, 1, 2), (3, 5, 7), (2, 5, 3), (6, 6, 6)]:
        system = bodge.Hamiltonian(bodge.CubicLattice(shape))
        tag = "skel_%d_%d_%d" % shape
        structures[tag + "_indptr"] = system._matrix.indptr.copy()
        structures[tag + "_indices"] = system._matrix.indices.copy()
        assert system._matrix.indptr.dtype == np.int32 and system._matrix.indices.dtype == np.int32

    # ---- full structures + values of small systems -------------------------------------
    small = {
        "random_3_5_7": lambda: cases.random_periodic(REF, (3, 5, 7), seed=11),
        "random_2_5_3": lambda: cases.random_periodic(REF, (2, 5, 3), seed=12),
        "random_5_5_2": lambda: cases.random_periodic(REF, (5, 5, 2), seed=13),
        "kat_3_5_7": lambda: cases.kat_export(REF),
        "readme_12_12_1": lambda: cases.readme_swave(REF, (12, 12, 1)),
        "dwave_9_8_1": lambda: cases.dwave_rashba(REF, (9, 8, 1)),
        "swave3d_5_4_6": lambda: cases.swave_3d(REF, (5, 4, 6)),
        "junction_30_10_1": lambda: cases.junction(REF, (30, 10, 1)),
        "snf_10_7_3": lambda: cases.snf_free_energy(REF),
    }
    systems = {}
    for tag, make in small.items():
        system = make()
        systems[tag] = system
        for key, val in struct_of(system).items():
            structures[f"{tag}_{key}"] = val
        digests["formats_" + tag] = format_digests(system)

    # ---- digests of the named configs ----------------------------------------------------
    big = {
        "C1_readme_40_40_1": lambda: cases.readme_swave(REF, (40, 40, 1)),
        "C2_readme_100_100_1": lambda: cases.readme_swave(REF, (100, 100, 1)),
        "C3_dwave_100_100_1": lambda: cases.dwave_rashba(REF, (100, 100, 1)),
        "C4s_swave3d_16_16_16": lambda: cases.swave_3d(REF, (16, 16, 16)),
        "C5s_junction_90_40_1": lambda: cases.junction(REF, (90, 40, 1)),
    }
    for tag, make in big.items():
        system = make()
        systems[tag] = system
        s = struct_of(system)
        digests[tag] = dict(
            sk_nb=int(len(s["sk_indices"])), ex_nb=int(len(s["ex_indices"])),
            sk_structure=digest(s["sk_indptr"], s["sk_indices"]),
            sk_data=digest(s["sk_data"]),
            ex_structure=digest(s["ex_indptr"], s["ex_indices"]),
            ex_data=digest(s["ex_data"]),
            norm_inf=float(abs(system.matrix("csr")).sum(axis=1).max()),
        )
        print(tag, digests[tag]["ex_nb"], digests[tag]["ex_structure"][:16], digests[tag]["ex_data"][:16])

    # ---- known answers (tests/test_hamiltonian.py:86-93) ---------------------------------
    Hd = np.asarray(systems["kat_3_5_7"].matrix("dense"))
    obs["kat_row0"] = Hd[0, :8].copy()

    # ---- free energy (hamiltonian.py:253-321) --------------------------------------------
    temps = np.array([0.0, 0.01, 0.05, 0.1, 1.0])
    obs["temps"] = temps
    for tag in ["snf_10_7_3", "readme_12_12_1", "junction_30_10_1", "dwave_9_8_1"]:
        obs[f"F_{tag}"] = np.array([systems[tag].free_energy(float(T)) for T in temps])
    if os.environ.get("GOLDEN_SKIP_C1") != "1":
        obs["F_C1"] = np.array([systems["C1_readme_40_40_1"].free_energy(float(T)) for T in temps])
        print("F_C1", obs["F_C1"])

    # ---- LDOS (hamiltonian.py:323-387) ---------------------------------------------------
    ldos_E = np.array([-0.9, -0.6, -0.3, 0.0, 0.3, 0.6, 0.9])
    obs["ldos_E"] = ldos_E
    obs["ldos_readme_12_12_1"] = systems["readme_12_12_1"].ldos((6, 6, 0), ldos_E)
    obs["ldos_random_5_5_2"] = systems["random_5_5_2"].ldos((2, 3, 1), ldos_E)
    obs["ldos_dwave_9_8_1"] = systems["dwave_9_8_1"].ldos((4, 4, 0), ldos_E)

    # ---- Chebyshev moments: scipy bsr_matvecs on the REFERENCE's matrix("bsr") -----------
    for tag in ["readme_12_12_1", "random_3_5_7", "dwave_9_8_1"]:
        H = systems[tag].matrix("bsr")
        scale = 1.01 * float(abs(H.tocsr()).sum(axis=1).max())
        n_rows = H.shape[0]
        site = n_rows // 8
        x0 = np.concatenate(
            [orc.probes(n_rows, [4 * site + a for a in range(4)]), orc.rademacher(1234, n_rows, np.arange(4))],
            axis=1,
        )
        obs[f"mu_{tag}"] = orc.cheb_moments(H, x0, 96, scale)
        obs[f"mu_{tag}_scale"] = np.array(scale)
        obs[f"mu_{tag}_site"] = np.array(site)

    np.savez_compressed(os.path.join(HERE, "structures.npz"), **structures)
    np.savez_compressed(os.path.join(HERE, "observables.npz"), **obs)
    with open(os.path.join(HERE, "digests.json"), "w") as fh:
        json.dump(digests, fh, indent=1, sort_keys=True)
    for name in ("structures.npz", "observables.npz", "digests.json"):
        print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")


def formats_only():
    """Add / refresh only the matrix("csr"/"csc"/"dense") digests in digests.json."""
    from test_oracle import SMALL

    path = os.path.join(HERE, "digests.json")
    with open(path) as fh:
        digests = json.load(fh)
    for tag, make in SMALL.items():
        digests["formats_" + tag] = format_digests(make(REF))
    with open(path, "w") as fh:
        json.dump(digests, fh, indent=1, sort_keys=True)
    print("formats:", ", ".join(sorted(SMALL)))


if __name__ == "__main__":
    formats_only() if "--formats-only" in sys.argv else main()
