"""The CPU oracle against the fixtures generated from the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare the
CUDA path with the oracle and with the same fixtures."""

import numpy as np
import pytest

import cases
from oracle import bdg_oracle as orc
from util import digest, oracle_assemble, rel_err, same_bits

SKELETON_SHAPES = [(1, 1, 1), (2, 1, 1), (2, 2, 2), (5, 1, 1), (1, 6, 1), (1, 1, 4), (4, 4, 1),
                   (2, 3, 1), (3, 1, 2), (3, 5, 7), (2, 5, 3), (6, 6, 6)]

SMALL = {
    "random_3_5_7": lambda api: cases.random_periodic(api, (3, 5, 7), seed=11),
    "random_2_5_3": lambda api: cases.random_periodic(api, (2, 5, 3), seed=12),
    "random_5_5_2": lambda api: cases.random_periodic(api, (5, 5, 2), seed=13),
    "kat_3_5_7": lambda api: cases.kat_export(api),
    "readme_12_12_1": lambda api: cases.readme_swave(api, (12, 12, 1)),
    "dwave_9_8_1": lambda api: cases.dwave_rashba(api, (9, 8, 1)),
    "swave3d_5_4_6": lambda api: cases.swave_3d(api, (5, 4, 6)),
    "junction_30_10_1": lambda api: cases.junction(api, (30, 10, 1)),
    "snf_10_7_3": lambda api: cases.snf_free_energy(api),
}

BIG = {
    "C1_readme_40_40_1": lambda api: cases.readme_swave(api, (40, 40, 1)),
    "C2_readme_100_100_1": lambda api: cases.readme_swave(api, (100, 100, 1)),
    "C3_dwave_100_100_1": lambda api: cases.dwave_rashba(api, (100, 100, 1)),
    "C4s_swave3d_16_16_16": lambda api: cases.swave_3d(api, (16, 16, 16)),
    "C5s_junction_90_40_1": lambda api: cases.junction(api, (90, 40, 1)),
}


def record(make):
    rec = make(cases.recorder_api())
    return rec.lattice.shape, [rec.packed(k) for k in range(len(rec.blocks))]


@pytest.mark.parametrize("shape", SKELETON_SHAPES)
def test_skeleton_matches_reference(structures, shape):
    indptr, indices = orc.cubic_skeleton(shape)
    tag = "skel_%d_%d_%d" % shape
    assert same_bits(indptr, structures[tag + "_indptr"])
    assert same_bits(indices, structures[tag + "_indices"])


def test_skeleton_known_rows():
    # 4x4x1 skeleton starts [0,1,3,4,12, 0,1,2,5,13], 5 blocks per row
    indptr, indices = orc.cubic_skeleton((4, 4, 1))
    assert indices[:10].tolist() == [0, 1, 3, 4, 12, 0, 1, 2, 5, 13]
    assert indptr.tolist() == list(range(0, 81, 5))


@pytest.mark.parametrize("tag", sorted(SMALL))
def test_assembly_matches_reference(structures, tag):
    shape, blocks = record(SMALL[tag])
    (sp, si, sd), (ep, ei, ed) = oracle_assemble(shape, blocks)
    assert same_bits(sp, structures[f"{tag}_sk_indptr"])
    assert same_bits(si, structures[f"{tag}_sk_indices"])
    assert np.array_equal(sd, structures[f"{tag}_sk_data"])
    assert same_bits(ep, structures[f"{tag}_ex_indptr"])
    assert same_bits(ei, structures[f"{tag}_ex_indices"])
    assert np.array_equal(ed, structures[f"{tag}_ex_data"])
    assert orc.hermitian_deviation(sp, si, sd) < 1e-12


@pytest.mark.parametrize("tag", sorted(BIG))
def test_config_digests(digests, tag):
    shape, blocks = record(BIG[tag])
    (sp, si, sd), (ep, ei, ed) = oracle_assemble(shape, blocks)
    want = digests[tag]
    assert len(si) == want["sk_nb"] and len(ei) == want["ex_nb"]
    assert digest(sp, si) == want["sk_structure"]
    assert digest(ep, ei) == want["ex_structure"]
    assert digest(sd) == want["sk_data"]
    assert digest(ed) == want["ex_data"]
    assert abs(orc.norm_inf(ep, ei, ed) - want["norm_inf"]) < 1e-12


def test_block_counts_formula(digests):
    # nb = N + 2[(Lx-1)LyLz + Lx(Ly-1)Lz + LxLy(Lz-1)] for open boundaries
    assert digests["C1_readme_40_40_1"]["ex_nb"] == 7840
    assert digests["C2_readme_100_100_1"]["ex_nb"] == 49600
    assert digests["C3_dwave_100_100_1"]["sk_nb"] == 50000


def test_known_answers(observables, structures):
    # reference tests/test_hamiltonian.py:86-93
    row0 = observables["kat_row0"]
    assert row0[0].real == 3 and row0[1].imag == 4 and row0[2].real == 2 and row0[3].imag == -5
    shape, blocks = record(SMALL["kat_3_5_7"])
    (sp, si, sd), _ = oracle_assemble(shape, blocks)
    dense = orc.to_scipy(sp, si, sd).toarray()
    assert np.array_equal(dense[0, :8], row0)


def test_not_neighbour_raises():
    indptr, indices = orc.cubic_skeleton((4, 4, 1))
    with pytest.raises(IndexError):
        orc.block_index(indptr, indices, [0], [10])


def test_non_hermitian_detected():
    indptr, indices = orc.cubic_skeleton((3, 3, 3))
    data = orc.zero_data(indices)
    z = np.zeros(0, dtype=np.int64)
    orc.scatter(indptr, indices, data, [13], [13], [1j * np.array([[0, 1], [1, 0]])], z, z, np.zeros((0, 2, 2)))
    assert orc.hermitian_deviation(indptr, indices, data) > 1e-6


def test_eliminate_zeros_semantics():
    indptr, indices = orc.cubic_skeleton((3, 1, 1))
    data = orc.zero_data(indices)
    data[0, 0, 0] = -0.0      # -0.0 == 0 -> dropped
    data[1, 2, 3] = np.nan    # NaN != 0 -> kept
    data[4, 1, 1] = 1e-300    # tiny but non-zero -> kept
    p, i, d = orc.eliminate_zeros(indptr, indices, data)
    assert len(i) == 2 and p[-1] == 2
    ref = orc.to_scipy(indptr, indices, data.copy())
    ref.eliminate_zeros()
    assert same_bits(p, ref.indptr) and same_bits(i, ref.indices)


# ---- Chebyshev restatement, pinned through the reference's observables --------------------
@pytest.mark.parametrize("tag", ["readme_12_12_1", "random_3_5_7", "dwave_9_8_1"])
def test_moments_match_fixture_and_doubling(observables, structures, tag):
    H = orc.to_scipy(structures[f"{tag}_ex_indptr"], structures[f"{tag}_ex_indices"], structures[f"{tag}_ex_data"])
    scale = float(observables[f"mu_{tag}_scale"])
    site = int(observables[f"mu_{tag}_site"])
    n_rows = H.shape[0]
    x0 = np.concatenate([orc.probes(n_rows, [4 * site + a for a in range(4)]),
                         orc.rademacher(1234, n_rows, np.arange(4))], axis=1)
    want = observables[f"mu_{tag}"]
    got = orc.cheb_moments(H, x0, want.shape[0], scale)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    dbl = orc.cheb_moments_doubling(H, x0, want.shape[0], scale)
    assert np.max(np.abs(dbl - want)) <= 1e-11 * np.max(np.abs(want))
    # the even-vector form the two-step CUDA kernel runs (E_{j+1} = 2 T_2(H~) E_j - E_{j-1}, product identities),
    # at every moment count modulo 4
    for n in (want.shape[0], want.shape[0] - 1, want.shape[0] - 2, want.shape[0] - 3, 1, 2, 3, 4, 5):
        even = orc.cheb_moments_even_vectors(H, x0, n, scale)
        assert even.shape == (n, x0.shape[1])
        assert np.max(np.abs(even - want[:n])) <= 1e-11 * np.max(np.abs(want))


@pytest.mark.parametrize("tag", ["snf_10_7_3", "readme_12_12_1", "junction_30_10_1", "dwave_9_8_1"])
def test_free_energy_from_moments_matches_reference(observables, structures, tag):
    """F from the exact Chebyshev trace == the reference's dense free_energy (T > 0: 1e-10)."""
    H = orc.to_scipy(structures[f"{tag}_ex_indptr"], structures[f"{tag}_ex_indices"], structures[f"{tag}_ex_data"])
    ptr, idx, dat = (structures[f"{tag}_ex_{k}"] for k in ("indptr", "indices", "data"))
    scale = 1.01 * orc.norm_inf(ptr, idx, dat)
    temps, want = observables["temps"], observables[f"F_{tag}"]
    n_rows = H.shape[0]
    n_mom = 900
    mu = orc.cheb_moments_doubling(H, np.eye(n_rows, dtype=np.complex128), n_mom, scale).sum(axis=1)
    for T, F_ref in zip(temps, want):
        F = orc.free_energy_from_moments(mu, float(T), scale)
        if T >= 0.1:
            assert abs(F - F_ref) <= 1e-10 * abs(F_ref), (tag, T, F, F_ref)
        elif T > 0:
            assert abs(F - F_ref) <= 1e-5 * abs(F_ref)
        else:
            assert abs(F - F_ref) <= 1e-4 * abs(F_ref)
    # and the dense restatement itself
    dense = H.toarray()
    for T, F_ref in zip(temps, want):
        assert abs(orc.free_energy_dense(dense, float(T)) - F_ref) <= 1e-11 * abs(F_ref)


@pytest.mark.parametrize("tag,site", [("readme_12_12_1", (6, 6, 0)), ("random_5_5_2", (2, 3, 1)), ("dwave_9_8_1", (4, 4, 0))])
def test_ldos_from_moments_matches_reference(observables, structures, tag, site):
    shape = tuple(int(v) for v in tag.split("_")[1:])
    ptr, idx, dat = (structures[f"{tag}_ex_{k}"] for k in ("indptr", "indices", "data"))
    H = orc.to_scipy(ptr, idx, dat)
    scale = 1.01 * orc.norm_inf(ptr, idx, dat)
    energies, want = observables["ldos_E"], observables[f"ldos_{tag}"]
    i = int(orc.cubic_index(shape, [site])[0])
    x0 = orc.probes(H.shape[0], [4 * i + a for a in range(4)])
    n_mom = int(np.ceil(32 * scale / 0.3))  # Γ = 0.3 for this energy grid
    mu = orc.cheb_moments_doubling(H, x0, n_mom + (n_mom & 1), scale)
    got = orc.ldos_from_moments(mu, energies, scale)
    assert rel_err(got, want) <= 1e-10


def test_rademacher_is_counter_based():
    a = orc.rademacher(7, 64, np.arange(8))
    b = orc.rademacher(7, 64, np.arange(4, 8))
    assert np.array_equal(a[:, 4:], b)           # sharding columns does not change them
    assert set(np.unique(a.real)) == {-1.0, 1.0} and not a.imag.any()
    assert abs(a.real.mean()) < 0.2
    assert not np.array_equal(a, orc.rademacher(8, 64, np.arange(8)))


@pytest.mark.parametrize("tag", sorted(SMALL))
def test_export_formats_match_reference(digests, tag):
    """matrix("csr") / ("csc") / ("dense") of the reference (hamiltonian.py:144-151), as sha256 of
    structure and values, against the scipy-free restatement in the oracle."""
    shape, blocks = record(SMALL[tag])
    (sp_, si, sd), _ = oracle_assemble(shape, blocks)
    want = digests["formats_" + tag]
    for fmt, transpose in (("csr", False), ("csc", True)):
        ptr, idx, val = orc.export_scalar(sp_, si, sd, transpose=transpose)
        assert ptr.dtype == np.int32 and idx.dtype == np.int32
        assert len(idx) == want[f"{fmt}_nnz"]
        assert digest(ptr, idx) == want[f"{fmt}_structure"]
        assert digest(val) == want[f"{fmt}_data"]
    assert digest(orc.export_dense(sp_, si, sd)) == want["dense"]
