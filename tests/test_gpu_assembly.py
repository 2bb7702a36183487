"""Parity of the CUDA assembly path (through the C ABI) with the reference's outputs (golden
fixtures) and with the oracle: structure bit-exact, values exact."""

import numpy as np
import pytest

import cases
from oracle import bdg_oracle as orc
from test_oracle import BIG, SKELETON_SHAPES, SMALL, record
from util import digest, oracle_assemble, same_bits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", SKELETON_SHAPES + [(40, 40, 1), (7, 1, 9), (2, 2, 1), (1, 2, 2), (33, 2, 5)])
def test_cubic_skeleton(gpu_api, structures, shape):
    system = gpu_api.Hamiltonian(gpu_api.CubicLattice(shape))
    sk = system._matrix
    indptr, indices = orc.cubic_skeleton(shape)
    assert sk.indptr.dtype == np.int32 and sk.indices.dtype == np.int32
    assert same_bits(sk.indptr, indptr) and same_bits(sk.indices, indices)
    tag = "skel_%d_%d_%d" % shape
    if tag + "_indptr" in structures:
        assert same_bits(sk.indptr, structures[tag + "_indptr"])
        assert same_bits(sk.indices, structures[tag + "_indices"])
    assert sk.blocksize == (4, 4) and sk.shape == (4 * np.prod(shape),) * 2
    assert not sk.data.any()


@pytest.mark.parametrize("tag", sorted(SMALL))
def test_assembly_matches_reference(gpu_api, structures, tag):
    system = SMALL[tag](gpu_api)
    sk, ex = system._matrix, system.matrix("bsr")
    assert same_bits(sk.indptr, structures[f"{tag}_sk_indptr"])
    assert same_bits(sk.indices, structures[f"{tag}_sk_indices"])
    assert np.array_equal(sk.data, structures[f"{tag}_sk_data"])
    assert same_bits(ex.indptr, structures[f"{tag}_ex_indptr"])
    assert same_bits(ex.indices, structures[f"{tag}_ex_indices"])
    assert np.array_equal(ex.data, structures[f"{tag}_ex_data"])
    # stronger than the reference's own check: bit patterns incl. signed zeros
    assert same_bits(sk.data, structures[f"{tag}_sk_data"])


@pytest.mark.parametrize("tag", sorted(BIG))
def test_config_digests(gpu_api, digests, tag):
    system = BIG[tag](gpu_api)
    sk, ex = system._matrix, system.matrix("bsr")
    want = digests[tag]
    assert len(sk.indices) == want["sk_nb"] and len(ex.indices) == want["ex_nb"]
    assert digest(sk.indptr, sk.indices) == want["sk_structure"]
    assert digest(ex.indptr, ex.indices) == want["ex_structure"]
    assert digest(sk.data) == want["sk_data"]
    assert digest(ex.data) == want["ex_data"]
    assert abs(system.spectral_bound() / 1.01 - want["norm_inf"]) < 1e-12


def test_export_formats_and_known_answers(gpu_api):
    # reference tests/test_hamiltonian.py:60-107
    system = cases.kat_export(gpu_api)
    H_DNS, H_BSR = system.matrix(format="dense"), system.matrix(format="bsr")
    H_CSR, H_CSC = system.matrix(format="csr"), system.matrix(format="csc")
    assert isinstance(H_DNS, np.ndarray)
    assert H_BSR.getformat() == "bsr" and H_CSR.getformat() == "csr" and H_CSC.getformat() == "csc"
    assert np.real(H_DNS[0, 0]) == 3 and np.imag(H_DNS[0, 1]) == 4
    assert np.real(H_DNS[0, 2]) == 2 and np.imag(H_DNS[0, 3]) == -5
    for M in (H_BSR, H_CSR, H_CSC):
        assert np.max(np.abs(M - H_DNS)) < 1e-6
    assert H_BSR.blocksize == (4, 4)
    with pytest.raises(Exception):
        system.matrix(format="blah")
    with pytest.raises(Exception):
        system.matrix(format=1)


@pytest.mark.parametrize("tag", ["random_3_5_7", "random_2_5_3", "kat_3_5_7", "dwave_9_8_1", "swave3d_5_4_6",
                                 "junction_30_10_1"])
def test_scalar_exports_match_reference_and_oracle(gpu_api, digests, tag):
    """matrix("csr") / ("csc") / ("dense") come from device kernels: bit-exact against
    the reference's digests and the oracle's restatement (reference hamiltonian.py:144-151)."""
    from test_oracle import SMALL
    from util import digest

    system = SMALL[tag](gpu_api)
    sk = system._matrix
    want = digests["formats_" + tag]
    for fmt, transpose in (("csr", False), ("csc", True)):
        M = system.matrix(fmt)
        assert M.getformat() == fmt and M.shape == system.shape
        assert M.indptr.dtype == np.int32 and M.indices.dtype == np.int32 and M.data.dtype == np.complex128
        optr, oidx, oval = orc.export_scalar(sk.indptr, sk.indices, sk.data, transpose=transpose)
        assert same_bits(M.indptr, optr) and same_bits(M.indices, oidx) and same_bits(M.data, oval)
        assert M.nnz == want[f"{fmt}_nnz"]
        assert digest(M.indptr, M.indices) == want[f"{fmt}_structure"] and digest(M.data) == want[f"{fmt}_data"]
    dense = np.asarray(system.matrix("dense"))
    assert digest(dense) == want["dense"]
    assert same_bits(dense, orc.export_dense(sk.indptr, sk.indices, sk.data))


def test_scalar_exports_zero_semantics(gpu_api):
    """Explicit zeros are dropped element-wise: -0.0 is zero, NaN and denormals are not (scipy's
    eliminate_zeros on the converted matrix); an empty Hamiltonian exports empty arrays."""
    system = gpu_api.Hamiltonian(gpu_api.CubicLattice((3, 2, 1)))
    for fmt in ("csr", "csc"):
        M = system.matrix(fmt)
        assert M.nnz == 0 and M.indptr.tolist() == [0] * 25
    data = system._data
    data[0, 0, 1] = -0.0
    data[1, 2, 3] = np.nan
    data[4, 1, 1] = 5e-324
    data[7, 3, 0] = 2.5 - 1j
    system._data = data
    ref = system._matrix
    for fmt, conv in (("csr", ref.tocsr), ("csc", ref.tocsc)):
        R = conv()
        R.eliminate_zeros()
        M = system.matrix(fmt)
        assert same_bits(M.indptr, R.indptr) and same_bits(M.indices, R.indices) and M.nnz == 3
        assert same_bits(M.data, R.data)


def test_hermitian_check_and_errors(gpu_api):
    # reference tests/test_hamiltonian.py:17-57
    system = cases.random_periodic(gpu_api, (3, 5, 7), seed=3)
    H = system._matrix.todense()
    assert np.allclose(H, H.T.conj())
    before = system._matrix.data.copy()
    with pytest.raises(RuntimeError, match="not Hermitian"):
        with system as (H, D):
            H[(1, 1, 1), (1, 1, 1)] = 1j * gpu_api.σ1
    # like the reference, the offending entry has been applied
    after = system._matrix.data
    k = system.index((1, 1, 1), (1, 1, 1))
    assert np.array_equal(after[k, 0:2, 0:2], 1j * gpu_api.σ1)
    assert np.array_equal(np.delete(after, k, axis=0), np.delete(before, k, axis=0))
    # pair outside the skeleton -> IndexError (reference: hamiltonian.py:170)
    with pytest.raises(IndexError):
        with system as (H, D):
            H[(0, 0, 0), (1, 1, 1)] = gpu_api.σ0
    with pytest.raises(IndexError):
        system.index((0, 0, 0), (2, 2, 2))
    # coordinate outside the lattice -> ValueError (reference: lattice.py:106)
    with pytest.raises(ValueError):
        with system as (H, D):
            H[(0, 0, 0), (0, 0, 7)] = gpu_api.σ0
    with pytest.raises(Exception):
        gpu_api.Hamiltonian("not a lattice")


def test_failing_entry_leaves_earlier_entries_applied(gpu_api):
    system = gpu_api.Hamiltonian(gpu_api.CubicLattice((4, 4, 1)))
    with pytest.raises(IndexError):
        with system as (H, D):
            H[(0, 0, 0), (0, 0, 0)] = 2 * gpu_api.σ0       # applied
            H[(0, 0, 0), (2, 2, 0)] = gpu_api.σ0           # fails
            H[(1, 1, 0), (1, 1, 0)] = 5 * gpu_api.σ0       # after the failure: not applied
    data = system._matrix.data
    assert np.array_equal(data[system.index((0, 0, 0), (0, 0, 0)), 0:2, 0:2], 2 * gpu_api.σ0)
    assert not data[system.index((1, 1, 0), (1, 1, 0))].any()


def test_index_matches_scipy_layout(gpu_api):
    lattice = gpu_api.CubicLattice((3, 4, 2))
    system = gpu_api.Hamiltonian(lattice)
    sk = system._matrix
    for ri, rj in list(lattice)[::7]:
        i, j = lattice[ri], lattice[rj]
        k = system.index(ri, rj)
        assert sk.indptr[i] <= k < sk.indptr[i + 1] and sk.indices[k] == j


def test_incremental_fill_and_overwrite(gpu_api):
    # later `with` blocks only touch the keys they set (reference tests/test_physics.py:155-160)
    shape = (6, 5, 1)
    rec_blocks = []
    api = cases.recorder_api()
    rec = cases.readme_swave(api, shape)
    system = cases.readme_swave(gpu_api, shape)
    extra = {((2, 2, 0), (2, 2, 0)): 0.7 * gpu_api.σ1 - 0.2 * gpu_api.σ3}
    with system as (H, D):
        for key, val in extra.items():
            H[key] = val
    with rec as (H, D):
        for key, val in extra.items():
            H[key] = val
    rec_blocks = [rec.packed(k) for k in range(len(rec.blocks))]
    (sp, si, sd), (ep, ei, ed) = oracle_assemble(shape, rec_blocks)
    assert np.array_equal(system._matrix.data, sd)
    ex = system.matrix("bsr")
    assert same_bits(ex.indptr, ep) and same_bits(ex.indices, ei) and np.array_equal(ex.data, ed)


def test_eliminate_zeros_semantics(gpu_api):
    system = gpu_api.Hamiltonian(gpu_api.CubicLattice((3, 1, 1)))
    data = system._data
    data[0, 0, 0] = -0.0
    data[1, 2, 3] = np.nan
    data[4, 1, 1] = 1e-300
    system._data = data
    ex = system.matrix("bsr")
    ref = system._matrix
    ref.eliminate_zeros()
    assert same_bits(ex.indptr, ref.indptr) and same_bits(ex.indices, ref.indices)
    assert len(ex.indices) == 2




def test_generic_lattice_path(gpu_api):
    """A Lattice subclass that is not the stock CubicLattice goes through the pair-list skeleton
    (bucket by row, per-row sort/unique) and must give the same structure and values."""
    import bodge_b200 as b

    class MyCubic(b.CubicLattice):
        def bonds(self, axis=None):  # same bonds, overridden -> generic path
            yield from super().bonds(axis)

    for shape in [(3, 5, 7), (2, 2, 2), (5, 1, 1), (4, 4, 1)]:
        lat = MyCubic(shape)
        system = b.Hamiltonian(lat)
        indptr, indices = orc.cubic_skeleton(shape)
        sk = system._matrix
        assert same_bits(sk.indptr, indptr) and same_bits(sk.indices, indices)
    api = cases.recorder_api()
    rec = cases.random_periodic(api, (3, 5, 7), seed=5)
    system = b.Hamiltonian(MyCubic((3, 5, 7)))
    system.fill(*rec.packed())
    (sp, si, sd), _ = oracle_assemble((3, 5, 7), [rec.packed()])
    assert np.array_equal(system._matrix.data, sd)


@pytest.mark.parametrize("name,shape", [("readme_swave", (7, 6, 1)), ("dwave_rashba", (6, 7, 1)),
                                         ("swave_3d", (4, 5, 3)), ("junction", (12, 5, 1))])
def test_packed_workloads_equal_dict_api(gpu_api, name, shape):
    """The vectorised builders used at 10^6 sites produce the same matrix as the dict API."""
    import bodge_b200 as b
    from bodge_b200 import workloads

    via_dict = getattr(cases, name)(gpu_api, shape)
    via_pack = b.Hamiltonian(b.CubicLattice(shape))
    via_pack.fill(*getattr(workloads, name)(shape))
    assert same_bits(via_dict._matrix.data, via_pack._matrix.data)
    a, c = via_dict.matrix("bsr"), via_pack.matrix("bsr")
    assert same_bits(a.indptr, c.indptr) and same_bits(a.indices, c.indices) and same_bits(a.data, c.data)


def test_oracle_parity_seeded_random(gpu_api):
    """CUDA scatter vs the oracle on fresh seeded inputs (not in the fixtures)."""
    import bodge_b200 as b

    for seed, shape in [(101, (4, 3, 5)), (102, (9, 2, 2)), (103, (1, 8, 3))]:
        rec = cases.random_periodic(cases.recorder_api(), shape, seed=seed)
        system = cases.random_periodic(gpu_api, shape, seed=seed)
        (sp, si, sd), (ep, ei, ed) = oracle_assemble(shape, [rec.packed()])
        sk, ex = system._matrix, system.matrix("bsr")
        assert same_bits(sk.indptr, sp) and same_bits(sk.indices, si) and same_bits(sk.data, sd)
        assert same_bits(ex.indptr, ep) and same_bits(ex.indices, ei) and same_bits(ex.data, ed)
        assert abs(system.spectral_bound() / 1.01 - orc.norm_inf(sp, si, sd)) < 1e-12


def test_full_size_properties(gpu_api):
    """C5 size (10^6 sites): block counts, Hermiticity and symmetry properties that do not need
    the (minutes-long) reference run."""
    import bodge_b200 as b
    from bodge_b200 import workloads

    shape = (1000, 1000, 1)
    system = b.Hamiltonian(b.CubicLattice(shape))
    assert system._sys.n_blocks == 5_000_000
    dev = system.fill(*workloads.junction(shape))
    assert dev == 0.0
    indptr, indices, data = system._sys.export_bsr(True)
    assert len(indices) == 4_996_000 and indptr[-1] == 4_996_000
    counts = np.diff(indptr)
    assert counts.min() == 3 and counts.max() == 5
    assert (np.diff(indices)[np.diff(np.repeat(np.arange(len(counts)), counts)) == 0] > 0).all()  # sorted rows
    # particle-hole structure of every block: lower-right = -conj(upper-left)
    assert np.array_equal(data[:, 2:4, 2:4], -data[:, 0:2, 0:2].conj())
    # small-lattice twin built through the same code agrees with the reference digest elsewhere;
    # here check translation invariance along y of the interior blocks
    row = lambda x, y: (x * 1000 + y)
    for x in (10, 500, 900):
        a = data[indptr[row(x, 400)] : indptr[row(x, 400) + 1]]
        c = data[indptr[row(x, 401)] : indptr[row(x, 401) + 1]]
        assert np.array_equal(a, c)
    row_sums = np.add.reduceat(np.abs(data).sum(axis=2), indptr[:-1], axis=0)
    assert abs(system.spectral_bound() / 1.01 - row_sums.max()) < 1e-12
