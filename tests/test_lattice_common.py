"""Host-side geometry and vocabulary, mirroring the reference's tests/test_lattice.py and
tests/test_common.py (site numbering, bond/edge counts, error behaviour, Pauli algebra)."""

import numpy as np
import pytest

import bodge_b200 as b
from oracle import bdg_oracle as orc


def test_pauli_algebra():
    # reference tests/test_common.py:8-18
    for s in (b.σ1, b.σ2, b.σ3):
        assert np.allclose(s @ s, b.σ0)
    assert np.allclose(b.σ1 @ b.σ2, b.jσ3)
    assert np.allclose(b.σ2 @ b.σ3, b.jσ1)
    assert np.allclose(b.σ3 @ b.σ1, b.jσ2)
    assert b.σ.shape == (3, 2, 2) and b.jσ.shape == (3, 2, 2)
    assert b.sigma2 is b.σ2 and b.jsigma0 is b.jσ0 and b.pi == np.pi
    assert b.σ0.dtype == np.complex128


def test_exports_match_reference_all():
    # bodge/__init__.py:13-51
    names = ["Lattice", "CubicLattice", "Hamiltonian", "Coord", "Coords", "Index", "Indices", "ssd", "swave", "pwave",
             "dwave", "π", "σ", "σ0", "σ1", "σ2", "σ3", "jσ", "jσ0", "jσ1", "jσ2", "jσ3", "pi", "sigma", "sigma0",
             "sigma1", "sigma2", "sigma3", "jsigma", "jsigma0", "jsigma1", "jsigma2", "jsigma3"]
    assert sorted(b.__all__) == sorted(names)
    for n in names:
        assert hasattr(b, n)


def test_lattice_is_abstract_and_typechecked():
    # reference tests/test_lattice.py:14-35
    with pytest.raises(ValueError):
        b.Lattice((1, 2, 3))

    class MyLattice(b.Lattice):
        pass

    lat = MyLattice((1, 2, 3))
    assert repr(lat) == "MyLattice(1, 2, 3)"
    assert lat.size == 6 and lat.dim == 2
    for call in (lambda: lat.index((0, 0, 0)), lambda: list(lat.sites()), lambda: list(lat.bonds()), lambda: list(lat.edges())):
        with pytest.raises(NotImplementedError):
            call()
    with pytest.raises(Exception):
        b.CubicLattice((1.5, 2, 3))
    with pytest.raises(Exception):
        b.CubicLattice((3, 3, 3))["a"]


def test_cubic_sites_are_index_ordered():
    # reference tests/test_lattice.py:38-68
    lat = b.CubicLattice((3, 5, 7))
    sites = list(lat.sites())
    assert len(sites) == 105 == lat.size and lat.dim == 3
    for ind, site in enumerate(sites):
        assert lat[site] == ind
    for bad in [(-1, 0, 0), (3, 0, 0), (0, 5, 0), (0, 0, 7), (0, -1, 0)]:
        with pytest.raises(ValueError):
            lat[bad]
    assert np.array_equal(lat.index_many(np.array(sites)), np.arange(105))
    assert np.array_equal(lat.sites_array(), np.array(sites))
    with pytest.raises(ValueError):
        lat.index_many(np.array([[0, 0, 7]]))


def test_cubic_bonds_and_edges():
    # reference tests/test_lattice.py:71-122
    lat = b.CubicLattice((2, 3, 5))
    bonds = list(lat.bonds())
    assert len(bonds) == 2 * ((2 - 1) * 3 * 5 + 2 * (3 - 1) * 5 + 2 * 3 * (5 - 1))
    for i, j in bonds:
        assert sum(abs(a - c) for a, c in zip(i, j)) == 1
    for axis in range(3):
        for i, j in lat.bonds(axis=axis):
            d = [abs(a - c) for a, c in zip(i, j)]
            assert d[axis] == 1 and sum(d) == 1
    edges = list(lat.edges())
    assert len(edges) == 2 * (2 * 3 + 3 * 5 + 5 * 2)
    for axis in range(3):
        for i, j in lat.edges(axis=axis):
            assert {i[axis], j[axis]} == {0, lat.shape[axis] - 1}
            assert all(i[a] == j[a] for a in range(3) if a != axis)
    with pytest.raises(ValueError):
        list(lat.bonds(axis=3))
    with pytest.raises(ValueError):
        list(lat.edges(axis=3))


@pytest.mark.parametrize("shape", [(3, 5, 7), (2, 2, 2), (4, 1, 1), (1, 6, 2), (1, 1, 1)])
def test_iteration_order_matches_reference_restatement(shape):
    """`for ri, rj in lattice` yields the pairs in the reference's order (oracle restates
    bodge/lattice.py:42-50,110-197 vectorised); the array companions agree with the generators."""
    lat = b.CubicLattice(shape)
    pi = [lat[i] for i, _ in lat]
    pj = [lat[j] for _, j in lat]
    oi, oj = orc.cubic_pairs(shape)
    assert pi == oi.tolist() and pj == oj.tolist()
    for axis in range(3):
        i, j = lat.bonds_array(axis)
        gen = [(lat[a], lat[c]) for a, c in lat.bonds(axis=axis)][::2]
        assert list(zip(i.tolist(), j.tolist())) == gen
        i, j = lat.edges_array(axis)
        gen = [(lat[a], lat[c]) for a, c in lat.edges(axis=axis)][::2]
        assert list(zip(i.tolist(), j.tolist())) == gen


def test_order_parameter_helpers():
    # reference tests/test_hamiltonian.py:132-315 (values of the helpers)
    assert np.array_equal(b.swave()((0, 0, 0), (1, 0, 0)), b.jσ2)
    sd = b.dwave()
    assert np.allclose(sd((0, 0, 0), (1, 0, 0)), b.jσ2)
    assert np.allclose(sd((0, 0, 0), (0, 1, 0)), -b.jσ2)
    assert np.allclose(sd((0, 0, 0), (0, 0, 1)), 0 * b.jσ2)
    sp = b.pwave("(p_x + jp_y) * (e_x + je_y)")
    for i, j in [((0, 0, 0), (1, 0, 0)), ((2, 2, 0), (2, 3, 0))]:
        assert np.allclose(sp(i, j), -sp(j, i))           # odd in momentum
        assert np.allclose(sp(i, j), sp(i, j).T)          # triplet: symmetric in spin
    dz = b.pwave("e_z * p_x")((0, 0, 0), (1, 0, 0))
    assert np.allclose(dz, b.σ3 @ b.jσ2 / 2)

    class Fake:
        lattice = b.CubicLattice((11, 11, 1))

    phi = b.ssd(Fake())
    assert np.isclose(phi((5, 5, 0), (5, 5, 0)), 1.0)
    assert phi((0, 0, 0), (0, 0, 0)) < phi((3, 3, 0), (3, 3, 0)) < 1.0
    assert phi((0, 0, 0), (0, 0, 0)) > 0
