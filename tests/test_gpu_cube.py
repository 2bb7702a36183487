"""The even-vector recursion on THREE-DIMENSIONAL lattices (csrc/cheb_cube.cu; VERDICT r1 next-4, BASELINE config C4):
kernel="t2" on a cubic lattice with Ly, Lz >= 2 -- two applications of H~ per launch on 4-column panels, a CTA per
(y, z) patch marching along x, the intermediate vector in shared memory.

Row arithmetic and order are those of the single-step dictionary kernel (on-site block by FP64 MMA, six real-diagonal
hopping blocks by DFMA in ascending block column), so E_j = T_2j(H~) x agrees with the three-term recursion's T_2j to
rounding for every patch shape / segment plan (halo values are recomputed, never exchanged); against the oracle
(scipy bsr_matvecs recursion) the moments hold the 1e-10 of BASELINE.json.
"""

import numpy as np
import pytest

import cases
from oracle import bdg_oracle as orc
from util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _patterned(api, shape, seed=5, cut=True):
    """s-wave model whose on-site block changes from site to site along z and y (three potentials, two gaps: z-adjacent
    sites of a warp's pair differ), with spin-dependent real-diagonal hopping that depends on the axis, and -- ``cut`` -- a few
    bonds left out inside the lattice (no block there: a zero coefficient, not a lattice boundary)."""
    rng = np.random.default_rng(seed)
    lattice = api.CubicLattice(shape)
    system = api.Hamiltonian(lattice)
    hop = {0: -1.0 * api.σ0 - 0.2 * api.σ3, 1: -0.8 * api.σ0 + 0.1 * api.σ3, 2: -1.1 * api.σ0}
    bonds = list(lattice.bonds())
    skip = set()
    if cut:
        for k in rng.choice(len(bonds), size=max(1, len(bonds) // 40), replace=False):
            i, j = bonds[int(k)]
            skip.add((i, j))
            skip.add((j, i))
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = (2.5 + 0.3 * ((i[2] + 2 * i[1]) % 3)) * api.σ0 - 0.05 * api.σ3
            D[i, i] = -(0.1 + 0.1 * ((i[1] + i[2]) % 2)) * api.jσ2 * np.exp(0.3j * (i[0] % 2))
        for i, j in bonds:
            if (i, j) in skip:
                continue
            axis = [a for a in range(3) if i[a] != j[a]][0]
            H[i, j] = hop[axis]
    return system


SYSTEMS = {
    "swave_6_5_4": lambda api: cases.swave_3d(api, (6, 5, 4)),            # one ragged patch
    "swave_5_2_2": lambda api: cases.swave_3d(api, (5, 2, 2)),            # the smallest planes it takes
    "swave_2_3_2": lambda api: cases.swave_3d(api, (2, 3, 2)),            # two planes: every plane is a boundary plane
    "junction_9_10_11": lambda api: cases.junction(api, (9, 10, 11)),     # odd Lz (a pair without its second site), codes change along x
    "junction_4_17_9": lambda api: cases.junction(api, (4, 17, 9)),       # three patches along y, two along z
    "swave_7_8_8": lambda api: cases.swave_3d(api, (7, 8, 8)),            # exactly one 8 x 8 patch
    "patterned_6_12_10": lambda api: _patterned(api, (6, 12, 10)),        # z-adjacent on-site blocks differ, cut bonds
    "patterned_5_9_18": lambda api: _patterned(api, (5, 9, 18), seed=8),
    "junction_13_20_3": lambda api: cases.junction(api, (13, 20, 3)),     # thin in z
}

# (BDG_CUBE_SHAPE, BDG_CUBE_SEG): None = planner's choice
PLANS = [(None, None), (0, 1), (1, 3), (2, 2), (0, 1000), (1, None), (2, None)]


def _set_plan(monkeypatch, plan):
    for name, value in zip(("BDG_CUBE_SHAPE", "BDG_CUBE_SEG"), plan):
        if value is None:
            monkeypatch.delenv(name, raising=False)
        else:
            monkeypatch.setenv(name, str(value))


@pytest.mark.parametrize("plan", PLANS)
@pytest.mark.parametrize("tag", sorted(SYSTEMS))
def test_cube_vectors_match_the_single_step_kernel(gpu_api, monkeypatch, tag, plan):
    system = SYSTEMS[tag](gpu_api)
    sysn = system._sys
    scale = system.spectral_bound()
    _set_plan(monkeypatch, plan)
    for n_cols, launches in ((8, 3), (4, 1), (5, 2), (3, 4), (13, 2), (1, 2)):
        sysn.cheb_begin(n_random=n_cols, seed=5, col_offset=3, scale=scale, kernel="t2")
        assert sysn.cheb_format()["kernel"] == "t2" and sysn.cheb_info()["panel_width"] == 4
        sysn.cheb_steps(2 * (launches - 1))                    # begin takes the first launch: T_{2 launches}
        got = sysn.cheb_vectors(n_cols, 0)
        moments = sysn.cheb_read(4 * launches, n_cols)
        sysn.cheb_begin(n_random=n_cols, seed=5, col_offset=3, scale=scale, kernel="dict_diag")
        sysn.cheb_steps(2 * launches - 1)
        want = sysn.cheb_vectors(n_cols, 0)
        ref = sysn.cheb_read(4 * launches, n_cols)
        sysn.cheb_end()
        assert np.max(np.abs(got - want)) <= 1e-13 * max(np.max(np.abs(want)), 1.0), (n_cols, launches)
        assert rel_err(moments, ref) <= 1e-12, (n_cols, launches)


@pytest.mark.parametrize("tag", sorted(SYSTEMS))
def test_cube_moments_match_the_oracle(gpu_api, monkeypatch, tag):
    system = SYSTEMS[tag](gpu_api)
    H = system.matrix("bsr")
    scale = system.spectral_bound()
    for plan in PLANS[:4]:
        _set_plan(monkeypatch, plan)
        for n_cols, n_moments in ((8, 48), (5, 47), (19, 50), (12, 4), (8, 2), (8, 5), (6, 1), (4, 33), (1, 20), (3, 7)):
            got = system.chebyshev_moments(n_moments, vectors=n_cols, seed=3, scale=scale, kernel="t2")
            assert system._sys.cheb_format()["kernel"] == "t2"
            want = orc.cheb_moments(H, orc.rademacher(3, H.shape[0], np.arange(n_cols)), n_moments, scale)
            assert got.shape == want.shape
            assert rel_err(got, want) <= TOL, (plan, n_cols, n_moments)
            assert np.array_equal(got, system.chebyshev_moments(n_moments, vectors=n_cols, seed=3, scale=scale, kernel="t2"))
    summed = system.chebyshev_moments(48, vectors=8, seed=3, scale=scale, kernel="t2", summed=True)
    assert rel_err(summed, system.chebyshev_moments(48, vectors=8, seed=3, scale=scale, kernel="t2").sum(axis=1)) <= 1e-13


def test_cube_probe_columns_observables_and_auto(gpu_api):
    """Unit start vectors spread one site per application of H~: halo errors show up as exact zeros / non-zeros in the
    wrong place.  The observables' default kernel is this one on three-dimensional lattices."""
    system = cases.junction(gpu_api, (9, 10, 11))
    H = system.matrix("bsr")
    scale = system.spectral_bound()
    sites = [(0, 0, 0), (8, 9, 10), (4, 7, 8), (4, 8, 7), (3, 0, 10), (5, 9, 0)]
    rows = [4 * system.lattice.index(s) + a for s in sites for a in (0, 3)]
    got = system.chebyshev_moments(64, rows=rows, scale=scale, kernel="t2")
    assert system._sys.cheb_format()["kernel"] == "t2" and system._sys.cheb_info()["panel_width"] == 4
    x0 = np.zeros((H.shape[0], len(rows)), dtype=np.complex128)
    x0[rows, np.arange(len(rows))] = 1.0
    assert rel_err(got, orc.cheb_moments(H, x0, 64, scale)) <= TOL
    E = np.linspace(-0.3, 0.3, 9)
    assert rel_err(system.ldos_map(sites[:2], E, kernel="t2"), system.ldos_map(sites[:2], E, kernel="dict_diag")) <= 1e-10
    small = cases.swave_3d(gpu_api, (5, 4, 4))
    F = small.free_energy(0.1, cuda=True, kernel="t2")         # exact trace: 320 columns = 80 panels of four
    assert small._sys.cheb_format()["kernel"] == "t2"
    assert abs(F - small.free_energy(0.1)) <= 1e-10 * abs(F)   # ... against the reference's dense algorithm
    # The observables' default picks it where its items fill the machine (>= 3/4 of the SMs get an 8 x 8 patch of a 4-column
    # panel without cutting x): 2000 columns on this lattice do, 8 do not.  The stepping API keeps T_n and T_{n-1}: single step.
    small.chebyshev_moments(16, vectors=2000, seed=1)
    assert small._sys.cheb_format()["kernel"] == "t2"
    small.chebyshev_moments(16, vectors=8, seed=1)
    assert small._sys.cheb_format()["kernel"] == "dict_diag"
    small._sys.cheb_begin(n_random=2000, seed=1, scale=10.0, kernel="auto")
    assert small._sys.cheb_format()["kernel"] == "dict_diag"
    small._sys.cheb_end()


def test_cube_declines_what_it_cannot_do(gpu_api, monkeypatch):
    scale = 10.0
    periodic = cases.random_periodic(gpu_api, (3, 5, 7), seed=11)._sys   # wrap-around bonds, no dictionary
    with pytest.raises(ValueError):
        periodic.cheb_begin(n_random=8, seed=1, scale=scale, kernel="t2")
    three_d = cases.swave_3d(gpu_api, (6, 5, 4))
    with pytest.raises(ValueError):                                     # the pair kernel (T_n and T_{n-1} kept) stays two-dimensional
        three_d._sys.cheb_begin(n_random=8, seed=1, scale=scale, kernel="pair")
    with three_d as (H, D):                                             # a wrap-around bond along y
        H[(2, 0, 1), (2, 4, 1)] = -1.0 * gpu_api.σ0
        H[(2, 4, 1), (2, 0, 1)] = -1.0 * gpu_api.σ0
    with pytest.raises(ValueError):
        three_d._sys.cheb_begin(n_random=8, seed=1, scale=scale, kernel="t2")
    Hm = three_d.matrix("bsr")
    got = three_d.chebyshev_moments(16, vectors=2000, seed=2)[:, :8]    # auto: the single-step kernel, however many columns
    assert three_d._sys.cheb_format()["kernel"] == "dict_diag"
    assert rel_err(got, orc.cheb_moments(Hm, orc.rademacher(2, Hm.shape[0], np.arange(8)), 16, three_d.spectral_bound())) <= TOL
    complex_hop = cases.swave_3d(gpu_api, (4, 4, 4))
    with complex_hop as (H, D):                                         # a hopping block that is not real-diagonal
        H[(1, 1, 1), (1, 1, 2)] = -1.0 * gpu_api.σ0 + 0.2j * gpu_api.σ1
        H[(1, 1, 2), (1, 1, 1)] = -1.0 * gpu_api.σ0 - 0.2j * gpu_api.σ1
    with pytest.raises(ValueError):
        complex_hop._sys.cheb_begin(n_random=8, seed=1, scale=scale, kernel="t2")
    plain = cases.swave_3d(gpu_api, (6, 5, 4))
    plain.chebyshev_moments(16, vectors=2000, seed=1)                  # 500 panels: enough items
    assert plain._sys.cheb_format()["kernel"] == "t2"
    monkeypatch.setenv("BDG_AUTO_CUBE", "0")                            # the preference can be switched off ...
    plain.chebyshev_moments(16, vectors=2000, seed=1)
    assert plain._sys.cheb_format()["kernel"] == "dict_diag"
    plain.chebyshev_moments(16, vectors=8, seed=1, kernel="t2")        # ... asking for the kernel by name still works
    assert plain._sys.cheb_format()["kernel"] == "t2"


def test_cube_follows_matrix_updates(gpu_api):
    """Incremental updates patch the direction codes of this kernel too (bdg_scatter -> ell_patch)."""
    system = cases.swave_3d(gpu_api, (6, 9, 10))
    scale = system.spectral_bound() * 1.3
    before = system.chebyshev_moments(32, vectors=8, seed=2, scale=scale, kernel="t2")
    builds = system._sys.stats()["native_builds"]
    with system as (H, D):
        H[(4, 4, 5), (4, 4, 5)] = 0.7 * gpu_api.σ0 + 0.2 * gpu_api.σ3
        H[(1, 8, 9), (1, 8, 9)] = 0.9 * gpu_api.σ0
    after = system.chebyshev_moments(32, vectors=8, seed=2, scale=scale, kernel="t2")
    assert system._sys.stats()["native_builds"] == builds              # patched in place
    assert not np.array_equal(before, after)
    Hm = system.matrix("bsr")
    assert rel_err(after, orc.cheb_moments(Hm, orc.rademacher(2, Hm.shape[0], np.arange(8)), 32, scale)) <= TOL


def test_cube_random_shapes_and_plans(monkeypatch):
    """Seeded sweep over lattice extents, patch shapes, segment lengths, column and launch counts."""
    import bodge_b200 as b
    from bodge_b200 import workloads

    rng = np.random.default_rng(4048)
    for case in range(16):
        shape = (int(rng.integers(2, 14)), int(rng.integers(2, 30)), int(rng.integers(2, 30)))
        build = workloads.junction if case % 2 else workloads.swave_3d
        system = b.Hamiltonian(b.CubicLattice(shape))
        assert system.fill(*build(shape)) == 0.0
        scale = system.spectral_bound()
        plan = (int(rng.integers(0, 3)) if rng.random() < 0.7 else None, int(rng.integers(1, shape[0] + 3)) if rng.random() < 0.7 else None)
        _set_plan(monkeypatch, plan)
        n_cols, n_mom = int(rng.integers(1, 20)), int(rng.integers(1, 40))
        ref = system.chebyshev_moments(n_mom, vectors=n_cols, seed=9, scale=scale, kernel="dict_diag")
        got = system.chebyshev_moments(n_mom, vectors=n_cols, seed=9, scale=scale, kernel="t2")
        assert system._sys.cheb_format()["kernel"] == "t2"
        assert rel_err(got, ref) <= 1e-12, (shape, plan, n_cols, n_mom)


def test_cube_full_size_C4():
    """C4 (64^3 sites, 8 columns = two 4-column panels): 40 moments against the single-step kernel, mu_0 = 4N exactly,
    bit-reproducible; the comparison with the CPU oracle is in test_gpu_fullsize.py."""
    import bodge_b200 as b
    from bodge_b200 import workloads

    c = workloads.CONFIGS["C4"]
    system = b.Hamiltonian(b.CubicLattice(c["shape"]))
    assert system.fill(*c["build"](c["shape"])) == 0.0
    scale = system.spectral_bound()
    single = system.chebyshev_moments(40, vectors=8, seed=1234, scale=scale, kernel="dict_diag")
    t2 = system.chebyshev_moments(40, vectors=8, seed=1234, scale=scale)
    assert system._sys.cheb_format()["kernel"] == "t2" and system._sys.cheb_info()["panel_width"] == 4
    assert rel_err(t2, single) <= 1e-12
    assert np.array_equal(t2[0], np.full(8, float(system.shape[0])))
    assert np.array_equal(t2, system.chebyshev_moments(40, vectors=8, seed=1234, scale=scale, kernel="t2"))
