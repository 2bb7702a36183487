"""BASELINE.json's configurations at FULL size, CUDA against the CPU oracle (not kernel against kernel).

Assembly: the exported BSR arrays of C5 (10^6 sites), C4 (64^3) and the two less repetitive C5 variants equal the
oracle's bit for bit (sha256 over ``indptr``, ``indices`` and ``data``).  Chebyshev: the first moments of every step
kernel the library can pick for that matrix agree with the oracle's literal three-term recursion (scipy
``bsr_matvecs`` on the oracle-assembled matrix, block rows split over the host cores) to <= 1e-10 norm-wise
(SURVEY H6).  Template: the reference's own CPU-vs-GPU test (tests/test_hamiltonian.py:389-425).
"""

import numpy as np
import pytest

from oracle import bdg_oracle as orc
from oracle import cpu_baseline as cb
from util import digest, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-10
N_MOMENTS = 18  # begin + 8 steps


@pytest.fixture(scope="module")
def full(request):
    """(system, oracle arrays, oracle moments) per config, built once per module run."""
    cache = {}

    def get(key, cols=8):
        if key not in cache:
            import bodge_b200 as b
            from bodge_b200 import workloads

            cache.clear()  # one 10^6-site system at a time (host and device memory)
            c = workloads.CONFIGS[key]
            packed = c["build"](c["shape"])
            system = b.Hamiltonian(b.CubicLattice(c["shape"]))
            assert system.fill(*packed) <= 1e-12
            ptr, idx, dat = cb.assemble(c["shape"], packed)
            scale = 1.01 * orc.norm_inf(ptr, idx, dat)
            x0 = orc.rademacher(1234, 4 * (len(ptr) - 1), np.arange(cols))
            want = cb.moments_parallel(ptr, idx, dat, scale, x0, N_MOMENTS)
            cache[key] = (system, (ptr, idx, dat), scale, want)
        return cache[key]

    yield get
    cache.clear()


@pytest.mark.parametrize("key", ["C5", "C4", "C5_disordered", "C5_random", "C5_periodic"])
def test_full_size_assembly_equals_the_oracle(full, key):
    system, (ptr, idx, dat), scale, _ = full(key)
    got_ptr, got_idx, got_dat = system._sys.export_bsr(True)
    assert got_ptr.dtype == np.int32 and got_idx.dtype == np.int32
    assert digest(got_ptr, got_idx) == digest(ptr, idx), "BSR structure differs from the oracle's"
    assert digest(got_dat) == digest(dat), "BSR values differ from the oracle's"
    assert abs(system.spectral_bound() - scale) <= 1e-12 * scale


@pytest.mark.parametrize("key,kernels", [
    ("C5", ["auto", "pair", "dict_diag", "dict", "ell", "dmma"]),
    ("C4", ["auto", "dict_diag", "dict", "ell", "dmma"]),   # auto = the even-vector recursion of cheb_cube.cu
    ("C5_disordered", ["auto", "pair", "dict_diag", "ell"]),
    ("C5_random", ["auto", "dmma"]),
    ("C5_periodic", ["auto", "pair", "dict_diag", "ell"]),   # wrap-around halos of 72 patches x 4 segments at full size
])
def test_full_size_moments_equal_the_oracle(full, key, kernels):
    system, _, scale, want = full(key)
    seen = set()
    for kernel in kernels:
        got = system.chebyshev_moments(N_MOMENTS, vectors=8, seed=1234, scale=scale, kernel=kernel)
        seen.add(system._sys.cheb_format()["kernel"])
        assert got.shape == want.shape
        assert rel_err(got, want) <= TOL, f"{key} kernel {kernel} ({system._sys.cheb_format()['kernel']})"
        assert np.array_equal(got[0], want[0])  # <x|x> = 4N exactly
    if key == "C5":
        assert {"t2", "pair", "dict_diag", "dict", "ell", "dmma"} <= seen
    if key == "C4":
        assert {"t2", "dict_diag", "dict", "ell", "dmma"} <= seen
    if key == "C5_random":
        assert "ell" in seen  # nothing repeats: the matrix is streamed
    if key == "C5_periodic":
        assert {"t2", "pair"} <= seen  # the periodic stencil runs the two-step kernels


def test_C2_256_columns_equal_the_oracle():
    """C2 (100x100 README s-wave), the 256 stochastic columns of SURVEY 8d, 32 moments."""
    import bodge_b200 as b
    from bodge_b200 import workloads

    c = workloads.CONFIGS["C2"]
    packed = c["build"](c["shape"])
    system = b.Hamiltonian(b.CubicLattice(c["shape"]))
    system.fill(*packed)
    ptr, idx, dat = cb.assemble(c["shape"], packed)
    H = orc.to_scipy(ptr, idx, dat)
    scale = system.spectral_bound()
    want = orc.cheb_moments(H, orc.rademacher(1234, H.shape[0], np.arange(256)), 32, scale)
    for kernel in ("auto", "pair", "dict_diag", "ell"):
        got = system.chebyshev_moments(32, vectors=256, seed=1234, scale=scale, kernel=kernel)
        assert rel_err(got, want) <= TOL, kernel


def test_C3_probe_columns_equal_the_oracle():
    """C3 (100x100 d-wave + Rashba): 64 probe columns (16 of the 1024 LDOS sites x 4 components), 32 moments."""
    import bodge_b200 as b
    from bodge_b200 import workloads

    c = workloads.CONFIGS["C3"]
    packed = c["build"](c["shape"])
    system = b.Hamiltonian(b.CubicLattice(c["shape"]))
    system.fill(*packed)
    ptr, idx, dat = cb.assemble(c["shape"], packed)
    H = orc.to_scipy(ptr, idx, dat)
    scale = system.spectral_bound()
    sites = [(3 * p + 2, 3 * q + 2, 0) for p in range(0, 32, 8) for q in range(0, 32, 8)]
    rows = system._probe_rows(sites)
    want = orc.cheb_moments(H, orc.probes(H.shape[0], rows), 32, scale)
    for kernel in ("auto", "pair", "t2", "dict", "ell", "dmma"):
        got = system.chebyshev_moments(32, rows=rows, scale=scale, kernel=kernel)
        assert rel_err(got, want) <= TOL, kernel
