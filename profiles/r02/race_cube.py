"""Tiny runs of the three-dimensional even-vector kernel (csrc/cheb_cube.cu) for compute-sanitizer racecheck / memcheck:
ragged patches, several patches per plane, every patch shape (BDG_CUBE_SHAPE), one-plane segments (BDG_CUBE_SEG), sites with
different on-site blocks in one warp pair, cut bonds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import types
import numpy as np
import bodge_b200 as b
import cases
import test_gpu_cube as t
from oracle import bdg_oracle as orc

api = types.SimpleNamespace(CubicLattice=b.CubicLattice, Hamiltonian=b.Hamiltonian, σ0=b.σ0, σ1=b.σ1, σ2=b.σ2, σ3=b.σ3, jσ2=b.jσ2, dwave=b.dwave)
for name, system in (("swave", cases.swave_3d(api, (4, 5, 4))), ("junction", cases.junction(api, (5, 10, 9))),
                     ("patterned", t._patterned(api, (3, 9, 10)))):
    H = system.matrix("bsr")
    scale = system.spectral_bound()
    want = orc.cheb_moments(H, orc.rademacher(3, H.shape[0], np.arange(8)), 12, scale)
    got = system.chebyshev_moments(12, vectors=8, seed=3, scale=scale, kernel="t2")
    err = np.max(np.abs(got - want)) / np.max(np.abs(want))
    print(name, system.lattice.shape, system._sys.cheb_format(), system._sys.cheb_info()["panel_width"], "rel err", err, flush=True)
    assert err < 1e-10 and system._sys.cheb_info()["panel_width"] == 4
print("race_cube ok")
