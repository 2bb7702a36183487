"""Tiny two-step-kernel runs for compute-sanitizer racecheck / memcheck (seconds under the tool): periodic and open
lattices, ragged patches (BDG_PAIR_P), one-plane segments (BDG_PAIR_SEG), both recursions, held and streamed on-site blocks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import bodge_b200 as b
import types
import test_gpu_pair as t
from oracle import bdg_oracle as orc

api = types.SimpleNamespace(CubicLattice=b.CubicLattice, Hamiltonian=b.Hamiltonian, σ0=b.σ0, σ1=b.σ1, σ2=b.σ2, σ3=b.σ3, jσ2=b.jσ2, dwave=b.dwave)
for name, kw in (("torus", dict(shape=(5, 7, 1))), ("torus disordered", dict(shape=(6, 9, 1), disorder=True)), ("open", dict(shape=(4, 1, 11), x=False, y=False)),
                 # MMA rows with real-diagonal on-site blocks (SD): held / streamed per row
                 ("torus normal random hop", dict(shape=(5, 8, 1), random_hop=True, onsite_pairing=False)),
                 ("open normal disordered", dict(shape=(6, 9, 1), x=False, y=False, disorder=True, random_hop=True, onsite_pairing=False))):
    shape = kw.pop("shape")
    system = t._periodic(api, shape, **kw)
    H = system.matrix("bsr")
    scale = system.spectral_bound()
    want = orc.cheb_moments(H, orc.rademacher(3, H.shape[0], np.arange(8)), 12, scale)
    for kernel in ("pair", "t2"):
        got = system.chebyshev_moments(12, vectors=8, seed=3, scale=scale, kernel=kernel)
        err = np.max(np.abs(got - want)) / np.max(np.abs(want))
        print(name, shape, kernel, system._sys.cheb_format(), "rel err", err, flush=True)
        assert err < 1e-10
print("race_small ok")
