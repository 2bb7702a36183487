"""Is the two-step kernel waiting for HBM?  The same CTA work (72 patches x 4 segments, 1000 sites per plane) on a lattice
whose two vector sets fit the 126 MB L2 (96 planes: 98 MB) against lattices that stream from HBM: time per plane-iteration."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bodge_b200 as b
from bodge_b200 import workloads

for Lx in (48, 96, 192, 400, 1000):
    shape = (Lx, 1000, 1)
    system = b.Hamiltonian(b.CubicLattice(shape))
    system.fill(*workloads.junction(shape))
    scale = system.spectral_bound()
    s = system._sys
    os.environ["BDG_PAIR_SEG"] = str(Lx // 4)
    for kernel in ("t2", "dict_diag"):
        s.cheb_begin(n_random=8, seed=1, scale=scale, kernel=kernel)
        s.cheb_steps(40, timed=True)
        steps = 2000 if Lx < 400 else 400
        ms = s.cheb_steps(steps, timed=True) / steps
        iters = Lx // 4 + 2
        print(json.dumps(dict(kernel=kernel, Lx=Lx, vectors_MB=round(2 * 64 * Lx * 1000 * 8 / 1e6), ms_per_step=round(ms, 5),
                              us_per_plane_iteration=round(2 * ms * 1e3 / iters, 4) if kernel == "t2" else None,
                              ns_per_site_step=round(ms * 1e6 / (Lx * 1000), 4))), flush=True)
    del system
