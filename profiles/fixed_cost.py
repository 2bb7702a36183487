"""Per-launch fixed cost of the two-step kernel: time per launch against the number of planes a CTA marches (Lx at fixed
plane size and a fixed number of segments), T = a + b * planes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bodge_b200 as b
from bodge_b200 import workloads

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rows = []
for Lx in (400, 1000, 2000, 4000):
    shape = (Lx, M, 1)
    system = b.Hamiltonian(b.CubicLattice(shape))
    system.fill(*workloads.junction(shape))
    scale = system.spectral_bound()
    s = system._sys
    for kernel in ("t2", "pair"):
        os.environ["BDG_PAIR_SEG"] = str(Lx // 4)          # always 4 segments x 72 patches = 288 CTAs
        s.cheb_begin(n_random=8, seed=1, scale=scale, kernel=kernel)
        s.cheb_steps(20, timed=True)
        ms = s.cheb_steps(400, timed=True) / 200            # per launch (two steps)
        rows.append((kernel, Lx // 4, ms))
        print(json.dumps(dict(kernel=kernel, Lx=Lx, planes_per_cta=Lx // 4, ms_per_launch=round(ms, 5))), flush=True)
    del system
for kernel in ("t2", "pair"):
    x = np.array([r[1] for r in rows if r[0] == kernel], float)
    y = np.array([r[2] for r in rows if r[0] == kernel], float)
    bfit, afit = np.polyfit(x, y, 1)
    print(json.dumps(dict(kernel=kernel, fixed_ms_per_launch=round(afit, 5), ms_per_plane=round(bfit, 7),
                          fixed_share_at_250_planes=round(afit / (afit + 250 * bfit), 4))))
