"""Where the end-to-end time of one [fill + chebyshev_moments] call goes (C5, pinned host inputs)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bodge_b200 as b
from bodge_b200 import workloads

cfg = workloads.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C5"]
packed = cfg["build"](cfg["shape"])
host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in packed]
system = b.Hamiltonian(b.CubicLattice(cfg["shape"]))
s = system._sys
def T(label, fn):
    s.sync(); t = time.perf_counter(); r = fn(); s.sync(); print(f"  {label:28s} {1e3 * (time.perf_counter() - t):9.2f} ms"); return r
for rep in range(3):
    print("call", rep)
    T("fill (H2D+scatter+check)", lambda: system.fill(*host))
    scale = T("spectral_bound", lambda: system.spectral_bound())
    T("cheb_begin (formats+init)", lambda: s.cheb_begin(n_random=8, seed=1234, scale=scale, kernel="auto_moments"))
    T("1024 steps", lambda: s.cheb_steps(1024))
    T("read moments", lambda: s.cheb_read(2050, 8, summed=True))
