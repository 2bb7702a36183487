"""Where the assembly time goes at C5 (packed arrays in pinned host memory -> device BSR)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bodge_b200 as b
from bodge_b200 import workloads

cfg = workloads.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C5"]
packed = cfg["build"](cfg["shape"])
host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in packed]
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    system = b.Hamiltonian(b.CubicLattice(cfg["shape"]))
    system._sys.sync(); t1 = time.perf_counter()
    system.fill(*host); system._sys.sync(); t2 = time.perf_counter()
    system.fill(*host); system._sys.sync(); t3 = time.perf_counter()
    info = system._sys.cheb_info(); system._sys.sync(); t4 = time.perf_counter()
    ex = system._sys.export_bsr(True); t5 = time.perf_counter()
    n = system.lattice.size
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms, first fill {1e3*(t2-t1):.1f} ms, second fill {1e3*(t3-t2):.1f} ms, "
          f"compaction {1e3*(t4-t3):.1f} ms, export D2H {1e3*(t5-t4):.1f} ms -> {n/(t2-t0)/1e6:.2f} M sites/s create+fill")
    del system
