import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bodge_b200 as b
from bodge_b200 import workloads
cfg = workloads.CONFIGS["C5"]
packed = cfg["build"](cfg["shape"])
host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in packed]
system = None
for rep in range(10):
    del system
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    system = b.Hamiltonian(b.CubicLattice(cfg["shape"])); system._sys.sync(); t1 = time.perf_counter()
    system.fill(*host); system._sys.sync(); t2 = time.perf_counter()
    system._sys.cheb_info(); system._sys.sync(); t3 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):6.2f}  fill {1e3*(t2-t1):6.2f}  compaction {1e3*(t3-t2):6.2f} ms", flush=True)
