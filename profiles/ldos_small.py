"""ldos() of one site (4 probe columns) on the README model at 100x100: the launch-bound small-lattice case."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bodge_b200 as b
from bodge_b200 import workloads

system = b.Hamiltonian(b.CubicLattice((100, 100, 1)))
system.fill(*workloads.readme_swave((100, 100, 1)))
E = np.linspace(-0.15, 0.15, 41)
for rep in range(3):
    t0 = time.perf_counter(); rho = system.ldos((50, 50, 0), E, moments=8192); dt = time.perf_counter() - t0
    print(f"ldos one site, 8192 moments: {1e3 * dt:.1f} ms  kernel={system._sys.cheb_format()['kernel']}  rho[20]={rho[20]:.12f}", flush=True)
