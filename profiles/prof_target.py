"""Minimal driver for ncu captures: python profiles/prof_target.py <config> <k> <kernel> <steps>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bodge_b200 as b
from bodge_b200 import workloads

cfg, k, kernel, steps = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
c = workloads.CONFIGS[cfg]
system = b.Hamiltonian(b.CubicLattice(c["shape"]))
system.fill(*c["build"](c["shape"]))
s = system._sys
s.cheb_begin(n_random=k, seed=1234, scale=system.spectral_bound(), kernel=kernel)
ms = s.cheb_steps(steps, timed=True)
print(cfg, k, kernel, "ms/step", ms / steps)
