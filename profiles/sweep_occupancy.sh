#!/bin/bash
# Occupancy scaling of the DICT / ELL step kernels: unused dynamic shared memory caps CTAs per SM.
for pad in 0 36000 44000 55000 74000 110000; do
  echo "PAD=$pad"
  BDG_ELL_PAD=$pad python profiles/quickperf2.py C5:8:dict,ell 2>&1 | cut -c1-175
done
echo "L2-resident lattice (256,256,1), k=8"
python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd())
import bodge_b200 as b
from bodge_b200 import workloads
for shape in ((256, 256, 1), (360, 360, 1)):
    system = b.Hamiltonian(b.CubicLattice(shape)); system.fill(*workloads.junction(shape))
    s = system._sys
    for kernel in ("dict", "ell"):
        s.cheb_begin(n_random=8, seed=1, scale=system.spectral_bound(), kernel=kernel)
        s.cheb_steps(20, timed=True); ms = s.cheb_steps(400, timed=True) / 400
        n = shape[0] * shape[1]
        print(shape, kernel, "ms/step", round(ms, 5), "ns per row", round(ms * 1e6 / n, 3), "vector GB/s", round(192 * n * 8 / ms / 1e6))
PY
