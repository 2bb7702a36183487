python profiles/assembly_phases.py C5 2>&1 | tail -3
python - <<'PY'
import time, sys, os
sys.path.insert(0, os.getcwd())
import bodge_b200 as b
from bodge_b200 import _native
for rep in range(4):
    t0=time.perf_counter(); s=_native.System.cubic((1000,1000,1),0); t1=time.perf_counter(); s.sync(); t2=time.perf_counter()
    del s; t3=time.perf_counter()
    print(f"native create {1e3*(t1-t0):.2f} ms, sync {1e3*(t2-t1):.2f}, destroy {1e3*(t3-t2):.2f}")
lat=b.CubicLattice((1000,1000,1))
t0=time.perf_counter(); h=b.Hamiltonian(lat); t1=time.perf_counter(); print(f"Hamiltonian() {1e3*(t1-t0):.2f} ms")
PY
