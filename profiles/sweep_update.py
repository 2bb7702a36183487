"""Cost of the reference's parameter-sweep idiom at 10^6 sites: re-enter `with` for the on-site terms of
half the lattice, then start the next recursion.  Patched in place vs rebuilt (BDG_NO_PATCH=1)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bodge_b200 as b
from bodge_b200 import workloads

cfg = sys.argv[1] if len(sys.argv) > 1 else "C5"
c = workloads.CONFIGS[cfg]
shape = c["shape"]
system = b.Hamiltonian(b.CubicLattice(shape))
system.fill(*c["build"](shape))
scale = system.spectral_bound()
s = system._sys
n = system.lattice.size
half = np.arange(n // 2, dtype=np.int32)


def begin():
    s.cheb_begin(n_random=8, seed=1, scale=scale, kernel="auto_moments")
    s.sync()


begin()
for mode in ("patch", "rebuild"):
    if mode == "rebuild":
        os.environ["BDG_NO_PATCH"] = "1"
    rows = []
    for rep in range(4):
        vals = np.broadcast_to((3.0 + 0.01 * rep) * b.σ0 - 0.3 * b.σ3, (len(half), 2, 2)).astype(np.complex128).copy()
        s.sync()
        t0 = time.perf_counter()
        system.fill(half, half, vals)
        s.sync()
        t1 = time.perf_counter()
        begin()
        t2 = time.perf_counter()
        rows.append((t1 - t0, t2 - t1))
    st = s.stats()
    fill_ms, begin_ms = (1e3 * float(np.median([r[k] for r in rows[1:]])) for k in (0, 1))
    print(json.dumps(dict(cfg=cfg, mode=mode, entries=len(half), h2d_MB=round(len(half) * 72 / 1e6, 1), fill_ms=round(fill_ms, 3),
                          next_begin_ms=round(begin_ms, 3), total_ms=round(fill_ms + begin_ms, 3), stats=st)), flush=True)
# a handful of terms
os.environ.pop("BDG_NO_PATCH", None)
few = np.arange(0, n, n // 16, dtype=np.int32)[:16]
for mode in ("patch", "rebuild"):
    if mode == "rebuild":
        os.environ["BDG_NO_PATCH"] = "1"
    rows = []
    for rep in range(4):
        vals = np.broadcast_to((3.0 + 0.02 * rep) * b.σ0, (len(few), 2, 2)).astype(np.complex128).copy()
        t0 = time.perf_counter()
        system.fill(few, few, vals)
        s.sync()
        t1 = time.perf_counter()
        begin()
        t2 = time.perf_counter()
        rows.append((t1 - t0, t2 - t1))
    fill_ms, begin_ms = (1e3 * float(np.median([r[k] for r in rows[1:]])) for k in (0, 1))
    print(json.dumps(dict(cfg=cfg, mode=mode, entries=len(few), fill_ms=round(fill_ms, 3), next_begin_ms=round(begin_ms, 3),
                          total_ms=round(fill_ms + begin_ms, 3))), flush=True)
