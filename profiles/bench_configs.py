"""The five BASELINE.json configs end to end through the public API (one JSON line each).
Parity references: tests/golden/observables.npz (the reference's own free_energy at C1)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bodge_b200 as b
from bodge_b200 import workloads

gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "observables.npz")))

def build(cfg):
    c = workloads.CONFIGS[cfg]
    t0 = time.perf_counter()
    system = b.Hamiltonian(b.CubicLattice(c["shape"]))
    system.fill(*c["build"](c["shape"]))
    system._sys.sync()
    return system, time.perf_counter() - t0

def timed(fn):
    t0 = time.perf_counter(); r = fn(); return r, time.perf_counter() - t0

def emit(**kw):
    print(json.dumps(kw), flush=True)

# C1: exact-trace free energy vs the reference's dense eigvalsh (23 s on 8 cores)
system, t_asm = build("C1")
for T, F_ref in zip(gold["temps"], gold["F_C1"]):
    if T < 0.05:
        continue
    system.free_energy(float(T), cuda=True)  # warm-up (buffers, formats)
    F, dt = timed(lambda: system.free_energy(float(T), cuda=True))
    emit(config="C1", what=f"free_energy(T={T}, cuda=True), exact trace over 6400 columns", seconds=dt, F=F, F_reference=float(F_ref),
         rel_err=abs(F - F_ref) / abs(F_ref), kernel=system._sys.cheb_format()["kernel"], reference_seconds="~23 (dense eigvalsh, 8 cores, SURVEY 6.2)")

# C2: 256 stochastic columns, 1024 moments
system, t_asm = build("C2")
system.chebyshev_moments(64, vectors=256, summed=True)
mu, dt = timed(lambda: system.chebyshev_moments(1024, vectors=256, summed=True))
emit(config="C2", what="chebyshev_moments(1024, vectors=256)", seconds=dt, steps_per_s=511 / dt, assembly_seconds=t_asm, kernel=system._sys.cheb_format()["kernel"])

# C3: LDOS by KPM at 1024 probe sites x 101 energies (the reference: one SuperLU solve per energy and site)
system, t_asm = build("C3")
sites = [(3 * p + 2, 3 * q + 2, 0) for p in range(32) for q in range(32)]
energies = np.linspace(-0.15, 0.15, 101)
system.ldos_map(sites[:8], energies, moments=256)
rho, dt = timed(lambda: system.ldos_map(sites, energies, moments=4096))
emit(config="C3", what="ldos_map(1024 sites, 101 energies, moments=4096): 4096 probe columns", seconds=dt, column_steps_per_s=4096 * 2047 / dt,
     min_ldos=float(rho.min()), shape=list(rho.shape), kernel=system._sys.cheb_format()["kernel"])
rho_default, dt = timed(lambda: system.ldos_map(sites[:64], energies))
emit(config="C3", what="ldos_map(64 sites, 101 energies), default moments from the broadening", seconds=dt, min_ldos=float(rho_default.min()))

# C4 / C5: stochastic-trace free energy, R = 64 columns, 2048 moments
for cfg, T in (("C4", 0.1), ("C5", 0.05)):
    system, t_asm = build(cfg)
    system.free_energy(T, cuda=True, vectors=8, moments=64)
    F, dt = timed(lambda: system.free_energy(T, cuda=True, vectors=64, moments=2048))
    emit(config=cfg, what=f"free_energy(T={T}, cuda=True, vectors=64, moments=2048)", seconds=dt, F_per_site=F / system.lattice.size,
         column_steps_per_s=64 * 1023 / dt, assembly_seconds=t_asm, kernel=system._sys.cheb_format()["kernel"])
    tight, dt = timed(lambda: system.spectral_bound("lanczos"))
    emit(config=cfg, what="spectral_bound('lanczos')", seconds=dt, norm_bound=system.spectral_bound(), lanczos_bound=tight)
    del system
