"""Quick device-only timing of the fused Chebyshev step for both kernel formulations
(development aid; bench.py is the judged harness)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bodge_b200 as b
from bodge_b200 import workloads

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
def run(cfg, k, kernels=tuple(os.environ.get("QP_KERNELS", "ell,dmma").split(",")), steps=50):
    c = workloads.CONFIGS[cfg]
    t0 = time.time()
    packed = c["build"](c["shape"])
    t1 = time.time()
    system = b.Hamiltonian(b.CubicLattice(c["shape"]))
    t2 = time.time()
    system.fill(*packed)
    system._sys.sync()
    t3 = time.time()
    scale = system.spectral_bound()
    for kernel in kernels:
        s = system._sys
        s.cheb_begin(n_random=k, seed=1234, scale=scale, kernel=kernel)
        s.cheb_steps(5, timed=True)
        ms = s.cheb_steps(steps, timed=True)
        info = s.cheb_info()
        gbs = info["bytes_per_step"] * steps / (ms * 1e-3) / 1e9
        print(json.dumps(dict(cfg=cfg, k=k, kernel=kernel, np=os.environ.get("BDG_ELL_NP"), pb=os.environ.get("BDG_ELL_PB"), ms_per_step=ms / steps, steps_per_s=steps / (ms * 1e-3),
                              GBps=gbs, frac=gbs / PEAK, nb=info["n_blocks"], panels=info["n_panels"], pw=info["panel_width"],
                              host_gen_s=t1 - t0, skeleton_s=t2 - t1, fill_s=t3 - t2)), flush=True)
        s.cheb_end()

if __name__ == "__main__":
    sel = sys.argv[1:] or ["C5:8", "C4:8", "C2:256", "C3:512", "C5:1", "C5:4", "C5:16", "C5:32", "C5:64", "C4:64"]
    for item in sel:
        cfg, k = item.split(":")
        run(cfg, int(k))
