"""Device-only timing of the fused Chebyshev step per kernel / matrix format (development aid;
bench.py is the judged harness).  usage: quickperf2.py C5:8:dict,ell C4:8:dict ..."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bodge_b200 as b
from bodge_b200 import workloads

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
systems = {}
def get(cfg):
    if cfg not in systems:
        systems.clear()
        c = workloads.CONFIGS[cfg]
        system = b.Hamiltonian(b.CubicLattice(c["shape"]))
        system.fill(*c["build"](c["shape"]))
        systems[cfg] = system
    return systems[cfg]

for item in sys.argv[1:]:
    cfg, k, kernels = item.split(":")
    system = get(cfg)
    scale = system.spectral_bound()
    s = system._sys
    for kernel in kernels.split(","):
        t0 = time.time()
        s.cheb_begin(n_random=int(k), seed=1234, scale=scale, kernel=kernel)
        s.sync()
        t_begin = time.time() - t0
        s.cheb_steps(10, timed=True)
        steps = int(os.environ.get("QP_STEPS", "100"))  # 100 = burst clocks; >= 2000 = sustained under the power cap
        ms = s.cheb_steps(steps, timed=True)
        info, fmt = s.cheb_info(), s.cheb_format()
        gbs = info["bytes_per_step"] * steps / (ms * 1e-3) / 1e9
        passes = {"pair": 128, "t2": 96}.get(fmt["kernel"], 192)   # pair: 4 vector passes per 2 steps, t2: 3
        actual = (fmt["matrix_bytes_per_step"] + passes * system.lattice.size * int(k)) * steps / (ms * 1e-3) / 1e9
        print(json.dumps(dict(cfg=cfg, k=int(k), kernel=fmt["kernel"], np=os.environ.get("BDG_ELL_NP"), ms_per_step=round(ms / steps, 5),
                              steps_per_s=round(steps / (ms * 1e-3), 1), alg_GBps=round(gbs), frac=round(gbs / PEAK, 4),
                              actual_GBps=round(actual), actual_frac=round(actual / PEAK, 4), distinct=fmt["distinct_blocks"],
                              begin_s=round(t_begin, 4))), flush=True)
        s.cheb_end()
