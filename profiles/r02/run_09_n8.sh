# round 2, call 9 (8 GPUs): the bench line at N = 8 exactly as the driver launches it (weak scaling headline, strong scaling,
# sharded C3 LDOS with gather check, parity of the all-reduce), and the C ABI's multi-GPU entry point on 8 GPUs
set -x
mkdir -p gpurun_out/r02
nvidia-smi -L | wc -l
( time NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02/09_bench_n8.json 2> gpurun_out/r02/09_bench_n8.err ); tail -2 gpurun_out/r02/09_bench_n8.err | cut -c1-300; grep -c "nranks 8" gpurun_out/r02/09_bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02/09_bench_n8.json").read())
print({k: d[k] for k in ("value", "n_gpus", "repeats", "timed_s", "parity_check")})
print("strong", {k: d["strong_scaling"][k] for k in ("cols_per_gpu", "steps_per_s", "frac")})
print("C3", {k: d["other_configs"]["C3"][k] for k in ("cols_per_gpu", "steps_per_s", "gather_check")})
print("C4", d["other_configs"]["C4"]["steps_per_s"], "C2", d["other_configs"]["C2"]["steps_per_s"], "e2e", d["e2e"]["value"])
PY
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, ".")
import bodge_b200 as b
from bodge_b200 import workloads
shape = (1000, 1000, 1)
packed = workloads.junction(shape)
for n in (1, 2, 4, 8):
    reps = b.Replicas(b.CubicLattice(shape), devices=list(range(n)))
    reps.fill(*packed)
    scale = reps.spectral_bound()
    reps.chebyshev_moments(64, vectors=64, scale=scale, summed=True)      # builds, communicators
    t0 = time.perf_counter()
    mu = reps.chebyshev_moments(2048, vectors=64, scale=scale, summed=True)
    dt = time.perf_counter() - t0
    print(f"C ABI multi: {n} GPUs, 64 columns x 2048 moments in {dt:.3f} s  mu0={mu[0]:.1f} mu2={mu[2]:.6e}", flush=True)
    del reps
PY
