# round 2, call 18 (4 GPUs): the missing point of the scaling curve, as the driver launches it
set -x
mkdir -p gpurun_out/r02
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02/18_bench_n4.json 2> gpurun_out/r02/18_bench_n4.err ); tail -2 gpurun_out/r02/18_bench_n4.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02/18_bench_n4.json").read())
print({k: d[k] for k in ("value", "n_gpus", "repeats", "timed_s", "parity_check")})
print("strong", {k: d["strong_scaling"][k] for k in ("cols_per_gpu", "steps_per_s", "frac")})
print("C3", {k: d["other_configs"]["C3"][k] for k in ("cols_per_gpu", "steps_per_s", "gather_check")}, "e2e", d["e2e"]["value"])
PY
