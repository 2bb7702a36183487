# round 2, call 76 (2 GPUs): multi-GPU tests and the bench at N = 2 on the final tree
set -x
mkdir -p gpurun_out/r02
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r02/76_pytest_multi.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r02/76_bench_n2.json 2> gpurun_out/r02/76_bench_n2.err ); tail -3 gpurun_out/r02/76_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02/76_bench_n2.json").read())
print({k: d[k] for k in ("value", "n_gpus", "repeats", "timed_s", "parity_check", "strong_scaling")})
for k, v in d["other_configs"].items():
    print(k, {x: v[x] for x in ("kernel", "steps_per_s", "frac", "gather_check") if x in v})
PY
