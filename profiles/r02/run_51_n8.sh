# round 2, call 51 (8 GPUs): the bench line at N = 8 on the final tree, launched as the driver launches it
set -x
mkdir -p gpurun_out/r02
nvidia-smi -L | wc -l
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02/51_bench_n8.json 2> gpurun_out/r02/51_bench_n8.err ); tail -2 gpurun_out/r02/51_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02/51_bench_n8.json").read())
print({k: d[k] for k in ("value", "n_gpus", "repeats", "timed_s", "parity_check")})
print("strong", {k: d["strong_scaling"][k] for k in ("cols_per_gpu", "steps_per_s", "frac")})
for k, v in d["other_configs"].items():
    print(k, {x: v[x] for x in ("kernel", "cols_per_gpu", "steps_per_s", "frac", "gather_check") if x in v})
PY
