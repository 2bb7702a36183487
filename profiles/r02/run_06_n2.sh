# round 2, call 6 (2 GPUs): torch.distributed column sharding + the C ABI's own multi-GPU entry point; bench at N = 2
set -x
mkdir -p gpurun_out/r02
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r02/06_pytest_multi.log
( time NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r02/06_bench_n2.json 2> gpurun_out/r02/06_bench_n2.err ); tail -3 gpurun_out/r02/06_bench_n2.err; grep -c "NCCL INFO" gpurun_out/r02/06_bench_n2.err; grep -m2 "nranks" gpurun_out/r02/06_bench_n2.err; cut -c1-400 gpurun_out/r02/06_bench_n2.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02/06_bench_n2.json").read())
print({k: d[k] for k in ("value", "n_gpus", "repeats", "timed_s", "parity_check", "strong_scaling")})
print(d["other_configs"]["C3"])
PY
