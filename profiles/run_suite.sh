set -x
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 2000 --warmup 20 > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; tail -3 gpurun_out/bench_pair.err; cat gpurun_out/bench_pair.json
