# round 2, call 46: the five BASELINE configs end to end through the public API on the final tree
set -x
mkdir -p gpurun_out/r02
timeout 900 python -W ignore profiles/bench_configs.py 2>&1 | tee gpurun_out/r02/46_configs_end_to_end.jsonl | cut -c1-330
