# round 2, call 20: soak -- the two-step kernel and incremental-update tests five times over (flakiness check)
set -x
mkdir -p gpurun_out/r02
for i in 1 2 3 4 5; do
  timeout 600 python -m pytest tests/test_gpu_pair.py tests/test_gpu_incremental.py -x -q -p no:cacheprovider 2>&1 | tail -1 | tee -a gpurun_out/r02/20_soak.log
done
BDG_PAIR_WARPS=16 timeout 600 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -1 | tee -a gpurun_out/r02/20_soak.log
BDG_PAIR_WARPS=12 timeout 600 python -m pytest tests/test_gpu_pair.py tests/test_gpu_incremental.py -x -q 2>&1 | tail -1 | tee -a gpurun_out/r02/20_soak.log
