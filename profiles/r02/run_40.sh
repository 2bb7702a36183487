set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_incremental.py -x -q 2>&1 | tail -40 > gpurun_out/r02/40_pytest_fail.log
