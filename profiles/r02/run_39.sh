# round 2, call 39: real-diagonal on-site blocks by two multiplications (SD) in the single-step dictionary kernel: tests, C3 timing
set -x
mkdir -p gpurun_out/r02
( timeout 1200 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_incremental.py -x -q 2>&1 | tail -3 ) | tee gpurun_out/r02/39_pytest.log
for sd in 0 1; do
  echo "== BDG_ELL_SD=$sd"
  BDG_ELL_SD=$sd QP_STEPS=400 timeout 300 python profiles/quickperf2.py C3:512:dict C3:4096:dict 2>&1 | cut -c1-200
  BDG_ELL_SD=$sd QP_STEPS=3000 timeout 300 python profiles/quickperf2.py C3:512:dict 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/r02/39_quickperf_c3_sd.log
