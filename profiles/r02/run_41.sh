# round 2, call 41: SD (real-diagonal on-site blocks by multiplication) in the single-step and two-step kernels: tests, C3 timing of every kernel
set -x
mkdir -p gpurun_out/r02
( timeout 1500 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_pair.py tests/test_gpu_incremental.py -q 2>&1 | tail -5 ) | tee gpurun_out/r02/41_pytest.log
for sd in 0 1; do
  echo "== BDG_ELL_SD=$sd"
  BDG_ELL_SD=$sd QP_STEPS=400 timeout 300 python profiles/quickperf2.py C3:512:dict,pair,t2 C3:4096:dict,t2 2>&1 | cut -c1-200
  BDG_ELL_SD=$sd QP_STEPS=3000 timeout 300 python profiles/quickperf2.py C3:512:dict,t2 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/r02/41_quickperf_c3_sd.log
