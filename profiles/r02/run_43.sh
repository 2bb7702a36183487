# round 2, call 43: tests after the auto policy for SD matrices; two-step vs single-step for ten-MMA rows (no SD) at other sizes
set -x
mkdir -p gpurun_out/r02
( timeout 1500 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_pair.py tests/test_gpu_incremental.py tests/test_gpu_fullsize.py -q 2>&1 | grep -v Warning | tail -6 ) | tee gpurun_out/r02/43_pytest.log
BDG_ELL_SD=0 QP_STEPS=400 timeout 300 python profiles/quickperf2.py C3:8:dict,t2 C3:64:dict,t2 C5_dwave:8:dict,pair,t2 C5_dwave:64:dict,t2 2>&1 | cut -c1-200 | tee gpurun_out/r02/43_quickperf_ten_mma_rows.log
