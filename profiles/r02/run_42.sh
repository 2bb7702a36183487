# round 2, call 42: which test failed; t2 vs dict for complex hopping blocks at other sizes (policy of auto_moments)
set -x
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_pair.py tests/test_gpu_incremental.py -q -x 2>&1 | grep -v Warning | tail -30 > gpurun_out/r02/42_pytest_fail.log
QP_STEPS=400 timeout 300 python profiles/quickperf2.py C3:8:dict,t2 C3:64:dict,t2 C5_dwave:8:dict,pair,t2 C5_dwave:64:dict,t2 2>&1 | cut -c1-200 | tee gpurun_out/r02/42_quickperf_dwave_policy.log
