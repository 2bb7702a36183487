# round 2, call 47: rim patches own their rim site on planes with open ends: two-step kernel tests, C3 / C2 / C5 timing with and without
set -x
mkdir -p gpurun_out/r02
( timeout 1500 python -m pytest tests/test_gpu_pair.py tests/test_gpu_incremental.py -q -x 2>&1 | grep -v Warning | tail -5 ) | tee gpurun_out/r02/47_pytest.log
for open in 0 1; do
  echo "== BDG_PAIR_OPEN=$open"
  BDG_PAIR_OPEN=$open QP_STEPS=400 timeout 300 python profiles/quickperf2.py C3:512:t2 C3:4096:t2 C2:256:t2 C5:8:t2 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/r02/47_quickperf_open_rim.log
