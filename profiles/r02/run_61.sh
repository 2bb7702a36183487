# round 2, call 61: work lists in the two-step kernel (classic plan as lists; balanced plan = one contiguous chunk of the (panel, patch, x) space per CTA slot): tests, racecheck, timing
set -x
mkdir -p gpurun_out/r02
( timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/61_pytest_pair.log
( BDG_PAIR_BALANCE=1 timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/61_pytest_pair_balanced.log
for bal in 0 1 auto; do
  echo "== BDG_PAIR_BALANCE=$bal"
  if [ $bal = auto ]; then unset BDG_PAIR_BALANCE; else export BDG_PAIR_BALANCE=$bal; fi
  QP_STEPS=400 timeout 300 python profiles/quickperf2.py C2:256:t2 C3:512:t2 C3:4096:t2 C5:8:t2 C5:64:t2 C5_dwave:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee gpurun_out/r02/61_quickperf_work_lists.log
unset BDG_PAIR_BALANCE
BDG_PAIR_BALANCE=1 BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/61_racecheck_small_balanced.log 2>&1; echo "racecheck balanced rc=$?"; tail -2 gpurun_out/r02/61_racecheck_small_balanced.log
BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/61_racecheck_small.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02/61_racecheck_small.log
