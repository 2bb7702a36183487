# round 2, call 69: group-balanced plans (sibling CTAs march the same (panel, x) chunk side by side): tests in forced-balanced mode, racecheck, timing
set -x
mkdir -p gpurun_out/r02
( BDG_PAIR_BALANCE=1 timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/69_pytest_pair_balanced.log
( BDG_CUBE_BALANCE=1 timeout 900 python -m pytest tests/test_gpu_cube.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/69_pytest_cube_balanced.log
for bal in 0 auto; do
  echo "== balance=$bal"
  if [ $bal = auto ]; then unset BDG_PAIR_BALANCE BDG_CUBE_BALANCE; else export BDG_PAIR_BALANCE=$bal BDG_CUBE_BALANCE=$bal; fi
  QP_STEPS=400 timeout 300 python profiles/quickperf2.py C2:256:t2 C3:512:t2 C3:4096:t2 C5:8:t2 C5:64:t2 C4:8:t2 C4:64:t2 2>&1 | cut -c1-120
done 2>&1 | tee gpurun_out/r02/69_quickperf_grouped.log
unset BDG_PAIR_BALANCE BDG_CUBE_BALANCE
BDG_PAIR_BALANCE=1 BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/69_racecheck_small_balanced.log 2>&1; echo "racecheck pair balanced rc=$?"; tail -2 gpurun_out/r02/69_racecheck_small_balanced.log
BDG_CUBE_BALANCE=1 BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_cube.py > gpurun_out/r02/69_racecheck_cube_balanced.log 2>&1; echo "racecheck cube balanced rc=$?"; tail -2 gpurun_out/r02/69_racecheck_cube_balanced.log
