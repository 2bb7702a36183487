# round 2, call 65: balanced plan for the three-dimensional kernel: tests in both plans, racecheck, C4 timing
set -x
mkdir -p gpurun_out/r02
( timeout 900 python -m pytest tests/test_gpu_cube.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/65_pytest_cube.log
( BDG_CUBE_BALANCE=1 timeout 900 python -m pytest tests/test_gpu_cube.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/65_pytest_cube_balanced.log
for bal in 0 1 auto; do
  echo "== BDG_CUBE_BALANCE=$bal"
  if [ $bal = auto ]; then unset BDG_CUBE_BALANCE; else export BDG_CUBE_BALANCE=$bal; fi
  QP_STEPS=400 timeout 300 python profiles/quickperf2.py C4:8:t2 C4:64:t2 2>&1 | cut -c1-120
done 2>&1 | tee gpurun_out/r02/65_quickperf_c4_balanced.log
unset BDG_CUBE_BALANCE
BDG_CUBE_BALANCE=1 BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_cube.py > gpurun_out/r02/65_racecheck_cube_balanced.log 2>&1; echo "racecheck balanced rc=$?"; tail -2 gpurun_out/r02/65_racecheck_cube_balanced.log
BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_cube.py > gpurun_out/r02/65_racecheck_cube.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02/65_racecheck_cube.log
