# round 2, call 54: lazy hand-over in the two-step kernel (arrive at the end of an iteration, wait after [A] of the next): tests, A/B timing, racecheck
set -x
mkdir -p gpurun_out/r02
export L=$PWD/bodge_b200/libbdg_lazy.so
( BDG_LIB=$L timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/54_pytest_pair_lazy.log
for rep in 1 2; do for lib in libbdg.so libbdg_lazy.so; do
  echo "== $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 timeout 300 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2 C3:512:t2 C2:256:t2 2>&1 | cut -c1-120
done; done 2>&1 | tee gpurun_out/r02/54_lazy_handover.log
for lib in libbdg.so libbdg_lazy.so; do
  echo "== sustained $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=8000 timeout 300 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee -a gpurun_out/r02/54_lazy_handover.log
BDG_LIB=$L BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/54_racecheck_small_lazy.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02/54_racecheck_small_lazy.log
