# round 2, call 58: lazy hand-over with one arrival / one poller per warp: sustained and burst A/B against the barrier build, racecheck, pair tests
set -x
mkdir -p gpurun_out/r02
for lib in libbdg_barrier.so libbdg.so libbdg_barrier.so libbdg.so; do
  echo "== sustained $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=8000 timeout 300 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee gpurun_out/r02/58_lazy_warp_ab.log
for lib in libbdg_barrier.so libbdg.so; do
  echo "== burst $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 timeout 300 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2 C2:256:t2 C3:512:t2 2>&1 | cut -c1-120
done 2>&1 | tee -a gpurun_out/r02/58_lazy_warp_ab.log
BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/58_racecheck_small.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02/58_racecheck_small.log
( timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/58_pytest_pair.log
