# round 2, call 56: lazy hand-over with a guard record behind every ring slot: racecheck (three plans), tests, timing
set -x
mkdir -p gpurun_out/r02
export L=$PWD/bodge_b200/libbdg_lazy.so
BDG_LIB=$L BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/56_racecheck_small_lazy.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02/56_racecheck_small_lazy.log
BDG_PAIR_SEG=1 BDG_PAIR_P=3 BDG_LIB=$L BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/56_racecheck_small_lazy_seg1_p3.log 2>&1; echo "racecheck seg1 p3 rc=$?"; tail -2 gpurun_out/r02/56_racecheck_small_lazy_seg1_p3.log
( BDG_LIB=$L timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/56_pytest_pair_lazy.log
for lib in libbdg.so libbdg_lazy.so; do
  echo "== $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 timeout 300 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2 C2:256:t2 2>&1 | cut -c1-120
done 2>&1 | tee gpurun_out/r02/56_lazy_handover.log
