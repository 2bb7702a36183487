# round 2, call 64: work lists, final form (classic loop identical to HEAD's): pair tests in both plans, racecheck both, A/B against HEAD, timing of the small-lattice configs
set -x
mkdir -p gpurun_out/r02
( timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/64_pytest_pair.log
( BDG_PAIR_BALANCE=1 timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | grep -v Warning | tail -3 ) | tee gpurun_out/r02/64_pytest_pair_balanced.log
for rep in 1 2; do for lib in libbdg_head.so libbdg.so; do
  echo "== $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 timeout 300 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-120
done; done 2>&1 | tee gpurun_out/r02/64_head_vs_lists.log
for lib in libbdg_head.so libbdg.so; do
  echo "== $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 timeout 300 python profiles/quickperf2.py C2:256:t2 C3:512:t2 C3:4096:t2 C5:64:t2 C5_disordered:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee -a gpurun_out/r02/64_head_vs_lists.log
BDG_PAIR_BALANCE=1 BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/64_racecheck_small_balanced.log 2>&1; echo "racecheck balanced rc=$?"; tail -2 gpurun_out/r02/64_racecheck_small_balanced.log
BDG_CACHE_MB=0 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/64_racecheck_small.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02/64_racecheck_small.log
