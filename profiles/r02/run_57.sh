# round 2, call 57: lazy hand-over as the default: whole GPU suite, smoke, sanitizers, sustained A/B against the barrier build, bench
set -x
mkdir -p gpurun_out/r02
( time timeout 1700 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -4 ) 2>&1 | tee gpurun_out/r02/57_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02/57_smoke.log
for lib in libbdg_barrier.so libbdg.so libbdg_barrier.so libbdg.so; do
  echo "== sustained $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=8000 timeout 300 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee gpurun_out/r02/57_lazy_sustained_ab.log
python bench.py > gpurun_out/r02/57_bench.json 2> gpurun_out/r02/57_bench.err; cut -c1-260 gpurun_out/r02/57_bench.json
export BDG_CACHE_MB=0
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/57_racecheck_small.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02/57_racecheck_small.log
BDG_PAIR_SEG=1 BDG_PAIR_P=3 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/57_racecheck_small_seg1_p3.log 2>&1; echo "racecheck seg1 p3 rc=$?"; tail -2 gpurun_out/r02/57_racecheck_small_seg1_p3.log
BDG_PAIR_WARPS=16 BDG_PAIR_SEG=2 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/57_racecheck_small_w16.log 2>&1; echo "racecheck w16 rc=$?"; tail -2 gpurun_out/r02/57_racecheck_small_w16.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/57_memcheck_small.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02/57_memcheck_small.log
