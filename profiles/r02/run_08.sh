# round 2, call 8: release candidate -- whole GPU suite, smoke, bench, launch list of the bench, racecheck / memcheck of the small cases
set -x
mkdir -p gpurun_out/r02
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r02/08_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02/08_smoke.log
( time python bench.py > gpurun_out/r02/08_bench.json 2> gpurun_out/r02/08_bench.err ); tail -3 gpurun_out/r02/08_bench.err; cut -c1-300 gpurun_out/r02/08_bench.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02/08_bench_ref.json 2> gpurun_out/r02/08_bench_ref.err; cut -c1-300 gpurun_out/r02/08_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02/08_launches_bench.csv python bench.py --steps 100 --warmup 6 --no-cpu-baseline --no-plain --no-extras --no-e2e > gpurun_out/r02/08_launches_bench.out 2>&1; tail -1 gpurun_out/r02/08_launches_bench.out | cut -c1-200
export BDG_CACHE_MB=0
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/08_racecheck_small.log 2>&1; echo "racecheck small rc=$?"; tail -3 gpurun_out/r02/08_racecheck_small.log
BDG_PAIR_SEG=1 BDG_PAIR_P=3 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/08_racecheck_small_seg1_p3.log 2>&1; echo "racecheck seg1 p3 rc=$?"; tail -3 gpurun_out/r02/08_racecheck_small_seg1_p3.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_incremental.py -x -q > gpurun_out/r02/08_memcheck_incremental.log 2>&1; echo "memcheck incremental rc=$?"; tail -3 gpurun_out/r02/08_memcheck_incremental.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/08_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/r02/08_memcheck_smoke.log
