# round 2, call 66: HEAD with the balanced plans: whole GPU suite, smoke, sanitizers, default bench, driver-style bench
set -x
mkdir -p gpurun_out/r02
( time timeout 1700 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -4 ) 2>&1 | tee gpurun_out/r02/66_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02/66_smoke.log
python bench.py > gpurun_out/r02/66_bench.json 2> gpurun_out/r02/66_bench.err; cut -c1-260 gpurun_out/r02/66_bench.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02/66_bench_driver_args.json 2> gpurun_out/r02/66_bench_driver_args.err; cut -c1-260 gpurun_out/r02/66_bench_driver_args.json
export BDG_CACHE_MB=0
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/66_racecheck_small.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02/66_racecheck_small.log
BDG_PAIR_SEG=1 BDG_PAIR_P=3 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/66_racecheck_small_seg1_p3.log 2>&1; echo "racecheck seg1 p3 rc=$?"; tail -2 gpurun_out/r02/66_racecheck_small_seg1_p3.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/66_memcheck_small.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02/66_memcheck_small.log
