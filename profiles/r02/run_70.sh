# round 2, call 70: planner penalties per kernel: C4 timing; then the whole validation on this tree
set -x
mkdir -p gpurun_out/r02
QP_STEPS=400 timeout 300 python profiles/quickperf2.py C4:8:t2 C4:64:t2 C2:256:t2 C3:512:t2 2>&1 | cut -c1-120 | tee gpurun_out/r02/70_quickperf.log
( time timeout 1700 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -4 ) 2>&1 | tee gpurun_out/r02/70_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02/70_smoke.log
python bench.py > gpurun_out/r02/70_bench.json 2> gpurun_out/r02/70_bench.err; cut -c1-200 gpurun_out/r02/70_bench.json
export BDG_CACHE_MB=0
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/70_racecheck_small.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02/70_racecheck_small.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_cube.py > gpurun_out/r02/70_racecheck_cube.log 2>&1; echo "racecheck cube rc=$?"; tail -2 gpurun_out/r02/70_racecheck_cube.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/70_memcheck_small.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02/70_memcheck_small.log
