# round 2, call 1: state of HEAD~ on the box -- full GPU test suite incl. the new full-size oracle tests, the new bench line,
# first numbers for the less repetitive workloads, sanitizer runs of the smoke path and the two-step kernels.
set -x
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -c "import os; print('cpus', os.cpu_count())"; free -g | head -2
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r02/01_pytest.log
( time python bench.py > gpurun_out/r02/01_bench.json 2> gpurun_out/r02/01_bench.err ); tail -3 gpurun_out/r02/01_bench.err; cut -c1-600 gpurun_out/r02/01_bench.json
python profiles/quickperf2.py C5_disordered:8:auto_moments,pair,dict_diag,dict,ell C5_bilayer:8:auto_moments,pair,dict_diag,ell C5_random:8:auto_moments,dmma 2>&1 | tee gpurun_out/r02/01_quickperf_repetition.log
# sanitizers: memcheck + racecheck on the smoke path and on the two-step kernels with ragged patches / short segments
export BDG_CACHE_MB=0
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/01_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/r02/01_memcheck_smoke.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/01_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 gpurun_out/r02/01_racecheck_smoke.log
BDG_PAIR_SEG=1 BDG_PAIR_P=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_pair.py -x -q -k "moments_match_the_oracle" > gpurun_out/r02/01_racecheck_pair_seg1_p1.log 2>&1; echo "racecheck pair rc=$?"; tail -3 gpurun_out/r02/01_racecheck_pair_seg1_p1.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_pair.py -x -q -k "random_shapes or moments_match" > gpurun_out/r02/01_memcheck_pair.log 2>&1; echo "memcheck pair rc=$?"; tail -3 gpurun_out/r02/01_memcheck_pair.log
