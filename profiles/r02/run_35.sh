# round 2, call 35: cube kernel (final form): affected tests, sanitizers, bench with extras
set -x
mkdir -p gpurun_out/r02
( time timeout 1500 python -m pytest tests/test_gpu_cube.py tests/test_gpu_pair.py tests/test_gpu_incremental.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r02/35_pytest.log
export BDG_CACHE_MB=0
for plan in "0 0" "1 1" "2 2"; do
  set -- $plan
  BDG_CUBE_SHAPE=$1 BDG_CUBE_SEG=$2 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_cube.py > gpurun_out/r02/35_racecheck_cube_shape$1_seg$2.log 2>&1; echo "racecheck shape $1 seg $2 rc=$?"; tail -2 gpurun_out/r02/35_racecheck_cube_shape$1_seg$2.log
done
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/r02/race_cube.py > gpurun_out/r02/35_memcheck_cube.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02/35_memcheck_cube.log
unset BDG_CACHE_MB
python bench.py > gpurun_out/r02/35_bench.json 2> gpurun_out/r02/35_bench.err; cut -c1-300 gpurun_out/r02/35_bench.json
