# round 2, call 30: cube kernel v3b (as v3, E_{j-1} loaded at the top of the iteration again): tests, C4 timing per shape
set -x
mkdir -p gpurun_out/r02
( time timeout 900 python -m pytest tests/test_gpu_cube.py -x -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r02/30_pytest_cube.log
for shape in 0 1 2; do
  echo "== BDG_CUBE_SHAPE=$shape"
  BDG_CUBE_SHAPE=$shape QP_STEPS=400 timeout 300 python profiles/quickperf2.py C4:8:t2 C4:64:t2 2>&1 | cut -c1-230
done 2>&1 | tee gpurun_out/r02/30_quickperf_c4_cube.log
