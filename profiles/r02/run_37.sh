# round 2, call 37: cube kernel with the own-site E_j records carried in registers at 12 warps (168 registers) and 8 warps (241)
set -x
mkdir -p gpurun_out/r02
for shape in 3 4; do
  BDG_CUBE_SHAPE=$shape timeout 600 python -m pytest tests/test_gpu_cube.py -x -q -k "oracle or single_step" 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r02/37_pytest_cube_carry.log
for shape in 0 3 4; do
  echo "== BDG_CUBE_SHAPE=$shape"
  BDG_CUBE_SHAPE=$shape QP_STEPS=400 timeout 300 python profiles/quickperf2.py C4:8:t2 C4:64:t2 2>&1 | cut -c1-130
done 2>&1 | tee gpurun_out/r02/37_quickperf_c4_cube_carry.log
