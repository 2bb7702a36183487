# round 2, call 38: cube kernel, 8 warps x 254 registers, own-site E_j records carried, two bodies interleaved by hand (BDG_CUBE_SHAPE=3)
set -x
mkdir -p gpurun_out/r02
BDG_CUBE_SHAPE=3 timeout 600 python -m pytest tests/test_gpu_cube.py -x -q -k "oracle or single_step" 2>&1 | tail -2 | tee gpurun_out/r02/38_pytest_cube_ilp2.log
for shape in 0 3; do
  echo "== BDG_CUBE_SHAPE=$shape"
  BDG_CUBE_SHAPE=$shape QP_STEPS=400 timeout 300 python profiles/quickperf2.py C4:8:t2 C4:64:t2 2>&1 | cut -c1-130
done 2>&1 | tee gpurun_out/r02/38_quickperf_c4_cube_ilp2.log
