# round 2, call 24: ncu of the three-dimensional even-vector kernel at C4 (first version)
set -x
mkdir -p gpurun_out/r02
ncu --set full --clock-control none --import-source on -k regex:cheb_cube -s 1 -c 1 -f -o gpurun_out/r02/24_cube_c4k8_v1 python profiles/prof_target.py C4 8 t2 8 2>&1 | tail -1
