# round 2, call 29: ncu of cube kernel v3 at C4
set -x
mkdir -p gpurun_out/r02
BDG_CUBE_SHAPE=0 ncu --set full --clock-control none --import-source on -k regex:cheb_cube -s 1 -c 1 -f -o gpurun_out/r02/29_cube_c4k8_v3 python profiles/prof_target.py C4 8 t2 8 2>&1 | tail -1
