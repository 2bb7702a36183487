# round 2, call 36: ncu of the single-step dictionary kernel at C4 (where do its 27 % under the HBM peak go?)
set -x
mkdir -p gpurun_out/r02
ncu --set full --clock-control none --import-source on -k regex:cheb_step_ell -s 2 -c 1 -f -o gpurun_out/r02/36_dictdiag_c4k8 python profiles/prof_target.py C4 8 dict_diag 8 2>&1 | tail -1
