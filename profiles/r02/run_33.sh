# round 2, call 33: C4 sustained (12000 steps under the power cap): three-dimensional even-vector kernel vs the single-step kernel
set -x
mkdir -p gpurun_out/r02
QP_STEPS=12000 timeout 600 python profiles/quickperf2.py C4:8:t2 C4:8:dict_diag C4:8:t2 C4:8:dict_diag 2>&1 | cut -c1-230 | tee gpurun_out/r02/33_c4_sustained.log
QP_STEPS=3000 timeout 600 python profiles/quickperf2.py C4:64:t2 C4:64:dict_diag 2>&1 | cut -c1-230 | tee -a gpurun_out/r02/33_c4_sustained.log
