# round 2, call 5: dictionary numbered in slot order (sequential on-site table), stencil codes two planes ahead,
# incremental updates (patch in place), dict packer, accuracy warnings -- whole GPU suite, timings, sweep cost, ncu (disordered)
set -x
mkdir -p gpurun_out/r02
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) 2>&1 | tee gpurun_out/r02/05_pytest.log
QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair,dict_diag C5_disordered:8:t2,pair,dict_diag C5_bilayer:8:t2,pair,dict_diag C2:256:t2 C3:512:auto_moments,pair C4:8:auto_moments 2>&1 | tee gpurun_out/r02/05_quickperf.log
python profiles/sweep_update.py C5 2>&1 | tee gpurun_out/r02/05_sweep_update.log
python profiles/sweep_update.py C5_disordered 2>&1 | tee -a gpurun_out/r02/05_sweep_update.log
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/05_t2_c5dis_k8 python profiles/prof_target.py C5_disordered 8 t2 8 2>&1 | tail -1
