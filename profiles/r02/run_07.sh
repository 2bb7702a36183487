# round 2, call 7: register-resident own records (REG, 12 warps, one CTA per SM) vs the 8-warp shape, plain and split-phase barrier
set -x
mkdir -p gpurun_out/r02
for lib in libbdg.so libbdg_split.so; do
  for warps in 8 12; do
    echo "== $lib warps=$warps" | tee -a gpurun_out/r02/07_quickperf.log
    BDG_PAIR_WARPS=$warps BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2 C5_bilayer:8:t2 C2:256:t2 C5:64:t2 2>&1 | tee -a gpurun_out/r02/07_quickperf.log
  done
done
BDG_PAIR_WARPS=12 timeout 600 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02/07_pytest_pair_w12.log
BDG_PAIR_WARPS=12 BDG_LIB=$PWD/bodge_b200/libbdg_split.so timeout 600 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02/07_pytest_pair_w12_split.log
BDG_PAIR_WARPS=12 ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/07_t2_c5k8_w12 python profiles/prof_target.py C5 8 t2 8 2>&1 | tail -1
BDG_PAIR_WARPS=12 BDG_LIB=$PWD/bodge_b200/libbdg_split.so ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/07_t2_c5k8_w12_split python profiles/prof_target.py C5 8 t2 8 2>&1 | tail -1
