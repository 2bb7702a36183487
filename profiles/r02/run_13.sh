# round 2, call 13: ablation -- on-site product without FP64 MMAs (wrong numbers, timing only), burst and sustained clocks
set -x
mkdir -p gpurun_out/r02
for lib in libbdg.so libbdg_ablate.so; do
  echo "== $lib burst (400 steps)" | tee -a gpurun_out/r02/13_ablate_dmma.log
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair 2>&1 | tee -a gpurun_out/r02/13_ablate_dmma.log
  echo "== $lib sustained (12000 steps)" | tee -a gpurun_out/r02/13_ablate_dmma.log
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=12000 python profiles/quickperf2.py C5:8:t2 2>&1 | tee -a gpurun_out/r02/13_ablate_dmma.log
done
