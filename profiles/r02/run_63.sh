# round 2, call 63: A/B on one box: HEAD's library vs the work-list tree, classic plan (C5 headline), burst and sustained
set -x
mkdir -p gpurun_out/r02
for rep in 1 2; do for lib in libbdg_head.so libbdg.so; do
  echo "== $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 timeout 300 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2 2>&1 | cut -c1-120
done; done 2>&1 | tee gpurun_out/r02/63_head_vs_lists.log
for lib in libbdg_head.so libbdg.so; do
  echo "== sustained $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=8000 timeout 300 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee -a gpurun_out/r02/63_head_vs_lists.log
