# round 2, call 27: cube kernel, what fewer shared-memory reads per body would buy (ablation builds: wrong numbers)
set -x
mkdir -p gpurun_out/r02
for lib in libbdg.so libbdg_abl1.so libbdg_abl2.so; do
  echo "== $lib (abl1: [A] reads 5 instead of 7 records per body; abl2: and [B] 4 instead of 6)"
  BDG_LIB=$PWD/bodge_b200/$lib BDG_CUBE_SHAPE=0 QP_STEPS=400 timeout 300 python profiles/quickperf2.py C4:8:t2 C4:64:t2 2>&1 | cut -c1-230
done 2>&1 | tee gpurun_out/r02/27_cube_ablate_lds.log
