# round 2, call 31: cube kernel v3b, ablations (wrong numbers): abl1 = no read of the newest E_j plane in [A], abl2 = no read of u(xa) in [B], abl3 = both
set -x
mkdir -p gpurun_out/r02
for lib in libbdg.so libbdg_abl1.so libbdg_abl2.so libbdg_abl3.so; do
  echo "== $lib"
  BDG_LIB=$PWD/bodge_b200/$lib BDG_CUBE_SHAPE=0 QP_STEPS=400 timeout 300 python profiles/quickperf2.py C4:8:t2 2>&1 | cut -c1-130
done 2>&1 | tee gpurun_out/r02/31_cube_v3b_ablate.log
