# round 2, call 34: cube kernel, E_{j-1} single-buffered (loaded at the top of the iteration that uses it) vs one iteration ahead
set -x
mkdir -p gpurun_out/r02
for lib in libbdg.so libbdg_pv1.so; do
  echo "== $lib"
  BDG_LIB=$PWD/bodge_b200/$lib BDG_CUBE_SHAPE=0 QP_STEPS=400 timeout 300 python profiles/quickperf2.py C4:8:t2 C4:64:t2 2>&1 | cut -c1-130
done 2>&1 | tee gpurun_out/r02/34_cube_pv_single.log
