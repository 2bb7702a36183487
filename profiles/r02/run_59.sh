# round 2, call 59: lazy hand-over, waiting threads parked by the hardware (try_wait with a suspend-time hint): sustained / burst A/B
set -x
mkdir -p gpurun_out/r02
for lib in libbdg.so libbdg_lazy.so libbdg.so libbdg_lazy.so; do
  echo "== sustained $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=8000 timeout 300 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee gpurun_out/r02/59_lazy_parked_ab.log
for lib in libbdg.so libbdg_lazy.so; do
  echo "== burst $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 timeout 300 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee -a gpurun_out/r02/59_lazy_parked_ab.log
