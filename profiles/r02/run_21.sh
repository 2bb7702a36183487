# round 2, call 21: wait for the newest T_n plane only before the last term of the row (BDG_PAIR_LATEWAIT)
set -x
mkdir -p gpurun_out/r02
for lib in libbdg.so libbdg_latewait.so; do
  echo "== $lib" | tee -a gpurun_out/r02/21_latewait.log
  BDG_LIB=$PWD/bodge_b200/$lib timeout 300 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -1 | tee -a gpurun_out/r02/21_latewait.log
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2 C2:256:t2 2>&1 | cut -c1-200 | tee -a gpurun_out/r02/21_latewait.log
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=12000 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-200 | tee -a gpurun_out/r02/21_latewait.log
done
