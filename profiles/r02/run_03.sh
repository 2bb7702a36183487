# round 2, call 3: after restoring the one-plane-ahead E_{j-1} buffer -- pair tests, timing of both [B] variants, ncu
set -x
mkdir -p gpurun_out/r02
( time timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r02/03_pytest_pair.log
for lib in libbdg.so libbdg_spec.so; do
  echo "== $lib" | tee -a gpurun_out/r02/03_quickperf.log
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2,pair C5_bilayer:8:t2,pair C2:256:t2 2>&1 | tee -a gpurun_out/r02/03_quickperf.log
done
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/03_t2_c5k8 python profiles/prof_target.py C5 8 t2 8 2>&1 | tail -1
