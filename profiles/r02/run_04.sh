# round 2, call 4: variants of the two-step kernel -- base / E_{j-1} two planes ahead (pv3) / split-phase barrier with staged [B] (split) / both
set -x
mkdir -p gpurun_out/r02
for lib in libbdg.so libbdg_pv3.so libbdg_split.so libbdg_pv3_split.so; do
  echo "== $lib" | tee -a gpurun_out/r02/04_quickperf.log
  BDG_LIB=$PWD/bodge_b200/$lib timeout 300 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/r02/04_quickperf.log
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2 C5_bilayer:8:t2 C2:256:t2 2>&1 | tee -a gpurun_out/r02/04_quickperf.log
done
BDG_LIB=$PWD/bodge_b200/libbdg_pv3_split.so ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/04_t2_c5k8_pv3_split python profiles/prof_target.py C5 8 t2 8 2>&1 | tail -1
export BDG_CACHE_MB=0
BDG_LIB=$PWD/bodge_b200/libbdg_pv3_split.so timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/04_racecheck_small_pv3_split.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02/04_racecheck_small_pv3_split.log
