# round 2, call 2: the reworked two-step kernel (torus geometry, on-site blocks streamed per row for large dictionaries,
# post-barrier staging, [B] specialised on live sites) -- correctness first, then timing of two build variants, ncu, racecheck.
set -x
mkdir -p gpurun_out/r02
( time timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r02/02_pytest_pair.log
( time timeout 600 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r02/02_pytest_cheb_full.log
for lib in libbdg.so libbdg_nospec.so; do
  echo "== $lib"
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2,pair,dict_diag C5_bilayer:8:t2,pair C2:256:t2 C5:64:t2 2>&1 | tee -a gpurun_out/r02/02_quickperf.log
done
echo "== SELF forced off / on at the junction"
BDG_PAIR_SELF=1 QP_STEPS=400 python profiles/quickperf2.py C5:8:t2 2>&1 | tee -a gpurun_out/r02/02_quickperf.log
BDG_PAIR_SELF=0 QP_STEPS=400 python profiles/quickperf2.py C5_bilayer:8:t2 2>&1 | tee -a gpurun_out/r02/02_quickperf.log
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/02_t2_c5k8 python profiles/prof_target.py C5 8 t2 8 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/02_t2_c5dis_k8 python profiles/prof_target.py C5_disordered 8 t2 8 2>&1 | tail -1
export BDG_CACHE_MB=0
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/02_racecheck_small.log 2>&1; echo "racecheck small rc=$?"; tail -4 gpurun_out/r02/02_racecheck_small.log
BDG_PAIR_SEG=1 BDG_PAIR_P=3 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/02_racecheck_small_seg1_p3.log 2>&1; echo "racecheck small seg1 p3 rc=$?"; tail -4 gpurun_out/r02/02_racecheck_small_seg1_p3.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/02_memcheck_small.log 2>&1; echo "memcheck small rc=$?"; tail -3 gpurun_out/r02/02_memcheck_small.log
