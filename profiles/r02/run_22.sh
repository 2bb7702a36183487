# round 2, call 22: final tree -- whole GPU suite, smoke, default bench, driver-style bench, racecheck of the small cases, ncu of the final kernel
set -x
mkdir -p gpurun_out/r02
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r02/22_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02/22_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02/22_bench_driver_args.json 2> gpurun_out/r02/22_bench_driver_args.err; cut -c1-260 gpurun_out/r02/22_bench_driver_args.json
python bench.py > gpurun_out/r02/22_bench.json 2> gpurun_out/r02/22_bench.err; cut -c1-260 gpurun_out/r02/22_bench.json
QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair,dict_diag,ell C5_disordered:8:t2 C5_bilayer:8:t2 C5_random:8:auto_moments C2:256:t2 C3:512:auto_moments C4:8:auto_moments C5:64:t2 2>&1 | tee gpurun_out/r02/22_quickperf_final.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/22_t2_c5k8_final python profiles/prof_target.py C5 8 t2 8 2>&1 | tail -1
export BDG_CACHE_MB=0
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/22_racecheck_small.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02/22_racecheck_small.log
BDG_PAIR_SEG=1 BDG_PAIR_P=3 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/r02/race_small.py > gpurun_out/r02/22_racecheck_small_seg1_p3.log 2>&1; echo "racecheck seg1 p3 rc=$?"; tail -2 gpurun_out/r02/22_racecheck_small_seg1_p3.log
