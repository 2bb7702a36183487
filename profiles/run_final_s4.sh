set -x
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 2000 --warmup 20 > gpurun_out/bench_s4_final.json 2> gpurun_out/bench_s4_final.err; tail -2 gpurun_out/bench_s4_final.err; cut -c1-400 gpurun_out/bench_s4_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_s4_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_s4_ref.json
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/s4_pair_final_c5k8 python profiles/prof_target.py C5 8 pair 8 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s4_launches_bench_c5k8.csv python bench.py --steps 100 --warmup 6 --no-cpu-baseline --no-plain > gpurun_out/s4_launches_bench.out 2>&1; tail -1 gpurun_out/s4_launches_bench.out | cut -c1-200
QP_STEPS=100 python profiles/quickperf2.py C5:8:pair,dict_diag,ell C5:64:pair,dict_diag C4:8:auto C4:64:auto C2:256:pair,dict_diag C3:512:auto,pair 2>&1 | grep cfg > gpurun_out/s4_quickperf_burst.log
QP_STEPS=3000 python profiles/quickperf2.py C5:8:pair C5:64:pair C2:256:pair C3:512:auto 2>&1 | grep cfg > gpurun_out/s4_quickperf_sustained.log
cat gpurun_out/s4_quickperf_burst.log gpurun_out/s4_quickperf_sustained.log | cut -c1-160
