set -x
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 2000 --warmup 20 > gpurun_out/bench_s4_t2.json 2> gpurun_out/bench_s4_t2.err; tail -2 gpurun_out/bench_s4_t2.err; cut -c1-300 gpurun_out/bench_s4_t2.json
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/s4_t2_final_c5k8 python profiles/prof_target.py C5 8 t2 8 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s4_launches_bench_c5k8_t2.csv python bench.py --steps 100 --warmup 6 --no-cpu-baseline --no-plain > gpurun_out/s4_launches_bench_t2.out 2>&1; tail -1 gpurun_out/s4_launches_bench_t2.out | cut -c1-200
