# round 2, call 52: final-tree evidence -- ncu launch list of the bench command, full capture of the headline kernel (C5 t2) and of the MMA-row kernel (C3 t2, SD)
set -x
mkdir -p gpurun_out/r02
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02/52_launches_bench.csv python bench.py --steps 100 --warmup 6 --no-cpu-baseline --no-plain --no-extras --no-e2e > gpurun_out/r02/52_launches_bench.out 2>&1; tail -1 gpurun_out/r02/52_launches_bench.out | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/52_t2_c5k8_final python profiles/prof_target.py C5 8 t2 8 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/52_t2_c3k512_sd python profiles/prof_target.py C3 512 t2 8 2>&1 | tail -1
