# round 2, call 15: final state -- whole GPU suite, smoke, both bench arms as the driver runs them, the default bench
set -x
mkdir -p gpurun_out/r02
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r02/15_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02/15_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02/15_bench_ref_driver_args.json 2> gpurun_out/r02/15_bench_ref.err; cut -c1-200 gpurun_out/r02/15_bench_ref_driver_args.json
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02/15_bench_driver_args.json 2> gpurun_out/r02/15_bench_driver_args.err ); tail -2 gpurun_out/r02/15_bench_driver_args.err; cut -c1-300 gpurun_out/r02/15_bench_driver_args.json
( time python bench.py > gpurun_out/r02/15_bench.json 2> gpurun_out/r02/15_bench.err ); tail -2 gpurun_out/r02/15_bench.err; cut -c1-300 gpurun_out/r02/15_bench.json
