set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 10 --no-cpu-baseline > gpurun_out/bench_s4_n2.json 2> gpurun_out/bench_s4_n2.err; tail -2 gpurun_out/bench_s4_n2.err; cut -c1-330 gpurun_out/bench_s4_n2.json
