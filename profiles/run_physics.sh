set -x
mkdir -p gpurun_out/r02
nvidia-smi -L
(time timeout 400 python -m pytest tests/test_gpu_physics.py -q --durations=8 2>&1 | tail -40) 2>&1 | tee gpurun_out/r02/72_pytest_physics.log
