set -x
mkdir -p gpurun_out/r02
python profiles/fixed_cost.py 1000 2>&1 | tee gpurun_out/r02/11_fixed_cost.log
