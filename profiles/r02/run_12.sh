set -x
mkdir -p gpurun_out/r02
python profiles/l2_resident.py 2>&1 | tee gpurun_out/r02/12_l2_resident.log
