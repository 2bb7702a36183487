# round 2, call 16: whole GPU suite on the final tree
set -x
mkdir -p gpurun_out/r02
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r02/16_pytest.log
