# Final validation of the tree at the end of round 2: every GPU test, the smoke path, the bench at its defaults.
set -x
mkdir -p gpurun_out/r02
nvidia-smi -L
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) 2>&1 | tee gpurun_out/r02/75_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02/75_smoke.log
(time python bench.py > gpurun_out/r02/75_bench.json 2> gpurun_out/r02/75_bench.err) 2>&1 | tail -4
tail -3 gpurun_out/r02/75_bench.err
cut -c1-400 gpurun_out/r02/75_bench.json
