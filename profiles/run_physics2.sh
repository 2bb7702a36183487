set -x
mkdir -p gpurun_out/r02
(timeout 100 python -m pytest tests/test_gpu_physics.py tests/test_gpu_incremental.py::test_dict_api_sweep_stays_incremental "tests/test_gpu_cheb.py::test_free_energy_matches_reference" -q 2>&1 | tail -25) 2>&1 | tee gpurun_out/r02/78_pytest_physics.log
