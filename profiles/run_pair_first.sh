set -x
timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -30
timeout 600 python profiles/quickperf2.py C5:8:pair,dict_diag C5:64:pair,dict_diag C3:512:pair,dict C2:256:pair,dict_diag 2>&1 | tail -12
