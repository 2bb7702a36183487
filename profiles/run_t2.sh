set -x
timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -15
python profiles/quickperf2.py C5:8:t2,pair,dict_diag C5:64:t2 C2:256:t2 C3:512:t2 2>&1 | grep cfg | cut -c1-150
QP_STEPS=3000 python profiles/quickperf2.py C5:8:t2 2>&1 | grep cfg | cut -c1-150
