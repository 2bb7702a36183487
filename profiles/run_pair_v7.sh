set -x
timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -3
python profiles/quickperf2.py C5:8:pair,dict_diag,pair C5:64:pair C2:256:pair 2>&1 | tail -5 | cut -c1-140
QP_STEPS=3000 python profiles/quickperf2.py C5:8:pair 2>&1 | tail -1 | cut -c1-140
