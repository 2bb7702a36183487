timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -8
for xd in 1 0; do echo "== XDIAG=$xd"; BDG_PAIR_XDIAG=$xd python profiles/quickperf2.py C5:8:t2 C5:64:t2 C2:256:t2 2>&1 | grep cfg | cut -c1-120; BDG_PAIR_XDIAG=$xd QP_STEPS=3000 python profiles/quickperf2.py C5:8:t2 2>&1 | grep cfg | cut -c1-120; done
