for sh in 82 161 121 62 83 162; do
  echo "== SHAPE=$sh"
  BDG_PAIR_SHAPE=$sh timeout 600 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -1
  BDG_PAIR_SHAPE=$sh python profiles/quickperf2.py C5:8:pair C5:64:pair C2:256:pair 2>&1 | grep cfg | cut -c1-125
  BDG_PAIR_SHAPE=$sh QP_STEPS=3000 python profiles/quickperf2.py C5:8:pair 2>&1 | grep cfg | cut -c1-125
done
