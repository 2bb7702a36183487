set -x
timeout 900 python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -15
for w in 8; do
  echo "== WARPS=$w"; BDG_PAIR_WARPS=$w python profiles/quickperf2.py C5:8:pair 2>&1 | tail -1
done
python profiles/quickperf2.py C5:64:pair C3:512:pair C2:256:pair 2>&1 | tail -3
QP_STEPS=3000 python profiles/quickperf2.py C5:8:pair 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/pair_c5k8_v5 python profiles/prof_target.py C5 8 pair 8 2>&1 | tail -2
