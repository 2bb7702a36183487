set -x
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/pair_c5k8_v1 python profiles/prof_target.py C5 8 pair 8 2>&1 | tail -3
for w in 16 8; do for seg in 0 40 125 250; do
  echo "== WARPS=$w SEG=$seg"; BDG_PAIR_WARPS=$w BDG_PAIR_SEG=$seg python profiles/quickperf2.py C5:8:pair 2>&1 | tail -1
done; done
for p in 14 22 30; do echo "== P=$p"; BDG_PAIR_P=$p python profiles/quickperf2.py C5:8:pair 2>&1 | tail -1; done
