# round 2, call 48: C5 headline with / without the rim-owning patches, alternating (is the 3 % of call 47 noise?); sustained too
set -x
mkdir -p gpurun_out/r02
for rep in 1 2 3; do for open in 0 1; do
  echo "== BDG_PAIR_OPEN=$open"
  BDG_PAIR_OPEN=$open QP_STEPS=400 timeout 300 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-120
done; done 2>&1 | tee gpurun_out/r02/48_c5_open_rim_ab.log
for open in 0 1; do
  echo "== sustained BDG_PAIR_OPEN=$open"
  BDG_PAIR_OPEN=$open QP_STEPS=8000 timeout 300 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee -a gpurun_out/r02/48_c5_open_rim_ab.log
