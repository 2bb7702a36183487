# round 2, call 68: what a piece really costs beyond its planes: C2 (100 x 100, 32 panels) and C3 (64 panels) over segment lengths, classic plan
set -x
mkdir -p gpurun_out/r02
for seg in 100 50 34 25 20 13; do
  echo "== BDG_PAIR_SEG=$seg"
  BDG_PAIR_BALANCE=0 BDG_PAIR_SEG=$seg QP_STEPS=400 timeout 300 python profiles/quickperf2.py C2:256:t2 C3:512:t2 2>&1 | cut -c1-120
done 2>&1 | tee gpurun_out/r02/68_piece_cost.log
