# round 2, call 50: CTA shapes of the two-step kernel on MMA rows (C3) and on C2 -- 8 warps x 2 CTAs (default), 12 (REG), 16
set -x
mkdir -p gpurun_out/r02
for w in 8 12 16; do
  echo "== BDG_PAIR_WARPS=$w"
  BDG_PAIR_WARPS=$w QP_STEPS=400 timeout 300 python profiles/quickperf2.py C3:512:t2 C3:4096:t2 C2:256:t2 C5_dwave:8:t2 2>&1 | cut -c1-120
done 2>&1 | tee gpurun_out/r02/50_shapes_mma_rows.log
