# round 2, call 19: CTA shapes under the sustained power cap (12 000 steps) -- less halo recompute (16 warps) / fewer shared-memory reads (12 warps, REG) draw less power
set -x
mkdir -p gpurun_out/r02
for w in 8 12 16; do
  echo "== BDG_PAIR_WARPS=$w sustained" | tee -a gpurun_out/r02/19_shapes_sustained.log
  BDG_PAIR_WARPS=$w QP_STEPS=12000 python profiles/quickperf2.py C5:8:t2 2>&1 | cut -c1-220 | tee -a gpurun_out/r02/19_shapes_sustained.log
done
