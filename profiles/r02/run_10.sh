# round 2, call 10: ablation -- the 8-warp kernel without the shared-memory reads of the warp's own records (wrong numbers, timing only):
# upper bound of what keeping them elsewhere (registers, tensor memory) could buy at 16 warps per SM
set -x
mkdir -p gpurun_out/r02
for lib in libbdg.so libbdg_ablate.so; do
  echo "== $lib" | tee -a gpurun_out/r02/10_ablate_own_lds.log
  BDG_LIB=$PWD/bodge_b200/$lib QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2 2>&1 | tee -a gpurun_out/r02/10_ablate_own_lds.log
done
