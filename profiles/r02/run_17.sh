# round 2, call 17: knobs of the single-step dictionary kernel on the 3-D lattice (C4), where no two-step kernel applies
set -x
mkdir -p gpurun_out/r02
L=gpurun_out/r02/17_sweep_c4.log
for pz in 1 2 4; do for pf in 0 1 2 3; do
  echo "== PZ=$pz PREFETCH=$pf" | tee -a $L
  BDG_ELL_PZ=$pz BDG_ELL_PREFETCH=$pf QP_STEPS=400 python profiles/quickperf2.py C4:8:dict_diag C4:64:dict_diag 2>&1 | cut -c1-150 | tee -a $L
done; done
for seg in 8 16 32 64; do
  echo "== SEG=$seg" | tee -a $L
  BDG_ELL_SEG=$seg QP_STEPS=400 python profiles/quickperf2.py C4:8:dict_diag C4:64:dict_diag 2>&1 | cut -c1-150 | tee -a $L
done
for np in 1 2 4; do
  echo "== NP=$np" | tee -a $L
  BDG_ELL_NP=$np QP_STEPS=400 python profiles/quickperf2.py C4:64:dict_diag 2>&1 | cut -c1-150 | tee -a $L
done
echo "== walk off" | tee -a $L
BDG_ELL_WALK=0 QP_STEPS=400 python profiles/quickperf2.py C4:8:dict_diag C4:64:dict_diag 2>&1 | cut -c1-150 | tee -a $L
