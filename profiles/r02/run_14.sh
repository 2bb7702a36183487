# round 2, call 14: the DFMA on-site product (XD) -- whole GPU suite, burst + sustained timing with and without it, ncu
set -x
mkdir -p gpurun_out/r02
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r02/14_pytest.log
for xd in 1 0; do
  echo "== BDG_PAIR_XDIAG=$xd burst (400 steps)" | tee -a gpurun_out/r02/14_quickperf_xdiag.log
  BDG_PAIR_XDIAG=$xd QP_STEPS=400 python profiles/quickperf2.py C5:8:t2,pair C5_disordered:8:t2 C5_bilayer:8:t2 C2:256:t2 2>&1 | tee -a gpurun_out/r02/14_quickperf_xdiag.log
  echo "== BDG_PAIR_XDIAG=$xd sustained (12000 steps)" | tee -a gpurun_out/r02/14_quickperf_xdiag.log
  BDG_PAIR_XDIAG=$xd QP_STEPS=12000 python profiles/quickperf2.py C5:8:t2 C5_disordered:8:t2 2>&1 | tee -a gpurun_out/r02/14_quickperf_xdiag.log
done
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/14_t2_c5k8_xd python profiles/prof_target.py C5 8 t2 8 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:cheb_pair -s 1 -c 1 -f -o gpurun_out/r02/14_t2_c5dis_k8_xd python profiles/prof_target.py C5_disordered 8 t2 8 2>&1 | tail -1
