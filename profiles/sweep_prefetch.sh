#!/bin/bash
# A/B of the L1 prefetch distance (x-steps ahead) of the marching ELL / DICT kernels.
for d in 0 1 2 3 5; do
  echo "PREFETCH=$d"
  BDG_ELL_PREFETCH=$d python profiles/quickperf2.py C5:8:dict,ell C5:4:dict C4:8:dict,ell C2:8:dict 2>&1 | cut -c1-175
done
