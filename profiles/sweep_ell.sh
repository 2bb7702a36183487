#!/bin/bash
# A/B of the ELL kernel's panels-per-group (NP) / panel batch (PB) on the many-column workloads.
export QP_KERNELS=ell
for combo in "1 1" "2 1" "2 2" "4 1" "4 2" "8 1"; do
  set -- $combo
  BDG_ELL_NP=$1 BDG_ELL_PB=$2 python profiles/quickperf.py C5:32 C5:64 C2:256 C3:512 C4:64 2>&1 | cut -c1-175
done
