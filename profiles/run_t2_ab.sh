for w in 8 16; do echo "== WARPS=$w"; BDG_PAIR_WARPS=$w python profiles/quickperf2.py C5:8:t2 C5:64:t2 C2:256:t2 2>&1 | grep cfg | cut -c1-120; BDG_PAIR_WARPS=$w QP_STEPS=3000 python profiles/quickperf2.py C5:8:t2 2>&1 | grep cfg | cut -c1-120; done
for seg in 125 84 63 50; do echo "== SEG=$seg"; BDG_PAIR_SEG=$seg python profiles/quickperf2.py C5:8:t2 2>&1 | grep cfg | cut -c1-120; done
