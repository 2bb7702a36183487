for pa in 0 16 32 64 48 80 112; do
  echo "== DEBUG_SKIP=$pa (16: no T_n+1 store, 32: no T_n-1 load, 64: no T_n+2 store)"; BDG_PAIR_PREV_AHEAD=$pa python profiles/quickperf2.py C5:8:pair 2>&1 | tail -1 | cut -c1-120
done
