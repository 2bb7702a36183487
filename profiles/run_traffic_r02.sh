# ncu DRAM bytes per launch for the bench's remaining "format accounting" entries (profiles/traffic.json).
set -x
mkdir -p gpurun_out/r02
cap() {  # cap <tag> <kernel regex> <config> <k> <kernel>
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:$2 -s 1 -c 1 --csv --log-file gpurun_out/r02/74_traffic_$1.csv python profiles/prof_target.py $3 $4 $5 6 2>&1 | tail -1
}
cap C2_k256_t2 cheb_pair C2 256 t2
cap C3_k512_t2 cheb_pair C3 512 t2
cap C3_k4096_t2 cheb_pair C3 4096 t2
cap C5_k64_t2 cheb_pair C5 64 t2
cap C5_periodic_k8_t2 cheb_pair C5_periodic 8 t2
cap C5_random_k8_ell cheb_step_ell C5_random 8 ell
grep -h "cheb" gpurun_out/r02/74_traffic_*.csv | cut -c1-40,200-
