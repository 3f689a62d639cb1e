#!/bin/bash
# round 2, call A: the Cholesky variants written at the end of round 1 (kernel = 6 / 7 tcgen05 Gram, 8 split rows, 9 warp per
# system): parity vs the fp64 oracle + bench lines, then `ncu --set full` of kernels 6 and 7 (rank 128) and 4 (default, C2).
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvsmi.txt 2>&1
for KN in 4 6 7 8 9; do
  echo "== kernel $KN"
  CHOL_KERNEL=$KN timeout 60 python scripts/check_chol_rows.py 2>&1 | tail -8 | tee $OUT/check_chol_k$KN.txt
  WL=c3-chol; [ "$KN" = "9" ] && WL=c2      # kernel 9 is the rank-64 variant
  timeout 60 python bench.py --workload $WL --kernel $KN --steps 3 2>&1 | tail -1 | tee $OUT/bench_${WL}_k$KN.json | cut -c1-300
done
timeout 60 python bench.py --workload c2 --kernel 4 --steps 3 2>&1 | tail -1 | tee $OUT/bench_c2_k4.json | cut -c1-300
for KN in 6 7; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:als_chol_rows -s 2 -c 1 -f -o $OUT/prof_chol_k$KN \
      python bench.py --workload c3-chol --kernel $KN --steps 1 --warmup 3 > $OUT/prof_chol_k$KN.log 2>&1
done
ls -la $OUT
