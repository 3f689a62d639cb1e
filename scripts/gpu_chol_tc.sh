#!/bin/bash
# usage: gpurun --timeout 200 -- 'bash scripts/gpu_chol_tc.sh <tag>'  -- the tcgen05 per-row Gram variants of the row-per-thread
# Cholesky kernel (kernel = 6 single-buffered, 7 pipelined) the split-row variant (kernel = 8) and the rank-64 warp-per-system variant (kernel = 9): parity vs the fp64 oracle, then the rank-128 bench line of each
TAG=${1:-choltc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for KN in ${KERNELS:-6 7 8 9}; do
  echo "== kernel $KN"
  CHOL_KERNEL=$KN timeout 50 python scripts/check_chol_rows.py 2>&1 | tail -8 | tee $OUT/check_chol_tc_k$KN.txt
  WL=c3-chol; [ "$KN" = "9" ] && WL=c2      # kernel 9 is the rank-64 variant
  timeout 50 python bench.py --workload $WL --kernel $KN --steps 3 2>&1 | tail -1 | tee $OUT/bench_${WL}_k$KN.json | cut -c1-300
done
if [ "$NCU" = "1" ]; then
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:als_chol_rows -s 2 -c 1 -f -o $OUT/prof_chol_tc \
      python bench.py --workload c3-chol --kernel ${NCU_KERNEL:-6} --steps 1 --warmup 3 > $OUT/prof_chol_tc.log 2>&1
fi
