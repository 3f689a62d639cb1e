#!/bin/bash
# usage: gpurun --timeout 300 -- 'bash scripts/gpu_ncu_chol_rows.sh <tag>'   (one GPU; ncu replays the kernel ~40 times)
TAG=${1:-ncu_cholrows}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for WL in c2 c3-chol; do
  timeout 140 ncu --set full --clock-control none --import-source on -k regex:als_chol_rows -s 2 -c 1 -f -o $OUT/prof_rows_$WL \
      python bench.py --workload $WL --kernel 4 --steps 1 --warmup 3 > $OUT/prof_rows_$WL.log 2>&1
  tail -2 $OUT/prof_rows_$WL.log | cut -c1-300
done
ls -la $OUT
