#!/bin/bash
# usage: gpurun --timeout 240 -- 'bash scripts/gpu_chol_rows.sh <tag>'
TAG=${1:-cholrows}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== parity of the row-per-thread Cholesky kernel"
timeout 120 python scripts/check_chol_rows.py 2>&1 | tail -9 | tee $OUT/check_chol_rows.txt
for WL in c2 c3-chol; do
  for KN in ${KERNELS:-4 0}; do
    echo "== bench $WL kernel=$KN"
    timeout 100 python bench.py --workload $WL --kernel $KN --steps 4 2>&1 | tail -1 | tee $OUT/bench_${WL}_k$KN.json | cut -c1-700
  done
done
