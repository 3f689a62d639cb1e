#!/bin/bash
# usage: gpurun --timeout 150 -- 'bash scripts/gpu_chol_tc.sh <tag>'  -- first light of the tcgen05 per-row Gram (kernel = 6)
TAG=${1:-choltc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
CHOL_KERNEL=6 timeout 50 python scripts/check_chol_rows.py 2>&1 | tail -8 | tee $OUT/check_chol_tc.txt
timeout 50 python bench.py --workload c3-chol --kernel 6 --steps 3 2>&1 | tail -1 | tee $OUT/bench_c3-chol_k6.json | cut -c1-300
