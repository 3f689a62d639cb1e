#!/bin/bash
# round 2, call W (1 GPU): ncu --set full of the resident kernel at the FULL C3 size (10 M rows, item matrix 512 MB >> L2):
# the 1 M-row slice used so far keeps the item matrix in L2 and under-reports the memory stalls.
TAG=${1:-r2w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_cg_resident -s 2 -c 1 -f -o $OUT/prof_resident_c3 \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_resident_c3.log 2>&1
tail -2 $OUT/prof_resident_c3.log | cut -c1-200
ls -la $OUT
