#!/bin/bash
# round 2, call M (1 GPU): source-level ncu capture of the resident kernel (stall reasons per SASS line) on c3-small.
TAG=${1:-r2m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:als_cg_resident -s 2 -c 1 -f -o $OUT/prof_resident \
    python bench.py --workload c3-small --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_resident.log 2>&1
tail -3 $OUT/prof_resident.log | cut -c1-300
ls -la $OUT
