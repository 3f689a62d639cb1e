#!/bin/bash
# round 2, call R (1 GPU): source-level ncu capture of the tile kernel in its final launch shape on a rank-256 slice (c5-small).
TAG=${1:-r2r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:als_cg_tile -s 2 -c 1 -f -o $OUT/prof_tile_c5 \
    python bench.py --workload c5-small --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_tile_c5.log 2>&1
tail -2 $OUT/prof_tile_c5.log | cut -c1-300
echo "== bench c5-slice"; timeout 300 python bench.py --workload c5-slice --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c5slice.json | cut -c1-1200
ls -la $OUT
