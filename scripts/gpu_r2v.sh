#!/bin/bash
# round 2, call V (1 GPU): ncu launch list of the bench command on the final kernels (per-launch durations, cold-cache and
# serialised: the kernel's SHARE of a step is what must agree with bench.py's step_breakdown_ms), C3 and the C5 slice.
TAG=${1:-r2v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_bench_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench_c5slice.csv \
    python bench.py --workload c5-slice --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_bench_c5slice.log 2>&1
ls -la $OUT
