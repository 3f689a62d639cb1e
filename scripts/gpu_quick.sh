#!/bin/bash
# quick GPU round: tests + bench (+ optional extra command)
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== bench c3 default"; timeout 900 python bench.py 2>&1 | tail -2 | tee $OUT/bench_c3.json | cut -c1-3000
if [ -n "$2" ]; then echo "== extra: $2"; eval "$2" 2>&1 | tail -20 | tee $OUT/extra.txt; fi
