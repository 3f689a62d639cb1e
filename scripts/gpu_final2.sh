#!/bin/bash
TAG=${1:-final2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== bench c3-chol"; timeout 60 python bench.py --workload c3-chol --steps 4 2>&1 | tail -1 | tee $OUT/bench_c3-chol.json | cut -c1-300
