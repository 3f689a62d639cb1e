#!/bin/bash
# usage: gpurun --timeout 300 -- 'bash scripts/gpu_final.sh <tag>'  -- end-of-session verification of the committed state
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== rows kernel, 3-CTA build of rank 128"; CHOL_CTAS=3 timeout 60 python scripts/check_chol_rows.py 2>&1 | tail -7 | tee $OUT/check_chol_rows_ctas3.txt
echo "== bench c2"; timeout 60 python bench.py --workload c2 --steps 5 2>&1 | tail -1 | tee $OUT/bench_c2.json | cut -c1-400
for CT in 0 3; do
  echo "== bench c3-chol ctas=$CT"; timeout 60 python bench.py --workload c3-chol --ctas $CT --steps 4 2>&1 | tail -1 | tee $OUT/bench_c3-chol_ctas$CT.json | cut -c1-400
done
echo "== bench c3 (device-resident only)"; timeout 90 python bench.py --steps 5 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c3_device.json | cut -c1-600
