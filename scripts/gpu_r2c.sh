#!/bin/bash
# round 2, call C: full GPU test suite on the new defaults (tile CG kernel v2, tcgen05-Gram Cholesky at rank 128, warp-per-system
# at rank 64), bench lines and ncu of the tile kernel.
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
for WL in "c3 --kernel 10" "c3-k64" "c5-slice" "c2" "c3-chol"; do
  NAME=$(echo $WL | tr -d ' -')
  echo "== bench $WL"; timeout 300 python bench.py --workload $WL --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$NAME.json | cut -c1-300
done
for WL in "c5-small" "c3-small --kernel 10"; do
  NAME=$(echo $WL | tr -d ' -')
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:als_cg_tile -s 2 -c 1 -f -o $OUT/prof_tile_$NAME \
      python bench.py --workload $WL --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_tile_$NAME.log 2>&1
done
ls -la $OUT
