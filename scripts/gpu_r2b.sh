#!/bin/bash
# round 2, call B: first run of als_cg_tile_kernel (parity through the session and the stateless calls), the tcgen05 Cholesky
# variants with the grid forced to 3 CTAs/SM, and first bench lines of the tile kernel (C3 forced, C3 at rank 64, C5 slice).
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -25 | tee $OUT/pytest_parity.txt
for KN in 6 7; do
  timeout 60 python bench.py --workload c3-chol --kernel $KN --steps 3 2>&1 | tail -1 | tee $OUT/bench_c3-chol_k$KN.json | cut -c1-400
done
for WL in "c3 --kernel 10" "c3-k64" "c5-slice" "c5-small" "c3 --kernel 1"; do
  NAME=$(echo $WL | tr -d ' -')
  echo "== bench $WL"; timeout 300 python bench.py --workload $WL --steps 3 --no-e2e --no-cpu 2>&1 | tail -2 | tee $OUT/bench_$NAME.json | cut -c1-1500
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:als_cg_tile -s 2 -c 1 -f -o $OUT/prof_tile_c5 \
    python bench.py --workload c5-small --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_tile_c5.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:als_cg_tile -s 2 -c 1 -f -o $OUT/prof_tile_c3 \
    python bench.py --workload c3-small --kernel 10 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_tile_c3.log 2>&1
ls -la $OUT
