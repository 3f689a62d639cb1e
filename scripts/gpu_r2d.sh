#!/bin/bash
# round 2, call D (2 GPUs): multi-GPU worker (per-half tolerances, skewed shards, sharded transpose + fit with sharded item
# half-iterations) in both exchange modes; C5 slice on the tile kernel with 16 and 8 warps; the singular-rows test.
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for EX in p2p nccl; do
  echo "== multigpu worker, exchange=$EX"
  B200ALS_EXCHANGE=$EX timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_worker.py 2>&1 | tail -14 | tee $OUT/multigpu_check_$EX.txt
done
echo "== singular rows"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "singular or tile_cg" 2>&1 | tail -5 | tee $OUT/pytest_sel.txt
for WW in 16 8; do
  echo "== c5-slice, class-L warps $WW"; B200ALS_TILE_WARPS_L=$WW timeout 300 python bench.py --workload c5-slice --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c5slice_w$WW.json | cut -c1-200
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:als_cg_tile -s 2 -c 1 -f -o $OUT/prof_tile_c5 \
    python bench.py --workload c5-small --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_tile_c5.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"rotate_any|jacobi|gram_partial" -s 6 -c 3 -f -o $OUT/prof_prep_c5 \
    python bench.py --workload c5-small --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_prep_c5.log 2>&1
ls -la $OUT
