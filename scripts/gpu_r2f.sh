#!/bin/bash
# round 2, call F (2 GPUs): full GPU suite (cluster tile kernel, biased session, warm-started eigenbasis, Gram blocks);
# item half and ragged robustness point with clusters on / off; racecheck detail of one tile-kernel and one resident case;
# 2-GPU: multigpu worker, C3 items half, C4, C5 slice.
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
for CLU in 8 1; do
  for WL in "c3-small --half items" "c3-ragged-small"; do
    NAME=$(echo $WL | tr -d ' -')
    echo "== bench $WL cluster<=$CLU"; B200ALS_TILE_CLUSTER=$CLU timeout 300 python bench.py --workload $WL --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_${NAME}_clu$CLU.json | cut -c1-1500
  done
done
echo "== bench c3 --half items"; timeout 400 python bench.py --workload c3 --half items --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c3_items.json | cut -c1-1500
echo "== bench c3-ragged"; timeout 400 python bench.py --workload c3-ragged --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c3ragged.json | cut -c1-1500
echo "== bench c3 --fit-iters 3"; timeout 600 python bench.py --workload c3 --steps 3 --no-e2e --no-cpu --fit-iters 3 2>&1 | tail -1 | tee $OUT/bench_c3_fit.json | cut -c1-2500
echo "== racecheck detail"
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
   -k "test_tile_cg_kernel_in_session and synth_ragged_implicit_cg_k128 and 10" > $OUT/racecheck_tile_detail.txt 2>&1
tail -5 $OUT/racecheck_tile_detail.txt
echo "== 2 GPUs"
B200ALS_EXCHANGE=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_worker.py 2>&1 | tail -8 | tee $OUT/multigpu_check_p2p.txt
for WL in "c3 --half items" "c4" "c5-slice"; do
  NAME=$(echo $WL | tr -d ' -')
  echo "== bench --gpus 2 $WL"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --workload $WL --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_n2_$NAME.json | cut -c1-1500
done
ls -la $OUT
