#!/bin/bash
# round 2, call K (2 GPUs): multi-GPU changes of the last commits (one all-reduce for loss / regulariser / status, fixed matrix
# shared over NVLink in the stateless calls): worker in both exchange modes, C3 at N = 2 with e2e, reference arm under torchrun.
TAG=${1:-r2k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for EX in p2p nccl; do
  echo "== multigpu worker, exchange=$EX"
  B200ALS_EXCHANGE=$EX timeout 600 $TR --master-port 29611 tests/multigpu_worker.py 2>&1 | tail -8 | cut -c1-300 | tee $OUT/multigpu_check_$EX.txt
done
echo "== bench c3 --gpus 2"; timeout 600 $TR --master-port 29631 bench.py --gpus 2 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_n2_c3.json | cut -c1-2600
echo "== reference arm under torchrun (N = 2)"; timeout 300 $TR --master-port 29641 bench.py --gpus 2 --impl reference --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_n2_reference.json | cut -c1-700
echo "== bench c3 --gpus 1 (same box)"; timeout 600 python bench.py --no-cpu 2>&1 | tail -1 | tee $OUT/bench_n1_c3.json | cut -c1-1500
ls $OUT
