#!/bin/bash
# usage: gpurun --gpus 2 --timeout 300 -- 'bash scripts/gpu_p2p.sh <tag>'
# The peer-memory exchange on 2 GPUs: parity of the sharded half-iteration vs the oracle with the pushes FORCED
# (B200ALS_EXCHANGE=p2p turns a silent NCCL fallback into an error), then the 2-GPU bench line with either exchange.
TAG=${1:-p2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== parity, exchange forced to p2p"
B200ALS_EXCHANGE=p2p timeout 200 $TR --master-port 29611 tests/multigpu_worker.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM_THREADS" | tail -8 | tee $OUT/multigpu_check_p2p.txt
echo "== bench --gpus 2, default exchange"
timeout 150 $TR --master-port 29622 bench.py --gpus 2 --steps 8 --no-cpu --no-e2e 2>&1 | tail -1 | tee $OUT/bench_n2_p2p.json | cut -c1-1200
echo "== bench --gpus 2, NCCL exchange"
B200ALS_EXCHANGE=nccl timeout 150 $TR --master-port 29623 bench.py --gpus 2 --steps 8 --no-cpu --no-e2e 2>&1 | tail -1 | tee $OUT/bench_n2_nccl.json | cut -c1-1200
echo "== parity, exchange forced to nccl"
B200ALS_EXCHANGE=nccl timeout 200 $TR --master-port 29612 tests/multigpu_worker.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM_THREADS" | tail -8 | tee $OUT/multigpu_check_nccl.txt
