#!/bin/bash
# usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh <tag> N'
TAG=${1:-m}
N=${2:-2}
LIST=${3:-"1 2 4 8"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L | tee $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== correctness: sharded half-iteration vs oracle (world=$N)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_worker.py 2>&1 | tail -12 | tee $OUT/multigpu_check.txt
for G in $LIST; do
  if [ $G -le $N ]; then
    echo "== bench --gpus $G"
    if [ $G -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 --no-cpu $BENCH_ARGS 2>&1 | tail -1 | tee $OUT/bench_n$G.json | cut -c1-1500
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29620+G)) bench.py --gpus $G --no-cpu $BENCH_ARGS 2>&1 | tail -1 | tee $OUT/bench_n$G.json | cut -c1-1500
    fi
  fi
done
ls -la $OUT
