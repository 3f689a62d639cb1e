#!/bin/bash
# round 2, call Q (2 GPUs): final validation of the resident kernel's half-warp layout -- the whole -m gpu suite (multi-GPU tests
# included), the default bench line at N = 1 (value + e2e + CPU baseline) and at N = 2, smoke().
TAG=${1:-r2q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (N = 1, default line)"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench_n1_c3.json | cut -c1-1800
echo "== bench c3 --gpus 2"; timeout 600 $TR --master-port 29631 bench.py --gpus 2 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_n2_c3.json | cut -c1-600
ls $OUT
