#!/bin/bash
# round 2, final scaling record: the default bench line (value + e2e) and the reference arm at N GPUs, as the driver runs them.
# usage: gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_r2_scale.sh r2scale N'
TAG=${1:-r2scale}
N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== reference arm, N=$N"; timeout 300 $TR --master-port 29701 bench.py --gpus $N --impl reference --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_n${N}_reference.json | cut -c1-400
echo "== bench, N=$N"; timeout 600 $TR --master-port 29702 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_n${N}.json | cut -c1-300
python - <<PY
import json
d=json.load(open("$OUT/bench_n${N}.json"))
print(round(d["value"]/1e6,1), "M/s", round(d["ms_per_step"],2), "ms", d["step_breakdown_ms"], "e2e", d["e2e"])
PY
ls $OUT
