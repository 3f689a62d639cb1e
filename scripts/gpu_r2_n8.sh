#!/bin/bash
# round 2, 8-GPU call: parity at world = 8, then the BASELINE configs that name 8 GPUs -- C3 (user and item half), C4, C5
# (3xTF32 and bf16 Gram) -- plus the p2p / NCCL A/B of the exchange at N = 8.
# usage: gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_r2_n8.sh r2n8'
TAG=${1:-r2n8}
N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== parity world=$N"
timeout 600 $TR --master-port 29611 tests/multigpu_worker.py 2>&1 | tail -9 | cut -c1-400 | tee $OUT/multigpu_check_n$N.txt
run() {  # name, env, args
  echo "== $1"; env $2 timeout 600 $TR --master-port $((29620 + RANDOM % 200)) bench.py --gpus $N $3 2>&1 | tail -1 > $OUT/bench_n${N}_$1.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n${N}_$1.json")); r=d["roofline"]
    print("   ", round(d["ms_per_step"],2), "ms", round(d["value"]/1e6,1), "M/s", {k:round(v,2) for k,v in d["step_breakdown_ms"].items()}, "frac", round(r["frac"] or 0,4), "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("   ERR", open("$OUT/bench_n${N}_$1.json").read()[-400:])
PY
}
run c3 "X=1" "--workload c3 --steps 10 --no-cpu"
run c3_nccl "B200ALS_EXCHANGE=nccl" "--workload c3 --steps 5 --no-e2e --no-cpu"
run c3_items "X=1" "--workload c3 --half items --steps 5 --no-e2e --no-cpu"
run c4 "X=1" "--workload c4 --steps 5 --no-e2e --no-cpu"
run c5 "X=1" "--workload c5 --steps 3 --no-e2e --no-cpu"
run c5_bf16 "X=1" "--workload c5 --gram bf16 --steps 3 --no-e2e --no-cpu"
ls -la $OUT
