#!/bin/bash
# round 2, call G (1 GPU): full GPU suite; A/B of the tile kernel's switches (clusters off / >= 4 / all; single-buffered
# small classes) on the ragged robustness point, the item half, C3 forced onto the tile kernel and rank 64.
# (B200ALS_TILE_SINGLE was the experiment switch of this call; its outcome is now the default class layout, engine_solve.inl)
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
run() {  # name, env, args
  echo "== $1"; env $2 timeout 300 python bench.py $3 --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 > $OUT/bench_$1.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$1.json")); r=d["roofline"]
    print("   ", round(d["ms_per_step"],2), "ms solve", round(d["step_breakdown_ms"]["solve_ms"],2), "frac", round(r["frac"] or 0,4), {k:v for k,v in (r.get("rows_by_kernel") or {}).items() if v})
except Exception as e:
    print("   ERR", open("$OUT/bench_$1.json").read()[-300:])
PY
}
for S in 0 1; do
  run ragged_small_cluoff_s$S "B200ALS_TILE_CLUSTER=1 B200ALS_TILE_SINGLE=$S" "--workload c3-ragged-small"
  run ragged_small_clu4_s$S "B200ALS_TILE_CLUSTER_MIN=4 B200ALS_TILE_SINGLE=$S" "--workload c3-ragged-small"
  run ragged_small_cluall_s$S "B200ALS_TILE_SINGLE=$S" "--workload c3-ragged-small"
  run c3small_k10_s$S "B200ALS_TILE_SINGLE=$S" "--workload c3-small --kernel 10"
  run c3small_items_s$S "B200ALS_TILE_SINGLE=$S" "--workload c3-small --half items"
done
run c3k64_s1 "B200ALS_TILE_SINGLE=1" "--workload c3-k64"
run c3k64_s0 "B200ALS_TILE_SINGLE=0" "--workload c3-k64"
ls $OUT
