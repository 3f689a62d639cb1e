#!/bin/bash
# round 2, call J (1 GPU): Gram-rows kernel v2 (256 threads, symmetric operands, cp.async ring): parity, item half, ragged point, ncu.
TAG=${1:-r2j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest (long rows, golden, tile)"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "long_rows or golden or tile_cg or variants" 2>&1 | tail -8 | tee $OUT/pytest_sel.txt
run() {  # name, env, args
  echo "== $1"; env $2 timeout 400 python bench.py $3 --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 > $OUT/bench_$1.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$1.json")); r=d["roofline"]
    print("   ", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["step_breakdown_ms"].items()}, "frac", round(r["frac"] or 0,4), {k:v for k,v in (r.get("rows_by_kernel") or {}).items() if v}, "loss", d["config"]["loss"])
except Exception as e:
    print("   ERR", open("$OUT/bench_$1.json").read()[-300:])
PY
}
run c3_items "X=1" "--workload c3 --half items"
run c3_items_1persm "B200ALS_GRAM_ROWS_PER_SM=1" "--workload c3 --half items"
run c3_ragged "X=1" "--workload c3-ragged"
run ragged_small_gram105 "B200ALS_GRAM_ROWS_MIN=105" "--workload c3-ragged-small"
run ragged_small "X=1" "--workload c3-ragged-small"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:als_cg_gram -s 1 -c 1 -f -o $OUT/prof_gram_rows \
    python bench.py --workload c3-items-small --half items --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_gram_rows.log 2>&1
ls $OUT
