#!/bin/bash
# round 2, call H (1 GPU): full GPU suite (Gram-rows kernel, new tile classes); item half / ragged / C5 slice / rank 64 with the
# new defaults; Gram-rows kernel A/B (off; taking over from 105 / 209 entries).
TAG=${1:-r2h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
run() {  # name, env, args
  echo "== $1"; env $2 timeout 400 python bench.py $3 --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 > $OUT/bench_$1.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$1.json")); r=d["roofline"]
    print("   ", round(d["ms_per_step"],2), "ms solve", round(d["step_breakdown_ms"]["solve_ms"],2), "frac", round(r["frac"] or 0,4), {k:v for k,v in (r.get("rows_by_kernel") or {}).items() if v}, "loss", d["config"]["loss"])
except Exception as e:
    print("   ERR", open("$OUT/bench_$1.json").read()[-300:])
PY
}
run ragged_small_gramoff "B200ALS_GRAM_ROWS=0" "--workload c3-ragged-small"
run ragged_small_gram209 "X=1" "--workload c3-ragged-small"
run ragged_small_gram105 "B200ALS_GRAM_ROWS_MIN=105" "--workload c3-ragged-small"
run ragged_small_gram81 "B200ALS_GRAM_ROWS_MIN=81" "--workload c3-ragged-small"
run c3_items_gram "X=1" "--workload c3 --half items"
run c3_items_gram_1persm "B200ALS_GRAM_ROWS_PER_SM=1" "--workload c3 --half items"
run c3_ragged "X=1" "--workload c3-ragged"
run c3_k10 "X=1" "--workload c3 --kernel 10"
run c5slice "X=1" "--workload c5-slice"
run c5slice_wm8 "B200ALS_TILE_WARPS_M=8" "--workload c5-slice"
run ragged_small_wm8 "B200ALS_TILE_WARPS_M=8" "--workload c3-ragged-small"
run c3k64 "X=1" "--workload c3-k64"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:als_cg_gram -s 1 -c 1 -f -o $OUT/prof_gram_rows \
    python bench.py --workload c3-small --half items --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_gram_rows.log 2>&1
ls $OUT
