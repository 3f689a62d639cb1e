#!/bin/bash
# round 2, call I (1 GPU): full GPU suite + smoke; the default C3 bench line (with e2e and the CPU baseline); item half / ragged /
# C5 slice on the final defaults; ncu launch list of the bench command and a full capture of the Gram-rows kernel.
TAG=${1:-r2i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench c3 (default line)"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench_c3.json | cut -c1-2500
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-600
run() {  # name, env, args
  echo "== $1"; env $2 timeout 400 python bench.py $3 --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 > $OUT/bench_$1.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$1.json")); r=d["roofline"]
    print("   ", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["step_breakdown_ms"].items()}, "frac", round(r["frac"] or 0,4), {k:v for k,v in (r.get("rows_by_kernel") or {}).items() if v})
except Exception as e:
    print("   ERR", open("$OUT/bench_$1.json").read()[-300:])
PY
}
run c3_items "X=1" "--workload c3 --half items"
run c3_ragged "X=1" "--workload c3-ragged"
run c5slice "X=1" "--workload c5-slice"
run c4 "X=1" "--workload c4"
run c2 "X=1" "--workload c2"
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_bench.log 2>&1
echo "== ncu dram bytes of the resident kernel at C3 size"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:als_cg_resident -s 3 -c 1 --csv --log-file $OUT/resident_dram_c3.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/resident_dram_c3.log 2>&1
echo "== ncu --set full of the Gram-rows kernel (200k item rows of ~800 entries)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:als_cg_gram -s 1 -c 1 -f -o $OUT/prof_gram_rows \
    python bench.py --workload c3-items-small --half items --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_gram_rows.log 2>&1
ls $OUT
