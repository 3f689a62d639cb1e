#!/bin/bash
# round 2, call S (1 GPU): tile kernel with the warp-per-row copy loop -- parity tests, then the workloads that run on it.
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest (parity + bias)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bias.py -q -m gpu -x 2>&1 | tail -4 | tee $OUT/pytest_parity.txt
for WL in "c5-slice" "c3-ragged" "c3 --kernel 10" "c3-k64" "c3 --half items"; do
  NAME=$(echo $WL | tr -d ' -')
  echo "== bench $WL"; timeout 300 python bench.py --workload $WL --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$NAME.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', round(d['ms_per_step'],2),'ms', d.get('step_breakdown_ms'), 'frac', d['roofline']['frac'])"
done
ls $OUT
echo "== A/B: c5-slice with 8 warps in the 2-CTA class"; B200ALS_TILE_WARPS_M=8 timeout 300 python bench.py --workload c5-slice --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c5slice_w8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', round(d['ms_per_step'],2),'ms', d.get('step_breakdown_ms'), 'frac', d['roofline']['frac'])"
