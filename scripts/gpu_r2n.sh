#!/bin/bash
# round 2, call N (1 GPU): resident kernel with the shortened scalar chains -- parity tests that touch it, then the C3 bench line.
TAG=${1:-r2n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest (parity + bias + abi on gpu)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bias.py -q -m gpu -x 2>&1 | tail -6 | tee $OUT/pytest_parity.txt
for WL in "c3" "c3 --kernel 2" "c4" "c3-ragged"; do
  NAME=$(echo $WL | tr -d ' -')
  echo "== bench $WL"; timeout 300 python bench.py --workload $WL --steps 5 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$NAME.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', round(d['ms_per_step'],2),'ms', d.get('step_breakdown_ms'), 'frac', d['roofline']['frac'])"
done
ls $OUT
