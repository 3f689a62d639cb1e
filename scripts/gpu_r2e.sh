#!/bin/bash
# round 2, call E (1 GPU): full GPU suite; Gram modes (3xTF32 / bf16 / FFMA) at rank 256; robustness points (Zipf columns,
# log-normal row lengths); item half-iteration; compute-sanitizer memcheck + racecheck on the small golden cases.
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== gram modes"; timeout 600 python scripts/gram_modes.py 2>&1 | tail -1 | tee $OUT/gram_modes_c5small.json | cut -c1-1200
for WL in "c5-slice" "c3-zipf" "c3-ragged" "c3 --half items" "c3-small --half items"; do
  NAME=$(echo $WL | tr -d ' -')
  echo "== bench $WL"; timeout 400 python bench.py --workload $WL --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$NAME.json | cut -c1-1300
done
echo "== compute-sanitizer memcheck (small golden cases through every CG / Cholesky kernel)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
   -k "tile_cg_kernel_in_session or cg_kernel_variants or tiled_cholesky or f32_engine_vs_reference_golden" 2>&1 | tail -6 | tee $OUT/sanitizer_memcheck.txt
echo "== compute-sanitizer racecheck"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
   -k "tile_cg_kernel_in_session or tiled_cholesky" 2>&1 | tail -6 | tee $OUT/sanitizer_racecheck.txt
ls -la $OUT
