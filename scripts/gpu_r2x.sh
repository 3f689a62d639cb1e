#!/bin/bash
# round 2, call X (1 GPU): compute-sanitizer on the kernels changed at the end of the round (resident kernel's half-warp layout,
# tile kernel's warp-per-row copies and early issue of the next tile in single-buffer mode).
TAG=${1:-r2x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
SEL="tile_cg_kernel_in_session or cg_kernel_variants"
echo "== memcheck"
timeout 230 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "$SEL" 2>&1 | tail -6 | tee $OUT/sanitizer_memcheck_final.txt
echo "== racecheck"
timeout 260 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "$SEL" > $OUT/racecheck_full.txt 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Race reported|hazard" $OUT/racecheck_full.txt | sort | uniq -c | sort -rn | head -12 | tee $OUT/sanitizer_racecheck_final.txt
tail -3 $OUT/racecheck_full.txt | cut -c1-200
